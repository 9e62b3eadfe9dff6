"""GPU parity: jitter stream, traversal + ray/sphere test (rtds_trace), and the whole render path — against the
oracle port, the golden vectors of the unmodified reference, and (size-independent) cross-mode properties."""
import hashlib
import os

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
G = T.load_golden_json()
pytestmark = pytest.mark.gpu


def test_jitter_stream_bit_exact(gpu_ctx, oracle):
    g = G["jitter"]
    j = gpu_ctx.jitter_stream(0, 1000004)
    assert [float(x).hex() for x in j[:8]] == g["first8_hex"]
    assert [float(x).hex() for x in j[1000000:1000004]] == g["at_1000000_hex"]
    assert hashlib.sha256(j[:1000000].tobytes()).hexdigest() == g["sha256_first_1e6"]
    # a window far into the stream (4K x 4 spp needs 66.4M doubles), regenerated from snapshots
    first = 66_000_000
    assert np.array_equal(gpu_ctx.jitter_stream(first, 4096), oracle.jitter(4096, first))


def _rays(n, seed, sph):
    rng = np.random.default_rng(seed)
    o = np.zeros((n, 3), np.float32)
    tgt = sph[rng.integers(0, sph.shape[0], n), :3] + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.05)
    o[n // 2:] = rng.normal(size=(n - n // 2, 3)).astype(np.float32) * np.float32(20)   # second half: arbitrary origins
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


@pytest.mark.parametrize("bits", [30, 63])
def test_trace_lbvh_true_matches_oracle_and_none(gpu_ctx, oracle, bits):
    sph, mat = T.synthetic_scene(20000, 21)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE, morton_bits=bits)
    nodes, order = gpu_ctx.export_bvh()
    o, d = _rays(20000, 22, sph)
    h_o, t_o, cand = oracle.trace(sph, nodes, order, o, d, tie_by_objid=1)
    for exact in (True, False):
        h, t, st = gpu_ctx.trace(rt.LBVH, o, d, exact=exact)
        assert np.array_equal(h, h_o), f"hit ids differ (exact={exact}): {np.count_nonzero(h != h_o)}"
        assert t.tobytes() == t_o.tobytes()
        if exact:
            assert st["prim_tests"] == cand      # same candidate set as boxIntersect's collect-all traversal
    hn, tn, _ = gpu_ctx.trace(rt.NONE, o, d)
    hn_o, tn_o, _ = oracle.trace(sph, None, None, o, d)
    assert np.array_equal(hn, hn_o) and tn.tobytes() == tn_o.tobytes()
    assert (h_o >= 0).mean() > 0.3


def test_render_none_default_config_is_byte_exact(gpu_ctx):
    """dataStructure = NONE on the default config: the GPU frame equals the reference's PPM byte for byte."""
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    rgb, hit, _, st = gpu_ctx.render(rt.NONE, 640, 480, 1, want_hit=True)
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    assert np.array_equal(hit, gold["hit_none"])
    assert T.ppm_md5(rgb) == G["default_config"]["NONE"]["ppm_md5"]
    assert st["prim_tests"] == 640 * 480 * sph.shape[0]


@pytest.mark.parametrize("exact", [True, False])
def test_render_lbvh_true_bunny_vs_oracle(gpu_ctx, oracle, exact):
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    nodes, order = gpu_ctx.export_bvh()
    rgb, hit, accum, st = gpu_ctx.render(rt.LBVH, 640, 480, 1, want_hit=True, want_accum=True, exact=exact)
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, 640, 480, 1, tie_by_objid=1, want_accum=True)
    assert np.array_equal(hit, hit_o)
    assert accum.tobytes() == accum_o.tobytes(), "float RGB sums differ from the CPU restatement"
    assert np.array_equal(rgb, rgb_o)
    # any valid BVH over the same leaf boxes yields the candidate set of the reference BVH path:
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    diff = np.count_nonzero(hit != gold["hit_bvh"])
    assert diff <= 3, f"{diff} pixels differ from the reference BVH path's hit ids"
    assert st["primary_rays"] == 640 * 480


def test_render_multisample_and_ranks(gpu_ctx, oracle):
    """aa_samples = 4 accumulation order + interleaved scanline tiles: every rank's rows equal the 1-rank frame."""
    sph, mat = T.synthetic_scene(3000, 31)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    nodes, order = gpu_ctx.export_bvh()
    W, H, spp = 322, 203, 4          # ragged: not multiples of the 16x8 block or of tile_rows
    full, hit, accum, _ = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True)
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, tie_by_objid=1, want_accum=True)
    assert accum.tobytes() == accum_o.tobytes() and np.array_equal(full, rgb_o) and np.array_equal(hit, hit_o)
    for world in (2, 3, 8):
        frame = np.zeros((H, W, 3), np.uint8)
        rows_seen = 0
        for rank in range(world):
            part = np.zeros((H, W, 3), np.uint8)
            _, _, _, st = gpu_ctx.render(rt.LBVH, W, H, spp, out=part, rank=rank, world=world)
            rows = rt.owned_rows(H, 8, rank, world)
            assert st["rows"] == rows.size == rt.rows_for_rank(H, 8, rank, world)
            frame[rows] = part[rows]
            rows_seen += rows.size
        assert rows_seen == H and np.array_equal(frame, full)


def test_fused_strip_kernel_is_exact_too(gpu_ctx, oracle):
    """RTDS_STRIP=1: the fused render + in-block MT19937 kernel (a measured negative result for speed)
    must still produce the reference's frame: default config md5, ragged multi-sample multi-rank frames."""
    with T.option(gpu_ctx, "strip", 1):
        _strip_checks(gpu_ctx, oracle)


def _strip_checks(gpu_ctx, oracle):
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.BVH)
    rgb, hit, _, _ = gpu_ctx.render(rt.BVH, 640, 480, 1, want_hit=True)
    assert T.ppm_md5(rgb) == "c69c66375f2c6bda433f9f457a4b2b2e"
    nodes, order = gpu_ctx.export_bvh()
    W, H, spp = 322, 203, 5          # 5 does not divide the 1,248-sample chunk: pixels straddle chunk ends
    full, hit, accum, _ = gpu_ctx.render(rt.BVH, W, H, spp, want_hit=True, want_accum=True)
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, want_accum=True)
    assert accum.tobytes() == accum_o.tobytes() and np.array_equal(full, rgb_o) and np.array_equal(hit, hit_o)
    frame = np.zeros((H, W, 3), np.uint8)
    for rank in range(3):
        part = np.zeros((H, W, 3), np.uint8)
        gpu_ctx.render(rt.BVH, W, H, spp, out=part, rank=rank, world=3)
        rows = rt.owned_rows(H, 8, rank, 3)
        frame[rows] = part[rows]
    assert np.array_equal(frame, full)


def test_rtds_frame_equals_the_three_calls(gpu_ctx):
    sph, mat = T.bunny_scene()
    rgb, bst, rst = gpu_ctx.frame(sph, mat, rt.BVH, 640, 480, 1)
    assert T.ppm_md5(rgb) == "c69c66375f2c6bda433f9f457a4b2b2e" and bst["total_nodes"] == 71895
    sph2, mat2 = T.material_scene(1500, 11)
    rgb_a, _, _ = gpu_ctx.frame(sph2, mat2, rt.LBVH, 240, 180, 2, mode=rt.MODE_TRUE, shadows=1)
    gpu_ctx.set_spheres(sph2, mat2)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    rgb_b, _, _, _ = gpu_ctx.render(rt.LBVH, 240, 180, 2, shadows=1)
    assert np.array_equal(rgb_a, rgb_b)


@pytest.mark.parametrize("acc", ["BVH", "LBVH"])
def test_packet_traversal_equals_single_ray_and_reference_order(gpu_ctx, oracle, acc):
    """aa_samples % 4 == 0 takes the packet kernel (four samples of a pixel share every node load): hit ids, float sums
    and bytes must equal the unpruned reference traversal (exact=1), the single-ray ordered kernel (RTDS_PACKET=0) and
    the CPU restatement; 8 spp = two packets per pixel, image centre included (mixed-octant pixels fall back)."""
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    a = getattr(rt, acc)
    gpu_ctx.build(a, mode=rt.MODE_TRUE if acc == "LBVH" else rt.MODE_COMPAT)
    nodes, order = gpu_ctx.export_bvh()
    W, H, spp = 400, 300, 8
    pk = gpu_ctx.render(a, W, H, spp, want_hit=True, want_accum=True)     # default: interior boxes tested once per packet (hull)
    ex = gpu_ctx.render(a, W, H, spp, want_hit=True, want_accum=True, exact=True)
    with T.option(gpu_ctx, "hull", 0):                                      # interior boxes tested per ray
        pr = gpu_ctx.render(a, W, H, spp, want_hit=True, want_accum=True)
    with T.option(gpu_ctx, "packet", 0):
        sr = gpu_ctx.render(a, W, H, spp, want_hit=True, want_accum=True)
    assert pk[3]["node_tests"] < pr[3]["node_tests"] and pk[3]["node_visits"] <= 1.05 * pr[3]["node_visits"]
    for other in (ex, sr, pr):
        assert np.array_equal(pk[1], other[1]) and pk[2].tobytes() == other[2].tobytes() and np.array_equal(pk[0], other[0])
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, tie_by_objid=1 if acc == "LBVH" else 0, want_accum=True)
    assert np.array_equal(pk[1], hit_o) and pk[2].tobytes() == accum_o.tobytes() and np.array_equal(pk[0], rgb_o)
    assert pk[3]["primary_rays"] == W * H * spp == sr[3]["primary_rays"]
    assert pk[3]["node_visits"] < sr[3]["node_visits"]          # the point of the packet: fewer node loads


def test_bench_frame_full_size_packet_equals_reference_traversal(gpu_ctx):
    """BASELINE config 3 at FULL size (bunny x 30 clones = 1,078,411 spheres, 3840x2160, 4 spp, LBVH): the frame the bench
    times (packet kernel, ordered + pruned) is byte-identical to the unpruned reference traversal (exact=1) and has the
    same hit ids; ray count and rank partition properties hold."""
    v = np.fromfile(T.GOLDEN + "/bunny_vertices.f32", np.float32).reshape(-1, 3)
    sph, mat = rt.scene_from_vertices(v, 30)
    assert sph.shape[0] == 1078411
    gpu_ctx.set_spheres(sph, mat)
    st = gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    assert st["total_nodes"] == 2 * 1078411 - 1
    W, H, spp = 3840, 2160, 4
    fast, hit_f, _, st_f = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True)
    exact, hit_e, _, st_e = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, exact=True)
    assert st_f["primary_rays"] == st_e["primary_rays"] == W * H * spp
    assert np.array_equal(hit_f, hit_e) and np.array_equal(fast, exact)
    assert hashlib.sha256(fast.tobytes()).hexdigest() == hashlib.sha256(exact.tobytes()).hexdigest()
    assert (hit_f >= 0).mean() > 0.3                      # bunny + ground cover a good part of the frame
    # two ranks' tiles assemble to the same frame
    frame = np.zeros_like(fast)
    for rank in range(2):
        part = np.zeros_like(fast)
        gpu_ctx.render(rt.LBVH, W, H, spp, out=part, rank=rank, world=2)
        rows = rt.owned_rows(H, 8, rank, 2)
        frame[rows] = part[rows]
    assert np.array_equal(frame, fast)


def test_visible_clones_variant_packet_equals_reference_traversal(gpu_ctx, oracle):
    """SURVEY 8d's visible variant of config 3 (clone shift 2: the bunnies overlap in view, far more candidates per ray), reduced
    to 8 clones at 960x540x4spp: packet + hull traversal == unpruned reference traversal == CPU restatement (hit ids, sums, bytes)."""
    v = np.fromfile(T.GOLDEN + "/bunny_vertices.f32", np.float32).reshape(-1, 3)
    sph, mat = rt.scene_from_vertices(v, 8, clone_shift=2)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    nodes, order = gpu_ctx.export_bvh()
    W, H, spp = 960, 540, 4
    fast, hit_f, acc_f, st_f = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True)
    exact, hit_e, acc_e, st_e = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, exact=True)
    assert np.array_equal(hit_f, hit_e) and acc_f.tobytes() == acc_e.tobytes() and np.array_equal(fast, exact)
    y0, y1 = 200, 330                                   # the CPU restatement on the densest rows
    rgb_o, hit_o, acc_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, y0, y1, tie_by_objid=1, want_accum=True)
    assert np.array_equal(hit_f[y0:y1], hit_o) and acc_f[y0:y1].tobytes() == acc_o.tobytes() and np.array_equal(fast[y0:y1], rgb_o)
    assert (hit_f >= 0).mean() > 0.4
    sh, _, _, st_s = gpu_ctx.render(rt.LBVH, W, H, spp, shadows=1)
    sh_e, _, _, _ = gpu_ctx.render(rt.LBVH, W, H, spp, shadows=1, exact=True)
    assert np.array_equal(sh, sh_e) and st_s["shadow_rays"] > 0


@pytest.mark.parametrize("spp,shadows", [(4, 0), (1, 0), (4, 1)])
def test_scheduling_options_do_not_change_the_frame(gpu_ctx, spp, shadows):
    """frame_graph (the whole device-buffer frame as ONE cudaGraphLaunch), l2_prefetch (the tree streamed into L2 beside the
    direction kernel) and lpt (the previous frame's heaviest blocks launched first) are pure scheduling: bytes, counters and ray counts equal the plain stream path's, frame after frame, also
    when the arguments change between frames (jitter offset, rank, resolution: SetParams / rebuild paths of the graph)."""
    import torch
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    cases = [(400, 300, 0, 1, 0), (400, 300, 0, 1, 0), (400, 300, 0, 1, 4321), (400, 300, 1, 3, 0), (400, 300, 1, 3, 0), (322, 203, 0, 1, 0),
             (400, 300, 0, 1, 0)]
    buf = torch.zeros((300, 400, 3), dtype=torch.uint8, device="cuda")

    def run():
        out = []
        for (W, H, rank, world, joff) in cases:
            buf.zero_()
            p = gpu_ctx.render_params(W, H, spp, rank=rank, world=world, shadows=shadows, jitter_offset=joff)
            st = gpu_ctx.render_device(rt.LBVH, p, buf.data_ptr())
            rows = rt.rows_for_rank(H, 8, rank, world)
            out.append((buf.cpu().numpy().reshape(-1)[: rows * W * 3].copy(), st["rays"], st["node_visits"], st["prim_tests"], st["rows"]))
        return out

    with T.option(gpu_ctx, "lpt", 0):
        plain = run()
    # lpt (default): a frame launches the blocks that were heaviest in the previous frame of the same geometry first (cases 0 -> 1,
    # and again across the two runs); another geometry or kernel drops the learned order
    for graph, pf, lpt in ((0, 0, 1), (1, 0, 0), (0, 1, 1), (1, 1, 0)):
        with T.option(gpu_ctx, "frame_graph", graph), T.option(gpu_ctx, "l2_prefetch", pf), T.option(gpu_ctx, "lpt", lpt):
            got = run() + run()
        for a, b in zip(plain + plain, got):
            # (render_heavy_kernel, which takes over the heaviest tiles once an order is learned, walks one ray per thread: its visit
            # and primitive-test counts are the single-ray traversal's, not the packet's - rays and rows must agree, the work
            # counters only without an order)
            assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[4:] == b[4:] and (lpt != 0 or a[2:4] == b[2:4]), (graph, pf, lpt)
    assert plain[0][1] >= 400 * 300 * spp and (shadows == 0) == (plain[0][1] == 400 * 300 * spp)


@pytest.mark.parametrize("kind", ["lbvh", "median", "sah", "median_mixed_radii", "triangles", "tiny"])
def test_wide_collapse_returns_the_binary_trees_hits(gpu_ctx, oracle, kind):
    """`wide` option: every BVH build is followed by the 4-wide collapse (wide_collapse_kernel: even-depth nodes gather their
    grandchildren), and the one-ray-per-thread primary rays walk it (traverse_wide_loop). Interior tests only steer, so hit ids, float
    sums and bytes must equal the binary walk's - every builder, trees with dropped ranges (leaf boxes from the parent's record),
    triangles, trees of 1..4 leaves - with fewer node visits."""
    if kind == "triangles":
        tris, mat = T.triangle_scene(3000, 9)
    elif kind == "median_mixed_radii":
        sph, mat = T.material_scene(2500, 4, frac_rr=0.0, frac_refl=0.0)      # one big sphere among small ones: dropped ranges
    else:
        sph, mat = T.synthetic_scene(5000, 21)
    sizes = (2, 3, 4, 5) if kind == "tiny" else (None,)
    for m in sizes:
        with T.option(gpu_ctx, "wide", 1):
            if kind == "triangles":
                gpu_ctx.set_triangles(tris, mat)
            elif m is not None:
                gpu_ctx.set_spheres(sph[:m], mat[:m])
            else:
                gpu_ctx.set_spheres(sph, mat)
            acc, mode = {"lbvh": (rt.LBVH, rt.MODE_TRUE), "median": (rt.BVH, rt.MODE_COMPAT), "sah": (rt.BVH, rt.MODE_SAH),
                         "median_mixed_radii": (rt.BVH, rt.MODE_COMPAT), "triangles": (rt.LBVH, rt.MODE_TRUE), "tiny": (rt.LBVH, rt.MODE_TRUE)}[kind]
            gpu_ctx.build(acc, mode=mode)
            for (W, H, spp) in ((322, 203, 1), (161, 97, 3)):
                wide = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True)
                with T.option(gpu_ctx, "wide", 0):
                    binary = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True)
                assert np.array_equal(wide[1], binary[1]) and wide[2].tobytes() == binary[2].tobytes() and np.array_equal(wide[0], binary[0])
                assert wide[3]["rays"] == binary[3]["rays"] and wide[3]["prim_tests"] <= binary[3]["prim_tests"] * 1.05 + 8
                if m is None:
                    assert wide[3]["node_visits"] < 0.75 * binary[3]["node_visits"]
                    assert np.count_nonzero(wide[1] >= 0) > 100
    gpu_ctx.set_spheres(*T.bunny_scene())


def test_grazing_rays_from_arbitrary_origins_fast_equals_exact(gpu_ctx):
    """Rays with a non-zero origin (what shadow and secondary rays are) test interior boxes with one FFMA per plane and an absolute
    widening (traverse.cuh, ray_affine). Hard cases for it: rays aimed EXACTLY at corners / edge points of leaf boxes and at
    tangent points of spheres, from origins on other spheres' surfaces, inside the scene and far outside it, with tiny direction
    components. The ordered traversal must return the unpruned reference traversal's hit id and t bits on every one."""
    sph, mat = T.synthetic_scene(20000, 31)
    gpu_ctx.set_spheres(sph, mat)
    rng = np.random.default_rng(32)
    n = 60000
    tgt_s = sph[rng.integers(0, sph.shape[0], n)]
    sign = rng.choice(np.float32([-1, 1]), size=(n, 3))
    tgt = tgt_s[:, :3] + sign * tgt_s[:, 3:4]                                            # a corner of the sphere's own box
    edge = rng.random(n) < 0.4
    ax = rng.integers(0, 3, n)
    tgt[edge, ax[edge]] = (tgt_s[edge, ax[edge]] + rng.uniform(-1, 1, edge.sum()) * tgt_s[edge, 3]).astype(np.float32)   # edge point
    tang = rng.random(n) < 0.3                                                           # tangent point: centre + r * unit vector
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    tgt[tang] = (tgt_s[tang, :3] + u[tang] * tgt_s[tang, 3:4]).astype(np.float32)
    org_s = sph[rng.integers(0, sph.shape[0], n)]
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (org_s[:, :3] + v * org_s[:, 3:4] * 1.001).astype(np.float32)                    # just off another sphere's surface
    far = rng.random(n) < 0.15
    o[far] = (o[far] * rng.choice(np.float32([20, 1e3, 1e5]), size=(far.sum(), 1))).astype(np.float32)
    d = (tgt - o).astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    tiny = rng.random(n) < 0.1
    d[tiny, ax[tiny]] = rng.choice([1e-6, 1e-9, 1e-13], tiny.sum())
    d = d.astype(np.float32)
    for acc, mode in ((rt.LBVH, rt.MODE_TRUE), (rt.BVH, rt.MODE_SAH)):
        gpu_ctx.build(acc, mode=mode)
        he, te, _ = gpu_ctx.trace(acc, o, d, exact=True)
        hf, tf, st = gpu_ctx.trace(acc, o, d, exact=False)
        assert np.array_equal(he, hf), np.count_nonzero(he != hf)
        assert te.tobytes() == tf.tobytes()
        assert 0.3 < (he >= 0).mean() < 1.0


@pytest.mark.parametrize("spp", [4, 8])
def test_heavy_tiles_rendered_ray_per_thread_give_the_same_frame(gpu_ctx, spp):
    """lpt_split: with a learned block order in use, the heaviest 16 x 8 tiles of the previous frame are rendered by
    render_heavy_kernel (one ray per thread, four lanes per pixel, sums in sample order through shuffles) while the packet kernel
    skips them. Hit ids, float sums, bytes and ray counts must equal the plain packet kernel's, frame after frame, for any number
    of handed-over tiles, on ragged frames and a rank's tiles."""
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    for (W, H, rank, world) in ((322, 203, 0, 1), (400, 300, 1, 3)):
        with T.option(gpu_ctx, "lpt", 0):
            plain = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, rank=rank, world=world)
        for split in (0, 7, 256):
            with T.option(gpu_ctx, "lpt", 2), T.option(gpu_ctx, "lpt_split", split):
                frames = [gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, rank=rank, world=world) for _ in range(4)]
            for f in frames:
                assert np.array_equal(f[1], plain[1]) and f[2].tobytes() == plain[2].tobytes() and np.array_equal(f[0], plain[0]), split
                assert f[3]["rays"] == plain[3]["rays"] == f[3]["primary_rays"]
            # once an order is learned the frame launches render_heavy_kernel as well
            assert frames[-1][3]["kernel_launches"] == plain[3]["kernel_launches"] + (1 if split else 0)
