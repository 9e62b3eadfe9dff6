"""GPU parity for the rows SURVEY.md §8f marks "next": castRay's material branches (REFLECTION_AND_REFRACTION,
REFLECTION, recursion depth) and the shadow query.  Oracle: the CPU restatement (pinned against the reference's own
castRay for materials in tests/test_oracle_vs_reference.py; shadows have no executable reference behaviour —
trace_more is a stub — PARITY UNPINNED for shadows, checked against the restatement of the contract of
main.cpp:468-473 only)."""
import os

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
pytestmark = pytest.mark.gpu
LIGHTS3 = np.asarray([[0, 3, 30, 10, 1, 1, 1], [20, 30, -40, 1, 0.5, 0.4, 0.3], [-30, 5, -70, 1, 0.2, 0.3, 0.6]], np.float32)


@pytest.mark.parametrize("acc,exact", [(rt.BVH, True), (rt.BVH, False), (rt.LBVH, False), (rt.NONE, False)])
@pytest.mark.parametrize("shadows", [0, 1])
def test_full_castray_matches_restatement(gpu_ctx, oracle, acc, exact, shadows):
    n = 1500 if acc != rt.NONE else 400
    sph, mat = T.material_scene(n, 11)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.set_lights(LIGHTS3)
    try:
        if acc == rt.NONE:
            nodes = order = None
            tie = 0
        else:
            gpu_ctx.build(acc, mode=rt.MODE_TRUE if acc == rt.LBVH else rt.MODE_COMPAT)
            nodes, order = gpu_ctx.export_bvh()
            tie = 1 if acc == rt.LBVH else 0
        W, H, spp = 240, 180, 2
        rgb, hit, accum, st = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True, exact=exact, shadows=shadows)
        rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, tie_by_objid=tie, lights=LIGHTS3,
                                                       want_accum=True, shadows=shadows)
        rays, sh, sec = oracle.last_ray_counts
        assert (st["rays"], st["shadow_rays"], st["secondary_rays"]) == (rays, sh, sec)
        assert st["primary_rays"] == W * H * spp and sec > 10 and (sh > 1000 if shadows else sh == 0)
        assert accum.tobytes() == accum_o.tobytes(), f"{np.count_nonzero((accum != accum_o).any(-1))} pixels differ"
        assert np.array_equal(rgb, rgb_o)
    finally:
        gpu_ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32))


def test_materials_vs_compiled_reference_castray(gpu_ctx):
    """The reference's own castRay (materials set on Sphere::materialType) — within 1 LSB (glibc powf rounding)."""
    if not os.path.exists(os.path.join(T.ROOT, "oracle", "_ref", "libref_oracle.so")):
        pytest.skip("compiled reference not on this box")
    ref = T.Ref()
    sph, mat = T.material_scene(1500, 11)
    ref.scene_from_spheres(sph, mat)
    ref.lib.ref_set_lights(LIGHTS3.ctypes.data_as(T.C.c_void_p), 3)
    ref.build(rt.BVH)
    rgb_r, _, acc_r, _ = ref.render_rows(rt.BVH, 240, 180, 2, want_accum=True)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.set_lights(LIGHTS3)
    try:
        gpu_ctx.build(rt.BVH)
        rgb, _, accum, _ = gpu_ctx.render(rt.BVH, 240, 180, 2, want_accum=True)
        assert np.allclose(accum, acc_r, rtol=3e-7, atol=1e-7)
        assert np.abs(rgb.astype(int) - rgb_r.astype(int)).max() <= 1
        assert np.count_nonzero(rgb != rgb_r) <= 3
    finally:
        gpu_ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32))


def test_shadows_darken_only(gpu_ctx):
    """Property at BASELINE config 2's size (1920x1080): shadows never brighten a pixel, sky pixels are unchanged,
    and some bunny/ground pixels get darker."""
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    a, hit, acc_a, st_a = gpu_ctx.render(rt.LBVH, 1920, 1080, 1, want_hit=True, want_accum=True)
    b, _, acc_b, st_b = gpu_ctx.render(rt.LBVH, 1920, 1080, 1, want_accum=True, shadows=1)
    assert np.all(acc_b <= acc_a)
    assert np.array_equal(a[hit < 0], b[hit < 0])
    assert np.count_nonzero((acc_b < acc_a).any(-1)) > 1000
    assert st_b["shadow_rays"] == np.count_nonzero(hit >= 0) and st_b["rays"] == 1920 * 1080 + st_b["shadow_rays"]
    print("1920x1080 LBVH: %.3f ms primary only, %.3f ms with shadows (%d shadow rays)" % (st_a["ms_kernel"], st_b["ms_kernel"], st_b["shadow_rays"]))


@pytest.mark.parametrize("acc", [rt.BVH, rt.LBVH])
def test_shadowed_packet_kernel_matches_restatement_and_full_kernel(gpu_ctx, oracle, acc):
    """Shadows on, diffuse materials only, aa_samples % 4 == 0: the packet kernel traces the primary rays (four samples of a
    pixel together) and single shadow rays; float sums, bytes and ray counts must equal the CPU restatement and the full
    castRay kernel (RTDS_PACKET=0)."""
    sph, mat = T.synthetic_scene(4000, 77)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.set_lights(LIGHTS3)
    try:
        gpu_ctx.build(acc, mode=rt.MODE_TRUE if acc == rt.LBVH else rt.MODE_COMPAT)
        nodes, order = gpu_ctx.export_bvh()
        W, H, spp = 322, 203, 4
        rgb, hit, accum, st = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True, shadows=1)
        rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, tie_by_objid=1 if acc == rt.LBVH else 0,
                                                       lights=LIGHTS3, want_accum=True, shadows=1)
        rays, sh, sec = oracle.last_ray_counts
        assert (st["rays"], st["shadow_rays"], st["secondary_rays"]) == (rays, sh, sec) and sh > 1000 and sec == 0
        assert np.array_equal(hit, hit_o) and accum.tobytes() == accum_o.tobytes() and np.array_equal(rgb, rgb_o)
        with T.option(gpu_ctx, "packet", 0):
            rgb_f, hit_f, accum_f, st_f = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True, shadows=1)
        assert accum.tobytes() == accum_f.tobytes() and np.array_equal(rgb, rgb_f) and np.array_equal(hit, hit_f)
        assert st_f["shadow_rays"] == st["shadow_rays"]
    finally:
        gpu_ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32))


@pytest.mark.parametrize("acc", [rt.BVH, rt.LBVH])
@pytest.mark.parametrize("shadows,spp", [(0, 4), (1, 4), (1, 8)])
def test_material_scenes_through_the_packet_kernel(gpu_ctx, oracle, acc, shadows, spp):
    """Scenes with REFLECTION_AND_REFRACTION / REFLECTION primitives and aa_samples % 4 == 0: the primary rays go as packets
    (render_packet_kernel<.., MATERIALS>), a ray that hits such a primitive continues through castRay's material branches alone.
    Result = the single-ray castRay kernel's (packet option off) = the CPU restatement's: hit ids, float sums, bytes, ray counts."""
    sph, mat = T.material_scene(1500, 13, frac_rr=0.15, frac_refl=0.15)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.set_lights(LIGHTS3)
    try:
        gpu_ctx.build(acc, mode=rt.MODE_TRUE if acc == rt.LBVH else rt.MODE_COMPAT)
        nodes, order = gpu_ctx.export_bvh()
        W, H = 240, 180
        pk = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True, shadows=shadows)
        with T.option(gpu_ctx, "packet", 0):
            sr = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True, shadows=shadows)
        assert np.array_equal(pk[1], sr[1]) and pk[2].tobytes() == sr[2].tobytes() and np.array_equal(pk[0], sr[0])
        for k in ("rays", "primary_rays", "shadow_rays", "secondary_rays"):
            assert pk[3][k] == sr[3][k], k
        assert pk[3]["secondary_rays"] > 100
        # (the median split over spheres of mixed size drops ranges, accelerators.h:321-327: leaves then carry range boxes and the
        # packet kernels do not apply - same kernel, same counters)
        # and so do material scenes WITH shadow rays (measured slower through the packets: the single-ray castRay kernel stays)
        assert pk[3]["node_visits"] < sr[3]["node_visits"] if (acc == rt.LBVH and not shadows) else pk[3]["node_visits"] <= sr[3]["node_visits"]
        rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, tie_by_objid=1 if acc == rt.LBVH else 0,
                                                       lights=LIGHTS3, want_accum=True, shadows=shadows)
        assert np.array_equal(pk[1], hit_o) and pk[2].tobytes() == accum_o.tobytes() and np.array_equal(pk[0], rgb_o)
        assert tuple(int(x) for x in oracle.last_ray_counts) == (pk[3]["rays"], pk[3]["shadow_rays"], pk[3]["secondary_rays"])
    finally:
        gpu_ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32))


@pytest.mark.parametrize("scene", ["materials", "diffuse"])
@pytest.mark.parametrize("spp", [4, 8])
def test_wavefront_shadow_frames_equal_the_single_kernel_form(gpu_ctx, oracle, scene, spp):
    """Frames with shadow rays and aa_samples % 4 == 0 can run as a wavefront of two kernels (primary packets -> one thread per
    sample: its shadow rays + shading, per-pixel sums in sample order; wavefront option 2 = always): hit ids, float sums, bytes and
    ray counts must equal the single-kernel forms'
    (wavefront off: packet kernel with single shadow rays / castRay kernel; packet off: the one-ray-at-a-time kernel) and the CPU
    restatement's - ragged frame, a rank's interleaved tiles and row bands (a tall frame) included."""
    sph, mat = T.material_scene(1500, 17, frac_rr=0.15, frac_refl=0.15) if scene == "materials" else T.synthetic_scene(1500, 17)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.set_lights(LIGHTS3)
    try:
        gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        nodes, order = gpu_ctx.export_bvh()
        for (W, H) in ((241, 179), (64, 600)):
            with T.option(gpu_ctx, "wavefront", 2):
                wv = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, shadows=1)
            with T.option(gpu_ctx, "wavefront", 0):
                mk = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, shadows=1)
            with T.option(gpu_ctx, "packet", 0):
                sr = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, shadows=1)
            for other in (mk, sr):
                assert np.array_equal(wv[1], other[1]) and wv[2].tobytes() == other[2].tobytes() and np.array_equal(wv[0], other[0])
                for k in ("rays", "primary_rays", "shadow_rays", "secondary_rays"):
                    assert wv[3][k] == other[3][k], k
            assert wv[3]["kernel_launches"] > mk[3]["kernel_launches"] and wv[3]["shadow_rays"] > 1000
        W, H = 241, 179
        with T.option(gpu_ctx, "wavefront", 2):
            wv = gpu_ctx.render(rt.LBVH, W, H, spp, want_hit=True, want_accum=True, shadows=1)
            part = gpu_ctx.render(rt.LBVH, W, H, spp, shadows=1, rank=1, world=3)[0]
        rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, tie_by_objid=1, lights=LIGHTS3, want_accum=True, shadows=1)
        assert np.array_equal(wv[1], hit_o) and wv[2].tobytes() == accum_o.tobytes() and np.array_equal(wv[0], rgb_o)
        assert tuple(int(x) for x in oracle.last_ray_counts) == (wv[3]["rays"], wv[3]["shadow_rays"], wv[3]["secondary_rays"])
        rows = rt.owned_rows(H, 8, 1, 3)
        assert np.array_equal(part[rows], wv[0][rows])
    finally:
        gpu_ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32))


def test_wavefront_form_is_chosen_by_measurement(gpu_ctx):
    """wavefront = 1 (default): the first four timed frames of a frame geometry alternate between the two forms (single kernel,
    wavefront, ...), later frames use the faster one; every frame is the same bytes. aa_samples 12 (not a power of two) goes through
    the block-sum variant of the per-sample kernel."""
    sph, mat = T.synthetic_scene(3000, 5)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.set_lights(LIGHTS3)
    try:
        gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        for spp in (8, 12):
            W, H = 203 + spp, 157
            with T.option(gpu_ctx, "wavefront", 0):
                base = gpu_ctx.render(rt.LBVH, W, H, spp, want_accum=True, shadows=1)
            frames = [gpu_ctx.render(rt.LBVH, W, H, spp, want_accum=True, shadows=1) for _ in range(7)]
            launches = [f[3]["kernel_launches"] for f in frames]
            assert launches[0] == launches[2] == base[3]["kernel_launches"] and launches[1] == launches[3] == launches[0] + 1
            assert len(set(launches[4:])) == 1
            for f in frames:
                assert np.array_equal(f[0], base[0]) and f[2].tobytes() == base[2].tobytes()
                assert f[3]["rays"] == base[3]["rays"] and f[3]["shadow_rays"] == base[3]["shadow_rays"]
    finally:
        gpu_ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32))
