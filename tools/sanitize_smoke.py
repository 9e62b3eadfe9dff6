"""Small pass over EVERY kernel of librtds.so, meant to run under compute-sanitizer (tools/sanitize.sh):
all four builders (median split incl. the cooperative top-level kernel, Morton/onesweep/Karras/refit LBVH at 30 and 63
bits, binned SAH, KD), every render kernel (exact, ordered, packet, NONE, KD any-hit / closest hit, shadows, materials,
strip), the probes (trace, jitter stream, Morton) and triangle scenes. No torch: ctypes binding only.
RTDS_SAN_SCALE scales the scene sizes (default 1; racecheck runs use a fraction)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402

rt = entry.load_rtds()
S = float(os.environ.get("RTDS_SAN_SCALE", "1"))


def scene(n, seed, radius=0.05):
    rng = np.random.default_rng(seed)
    sph = np.zeros((n + 1, 4), np.float32)
    sph[:n, :3] = rng.normal(size=(n, 3)).astype(np.float32) * np.float32(4) + np.float32([0, 0, -60])
    sph[:n, 3] = radius
    sph[n] = np.asarray(rt.GROUND, np.float32)
    mat = np.zeros_like(sph)
    mat[:n, :3] = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    return sph, mat


def main():
    ctx = rt.Rtds(0)
    W, H = 96, 64
    n = max(200, int(3000 * S))
    sph, mat = scene(n, 1)
    ctx.set_spheres(sph, mat)
    ctx.set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1], [20, 30, -40, 1, 0.5, 0.4, 0.3]], np.float32))
    rays_d = np.random.default_rng(2).normal(size=(2000, 3)).astype(np.float32) * np.float32([0.1, 0.1, 0]) + np.float32([0, 0, -1])
    rays_d /= np.linalg.norm(rays_d, axis=1, keepdims=True)
    rays_o = np.zeros((1, 3), np.float32)
    done = []
    for acc, kw in ((rt.BVH, {}), (rt.LBVH, {}), (rt.LBVH, {"mode": rt.MODE_TRUE}), (rt.LBVH, {"mode": rt.MODE_TRUE, "morton_bits": 63}),
                    (rt.BVH, {"mode": rt.MODE_SAH})):
        try:
            ctx.build(acc, **kw)
        except rt.RtdsError as e:         # RTDS_ERR_DEGENERATE: an input the reference builder itself does not survive
            if e.code != -6:
                raise
            print("skipped", acc, kw, e)
            continue
        ctx.export_bvh()
        if kw.get("mode") == rt.MODE_TRUE:
            ctx.export_morton()
        for exact in (True, False):
            ctx.render(acc, W, H, 1, exact=exact, want_hit=True, want_accum=True)
            ctx.trace(acc, rays_o, rays_d, exact=exact)
        ctx.render(acc, W, H, 4)                      # packet kernel
        ctx.render(acc, W, H, 4, shadows=1)           # packet kernel + single shadow rays
        ctx.render(acc, W + 5, H + 3, 3, shadows=1)   # ragged frame, single-ray kernels
        ctx.render(acc, W, H, 2, rank=1, world=3)     # a rank's interleaved tiles
        done.append((acc, kw))
    ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    ctx.set_option("strip", 1)
    ctx.render(rt.LBVH, W, H, 2)                      # fused strip kernel (opt-in)
    ctx.set_option("strip", 0)
    ctx.build(rt.KDTREE)
    ctx.export_kd()
    ctx.render(rt.KDTREE, W, H, 2)
    ctx.render(rt.KDTREE, W, H, 2, kd_closest=1)
    ctx.trace(rt.KDTREE, rays_o, rays_d)
    ctx.trace(rt.KDTREE, rays_o, rays_d, kd_closest=True)
    ctx.render(rt.NONE, 48, 32, 1)
    ctx.trace(rt.NONE, rays_o, rays_d[:500])
    ctx.jitter_stream(12345, 5000)
    ctx.morton30(sph[:, :3])
    # materials: reflection / refraction branches of castRay (render_full_kernel)
    mat2 = mat.copy()
    mat2[: n // 5, 3] = 1.0
    mat2[n // 5: 2 * n // 5, 3] = 2.0
    sph2 = sph.copy()
    sph2[:16, 3] = 1.0
    ctx.set_spheres(sph2, mat2)
    ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    for exact in (True, False):
        ctx.render(rt.LBVH, W, H, 2, shadows=1, exact=exact)
    for sh in (0, 1):
        ctx.render(rt.LBVH, W, H, 4, shadows=sh)      # materials through the packet kernel (no shadows) / castRay kernel (shadows)
    # 4-wide collapse behind the build + the wide walk of the one-ray-per-thread kernel (opt-in)
    ctx.set_option("wide", 1)
    ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    ctx.render(rt.LBVH, W + 3, H + 1, 1, want_hit=True)
    ctx.render(rt.LBVH, W, H, 3)
    ctx.set_option("wide", 0)
    ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    # wavefront form of shadowed frames (wave_primary_kernel + wave_shade_kernel<warp sums / block sums>), materials in the scene
    ctx.set_option("wavefront", 2)
    for spp in (4, 16, 12):
        ctx.render(rt.LBVH, W + 3, H + 1, spp, shadows=1, want_hit=True, want_accum=True)
    ctx.render(rt.LBVH, W, H, 8, shadows=1, rank=1, world=3)
    ctx.set_option("wavefront", 1)
    ctx.frame(sph, mat, rt.LBVH, W, H, 4, mode=rt.MODE_TRUE)      # rtds_frame: overlapped upload + build + render
    # round 2: scheduling options on device-buffer renders - lpt (block_order_kernel: needs consecutive frames of one geometry),
    # one CUDA graph per frame, L2 prefetch of the tree; rtds_prepare_frame; the in-process shared frame (flag + wait kernels)
    import ctypes as C
    dbuf = C.c_void_p()
    cudart = C.CDLL("libcudart.so")
    assert cudart.cudaMalloc(C.byref(dbuf), W * H * 3) == 0
    ctx.set_option("lpt", 2)           # (lpt_split default: render_heavy_kernel takes the heaviest tiles from the second frame on)
    for graph, pf in ((0, 0), (0, 1), (1, 0), (1, 1)):
        ctx.set_option("frame_graph", graph); ctx.set_option("l2_prefetch", pf)
        for rep in range(3):
            ctx.render_device(rt.LBVH, ctx.render_params(W, H, 4), dbuf.value)
            ctx.render_device(rt.LBVH, ctx.render_params(W, H, 1), dbuf.value)
    ctx.set_option("frame_graph", 0); ctx.set_option("l2_prefetch", 0); ctx.set_option("lpt", 1)
    for rep in range(8):
        ctx.render(rt.LBVH, W, H, 4)                  # lpt = 1: baseline / trial frames, then the decision
    ctx.prepare_frame(ctx.render_params(W, H, 4))
    ctx.render(rt.LBVH, W, H, 4)
    other = rt.Rtds(0)
    other.set_spheres(sph, mat)
    other.build(rt.LBVH, mode=rt.MODE_TRUE)
    ctx.shared_frame_create(W, H, 2)
    other.shared_frame_attach(ctx, 1)
    for graph in (0, 1):
        ctx.set_option("frame_graph", graph); other.set_option("frame_graph", graph)
        for seq in (1 + 2 * graph, 2 + 2 * graph):
            other.render_shared(rt.LBVH, other.render_params(W, H, 4, rank=1, world=2), seq)
            ctx.render_shared(rt.LBVH, ctx.render_params(W, H, 4, rank=0, world=2), seq)
    ctx.shared_frame_read(W, H)
    other.shared_frame_close(); other.close(); ctx.shared_frame_close()
    ctx.set_option("frame_graph", 0)
    # rtds_set_spheres_device (D2D upload)
    dsph, dmat = C.c_void_p(), C.c_void_p()
    assert cudart.cudaMalloc(C.byref(dsph), sph.nbytes) == 0 and cudart.cudaMalloc(C.byref(dmat), mat.nbytes) == 0
    assert cudart.cudaMemcpy(dsph, sph.ctypes.data_as(C.c_void_p), sph.nbytes, 1) == 0 and cudart.cudaMemcpy(dmat, mat.ctypes.data_as(C.c_void_p), mat.nbytes, 1) == 0
    ctx.set_spheres_device(dsph.value, dmat.value, sph.shape[0])
    ctx.build(rt.BVH, mode=rt.MODE_SAH)               # SAH: large / warp / thread task kernels all occur at this size
    ctx.render(rt.BVH, W, H, 4)
    for pbuf in (dbuf, dsph, dmat):
        cudart.cudaFree(pbuf)
    # triangles
    rng = np.random.default_rng(3)
    nt = max(200, int(2000 * S))
    c = rng.normal(size=(nt, 1, 3)).astype(np.float32) * np.float32(4) + np.float32([0, 0, -60])
    tris = np.ascontiguousarray((c + rng.normal(size=(nt, 3, 3)).astype(np.float32) * np.float32(0.15)).reshape(nt, 9))
    ctx.set_triangles(tris, None)
    for acc, kw in ((rt.LBVH, {"mode": rt.MODE_TRUE}), (rt.BVH, {"mode": rt.MODE_SAH}), (rt.KDTREE, {})):
        ctx.build(acc, **kw)
        ctx.render(acc, W, H, 2)
        ctx.trace(acc, rays_o, rays_d, exact=False)
        ctx.render(acc, W, H, 2, tri_geometric=1)     # the triangle test the reference compiles
        ctx.trace(acc, rays_o, rays_d, exact=False, tri_geometric=True)
    ctx.render(rt.KDTREE, W, H, 1, kd_closest=1)
    # the cooperative top-level median kernel needs a range > 65,536 objects
    if S >= 1:
        big, bmat = scene(70000, 4)
        ctx.set_spheres(big, bmat)
        ctx.build(rt.BVH)
        ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        ctx.render(rt.LBVH, W, H, 4)
    ctx.close()
    print("sanitize_smoke: every kernel family ran (%d sphere prims, %d triangles)" % (n, nt))


if __name__ == "__main__":
    main()
