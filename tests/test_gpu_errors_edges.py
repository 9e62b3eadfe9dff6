"""GPU: error behaviour at the C ABI and edge-case inputs (ragged / minimal / large sample counts)."""
import ctypes as C

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
pytestmark = pytest.mark.gpu


def test_error_codes(oracle):
    ctx = rt.Rtds(0)
    try:
        with pytest.raises(rt.RtdsError) as e:
            ctx.build(rt.BVH)
        assert e.value.code == -4                                   # NO_SCENE
        with pytest.raises(rt.RtdsError) as e:
            ctx.set_spheres(np.zeros((0, 4), np.float32))
        assert e.value.code == -1                                   # INVALID
        sph, mat = T.synthetic_scene(50, 1)
        ctx.set_spheres(sph, mat)
        with pytest.raises(rt.RtdsError) as e:
            ctx.render(rt.BVH, 32, 32, 1)
        assert e.value.code == -5                                   # NOT_BUILT
        ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        with pytest.raises(rt.RtdsError) as e:
            ctx.render(rt.BVH, 32, 32, 1)                           # built as LBVH, asked for BVH
        assert e.value.code == -5
        with pytest.raises(rt.RtdsError) as e:
            ctx.render(rt.KDTREE, 32, 32, 1)
        assert e.value.code == -5
        with pytest.raises(rt.RtdsError) as e:
            ctx.render(rt.LBVH, 0, 32, 1)
        assert e.value.code == -1
        with pytest.raises(rt.RtdsError) as e:
            ctx.render(rt.LBVH, 32, 32, 1, rank=3, world=2)
        assert e.value.code == -1
        nn, npr = C.c_int(), C.c_int()
        small = np.zeros(3, rt.LINEAR_NODE_DTYPE)
        rc = ctx.lib.rtds_export_bvh(ctx.ctx, small.ctypes.data_as(C.c_void_p), 3, C.byref(nn), None, 0, C.byref(npr))
        assert rc == -7 and nn.value == 2 * sph.shape[0] - 1        # CAPACITY, and the needed size is reported
        with pytest.raises(rt.RtdsError) as e:
            ctx.build(rt.LBVH, mode=rt.MODE_TRUE, morton_bits=17)
        assert e.value.code == -1
        ctx.set_spheres(sph, mat)                                   # a new scene invalidates the old structure
        with pytest.raises(rt.RtdsError) as e:
            ctx.render(rt.LBVH, 32, 32, 1)
        assert e.value.code == -5
        h, t, _ = ctx.trace(rt.NONE, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
        assert h.size == 0
    finally:
        ctx.close()


@pytest.mark.parametrize("W,H,spp", [(1, 1, 1), (17, 9, 3), (5, 64, 16), (640, 1, 1)])
def test_ragged_frames_match_oracle(gpu_ctx, oracle, W, H, spp):
    sph, mat = T.synthetic_scene(800, 77)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.BVH)
    nodes, order = gpu_ctx.export_bvh()
    for acc, nd, od in ((rt.BVH, nodes, order), (rt.NONE, None, None)):
        rgb, hit, accum, st = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True)
        rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nd, od, W, H, spp, want_accum=True)
        assert np.array_equal(hit, hit_o) and accum.tobytes() == accum_o.tobytes() and np.array_equal(rgb, rgb_o)
        assert st["primary_rays"] == W * H * spp


def test_rays_with_zero_direction_components(gpu_ctx, oracle):
    """Axis-parallel rays: divisions by +-0 give +-inf / NaN in the reference's slab test; both traversal modes must
    reproduce what IEEE arithmetic does there."""
    sph, mat = T.synthetic_scene(3000, 78)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.BVH)
    nodes, order = gpu_ctx.export_bvh()
    c = sph[:600, :3]
    o = np.concatenate([c + np.float32([0, 0, 30]), c + np.float32([25, 0, 0]), c + np.float32([0, -40, 0])]).astype(np.float32)
    d = np.concatenate([np.tile(np.float32([0, 0, -1]), (600, 1)), np.tile(np.float32([-1, 0, 0]), (600, 1)),
                        np.tile(np.float32([0, 1, -0.0]), (600, 1))]).astype(np.float32)
    h_o, t_o, _ = oracle.trace(sph, nodes, order, o, d)
    for exact in (True, False):
        h, t, _ = gpu_ctx.trace(rt.BVH, o, d, exact=exact)
        assert np.array_equal(h, h_o) and t.tobytes() == t_o.tobytes()
    assert (h_o >= 0).mean() > 0.5


def test_render_is_deterministic_and_rebuild_stable(gpu_ctx):
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    frames, trees = [], []
    for _ in range(3):
        gpu_ctx.build(rt.BVH)
        nodes, order = gpu_ctx.export_bvh()
        trees.append(nodes.tobytes() + order.tobytes())
        frames.append(gpu_ctx.render(rt.BVH, 320, 240, 2)[0].tobytes())
    assert len(set(trees)) == 1 and len(set(frames)) == 1


def test_failing_frame_call_is_synchronous_and_leaves_the_context_usable():
    """rtds_frame that fails AFTER it queued its uploads (here: a KD depth beyond the 64-entry traversal stack, refused by the
    builder) must not leave copies or the direction kernel running behind the error return, and the context must work afterwards;
    kd_max_depth > 64 is an error, not a silently wrong frame (kdtreeIntersect's todo[64], accelerators.h:1008)."""
    import torch
    ctx = rt.Rtds(0)
    try:
        sph, mat = T.synthetic_scene(20000, 3)
        sph_pin, mat_pin = torch.from_numpy(sph).pin_memory(), torch.from_numpy(mat).pin_memory()
        out = np.zeros((120, 160, 3), np.uint8)
        bp = rt.BuildParams()
        bp.kd_max_depth = 65
        rp = ctx.render_params(160, 120, 4)
        for _ in range(3):
            rc = ctx.lib.rtds_frame(ctx.ctx, C.c_void_p(sph_pin.data_ptr()), C.c_void_p(mat_pin.data_ptr()), sph.shape[0], rt.KDTREE, C.byref(bp),
                                    C.byref(rp), out.ctypes.data_as(C.c_void_p), None, None)
            assert rc == -8 and b"64-entry" in ctx.lib.rtds_last_error()          # UNSUPPORTED
            sph_pin.fill_(0.0)                      # would corrupt an upload still in flight ...
            sph_pin.copy_(torch.from_numpy(sph))    # ... and is restored for the next round
        with pytest.raises(rt.RtdsError):
            ctx.build(rt.KDTREE, kd_max_depth=100)
        # the context is fine: a normal frame through the same call, equal to the three-call sequence
        rgb, bst, rst = ctx.frame(sph, mat, rt.LBVH, 160, 120, 4, mode=rt.MODE_TRUE)
        ctx.set_spheres(sph, mat)
        ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        assert np.array_equal(rgb, ctx.render(rt.LBVH, 160, 120, 4)[0])
        tris, tmat = T.triangle_scene(500, 5)
        ctx.set_triangles(tris, tmat)               # right after an asynchronous-material frame: no stale pending state
        ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        assert ctx.render(rt.LBVH, 64, 48, 1)[3]["primary_rays"] == 64 * 48
    finally:
        ctx.close()


def test_set_spheres_device_equals_host_upload(gpu_ctx):
    """rtds_set_spheres_device (tables already in device memory, e.g. all-gathered over NVLink from per-rank parts) gives the same
    structure and frame as rtds_set_spheres, also for scenes with materials (the material flag is computed on the device either way)."""
    import torch
    for sph, mat in (T.synthetic_scene(5000, 9), T.material_scene(3000, 10)):
        gpu_ctx.set_spheres(sph, mat)
        gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        nodes, order = gpu_ctx.export_bvh()
        rgb = gpu_ctx.render(rt.LBVH, 200, 150, 4)[0]
        d_sph, d_mat = torch.from_numpy(sph).cuda(), torch.from_numpy(mat).cuda()
        torch.cuda.synchronize()
        gpu_ctx.set_spheres_device(d_sph.data_ptr(), d_mat.data_ptr(), sph.shape[0])
        gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        nodes2, order2 = gpu_ctx.export_bvh()
        assert nodes2.tobytes() == nodes.tobytes() and np.array_equal(order2, order)
        assert np.array_equal(gpu_ctx.render(rt.LBVH, 200, 150, 4)[0], rgb)
    with pytest.raises(rt.RtdsError):
        gpu_ctx.set_spheres_device(d_sph.data_ptr(), 0, sph.shape[0])
