"""GPU: reference parity ON THE BENCH WORKLOAD ITSELF (BASELINE config 3 at full size: bunny x 30 clones = 1,078,411 spheres,
3840x2160, 4 spp). The compiled, unmodified reference (oracle/_ref) renders a handful of row bands of exactly this frame; the GPU
frame of the same structure must have the same bytes. bench.py stamps every run with the same check over 60 bands (`parity`)."""
import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
pytestmark = pytest.mark.gpu
W, H, SPP = 3840, 2160, 4
BANDS = [(300, 302), (700, 702), (1000, 1002), (1079, 1081), (1300, 1302), (1900, 1902)]


def test_config3_bands_equal_the_unmodified_reference(gpu_ctx, ref):
    sph, mat = rt.scene_from_vertices(T.bunny_vertices(), 30)
    ref.scene_from_spheres(sph, mat)
    ref.build(rt.LBVH)                                   # constructLBVHTree (main.cpp:832): median split without the last object
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_COMPAT)          # the same tree, bit-exact
    compat = gpu_ctx.render(rt.LBVH, W, H, SPP)[0]
    hits = 0
    for y0, y1 in BANDS:
        rgb_ref = ref.render_rows(rt.LBVH, W, H, SPP, y0, y1)[0]
        assert np.array_equal(compat[y0:y1], rgb_ref), (y0, int(np.abs(compat[y0:y1].astype(int) - rgb_ref.astype(int)).max()))
        hits += int((rgb_ref != rgb_ref[0, 0]).any(axis=2).sum())
    assert hits > 1000                                   # the bands do cross the model
    # the frame bench.py times (LBVH true mode: Morton + Karras, keeps the ground sphere) against the reference's BVH rows
    # (constructBVHNew keeps it too): another tree, same leaf-local candidate criterion -> same pixels up to exact-t ties
    ref.scene_from_spheres(sph, mat)
    ref.build(rt.BVH)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    timed = gpu_ctx.render(rt.LBVH, W, H, SPP)[0]
    for y0, y1 in BANDS[1:4]:
        rgb_ref = ref.render_rows(rt.BVH, W, H, SPP, y0, y1)[0]
        d = np.abs(timed[y0:y1].astype(np.int16) - rgb_ref.astype(np.int16))
        assert d.max() <= 1 and (d.max(axis=2) > 0).mean() < 1e-4, (y0, int(d.max()), float((d.max(axis=2) > 0).mean()))
