"""Host-clock timeline of rtds_frame's stages on the bench workload (RTDS_TRACE_FRAME=1 -> stderr). GPU box only."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTDS_TRACE_FRAME", "1")
import __graft_entry__ as entry  # noqa: E402

rt = entry.load_rtds()
import torch  # noqa: E402  (pinned host buffers only)

v = np.fromfile(os.path.join(ROOT, "tests", "golden", "bunny_vertices.f32"), np.float32).reshape(-1, 3)
sph, mat = rt.scene_from_vertices(v, 30)
sph_pin, mat_pin = torch.from_numpy(sph).pin_memory(), torch.from_numpy(mat).pin_memory()
W, H, SPP = 3840, 2160, 4
frame = torch.zeros((H, W, 3), dtype=torch.uint8).pin_memory()
ctx = rt.Rtds(0)
bp = rt.BuildParams()
bp.mode = rt.MODE_TRUE
rp = ctx.render_params(W, H, SPP)
bst, rst = rt.BuildStats(), rt.RenderStats()
for i in range(8):
    t0 = time.perf_counter()
    rc = ctx.lib.rtds_frame(ctx.ctx, C.c_void_p(sph_pin.data_ptr()), C.c_void_p(mat_pin.data_ptr()), sph.shape[0], rt.LBVH, C.byref(bp),
                            C.byref(rp), C.c_void_p(frame.data_ptr()), C.byref(bst), C.byref(rst))
    dt = (time.perf_counter() - t0) * 1e3
    assert rc == 0, ctx.lib.rtds_last_error()
    print("frame %d: %.3f ms wall; build %.3f ms device, render kernel %.3f ms, render total %.3f ms device" %
          (i, dt, bst.ms, rst.ms_kernel, rst.ms_total), file=sys.stderr)
if os.environ.get("RTDS_CHECK_FRAME"):          # compare the last frame with a plain render of the same scene
    ref_rgb, _, _, _ = ctx.render(rt.LBVH, W, H, SPP)
    print("frame equals rtds_render's:", bool((frame.numpy() == ref_rgb).all()), file=sys.stderr)
