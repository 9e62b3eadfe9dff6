"""Per-block timeline of render_packet_kernel (profiling variant: tools/build_variant.sh bt -DRTDS_BLOCK_TIMING=1):
where does a rank's kernel at world 8 spend the ~0.1 ms that does not shrink with the number of GPUs?
usage: RTDS_LIB=.../variants/librtds_bt.so [WORLD=8] [RANK=5] python tools/block_timeline.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rt = bench.entry.load_rtds()
wl = bench.workloads(rt)["config3"]
world, rank = int(os.environ.get("WORLD", "8")), int(os.environ.get("RANK", "5"))
ctx = rt.Rtds(0)
if os.environ.get("ORDER"):
    ctx.set_option("block_order", int(os.environ["ORDER"]))
if os.environ.get("LPT"):
    ctx.set_option("lpt", int(os.environ["LPT"]))
if os.environ.get("LPT_SPLIT"):
    ctx.set_option("lpt_split", int(os.environ["LPT_SPLIT"]))
sph, mat = wl.scene()
ctx.set_spheres(sph, mat)
ctx.build(wl.acc, mode=wl.mode)
rows = rt.rows_for_rank(wl.H, 8, rank, world)
buf = torch.zeros((rows, wl.W, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p = ctx.render_params(wl.W, wl.H, wl.spp, rank=rank, world=world)
for it in range(8):
    flush.fill_(1)
    torch.cuda.synchronize()
    st = ctx.render_device(wl.acc, p, buf.data_ptr())
nb = ((wl.W + 15) // 16) * ((rows + 7) // 8)
t = np.zeros((nb, 3), np.uint64)
assert ctx.lib.rtds_debug_block_times(t.ctypes.data_as(C.c_void_p), nb) == 0
start, end = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64)
sm, visits = (t[:, 2] >> np.uint64(32)).astype(np.int64), (t[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
# blocks that handed their tile to render_heavy_kernel (lpt_split) leave no record in the last frame: drop stale entries
live = start > end.max() - 2_000_000
if not live.all():
    print("(%d tiles rendered by render_heavy_kernel in this frame: not in the statistics below)" % int((~live).sum()))
    start, end, sm, visits = start[live], end[live], sm[live], visits[live]
t0 = start.min()
start, end = (start - t0) / 1e3, (end - t0) / 1e3     # microseconds
dur = end - start
print("lpt %d block_order %d | world %d rank %d: %d blocks, kernel %.1f us by CUDA events, %.1f us first block start -> last block end" % (ctx.get_option("lpt"), ctx.get_option("block_order"), world, rank, nb, 1e3 * st["ms_kernel"], end.max()))
print("block duration us: min %.1f  median %.1f  mean %.1f  p90 %.1f  p99 %.1f  max %.1f" %
      (dur.min(), np.median(dur), dur.mean(), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
print("sum of block durations / (SMs x 6 slots): %.1f us (the kernel's length if every slot were always busy)" % (dur.sum() / (148 * 6)))
order = np.argsort(end)[::-1][:8]
print("last blocks to finish: " + ", ".join("#%d start %.0f dur %.0f (thread 0: %d visits)" % (b, start[b], dur[b], visits[b]) for b in order))
edges = np.arange(0, end.max() + 10, 10.0)
busy = [(int(((start < e + 10) & (end > e)).sum())) for e in edges]
print("resident blocks per 10 us slice (of %d slots): " % (148 * 6) + " ".join(str(x) for x in busy))
late = start > 0.5 * end.max()
print("blocks started in the second half: %d, their mean duration %.1f us; blocks started in the first 20 us: %d, mean duration %.1f us" %
      (late.sum(), dur[late].mean() if late.any() else 0, (start < 20).sum(), dur[start < 20].mean()))
per_sm = np.bincount(sm, minlength=148)
print("blocks per SM: min %d max %d" % (per_sm.min(), per_sm.max()))
