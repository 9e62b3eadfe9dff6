"""One rank's end-to-end call (rtds_frame with rank r of WORLD: full upload + build + its tiles rendered + downloaded) on one GPU, for
different numbers of row bands - the per-GPU critical path of the N-GPU end-to-end step without paying for N GPUs.
usage: WORLD=4 python tools/e2e_rank_bands.py 1 2 4 6"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rt = bench.entry.load_rtds()
wl = bench.workloads(rt)["config3"]
world = int(os.environ.get("WORLD", "4"))
ctx = rt.Rtds(0)
sph, mat = wl.scene()
sph_pin = torch.from_numpy(sph).pin_memory()
mat_pin = torch.from_numpy(mat).pin_memory()
out = torch.zeros((wl.H, wl.W, 3), dtype=torch.uint8).pin_memory().numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(2):
    for bands in [int(a) for a in sys.argv[1:]]:
        ctx.set_option("bands", bands)
        worst = 0.0
        for rank in range(world):
            ts = []
            for it in range(12):
                flush.fill_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.frame(sph_pin.numpy(), mat_pin.numpy(), wl.acc, wl.W, wl.H, wl.spp, mode=wl.mode, out=out, rank=rank, world=world)
                ts.append((time.perf_counter() - t0) * 1e3)
            worst = max(worst, float(np.median(ts[7:])))
        print("world %d bands %d: slowest rank's rtds_frame call %.3f ms" % (world, bands, worst), flush=True)
