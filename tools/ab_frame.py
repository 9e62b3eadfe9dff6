"""A/B of whole-frame options on the RESIDENT path (rtds_render_device: what bench.py's `value` times), L2 flushed before every
frame, optionally as rank r of WORLD on one GPU (the per-GPU critical path of the strong-scaling run without paying for N GPUs).
Reports per option combination: wall time of the synchronous call (host clock), device total (ev0..ev1), render kernel.
usage: [WORLD=8] [TILE_ROWS=8] [ITERS=12] [WORKLOAD=config3] python tools/ab_frame.py frame_graph=0,1 l2_prefetch=0,1"""
import hashlib
import itertools
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rt = bench.entry.load_rtds()
wl = bench.workloads(rt)[os.environ.get("WORKLOAD", "config3")]
world, tile_rows, iters = int(os.environ.get("WORLD", "1")), int(os.environ.get("TILE_ROWS", "8")), int(os.environ.get("ITERS", "12"))
opts = [a.split("=") for a in sys.argv[1:]]
names = [o[0] for o in opts]
values = [[int(v) for v in o[1].split(",")] for o in opts]
ctx = rt.Rtds(0)
sph, mat = wl.scene()
ctx.set_spheres(sph, mat)
if wl.lights is not None:
    ctx.set_lights(wl.lights)
ctx.build(wl.acc, mode=wl.mode, **wl.build_kw)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows_max = max(rt.rows_for_rank(wl.H, tile_rows, r, world) for r in range(world))
buf = torch.zeros((rows_max, wl.W, 3), dtype=torch.uint8, device=dev)
ref_sha = {}
for rep in range(2):
    for combo in itertools.product(*values):
        for n, v in zip(names, combo):
            ctx.set_option(n, v)
        per_rank = []
        for r in range(world):
            p = ctx.render_params(wl.W, wl.H, wl.spp, rank=r, world=world, tile_rows=tile_rows, shadows=wl.shadows)
            wall, tot, ker = [], [], []
            for it in range(iters):
                flush.fill_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                st = ctx.render_device(wl.acc, p, buf.data_ptr())
                t1 = time.perf_counter()
                if it >= 3:
                    wall.append((t1 - t0) * 1e3); tot.append(st["ms_total"]); ker.append(st["ms_kernel"])
            h = hashlib.md5(buf[: rt.rows_for_rank(wl.H, tile_rows, r, world)].cpu().numpy().tobytes()).hexdigest()
            same = ref_sha.setdefault(r, h) == h
            per_rank.append((float(np.median(wall)), float(np.median(tot)), float(np.median(ker)), same, st["kernel_launches"]))
        print(json.dumps({"options": dict(zip(names, combo)), "world": world, "tile_rows": tile_rows,
                          "wall_ms_max": round(max(x[0] for x in per_rank), 4), "device_total_ms_max": round(max(x[1] for x in per_rank), 4),
                          "kernel_ms_max": round(max(x[2] for x in per_rank), 4), "kernel_ms_mean": round(float(np.mean([x[2] for x in per_rank])), 4),
                          "frames_identical": all(x[3] for x in per_rank), "launches": per_rank[0][4]}), flush=True)
ctx.close()
