"""Simulates the N-GPU split of the bench frame on ONE B200: renders rank r of world N (resident scene + tree, L2 flushed
before every launch) for several tile heights and prints the slowest rank's kernel / total time — the per-GPU critical
path of the strong-scaling run without paying for N GPUs.  python tools/rank_sim.py [world] [tile_rows,...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

rt = g.load_rtds()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = np.fromfile(os.path.join(ROOT, "tests", "golden", "bunny_vertices.f32"), np.float32).reshape(-1, 3)
W, H, SPP = 3840, 2160, 4


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    tiles = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32, 64]
    flush_on = os.environ.get("FLUSH", "1") != "0"
    ctx = rt.Rtds(0)
    sph, mat = rt.scene_from_vertices(V, 30)
    ctx.set_spheres(sph, mat)
    ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for tr in tiles:
        rows_max = max(rt.rows_for_rank(H, tr, r, world) for r in range(world))
        buf = torch.zeros((rows_max, W, 3), dtype=torch.uint8, device=dev)
        per_rank = []
        for r in range(world):
            p = ctx.render_params(W, H, SPP, rank=r, world=world, tile_rows=tr)
            ks, ts = [], []
            for it in range(int(os.environ.get("ITERS", "6"))):
                if flush_on:
                    flush.fill_(1)
                torch.cuda.synchronize()
                st = ctx.render_device(rt.LBVH, p, buf.data_ptr())
                if it >= 2:
                    ks.append(st["ms_kernel"]); ts.append(st["ms_total"])
            per_rank.append((float(np.mean(ks)), float(np.mean(ts))))
        if os.environ.get("PER_RANK"):
            print(json.dumps({"tile_rows": tr, "kernel_ms_per_rank": [round(a, 4) for a, _ in per_rank]}), flush=True)
        k = [a for a, _ in per_rank]
        t = [b for _, b in per_rank]
        print(json.dumps({"world": world, "tile_rows": tr, "flush": flush_on, "kernel_ms_max": round(max(k), 4), "kernel_ms_min": round(min(k), 4),
                          "kernel_ms_mean": round(float(np.mean(k)), 4), "total_ms_max": round(max(t), 4)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
