"""Runs every BASELINE.json config (or its stand-in) on ONE B200 and prints one JSON line per case:
build ms, render-kernel ms, Mrays/s (all rays traced: primary + shadow + secondary). Results go to profiles/.
The bench line (config 3) is bench.py's job; this script is the side table."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

rt = g.load_rtds()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = np.fromfile(os.path.join(ROOT, "tests", "golden", "bunny_vertices.f32"), np.float32).reshape(-1, 3)
import rtds_b200.standins as standins  # noqa: E402
from rtds_b200.standins import LIGHTS3, torus_knot_scene, city_trees_scene  # noqa: E402,F401


def run(ctx, name, acc, mode, W, H, spp, shadows=0, reps=int(os.environ.get("REPS", "3")), **bk):
    st = [ctx.build(acc, mode=mode, **bk) for _ in range(reps)][-1] if acc != rt.NONE else {"ms": 0.0, "total_nodes": 0, "kernel_launches": 0}
    # the GPU clocks ramp up over the first few hundred ms of work: repeat until the kernel time settles
    rs, prev, best = None, None, None
    for i in range(400):
        _, _, _, rs = ctx.render(acc, W, H, spp, shadows=shadows)
        t = rs["ms_kernel"]
        best = rs if best is None or t < best["ms_kernel"] else best
        if i >= reps and prev is not None and abs(t - prev) <= 0.01 * t and t <= 1.02 * best["ms_kernel"]:
            break
        prev = t
    rs = best
    n = ctx.n
    line = {"case": name, "n_prims": int(n), "width": W, "height": H, "aa_samples": spp, "shadows": shadows,
            "build_ms": round(st["ms"], 4), "build_ms_per_mprim": round(st["ms"] / (n / 1e6), 4), "nodes": st["total_nodes"],
            "render_kernel_ms": round(rs["ms_kernel"], 4), "render_total_ms": round(rs["ms_total"], 4), "rays": int(rs["rays"]),
            "primary": int(rs["primary_rays"]), "shadow": int(rs["shadow_rays"]), "secondary": int(rs["secondary_rays"]),
            "mrays_per_s_kernel": round(rs["rays"] / (rs["ms_kernel"] * 1e-3) / 1e6, 1),
            "mrays_per_s_total": round(rs["rays"] / (rs["ms_total"] * 1e-3) / 1e6, 1),
            "slab_tests_per_ray": round(rs["node_tests"] / max(rs["rays"], 1), 2), "prim_tests_per_ray": round(rs["prim_tests"] / max(rs["rays"], 1), 3)}
    print(json.dumps(line), flush=True)
    return line


def main():
    ctx = rt.Rtds(0)
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2", "3", "4", "5"]
    out = []
    if "1" in which:
        sph, mat = rt.scene_from_vertices(V, 1)
        ctx.set_spheres(sph, mat)
        out.append(run(ctx, "config1: bunny 640x480 aa1 BVH (settings.h defaults)", rt.BVH, rt.MODE_COMPAT, 640, 480, 1))
        out.append(run(ctx, "config1: same, NONE brute force", rt.NONE, 0, 640, 480, 1))
    if "2" in which:
        sph, mat = rt.scene_from_vertices(V, 1)
        ctx.set_spheres(sph, mat)
        for sh in (0, 1):
            out.append(run(ctx, "config2: bunny 1920x1080 LBVH(true)", rt.LBVH, rt.MODE_TRUE, 1920, 1080, 1, shadows=sh))
            out.append(run(ctx, "config2: bunny 1920x1080 BVH(median, bit-exact)", rt.BVH, rt.MODE_COMPAT, 1920, 1080, 1, shadows=sh))
            out.append(run(ctx, "config2: bunny 1920x1080 BVH(SAH)", rt.BVH, rt.MODE_SAH, 1920, 1080, 1, shadows=sh))
        out.append(run(ctx, "config2: bunny 1920x1080 KDTREE (any-hit, unshaded like the reference)", rt.KDTREE, 0, 1920, 1080, 1))
    if "3" in which:
        sph, mat = rt.scene_from_vertices(V, 30)
        ctx.set_spheres(sph, mat)
        for sh in (0, 1):
            out.append(run(ctx, "config3: bunny x30 clones 3840x2160 aa4 LBVH(true)", rt.LBVH, rt.MODE_TRUE, 3840, 2160, 4, shadows=sh))
        out.append(run(ctx, "config3: same, BVH(median, bit-exact)", rt.BVH, rt.MODE_COMPAT, 3840, 2160, 4))
    if "3v" in which or "3" in which:
        # the VISIBLE variant of config 3 - clone shift 2 instead of 20, every clone inside the view
        sph, mat = rt.scene_from_vertices(V, 30, clone_shift=2)
        ctx.set_spheres(sph, mat)
        for sh in (0, 1):
            out.append(run(ctx, "config3-visible: bunny x30 clones, clone shift 2 (all in view), 3840x2160 aa4 LBVH(true)", rt.LBVH, rt.MODE_TRUE,
                           3840, 2160, 4, shadows=sh))
    if "4" in which:
        sph, mat = torus_knot_scene(7_000_000)
        ctx.set_spheres(sph, mat)
        out.append(run(ctx, "config4: 7M-prim torus-knot stand-in 3840x2160 aa1 BVH(SAH)", rt.BVH, rt.MODE_SAH, 3840, 2160, 1))
        out.append(run(ctx, "config4: same, LBVH(true, 63-bit)", rt.LBVH, rt.MODE_TRUE, 3840, 2160, 1, morton_bits=63))
        out.append(run(ctx, "config4: same, LBVH(true, 30-bit)", rt.LBVH, rt.MODE_TRUE, 3840, 2160, 1))
    if "5" in which:
        sph, mat = city_trees_scene()
        ctx.set_spheres(sph, mat)
        ctx.set_lights(LIGHTS3)
        out.append(run(ctx, "config5: city+trees stand-in 342,990 prims, 3 lights, shadows, 10% R&R + 10% REFLECTION, 3840x2160 aa16, LBVH(true)",
                       rt.LBVH, rt.MODE_TRUE, 3840, 2160, 16, shadows=1, reps=2))
        out.append(run(ctx, "config5: same, BVH(SAH)", rt.BVH, rt.MODE_SAH, 3840, 2160, 16, shadows=1, reps=2))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
