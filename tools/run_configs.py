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
LIGHTS3 = np.asarray([[0, 3, 30, 10, 1, 1, 1], [20, 30, -40, 1, 0.5, 0.4, 0.3], [-30, 5, -70, 1, 0.2, 0.3, 0.6]], np.float32)


def torus_knot_scene(n, seed=7):
    """config 4 stand-in: n spheres on a displaced (2,3) torus-knot tube inside the camera frustum."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 2 * np.pi, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    R, r, tube = 9.0, 3.5, 1.2 + 0.25 * np.sin(7 * t)
    cx = (R + r * np.cos(3 * t)) * np.cos(2 * t)
    cy = (R + r * np.cos(3 * t)) * np.sin(2 * t)
    cz = r * np.sin(3 * t)
    # tube cross-section in a frame that is good enough for a point cloud
    nx, ny, nz = np.cos(2 * t) * np.cos(phi), np.sin(2 * t) * np.cos(phi), np.sin(phi)
    p = np.stack([cx + tube * nx, cy + tube * ny, cz + tube * nz], 1) + rng.normal(size=(n, 3)) * 0.01
    sph = np.zeros((n + 1, 4), np.float32)
    sph[:n, :3] = p.astype(np.float32) * np.float32(0.9) + np.asarray([0, 0, -70], np.float32)
    sph[:n, 3] = 0.02
    sph[n] = np.asarray(rt.GROUND, np.float32)
    mat = np.zeros_like(sph)
    mat[:n, 0], mat[:n, 1] = 0.8, 0.7
    return sph, mat


def city_trees_scene(seed=5):
    """config 5 stand-in: 90,811 'city' prims (axis-aligned boxes of points) + 252,178 'tree' prims (clustered blobs),
    10 % REFLECTION_AND_REFRACTION and 10 % REFLECTION materials."""
    rng = np.random.default_rng(seed)
    n_city, n_tree = 90811, 252178
    b = rng.integers(0, 60, n_city)
    origin = np.stack([(b % 10) * 4.0 - 20, np.full(60, -8.0)[b], -(b // 10) * 6.0 - 50], 1)
    size = np.stack([rng.uniform(1, 3, 60), rng.uniform(2, 14, 60), rng.uniform(1, 3, 60)], 1)[b]
    city = origin + rng.uniform(0, 1, (n_city, 3)) * size
    k = rng.integers(0, 300, n_tree)
    tc = np.stack([rng.uniform(-25, 25, 300), rng.uniform(-6, 2, 300), rng.uniform(-95, -45, 300)], 1)[k]
    trees = tc + rng.normal(size=(n_tree, 3)) * rng.uniform(0.3, 1.2, (300, 1))[k]
    p = np.concatenate([city, trees]).astype(np.float32)
    n = p.shape[0]
    sph = np.zeros((n + 1, 4), np.float32)
    sph[:n, :3] = p
    sph[:n, 3] = 0.05
    sph[n] = np.asarray(rt.GROUND, np.float32)
    mat = np.zeros_like(sph)
    mat[:n, :3] = rng.uniform(0.1, 0.9, (n, 3)).astype(np.float32)
    u = rng.uniform(size=n)
    mat[:n, 3] = np.where(u < 0.1, 1.0, np.where(u < 0.2, 2.0, 0.0))
    return sph, mat


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
