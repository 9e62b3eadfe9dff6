"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref/libref_oracle.so, compiled from
/root/reference where it lies) — run in the build container only:  python tests/golden/make_golden.py

Outputs (committed):
  bunny_vertices.f32        the 35,947 `v` triples of models/bunny.obj as parsed by the reference's loader
                            semantics (main.cpp:663-698); input fixture for the GPU box, where /root/reference
                            does not exist
  golden.json               PPM md5s / node counts / sphere-test counts of the reference's four dataStructure
                            settings on the default config, Morton KATs, jitter KATs, sha256 of the reference's
                            BVH and LBVH trees (pre-order LinearBVHNode bytes + primitive order) on several scenes
  bunny_hits_640x480.npz    per-pixel hit objId of the reference BVH path and of its NONE brute-force loop,
                            and the BVH path's tnear, default config
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest as T  # noqa: E402

rt = T.rtds_b200


def tree_sha(nodes, order):
    return hashlib.sha256(nodes.tobytes() + np.ascontiguousarray(order, np.int32).tobytes()).hexdigest()


def quantised_scene(n, seed, q):
    """Heavy coordinate ties: centres snapped to a grid of step q (exercises libstdc++'s tie handling)."""
    sph, mat = T.synthetic_scene(n, seed)
    sph[:n, :3] = np.round(sph[:n, :3] / np.float32(q)) * np.float32(q)
    return sph, mat


def golden_kd(ref):
    """KDTREE on the default config: the any-hit mask (pixel is black <=> kdtreeIntersect returned true)."""
    sph, mat = T.bunny_scene()
    ref.scene_from_spheres(sph, mat)
    total, _ = ref.build(rt.KDTREE)
    rgb, dirs, _, _ = ref.render_rows(rt.KDTREE, 640, 480, 1, want_dirs=True)
    mask = (rgb.astype(np.int32).sum(-1) == 0)
    h, _, _ = ref.trace(rt.KDTREE, np.zeros((1, 3), np.float32), dirs.reshape(-1, 3))
    assert np.array_equal(h.reshape(480, 640) > 0, mask)
    np.save(os.path.join(HERE, "bunny_kd_mask_640x480.npy"), np.packbits(mask))
    return {"total_nodes": int(total), "hits": int(mask.sum()), "ppm_md5": T.ppm_md5(rgb)}


def main():
    ref = T.Ref()
    if len(sys.argv) > 1 and sys.argv[1] == "kd":
        with open(os.path.join(HERE, "golden.json")) as f:
            G = json.load(f)
        G["kd"] = golden_kd(ref)
        print(G["kd"])
        with open(os.path.join(HERE, "golden.json"), "w") as f:
            json.dump(G, f, indent=1, sort_keys=True)
        return
    G = {}
    v = rt.parse_obj_vertices(os.path.join(T.REF_TREE, "models", "bunny.obj"))
    sph_ref, mat_ref = ref.scene_from_obj(rt.BUNNY, 1)
    sph, mat = rt.scene_from_vertices(v, 1)
    assert sph.tobytes() == sph_ref.tobytes() and mat.tobytes() == mat_ref.tobytes(), "loader restatement != reference"
    v.tofile(os.path.join(HERE, "bunny_vertices.f32"))
    G["bunny"] = {"n_vertices": int(v.shape[0]), "n_prims": int(sph.shape[0]),
                  "vertices_sha256": hashlib.sha256(v.tobytes()).hexdigest(),
                  "scene_sha256": hashlib.sha256(sph.tobytes()).hexdigest()}

    W, H = 640, 480
    names = {rt.BVH: "BVH", rt.LBVH: "LBVH", rt.KDTREE: "KDTREE", rt.NONE: "NONE"}
    G["default_config"] = {}
    hits = {}
    for acc in (rt.BVH, rt.LBVH, rt.KDTREE, rt.NONE):
        total, secs = ref.build(acc)
        ref.lib.ref_reset_counters()
        rgb, dirs, _, rsecs = ref.render_rows(acc, W, H, 1, want_dirs=(acc == rt.BVH))
        tests = int(ref.lib.ref_sphere_tests())
        entry = {"total_nodes": int(total), "ppm_md5": T.ppm_md5(rgb), "sphere_tests_int32": int(np.int32(np.uint32(tests & 0xffffffff)))}
        if acc in (rt.BVH, rt.LBVH):
            nodes, objs, order = ref.bvh_linear()
            entry["tree_sha256"] = tree_sha(nodes, objs)
            entry["n_leaves"] = int(len(objs))
            assert np.array_equal(objs, order[:len(objs)])
        if acc == rt.BVH:
            primary = dirs.reshape(-1, 3)
            h, t, cand = ref.trace(rt.BVH, np.zeros((1, 3), np.float32), primary)
            hits["hit_bvh"], hits["t_bvh"] = h.reshape(H, W), t.reshape(H, W)
            entry["candidates"] = int(cand)
            np.save(os.path.join(HERE, "bunny_bvh_prim_order.npy"), order.astype(np.int32))
        if acc == rt.NONE:
            h, t, _ = ref.trace(rt.NONE, np.zeros((1, 3), np.float32), primary)
            hits["hit_none"] = h.reshape(H, W)
        G["default_config"][names[acc]] = entry
        print(names[acc], entry, flush=True)
    np.savez_compressed(os.path.join(HERE, "bunny_hits_640x480.npz"), **hits)

    # Morton KATs (accelerators.h:374-394)
    G["morton"] = {"expandBits_1023": int(ref.lib.ref_expandBits(1023)),
                   "half": int(ref.lib.ref_morton3D(T.C.c_float(.5), T.C.c_float(.5), T.C.c_float(.5))),
                   "ones": int(ref.lib.ref_morton3D(T.C.c_float(1), T.C.c_float(1), T.C.c_float(1))),
                   "bunny_v0_refnorm": int(ref.lib.ref_morton3D(*[T.C.c_float(float((np.float32(x) + np.float32(30)) / np.float32(1000))) for x in sph[0, :3]]))}
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.1, 1.1, size=(64, 3)).astype(np.float32)
    G["morton"]["points_seed5"] = [int(ref.lib.ref_morton3D(T.C.c_float(float(p[0])), T.C.c_float(float(p[1])), T.C.c_float(float(p[2])))) for p in pts]

    # jitter KATs (main.cpp:503-508)
    j = ref.jitter(1000004)
    G["jitter"] = {"first8_hex": [float(x).hex() for x in j[:8]], "at_1000000_hex": [float(x).hex() for x in j[1000000:1000004]],
                   "sha256_first_1e6": hashlib.sha256(j[:1000000].tobytes()).hexdigest()}

    # reference BVH trees on other scenes (tie-heavy ones included)
    G["trees"] = {}
    scenes = {"bunny_clones3": rt.scene_from_vertices(v, 3),
              "synthetic_2000_seed3": T.synthetic_scene(2000, 3),
              "quantised_3000_seed4_q0.5": quantised_scene(3000, 4, 0.5),
              "quantised_5000_seed6_q2": quantised_scene(5000, 6, 2.0),
              "synthetic_17_seed9": T.synthetic_scene(17, 9),
              "synthetic_2_seed1": T.synthetic_scene(1, 1)}
    for model, name in ((rt.ARMADILLO, "armadillo"), (rt.IGEA, None)):
        if name:
            s_, m_ = ref.scene_from_obj(model, 1)
            scenes[name] = (s_, m_)
    for name, (s_, m_) in scenes.items():
        ref.scene_from_spheres(s_, m_)
        e = {"n": int(s_.shape[0]), "scene_sha256": hashlib.sha256(s_.tobytes()).hexdigest()}
        for acc in (rt.BVH, rt.LBVH):
            if acc == rt.LBVH and s_.shape[0] < 3:
                continue
            total, _ = ref.build(acc)
            nodes, objs, order = ref.bvh_linear()
            e[names[acc]] = {"total_nodes": int(total), "n_leaves": int(len(objs)), "tree_sha256": tree_sha(nodes, objs)}
        G["trees"][name] = e
        print(name, e, flush=True)

    G["kd"] = golden_kd(ref)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
