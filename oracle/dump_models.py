"""TEST INFRASTRUCTURE. Build container only: parses the reference's vertex-only OBJ models with the reference's loader
semantics, writes their sphere tables to oracle/_ref/models/<name>.f32 (git-ignored, travels to the GPU box) and the
sha256 of the UNMODIFIED reference's BVH / LBVH trees on them to tests/golden/model_trees.json (committed)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as T  # noqa: E402

rt = T.rtds_b200
MODELS = ["Igea", "armadillo", "dragon", "lucy", "teapot", "woody"]


def tree_sha(nodes, order):
    return hashlib.sha256(nodes.tobytes() + np.ascontiguousarray(order, np.int32).tobytes()).hexdigest()


def main():
    ref, oracle = T.Ref(), T.Oracle()
    out_dir = os.path.join(ROOT, "oracle", "_ref", "models")
    os.makedirs(out_dir, exist_ok=True)
    G = {}
    for name in MODELS:
        v = rt.parse_obj_vertices(os.path.join(T.REF_TREE, "models", name + ".obj"))
        sph, mat = rt.scene_from_vertices(v, 1)
        sph.tofile(os.path.join(out_dir, name + ".f32"))
        e = {"n": int(sph.shape[0]), "scene_sha256": hashlib.sha256(sph.tobytes()).hexdigest()}
        for acc, key, n_use in ((rt.BVH, "BVH", sph.shape[0]), (rt.LBVH, "LBVH", sph.shape[0] - 1)):
            rc = oracle.build_bvh(sph, n_use)[0]
            if rc != 0:                      # the reference would recurse forever / misbehave: do not call it
                e[key] = {"reference_status": int(rc)}
                continue
            ref.scene_from_spheres(sph, mat)
            total, secs = ref.build(acc)
            nodes, objs, _ = ref.bvh_linear()
            e[key] = {"total_nodes": int(total), "n_leaves": int(len(objs)), "tree_sha256": tree_sha(nodes, objs), "ref_build_s": secs}
        G[name] = e
        print(name, e, flush=True)
    # the committed golden file only changes when the trees do: a re-run (fresh checkout, build()) keeps the reference build
    # times that are already recorded instead of rewriting them with this machine's
    path = os.path.join(ROOT, "tests", "golden", "model_trees.json")
    if os.path.exists(path):
        with open(path) as f:
            old = json.load(f)
        strip = lambda g: {m: {k: ({kk: vv for kk, vv in v.items() if kk != "ref_build_s"} if isinstance(v, dict) else v) for k, v in e.items()}
                           for m, e in g.items()}
        if strip(old) == strip(G):
            return
    with open(path, "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
