"""Applies the librtds.so binding (integration/rtds_binding.inc) to a COPY of the reference's main.cpp.

    python integration/apply_patch.py /path/to/reference/project/raytracer/main.cpp /tmp/main_rtds.cpp
    g++ -O3 -w -include cstdint -include memory -I include -I integration -I /path/to/reference/project/raytracer \\
        -o output_rtds /tmp/main_rtds.cpp -L raytracer-data-structures_b200 -lrtds -Wl,-rpath,$PWD/raytracer-data-structures_b200

Six textual edits, each replacing one reference call site (main.cpp line numbers of the reference as surveyed):
    :723  int main(                       <- #include "rtds_binding.inc" inserted above it
    :751  scene = createScene_new(...)    <- + rtds_upload_scene(scene);
    :800  root = constructBVHNew(...)     -> totalNodes = rtds_build_for(BVH);
    :816  constructKDTreeNew(...)         -> totalKdNodes = rtds_build_for(KDTREE);
    :832  root = constructLBVHTree(...)   -> rtds_build_for(LBVH);
    :808,:824,:835,:840  render(...)      -> rtds_render_frame(settings, lights);
The patched file is an OUTPUT (never committed: it holds the reference's source). tests/test_integration_patch.py compiles and
links it; oracle/Makefile builds it into oracle/_ref/ref_patched_main for the GPU box, where tests/test_gpu_host_main.py runs it.
"""
import re
import sys


def apply(src: str) -> str:
    def sub(pattern, repl, text, count=1, flags=re.S):
        out, n = re.subn(pattern, repl, text, count=0 if count == 0 else count, flags=flags)
        if n == 0:
            raise SystemExit(f"apply_patch: call site not found: {pattern}")
        return out

    s = src
    s = sub(r"\nint main\(int argc", '\n#include "rtds_binding.inc"\n\nint main(int argc', s)
    s = sub(r"(std::vector<SceneObject> scene = createScene_new\(settings\);)", r"\1\n\trtds_upload_scene(scene);", s)
    s = sub(r"root = constructBVHNew\(scene, 0, scene\.size\(\),\s*&totalNodes\);", "totalNodes = rtds_build_for(BVH);", s)
    s = sub(r"constructKDTreeNew\(\s*scene,\s*80, 1, 0\.5f,\s*1, -1\s*\);", "totalKdNodes = rtds_build_for(KDTREE);", s)
    s = sub(r"root = constructLBVHTree\(\s*scene, root, nodes\);", "rtds_build_for(LBVH);", s)
    s = sub(r"render\(settings, spheres, lights, triangles, frame, scene, root\);", "rtds_render_frame(settings, lights);", s, count=0)
    return s


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    with open(sys.argv[1]) as f:
        patched = apply(f.read())
    with open(sys.argv[2], "w") as f:
        f.write(patched)
