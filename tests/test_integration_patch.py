"""INTEGRATION.md section B, compiled: integration/apply_patch.py applies the librtds.so binding (integration/rtds_binding.inc) to a
copy of the reference's own main.cpp, and the result compiles and links against librtds.so. CPU-only (no kernel runs here); the
GPU box runs the resulting binary in tests/test_gpu_host_main.py. Skipped where the reference tree is absent."""
import os
import subprocess
import sys

import pytest

import conftest as T

sys.path.insert(0, os.path.join(T.ROOT, "integration"))
import apply_patch  # noqa: E402

REF_MAIN = os.path.join(T.REF_TREE, "main.cpp")


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="/root/reference not present")
def test_binding_applies_compiles_and_links(tmp_path):
    lib = os.path.join(T.PKG_DIR, "librtds.so")
    if not os.path.exists(lib):
        pytest.skip("librtds.so not built")
    src = open(REF_MAIN).read()
    patched = apply_patch.apply(src)
    # all five call sites + the include are in, nothing of the reference's build/render calls is left in main()
    main_body = patched[patched.index("int main(int argc"):]
    assert patched.count('#include "rtds_binding.inc"') == 1 and "rtds_upload_scene(scene);" in main_body
    assert main_body.count("rtds_render_frame(settings, lights);") == 4
    for gone in ("constructBVHNew(", "constructLBVHTree(", "constructKDTreeNew(", "render(settings"):
        assert gone not in main_body, gone
    for call in ("rtds_build_for(BVH)", "rtds_build_for(KDTREE)", "rtds_build_for(LBVH)"):
        assert call in main_body
    out_cpp, out_bin = tmp_path / "main_rtds.cpp", tmp_path / "output_rtds"
    out_cpp.write_text(patched)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O1", "-w", "-include", "cstdint", "-include", "memory", "-I", os.path.join(T.ROOT, "include"),
           "-I", os.path.join(T.ROOT, "integration"), "-I", T.REF_TREE, "-o", str(out_bin), str(out_cpp),
           "-L", T.PKG_DIR, "-lrtds", "-Wl,-rpath," + T.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # the binary's undefined symbols that librtds.so must provide are exactly the binding's calls
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(out_bin)], capture_output=True, text=True).stdout
    used = sorted({l.split()[-1] for l in nm.splitlines() if " rtds_" in l})
    assert used == ["rtds_build", "rtds_create", "rtds_last_error", "rtds_render", "rtds_set_lights", "rtds_set_spheres"]
    # without a GPU the patched program must fail loudly through the binding (no CPU fallback), not crash
    if not T.has_gpu():
        os.makedirs(tmp_path / "models")
        with open(tmp_path / "models" / "bunny.obj", "w") as f:
            for x, y, z in T.bunny_vertices()[:100]:
                f.write("v %.9g %.9g %.9g\n" % (x, y, z))
        r = subprocess.run([str(out_bin)], cwd=tmp_path, capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "rtds_create" in r.stderr and "no CUDA device" in r.stderr
