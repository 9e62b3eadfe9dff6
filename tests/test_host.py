"""CPU: host-side logic — the C++ OBJ loader / scene creation (host/scene.cpp) against the numpy restatement, the
golden scene, and (in the build container) the unmodified reference's createScene_new."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
G = T.load_golden_json()


def _host():
    if not os.path.exists(rt.HOST_LIB_PATH):
        pytest.skip("librtds_host.so not built")
    return C.CDLL(rt.HOST_LIB_PATH)


def _load(lib, d, model, clones, cap):
    sph = np.zeros((cap, 4), np.float32)
    mat = np.zeros((cap, 4), np.float32)
    n = lib.rtds_host_scene_from_obj(str(d).encode(), model, clones, sph.ctypes.data_as(C.c_void_p), mat.ctypes.data_as(C.c_void_p), cap)
    return n, sph[:max(n, 0)], mat[:max(n, 0)]


def test_cpp_loader_matches_golden_scene(tmp_path):
    lib = _host()
    v = T.bunny_vertices()
    with open(tmp_path / "bunny.obj", "w") as f:
        for x, y, z in v:
            f.write("v %.9g %.9g %.9g\n" % (x, y, z))
        f.write("vn 0 0 1\nv 9 9 9\n")           # the reference's loader stops at the first non-"v" token
    for clones in (1, 3):
        n, sph, mat = _load(lib, tmp_path, rt.BUNNY, clones, v.shape[0] * clones + 1)
        assert n == v.shape[0] * clones + 1
        sph_n, mat_n = rt.scene_from_vertices(v, clones)
        assert sph.tobytes() == sph_n.tobytes() and mat.tobytes() == mat_n.tobytes()
        if clones == 1:
            assert hashlib.sha256(sph.tobytes()).hexdigest() == G["bunny"]["scene_sha256"]


def test_cpp_loader_missing_model_is_an_error(tmp_path):
    lib = _host()
    n, _, _ = _load(lib, tmp_path, rt.BUNNY, 1, 10)
    assert n == -1


def test_python_parser_stops_like_the_reference(tmp_path):
    with open(tmp_path / "m.obj", "w") as f:
        f.write("v 1 2 3\nv 4 5 6\n# comment\nv 7 8 9\n")
    assert rt.parse_obj_vertices(str(tmp_path / "m.obj")).tolist() == [[1, 2, 3], [4, 5, 6]]


@pytest.mark.parametrize("model,clones", [(rt.BUNNY, 1), (rt.BUNNY, 2), (rt.ARMADILLO, 1), (rt.IGEA, 1)])
def test_cpp_loader_matches_reference_createScene(ref, model, clones):
    """Build container only: the reference's own loader on its own model files."""
    lib = _host()
    if model == rt.IGEA:
        pytest.skip("the reference opens igea.obj but ships Igea.obj (fails on case-sensitive file systems)")
    sph_r, mat_r = ref.scene_from_obj(model, clones)
    n, sph, mat = _load(lib, os.path.join(T.REF_TREE, "models"), model, clones, sph_r.shape[0])
    assert n == sph_r.shape[0]
    assert sph.tobytes() == sph_r.tobytes() and mat.tobytes() == mat_r.tobytes()
