"""Shared test plumbing: loads the hyphen-named package as `rtds_b200`, the CPU oracles via ctypes, golden data.

Markers: `gpu` = needs a B200 (run by the driver on the GPU box with `-m gpu`); everything else runs on CPU.
Only tests (and smoke / bench's cpu_baseline) may touch oracle/ — the product never does.
"""
import ctypes as C
import hashlib
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "raytracer-data-structures_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_TREE = "/root/reference/project/raytracer"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "ref: needs the compiled reference oracle/_ref/libref_oracle.so")


def load_rtds():
    if "rtds_b200" in sys.modules:
        return sys.modules["rtds_b200"]
    spec = importlib.util.spec_from_file_location("rtds_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rtds_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


rtds_b200 = load_rtds()
LINEAR = rtds_b200.LINEAR_NODE_DTYPE
WIDE4 = np.dtype([("bmin", np.float32, (4, 3)), ("bmax", np.float32, (4, 3)), ("child", np.int32, 4), ("n_children", np.int32),
                  ("binary_node", np.int32), ("pad", np.int32, 2)])
assert WIDE4.itemsize == 128


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------------
# oracle port (oracle/liboracle.so) — built on demand (g++ only)
# ---------------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ROOT, "oracle", "oracle.cpp")):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(path)

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def scene_from_vertices(self, v, clones=1):
        v = np.ascontiguousarray(v, np.float32).reshape(-1, 3)
        n = v.shape[0] * clones + 1
        sph = np.zeros((n, 4), np.float32)
        mat = np.zeros((n, 4), np.float32)
        got = self.lib.orc_scene_from_vertices(self._p(v), v.shape[0], clones, self._p(sph), self._p(mat))
        assert got == n
        return sph, mat

    def build_bvh(self, sph, n_use=None, prim_type=0):
        sph = np.ascontiguousarray(sph, np.float32)
        n_use = sph.shape[0] if n_use is None else n_use
        nodes = np.zeros(2 * n_use - 1, LINEAR)
        order = np.zeros(n_use, np.int32)
        nn, md = C.c_int(), C.c_int()
        rc = self.lib.orc_build_bvh_p(self._p(sph), prim_type, n_use, self._p(nodes), self._p(order), C.byref(nn), C.byref(md))
        if rc != 0:
            return rc, None, None, 0
        nodes = nodes[:nn.value]
        # Export convention: primitivesOffset = rank of the leaf in DFS order, prim_order = the leaves' objIds in
        # that order. They coincide with positions in the reordered scene vector unless the reference dropped a
        # range (accelerators.h:321-327: `startIndex == midSplitIndex` keeps the first object only).
        leaf = nodes["nPrimitives"] > 0
        order = order[nodes["offset"][leaf]]
        nodes["offset"][leaf] = np.arange(order.shape[0], dtype=np.int32)
        return 0, nodes, order, md.value

    def build_lbvh(self, sph, bits=30, ref_norm=0, prim_type=0):
        sph = np.ascontiguousarray(sph, np.float32)
        n = sph.shape[0]
        nodes = np.zeros(2 * n - 1, LINEAR)
        order = np.zeros(n, np.int32)
        keys = np.zeros(n, np.uint64)
        nn, md = C.c_int(), C.c_int()
        rc = self.lib.orc_build_lbvh_p(self._p(sph), prim_type, n, bits, ref_norm, self._p(nodes), self._p(order), self._p(keys),
                                       C.byref(nn), C.byref(md))
        assert rc == 0
        return nodes[:nn.value], order, keys, md.value

    def build_sah(self, sph, bins=16, prim_type=0):
        sph = np.ascontiguousarray(sph, np.float32)
        n = sph.shape[0]
        nodes = np.zeros(2 * n - 1, LINEAR)
        order = np.zeros(n, np.int32)
        nn, md = C.c_int(), C.c_int()
        self.lib.orc_build_sah_p(self._p(sph), prim_type, n, bins, self._p(nodes), self._p(order), C.byref(nn), C.byref(md))
        return nodes[:nn.value], order, md.value

    def morton30(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        codes = np.zeros(xyz.shape[0], np.uint32)
        self.lib.orc_morton30(self._p(xyz), xyz.shape[0], self._p(codes))
        return codes

    def trace(self, sph, nodes, order, o, d, tie_by_objid=0, prim_type=0):
        sph = np.ascontiguousarray(sph, np.float32)
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = d.shape[0]
        if o.shape[0] == 1 and n > 1:
            o = np.ascontiguousarray(np.broadcast_to(o, (n, 3)))
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        cand = C.c_longlong()
        self.lib.orc_trace_p(self._p(sph), prim_type, sph.shape[0], self._p(nodes), self._p(order), 0 if nodes is None else nodes.shape[0],
                             tie_by_objid, self._p(o), self._p(d), n, self._p(hit), self._p(t), C.byref(cand))
        return hit, t, cand.value

    def kd_closest(self, prims, kd_nodes, kd_idx, bounds, o, d, prim_type=0):
        """Closest-hit KD traversal (extension) over an exported KdAccelNode[]: (hit objId | -1, tnear, prim tests)."""
        prims = np.ascontiguousarray(prims, np.float32)
        kd_nodes = np.ascontiguousarray(kd_nodes)
        kd_idx = np.ascontiguousarray(kd_idx, np.int32)
        bounds = np.ascontiguousarray(bounds, np.float32)
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = d.shape[0]
        if o.shape[0] == 1 and n > 1:
            o = np.ascontiguousarray(np.broadcast_to(o, (n, 3)))
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        tests = C.c_longlong()
        self.lib.orc_kd_closest(self._p(prims), prim_type, prims.shape[0], self._p(kd_nodes), self._p(kd_idx), self._p(bounds),
                                self._p(o), self._p(d), n, self._p(hit), self._p(t), C.byref(tests))
        return hit, t, tests.value

    def collapse4(self, nodes, rule=0):
        """4-wide collapse of a LinearBVHNode[] (oracle.cpp: Wide4Node, 128 bytes). rule 0 = the definition (largest area)."""
        nodes = np.ascontiguousarray(nodes)
        n = self.lib.orc_collapse4_rule(self._p(nodes), nodes.shape[0], None, 0, rule)
        wide = np.zeros(n, WIDE4)
        assert self.lib.orc_collapse4_rule(self._p(nodes), nodes.shape[0], self._p(wide), n, rule) == n
        return wide

    def trace_wide4(self, prims, wide, order, o, d, tie_by_objid=0, prim_type=0):
        prims = np.ascontiguousarray(prims, np.float32)
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = d.shape[0]
        if o.shape[0] == 1 and n > 1:
            o = np.ascontiguousarray(np.broadcast_to(o, (n, 3)))
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        tests = C.c_longlong()
        self.lib.orc_trace_wide4(self._p(prims), prim_type, prims.shape[0], self._p(wide), wide.shape[0], self._p(np.ascontiguousarray(order, np.int32)),
                                 tie_by_objid, self._p(o), self._p(d), n, self._p(hit), self._p(t), C.byref(tests))
        return hit, t, tests.value

    def jitter(self, n, first=0):
        out = np.zeros(n, np.float64)
        self.lib.orc_jitter(self._p(out), n, C.c_ulonglong(first))
        return out

    def render_rows(self, sph, mat, nodes, order, width, height, spp, y0=0, y1=None, tie_by_objid=0, lights=None,
                    want_hit=True, want_accum=False, want_dirs=False, shadows=0, prim_type=0):
        y1 = height if y1 is None else y1
        sph = np.ascontiguousarray(sph, np.float32)
        mat = np.ascontiguousarray(mat, np.float32)
        lights = np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32) if lights is None else np.ascontiguousarray(lights, np.float32)
        rows = y1 - y0
        rgb = np.zeros((rows, width, 3), np.uint8)
        hit = np.zeros((rows, width), np.int32) if want_hit else None
        accum = np.zeros((rows, width, 3), np.float32) if want_accum else None
        dirs = np.zeros((rows, width, spp, 3), np.float32) if want_dirs else None
        counts = np.zeros(3, np.int64)
        self.lib.orc_render_rows_ex(self._p(sph), self._p(mat), sph.shape[0], self._p(nodes), self._p(order),
                                    0 if nodes is None else nodes.shape[0], tie_by_objid | (prim_type << 8), self._p(lights), lights.shape[0],
                                    width, height, spp, y0, y1, int(shadows), self._p(rgb), self._p(hit), self._p(accum),
                                    self._p(dirs), self._p(counts))
        self.last_ray_counts = counts
        return rgb, hit, accum, dirs


# ---------------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref/libref_oracle.so) — present when it was built in this container
# ---------------------------------------------------------------------------------------------------
REF_DUMP = np.dtype([("isleaf", np.int32), ("nobjs", np.int32), ("first_obj", np.int32), ("axis", np.int32),
                     ("box", np.uint32, 6)])


class Ref:
    def __init__(self):
        path = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")
        if not os.path.exists(path):
            pytest.skip("compiled reference oracle/_ref/libref_oracle.so not present")
        self.lib = C.CDLL(path)
        self.lib.ref_render_rows.restype = C.c_double
        self.lib.ref_sphere_tests.restype = C.c_longlong

    _p = staticmethod(Oracle._p)

    def scene_from_obj(self, model, clones=1):
        if not os.path.isdir(REF_TREE):
            pytest.skip("/root/reference not present")
        n = self.lib.ref_scene_from_obj(REF_TREE.encode(), model, clones)
        assert n > 0
        return self.scene_get()

    def scene_from_spheres(self, sph, mat):
        sph = np.ascontiguousarray(sph, np.float32)
        mat = np.ascontiguousarray(mat, np.float32)
        self.lib.ref_scene_from_spheres(self._p(sph), self._p(mat), sph.shape[0])

    def scene_get(self):
        n = self.lib.ref_scene_size()
        sph = np.zeros((n, 4), np.float32)
        mat = np.zeros((n, 4), np.float32)
        self.lib.ref_scene_get(self._p(sph), self._p(mat))
        return sph, mat

    def build(self, acc):
        secs = C.c_double()
        total = self.lib.ref_build(acc, C.byref(secs))
        return total, secs.value

    def bvh_dump(self):
        n = self.lib.ref_bvh_dump(None, 0, None, 0, None)
        recs = np.zeros(n, REF_DUMP)
        objs = np.zeros(n, np.int32)
        nobj = C.c_int()
        self.lib.ref_bvh_dump(self._p(recs), n, self._p(objs), n, C.byref(nobj))
        order = np.zeros(self.lib.ref_scene_size(), np.int32)
        self.lib.ref_scene_order(self._p(order))
        return recs, objs[:nobj.value], order

    def bvh_linear(self):
        """The reference's pointer tree, flattened into the LinearBVHNode layout (accelerators.h:231-240)."""
        recs, objs, order = self.bvh_dump()
        return ref_dump_to_linear(recs), objs, order

    def trace(self, acc, o, d):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = d.shape[0]
        if o.shape[0] == 1 and n > 1:
            o = np.ascontiguousarray(np.broadcast_to(o, (n, 3)))
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        cand = C.c_longlong()
        self.lib.ref_trace(self._p(o), self._p(d), n, acc, self._p(hit), self._p(t), C.byref(cand))
        return hit, t, cand.value

    def triangle_trace(self, tris, o, d):
        """Brute-force closest hit with the reference's own class Triangle / rayTriangleIntersect (compiled geometric branch):
        (hit index | -1, t, number of triangles hit) per ray."""
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = d.shape[0]
        if o.shape[0] == 1 and n > 1:
            o = np.ascontiguousarray(np.broadcast_to(o, (n, 3)))
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        cnt = np.zeros(n, np.int32)
        self.lib.ref_triangle_trace(self._p(tris), tris.shape[0], self._p(o), self._p(d), n, self._p(hit), self._p(t), self._p(cnt))
        return hit, t, cnt

    def render_rows(self, acc, width, height, spp, y0=0, y1=None, want_dirs=False, want_accum=False):
        y1 = height if y1 is None else y1
        rows = y1 - y0
        rgb = np.zeros((rows, width, 3), np.uint8)
        dirs = np.zeros((rows, width, spp, 3), np.float32) if want_dirs else None
        accum = np.zeros((rows, width, 3), np.float32) if want_accum else None
        secs = self.lib.ref_render_rows(width, height, spp, acc, y0, y1, self._p(rgb), self._p(accum), self._p(dirs))
        return rgb, dirs, accum, secs

    def jitter(self, n):
        out = np.zeros(n, np.float64)
        self.lib.ref_jitter(self._p(out), n)
        return out

    def kd_dump(self):
        """The reference's own KdAccelNode[] (nextFreeNode entries), kdtreePrimitiveIndices and tree bounds."""
        n = self.lib.ref_kd_dump(None, 0, None, 0, None, None)
        assert n > 0, "no KD-tree built in the reference harness"
        nodes = np.zeros(n, rtds_b200.KD_NODE_DTYPE)
        nidx = C.c_int()
        self.lib.ref_kd_dump(None, 0, None, 0, C.byref(nidx), None)
        idx = np.zeros(max(nidx.value, 1), np.int32)
        bounds = np.zeros(6, np.float32)
        self.lib.ref_kd_dump(self._p(nodes), n, self._p(idx), idx.size, C.byref(nidx), self._p(bounds))
        return nodes, idx[:nidx.value], bounds



def ref_dump_to_linear(recs):
    """Pre-order DumpRec[] -> LinearBVHNode[] (second-child offsets from subtree sizes, leaf offsets = DFS leaf rank)."""
    n = recs.shape[0]
    out = np.zeros(n, LINEAR)
    box = recs["box"].view(np.float32)
    out["bmin"] = box[:, :3]
    out["bmax"] = box[:, 3:]
    isleaf = recs["isleaf"].astype(bool)
    out["nPrimitives"] = np.where(isleaf, 1, 0)
    out["axis"] = np.where(isleaf, 0, recs["axis"]).astype(np.uint8)
    # subtree sizes by a reverse scan with an explicit stack
    size = np.ones(n, np.int64)
    offset = np.zeros(n, np.int32)
    stack = []
    leaf_rank = np.cumsum(isleaf) - 1
    for i in range(n - 1, -1, -1):
        if isleaf[i]:
            stack.append(i)
            offset[i] = leaf_rank[i]
        else:
            l = stack.pop()
            r = stack.pop()
            assert l == i + 1
            size[i] = 1 + size[l] + size[r]
            offset[i] = r
            stack.append(i)
    out["offset"] = offset
    return out


def ppm_md5(rgb):
    h, w, _ = rgb.shape
    return hashlib.md5(b"P6\n%d %d\n255\n" % (w, h) + rgb.tobytes()).hexdigest()


def load_golden_json():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def bunny_vertices():
    return np.fromfile(os.path.join(GOLDEN, "bunny_vertices.f32"), np.float32).reshape(-1, 3)


def bunny_scene(clones=1):
    return rtds_b200.scene_from_vertices(bunny_vertices(), clones)


def synthetic_scene(n, seed=1, ground=True, radius=0.05):
    """Seeded blob of small spheres in front of the camera (+ the reference's ground sphere)."""
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(n, 3)).astype(np.float32) * np.float32(4.0) + np.asarray([0, 0, -60], np.float32)
    sph = np.zeros((n + (1 if ground else 0), 4), np.float32)
    sph[:n, :3] = c
    sph[:n, 3] = radius
    if ground:
        sph[n] = np.asarray(rtds_b200.GROUND, np.float32)
    mat = np.zeros_like(sph)
    mat[:n, 0], mat[:n, 1] = 0.8, 0.7
    return sph, mat


def triangle_scene(n, seed, ground=True):
    """Seeded soup of small triangles in front of the camera (+ two large ground triangles): (m, 9) and (m, 4)."""
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(n, 1, 3)).astype(np.float32) * np.float32(4.0) + np.asarray([0, 0, -60], np.float32)
    v = c + rng.normal(size=(n, 3, 3)).astype(np.float32) * np.float32(0.6)
    tris = v.reshape(n, 9)
    if ground:
        g = np.asarray([[-60, -8, -20, 60, -8, -20, 60, -8, -140], [-60, -8, -20, 60, -8, -140, -60, -8, -140]], np.float32)
        tris = np.concatenate([tris, g], 0)
    mat = np.zeros((tris.shape[0], 4), np.float32)
    mat[:, :3] = rng.uniform(0.1, 0.9, size=(tris.shape[0], 3)).astype(np.float32)
    return np.ascontiguousarray(tris, np.float32), mat


def material_scene(n, seed, frac_rr=0.1, frac_refl=0.1, big=True):
    """Synthetic scene with REFLECTION_AND_REFRACTION / REFLECTION spheres (unreachable through the reference's
    loader, reachable through its castRay) — a few large ones so that secondary rays hit things."""
    sph, mat = synthetic_scene(n, seed)
    rng = np.random.default_rng(seed + 1000)
    if big:
        k = max(4, n // 100)
        sph[:k, 3] = rng.uniform(0.5, 1.5, k).astype(np.float32)
    u = rng.uniform(size=n)
    mat[:n, 3] = np.where(u < frac_rr, 1.0, np.where(u < frac_rr + frac_refl, 2.0, 0.0)).astype(np.float32)
    mat[:n, :3] = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    return sph, mat


import contextlib


@contextlib.contextmanager
def option(ctx, name, value):
    """Set a tuning / test switch of the context (rtds_set_option) for the duration of a with-block."""
    old = ctx.get_option(name)
    ctx.set_option(name, value)
    try:
        yield
    finally:
        ctx.set_option(name, old)


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    return Ref()


@pytest.fixture(scope="session")
def gpu_ctx():
    if not has_gpu():
        pytest.skip("no GPU")
    ctx = rtds_b200.Rtds(0)
    yield ctx
    ctx.close()
