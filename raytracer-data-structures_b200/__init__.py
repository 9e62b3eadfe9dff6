"""rtds_b200 — Python (ctypes) binding of librtds.so, the B200-native build + traversal path of
Alhajras/Raytracer-Data-structures, plus a mirror of the reference's `Settings` surface.

The directory name carries a hyphen (it follows the reference repository's name), so import it with
``tests/conftest.py``'s ``load_rtds()`` / ``bench.py``'s loader, which register it as ``rtds_b200``.

There is NO CPU fallback here: importing works anywhere (so CPU-only tests can check the ABI), but every
compute call goes through the CUDA library and `Rtds()` raises when the library or a GPU is missing.

Reference surface mirrored (paths relative to /root/reference/project/raytracer/):
  AccType      accelerators.h:21      SceneModel / Settings   settings.h:4-17
  createScene_new  main.cpp:599-721   constructBVHNew / constructLBVHTree / constructKDTreeNew (main.cpp:800,832,816)
  render       main.cpp:541-566
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTDS_LIB", os.path.join(_HERE, "librtds.so"))   # RTDS_LIB: A/B-test another build of the library
HOST_LIB_PATH = os.path.join(_HERE, "librtds_host.so")

# enum AccType { BVH, KDTREE, UNIFORM_GRID, LBVH, NONE }  (accelerators.h:21)
BVH, KDTREE, UNIFORM_GRID, LBVH, NONE = 0, 1, 2, 3, 4
# enum SceneModel { IGEA, ARMADILLO, BUNNY, BUNNIES, TEST, GRASS, BUDDHA, CITY }  (settings.h:4)
IGEA, ARMADILLO, BUNNY, BUNNIES, TEST, GRASS, BUDDHA, CITY = range(8)
MODE_COMPAT, MODE_TRUE, MODE_SAH = 0, 1, 2
TRACE_KD_CLOSEST = 2      # rtds.h: or-ed into rtds_trace's `exact`
TRACE_TRI_GEOMETRIC = 4

STATUS = {0: "OK", -1: "INVALID", -2: "CUDA", -3: "NO_DEVICE", -4: "NO_SCENE", -5: "NOT_BUILT", -6: "DEGENERATE",
          -7: "CAPACITY", -8: "UNSUPPORTED"}


@dataclass
class Settings:
    """settings.h:7-17, same names and defaults."""
    width: int = 640
    height: int = 480
    fov: float = 90.0            # unused by the reference (render() hard-codes 30, main.cpp:545)
    backgroundColor: tuple = (1.0, 1.0, 1.0)   # unused by the reference (castRay hard-codes the sky)
    bias: float = 0.0001         # unused by the reference (castRay hard-codes 1e-4)
    aa_samples: int = 1
    dataStructure: int = BVH
    sceneModel: int = BUNNY


class RtdsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rtds error {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


class BuildParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("morton_bits", C.c_int), ("morton_ref_norm", C.c_int),
                ("kd_isect_cost", C.c_int), ("kd_traversal_cost", C.c_int), ("kd_empty_bonus", C.c_float),
                ("kd_max_prims", C.c_int), ("kd_max_depth", C.c_int), ("sah_bins", C.c_int), ("reserved", C.c_int * 6)]


class BuildStats(C.Structure):
    _fields_ = [("n_prims", C.c_int), ("total_nodes", C.c_int), ("alloc_nodes", C.c_int), ("max_depth", C.c_int),
                ("kernel_launches", C.c_int), ("ms", C.c_float), ("reserved", C.c_int * 8)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("aa_samples", C.c_int), ("fov", C.c_float),
                ("bg", C.c_float * 3), ("bias", C.c_float), ("max_depth", C.c_int), ("shadows", C.c_int),
                ("exact", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("tile_rows", C.c_int),
                ("jitter_offset", C.c_uint64), ("no_jitter_regen", C.c_int), ("kd_closest", C.c_int), ("tri_geometric", C.c_int),
                ("reserved", C.c_int * 5)]


class RenderStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("secondary_rays", C.c_uint64), ("node_tests", C.c_uint64), ("prim_tests", C.c_uint64),
                ("node_visits", C.c_uint64), ("ms_kernel", C.c_float), ("ms_total", C.c_float),
                ("kernel_launches", C.c_int), ("rows", C.c_int), ("reserved", C.c_int * 6)]


# 32-byte LinearBVHNode (accelerators.h:231-240)
LINEAR_NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("bmax", np.float32, 3), ("offset", np.int32),
                              ("nPrimitives", np.uint16), ("axis", np.uint8), ("pad", np.uint8)])
KD_NODE_DTYPE = np.dtype([("w0", np.uint32), ("w1", np.uint32), ("w2", np.uint32)])
assert LINEAR_NODE_DTYPE.itemsize == 32 and KD_NODE_DTYPE.itemsize == 12

# every symbol include/rtds.h declares
ABI_SYMBOLS = ["rtds_last_error", "rtds_version", "rtds_create", "rtds_destroy", "rtds_set_spheres",
               "rtds_set_triangles", "rtds_set_lights", "rtds_build", "rtds_export_bvh", "rtds_export_kd",
               "rtds_export_morton", "rtds_trace", "rtds_render", "rtds_render_device", "rtds_rows_for_rank",
               "rtds_jitter_stream", "rtds_morton30", "rtds_frame", "rtds_shared_frame_create", "rtds_shared_frame_open",
               "rtds_shared_frame_attach", "rtds_render_shared", "rtds_frame_shared", "rtds_shared_frame_ptr", "rtds_shared_frame_read",
               "rtds_shared_frame_close", "rtds_set_option", "rtds_get_option", "rtds_set_spheres_device", "rtds_prepare_frame"]

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen librtds.so. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RtdsError(-3, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.rtds_last_error.restype = C.c_char_p
    lib.rtds_version.restype = C.c_char_p
    vp = C.c_void_p
    lib.rtds_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.rtds_destroy.argtypes = [vp]
    lib.rtds_set_spheres.argtypes = [vp, vp, vp, C.c_int]
    lib.rtds_set_spheres_device.argtypes = [vp, vp, vp, C.c_int]
    lib.rtds_prepare_frame.argtypes = [vp, C.POINTER(RenderParams)]
    lib.rtds_set_triangles.argtypes = [vp, vp, vp, C.c_int]
    lib.rtds_set_lights.argtypes = [vp, vp, C.c_int]
    lib.rtds_build.argtypes = [vp, C.c_int, C.POINTER(BuildParams), C.POINTER(BuildStats)]
    lib.rtds_export_bvh.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), vp, C.c_int, C.POINTER(C.c_int)]
    lib.rtds_export_kd.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), vp, C.c_int, C.POINTER(C.c_int), vp]
    lib.rtds_export_morton.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
    lib.rtds_trace.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.POINTER(RenderStats)]
    lib.rtds_render.argtypes = [vp, C.c_int, C.POINTER(RenderParams), vp, vp, vp, C.POINTER(RenderStats)]
    lib.rtds_render_device.argtypes = [vp, C.c_int, C.POINTER(RenderParams), vp, C.POINTER(RenderStats)]
    lib.rtds_frame.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.POINTER(BuildParams), C.POINTER(RenderParams), vp, C.POINTER(BuildStats),
                               C.POINTER(RenderStats)]
    lib.rtds_rows_for_rank.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.rtds_jitter_stream.argtypes = [vp, C.c_uint64, C.c_int, vp]
    lib.rtds_morton30.argtypes = [vp, vp, C.c_int, vp]
    lib.rtds_shared_frame_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    lib.rtds_shared_frame_open.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.rtds_shared_frame_attach.argtypes = [vp, vp, C.c_int]
    lib.rtds_render_shared.argtypes = [vp, C.c_int, C.POINTER(RenderParams), C.c_uint32, C.POINTER(RenderStats)]
    lib.rtds_frame_shared.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.POINTER(BuildParams), C.POINTER(RenderParams), C.c_uint32,
                                      C.POINTER(BuildStats), C.POINTER(RenderStats)]
    lib.rtds_shared_frame_ptr.argtypes = [vp, C.POINTER(vp)]
    lib.rtds_shared_frame_read.argtypes = [vp, vp]
    lib.rtds_shared_frame_close.argtypes = [vp]
    lib.rtds_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.rtds_get_option.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _stats_dict(s):
    return {k: getattr(s, k) for k, _ in s._fields_ if k != "reserved"}


class Rtds:
    """One context = one GPU (rtds_create). Mirrors the reference's build/render call sites."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        self._check(self.lib.rtds_create(C.byref(self.ctx), device))
        self.n = 0

    def _check(self, rc):
        if rc != 0:
            raise RtdsError(rc, self.lib.rtds_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.rtds_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """Tuning / test switch of this context (rtds_set_option); defaults come from RTDS_<NAME>, read once at creation."""
        self._check(self.lib.rtds_set_option(self.ctx, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int()
        self._check(self.lib.rtds_get_option(self.ctx, name.encode(), C.byref(v)))
        return v.value

    # -- scene ---------------------------------------------------------------------------------
    def set_spheres(self, cxyz_r, rgb_mat=None):
        cxyz_r = np.ascontiguousarray(cxyz_r, np.float32).reshape(-1, 4)
        if rgb_mat is not None:
            rgb_mat = np.ascontiguousarray(rgb_mat, np.float32).reshape(-1, 4)
            assert rgb_mat.shape == cxyz_r.shape
        self._check(self.lib.rtds_set_spheres(self.ctx, _ptr(cxyz_r), _ptr(rgb_mat), cxyz_r.shape[0]))
        self.n = cxyz_r.shape[0]

    def set_spheres_device(self, d_cxyz_r_ptr, d_rgb_mat_ptr, n):
        """rtds_set_spheres_device: both tables already in this GPU's memory (raw device pointers)."""
        self._check(self.lib.rtds_set_spheres_device(self.ctx, C.c_void_p(d_cxyz_r_ptr), C.c_void_p(d_rgb_mat_ptr), n))
        self.n = n

    def set_triangles(self, v0v1v2, rgb_mat=None):
        """Triangle scene (extension; the reference never instantiates class Triangle): (n, 9) float32."""
        tris = np.ascontiguousarray(v0v1v2, np.float32).reshape(-1, 9)
        if rgb_mat is not None:
            rgb_mat = np.ascontiguousarray(rgb_mat, np.float32).reshape(-1, 4)
            assert rgb_mat.shape[0] == tris.shape[0]
        self._check(self.lib.rtds_set_triangles(self.ctx, _ptr(tris), _ptr(rgb_mat), tris.shape[0]))
        self.n = tris.shape[0]

    def set_lights(self, lights7):
        lights7 = np.ascontiguousarray(lights7, np.float32).reshape(-1, 7)
        self._check(self.lib.rtds_set_lights(self.ctx, _ptr(lights7), lights7.shape[0]))

    # -- build ---------------------------------------------------------------------------------
    def build(self, acc, mode=MODE_COMPAT, morton_bits=0, morton_ref_norm=0, **kd):
        p = BuildParams()
        p.mode, p.morton_bits, p.morton_ref_norm = mode, morton_bits, morton_ref_norm
        for k, v in kd.items():
            setattr(p, k, v)
        st = BuildStats()
        self._check(self.lib.rtds_build(self.ctx, acc, C.byref(p), C.byref(st)))
        return _stats_dict(st)

    def export_bvh(self):
        nn, npr = C.c_int(), C.c_int()
        self._check(self.lib.rtds_export_bvh(self.ctx, None, 0, C.byref(nn), None, 0, C.byref(npr)))
        nodes = np.zeros(nn.value, LINEAR_NODE_DTYPE)
        order = np.zeros(npr.value, np.int32)
        self._check(self.lib.rtds_export_bvh(self.ctx, _ptr(nodes), nn.value, C.byref(nn), _ptr(order), npr.value, C.byref(npr)))
        return nodes, order

    def export_kd(self):
        nn, ni = C.c_int(), C.c_int()
        self._check(self.lib.rtds_export_kd(self.ctx, None, 0, C.byref(nn), None, 0, C.byref(ni), None))
        nodes = np.zeros(nn.value, KD_NODE_DTYPE)
        idx = np.zeros(max(ni.value, 1), np.int32)
        bounds = np.zeros(6, np.float32)
        self._check(self.lib.rtds_export_kd(self.ctx, _ptr(nodes), nn.value, C.byref(nn), _ptr(idx), idx.size, C.byref(ni), _ptr(bounds)))
        return nodes, idx[:ni.value], bounds

    def export_morton(self):
        n = C.c_int()
        keys = np.zeros(self.n, np.uint64)
        ids = np.zeros(self.n, np.int32)
        self._check(self.lib.rtds_export_morton(self.ctx, _ptr(keys), _ptr(ids), self.n, C.byref(n)))
        return keys[:n.value], ids[:n.value]

    # -- trace / render ------------------------------------------------------------------------
    def trace(self, acc, o, d, exact=True, kd_closest=False, tri_geometric=False):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = d.shape[0]
        if o.shape[0] == 1 and n > 1:
            o = np.ascontiguousarray(np.broadcast_to(o, (n, 3)))
        hit = np.full(n, -2, np.int32)
        t = np.zeros(n, np.float32)
        st = RenderStats()
        self._check(self.lib.rtds_trace(self.ctx, acc, int(bool(exact)) | (TRACE_KD_CLOSEST if kd_closest else 0) | (TRACE_TRI_GEOMETRIC if tri_geometric else 0), _ptr(o), _ptr(d), n, _ptr(hit), _ptr(t), C.byref(st)))
        return hit, t, _stats_dict(st)

    def render_params(self, width, height, aa_samples=1, exact=False, rank=0, world=1, tile_rows=8, shadows=0,
                      jitter_offset=0, no_jitter_regen=0, max_depth=0, kd_closest=0, tri_geometric=0):
        p = RenderParams()
        p.width, p.height, p.aa_samples = width, height, aa_samples
        p.exact, p.rank, p.world, p.tile_rows, p.shadows = int(exact), rank, world, tile_rows, shadows
        p.jitter_offset, p.no_jitter_regen, p.max_depth = jitter_offset, no_jitter_regen, max_depth
        p.kd_closest = int(kd_closest)
        p.tri_geometric = int(tri_geometric)
        return p

    def render(self, acc, width, height, aa_samples=1, want_hit=False, want_accum=False, out=None, **kw):
        """render() + write_into_file's quantisation: returns (rgb uint8 [H,W,3], hit|None, accum|None, stats)."""
        p = self.render_params(width, height, aa_samples, **kw)
        rgb = out if out is not None else np.zeros((height, width, 3), np.uint8)
        hit = np.full((height, width), -2, np.int32) if want_hit else None
        accum = np.zeros((height, width, 3), np.float32) if want_accum else None
        st = RenderStats()
        self._check(self.lib.rtds_render(self.ctx, acc, C.byref(p), _ptr(rgb), _ptr(hit), _ptr(accum), C.byref(st)))
        return rgb, hit, accum, _stats_dict(st)

    def frame(self, cxyz_r, rgb_mat, acc, width, height, aa_samples=1, mode=MODE_COMPAT, out=None, **kw):
        """rtds_frame: upload + build + render + download in one call (what main() does per run)."""
        cxyz_r = np.ascontiguousarray(cxyz_r, np.float32).reshape(-1, 4)
        rgb_mat = None if rgb_mat is None else np.ascontiguousarray(rgb_mat, np.float32).reshape(-1, 4)
        bp = BuildParams()
        bp.mode = mode
        rp = self.render_params(width, height, aa_samples, **kw)
        rgb = out if out is not None else np.zeros((height, width, 3), np.uint8)
        bst, rst = BuildStats(), RenderStats()
        self._check(self.lib.rtds_frame(self.ctx, _ptr(cxyz_r), _ptr(rgb_mat), cxyz_r.shape[0], acc, C.byref(bp), C.byref(rp), _ptr(rgb),
                                        C.byref(bst), C.byref(rst)))
        self.n = cxyz_r.shape[0]
        return rgb, _stats_dict(bst), _stats_dict(rst)

    def prepare_frame(self, params):
        """rtds_prepare_frame: start generating the frame's ray directions now (side stream); returns at once."""
        self._check(self.lib.rtds_prepare_frame(self.ctx, C.byref(params)))

    def render_device(self, acc, params, device_ptr):
        st = RenderStats()
        self._check(self.lib.rtds_render_device(self.ctx, acc, C.byref(params), C.c_void_p(device_ptr), C.byref(st)))
        return _stats_dict(st)

    # -- multi-GPU frame assembled by peer stores (no collective) ----------------------------------
    def shared_frame_create(self, width, height, world):
        """Rank 0: allocate the frame every rank renders into; returns the 64-byte CUDA IPC handle for other processes."""
        h = (C.c_ubyte * 64)()
        self._check(self.lib.rtds_shared_frame_create(self.ctx, width, height, world, C.cast(h, C.c_void_p)))
        return bytes(h)

    def shared_frame_open(self, handle, width, height, world, rank):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self.lib.rtds_shared_frame_open(self.ctx, C.cast(buf, C.c_void_p), width, height, world, rank))

    def shared_frame_attach(self, owner, rank):
        self._check(self.lib.rtds_shared_frame_attach(self.ctx, owner.ctx, rank))

    def render_shared(self, acc, params, frame_seq):
        st = RenderStats()
        self._check(self.lib.rtds_render_shared(self.ctx, acc, C.byref(params), frame_seq, C.byref(st)))
        return _stats_dict(st)

    def frame_shared(self, cxyz_r, rgb_mat, acc, params, frame_seq, mode=MODE_COMPAT):
        """rtds_frame_shared: upload + build + render_shared in one call (one rank of a multi-GPU run)."""
        cxyz_r = np.ascontiguousarray(cxyz_r, np.float32).reshape(-1, 4)
        rgb_mat = None if rgb_mat is None else np.ascontiguousarray(rgb_mat, np.float32).reshape(-1, 4)
        bp = BuildParams()
        bp.mode = mode
        bst, rst = BuildStats(), RenderStats()
        self._check(self.lib.rtds_frame_shared(self.ctx, _ptr(cxyz_r), _ptr(rgb_mat), cxyz_r.shape[0], acc, C.byref(bp), C.byref(params),
                                               frame_seq, C.byref(bst), C.byref(rst)))
        self.n = cxyz_r.shape[0]
        return _stats_dict(bst), _stats_dict(rst)

    def shared_frame_ptr(self):
        p = C.c_void_p()
        self._check(self.lib.rtds_shared_frame_ptr(self.ctx, C.byref(p)))
        return p.value

    def shared_frame_read(self, width, height, out=None):
        rgb = out if out is not None else np.zeros((height, width, 3), np.uint8)
        self._check(self.lib.rtds_shared_frame_read(self.ctx, _ptr(rgb)))
        return rgb

    def shared_frame_close(self):
        self._check(self.lib.rtds_shared_frame_close(self.ctx))

    def jitter_stream(self, first, n):
        out = np.zeros(n, np.float64)
        self._check(self.lib.rtds_jitter_stream(self.ctx, first, n, _ptr(out)))
        return out

    def morton30(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        codes = np.zeros(xyz.shape[0], np.uint32)
        self._check(self.lib.rtds_morton30(self.ctx, _ptr(xyz), xyz.shape[0], _ptr(codes)))
        return codes


def rows_for_rank(height, tile_rows, rank, world):
    return load_library().rtds_rows_for_rank(height, tile_rows, rank, world)


def owned_rows(height, tile_rows, rank, world):
    """Global row indices rank owns, in its local (compact) order: interleaved scanline tiles."""
    rows = []
    for t in range(rank, (height + tile_rows - 1) // tile_rows, world):
        rows.extend(range(t * tile_rows, min(height, (t + 1) * tile_rows)))
    return np.asarray(rows, np.int64)


# ------------------------------------------------------------------------------------------------------
# host-side scene creation: createScene_new's arithmetic (main.cpp:650-717) on parsed vertices.
# (The C++ host, host/scene.cpp, is what main.cpp uses; this numpy version serves the Python callers.)
# ------------------------------------------------------------------------------------------------------
GROUND = (0.93591022, -105.47120094, -43.2363205, 100.0)   # main.cpp:703


def scene_from_vertices(v, clones=1, clone_shift=20):
    """(n*clones+1, 4) float32 {cx,cy,cz,r} and matching {r,g,b,material}, float arithmetic as main.cpp:679-682.
    clone_shift: the reference shifts clone k by 20*k on every axis (mostly off-screen); 2 is the report's experiment with
    the clones in view (Report/Performance.xlsx rows 10-12) - measurement variant only."""
    v = np.ascontiguousarray(v, np.float32).reshape(-1, 3)
    parts = []
    for clone in range(clones):
        shift = np.float32(clone * clone_shift)
        c = v * np.float32(100) + shift
        c[:, 1] += np.float32(-10)
        c[:, 2] += np.float32(-60)
        parts.append(c)
    c = np.concatenate(parts, 0)
    n = c.shape[0]
    sph = np.zeros((n + 1, 4), np.float32)
    sph[:n, :3] = c
    sph[:n, 3] = np.float32(0.01 * 5)
    sph[n] = np.asarray(GROUND, np.float32)
    mat = np.zeros((n + 1, 4), np.float32)
    mat[:n, 0], mat[:n, 1] = 0.8, 0.7
    return sph, mat


def parse_obj_vertices(path):
    """The reference's loader (main.cpp:663-698): whitespace tokens; "v" + 3 floats; any other token stops."""
    out = []
    with open(path, "r") as f:
        toks = f.read().split()
    i = 0
    while i < len(toks) and toks[i] == "v":
        out.append((np.float32(toks[i + 1]), np.float32(toks[i + 2]), np.float32(toks[i + 3])))
        i += 4
    return np.asarray(out, np.float32).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------------
# multi-GPU plumbing: the framebuffer gather (the path's only collective). torch.distributed is plumbing here —
# NCCL over NVLink on GPUs, gloo in the CPU tests.
# ------------------------------------------------------------------------------------------------------
def scene_slice(n, rank, world):
    """The part of an n-row scene table rank uploads when the ranks exchange the scene among themselves (bench.py, N >= 2):
    (rows per rank in the exchange buffer, first row, one past the last row). Every row belongs to exactly one rank; the last
    ranks may own fewer rows (or none)."""
    per = (n + world - 1) // world
    return per, min(n, rank * per), min(n, (rank + 1) * per)


def exchange_scene(part, full, group=None):
    """All-gather of the per-rank parts of a scene table (one padded [per, 4] float32 tensor per rank) into `full`
    ([per * world, 4]; its first n rows are the table). The one collective of the path - NCCL over NVLink on GPUs, gloo in the CPU
    tests."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(full, part, group=group)
    return full


def gather_frame(local_rows, height, width, tile_rows, rank, world, dst=0, scratch=None):
    """local_rows: uint8 tensor [>= rows_for_rank, width, 3] holding this rank's rows compactly (local tile j = global
    tile j*world + rank). Returns the assembled [height, width, 3] frame on `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_rows[:height]
    max_rows = max(len(owned_rows(height, tile_rows, r, world)) for r in range(world))
    if local_rows.shape[0] != max_rows:
        pad = torch.zeros((max_rows, width, 3), dtype=torch.uint8, device=local_rows.device)
        pad[: local_rows.shape[0]] = local_rows
        local_rows = pad
    if rank == dst:
        parts = scratch if scratch is not None else [torch.zeros_like(local_rows) for _ in range(world)]
        dist.gather(local_rows, parts, dst=dst)
        frame = torch.zeros((height, width, 3), dtype=torch.uint8, device=local_rows.device)
        for r in range(world):
            idx = torch.from_numpy(owned_rows(height, tile_rows, r, world)).to(local_rows.device)
            frame[idx] = parts[r][: idx.numel()]
        return frame
    dist.gather(local_rows, None, dst=dst)
    return None
