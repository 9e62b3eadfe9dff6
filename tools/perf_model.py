"""Loader of tools/perf_model.cpp (the CPU performance model of the ordered packet traversal; design evidence, not an oracle).
Builds tools/libperfmodel.so with g++ on demand."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class PerfModel:
    def __init__(self):
        src, lib = os.path.join(HERE, "perf_model.cpp"), os.path.join(HERE, "libperfmodel.so")
        dep = os.path.join(HERE, "..", "oracle", "oracle.cpp")
        if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(dep)):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", lib, src])
        self.lib = C.CDLL(lib)

    def packet_model(self, sph, nodes, wide, order, dirs, tie_by_objid=1, use_wide=False, order_mode=0, quant_bits=0):
        """Counts of the ordered packet traversal (orc_packet_model). dirs: (packets, rays per packet <= 16, 3)."""
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)   # noqa: E731
        sph = np.ascontiguousarray(sph, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        nr = dirs.shape[-2] if dirs.ndim == 3 else 4
        dirs = dirs.reshape(-1, nr, 3)
        hit = np.zeros((dirs.shape[0], nr), np.int32)
        st = np.zeros(6, np.int64)
        order = np.ascontiguousarray(order, np.int32)
        self.lib.orc_packet_model(p(sph), sph.shape[0], p(nodes), nodes.shape[0], p(wide), wide.shape[0], p(order), tie_by_objid, p(dirs),
                                  dirs.shape[0], int(use_wide) | (order_mode << 4) | (quant_bits << 8) | (nr << 16), p(hit), p(st))
        keys = ("packets", "interior_visits", "leaf_visits", "box_tests", "prim_tests", "max_stack")
        return hit, dict(zip(keys, (int(x) for x in st)))
