"""CPU: the C-ABI library loads and exports every symbol include/rtds.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

import conftest as T

rt = T.rtds_b200


def _declared_symbols():
    src = open(os.path.join(T.ROOT, "include", "rtds.h")).read()
    return sorted(set(re.findall(r"\b(rtds_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(rt.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(rt.LIB_PATH), "librtds.so not built: run __graft_entry__.build()"
    lib = C.CDLL(rt.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header_layouts():
    assert C.sizeof(rt.BuildParams) == 4 * 15
    assert C.sizeof(rt.BuildStats) == 4 * 14
    assert rt.LINEAR_NODE_DTYPE.itemsize == 32 and rt.KD_NODE_DTYPE.itemsize == 12


def test_no_cpu_fallback_without_gpu():
    if T.has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(rt.RtdsError) as e:
        rt.Rtds(0)
    assert e.value.code == -3          # RTDS_ERR_NO_DEVICE: the product path fails loudly


def test_rows_for_rank_partition():
    for h in (1, 7, 8, 480, 1080, 2160, 203):
        for world in (1, 2, 3, 4, 8):
            rows = [rt.owned_rows(h, 8, r, world) for r in range(world)]
            assert sorted(sum((r.tolist() for r in rows), [])) == list(range(h))
            for r in range(world):
                assert rt.rows_for_rank(h, 8, r, world) == rows[r].size


def test_ctypes_layouts_equal_the_headers(tmp_path):
    """sizeof / offsetof of every struct field as the C compiler lays include/rtds.h out == the ctypes mirror."""
    import subprocess
    structs = {"rtds_build_params": rt.BuildParams, "rtds_build_stats": rt.BuildStats, "rtds_render_params": rt.RenderParams,
               "rtds_render_stats": rt.RenderStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rtds.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['printf("rtds_linear_bvh_node %zu\\n", sizeof(rtds_linear_bvh_node));', 'printf("rtds_kd_node %zu\\n", sizeof(rtds_kd_node));',
              'printf("RTDS_TRACE_KD_CLOSEST %d\\n", RTDS_TRACE_KD_CLOSEST);', "return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(T.ROOT, "include"), "-o", str(exe), str(src)])   # the header is plain C
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, f"{cname}.{f}"
    assert int(got["rtds_linear_bvh_node"]) == 32 and int(got["rtds_kd_node"]) == 12
    assert int(got["RTDS_TRACE_KD_CLOSEST"]) == rt.TRACE_KD_CLOSEST
