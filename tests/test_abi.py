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
