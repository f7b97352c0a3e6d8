"""CPU suite: the C-ABI library loads and exports every symbol include/hc_b200.h declares; the
host-only entry points work without a GPU; the compute entry points fail loudly without one."""
import os
import re

import numpy as np
import pytest

from haploconduct_b200 import capi, formats as F
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built_lib):
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(built_lib, s), "libhc_b200.so does not export " + s
    assert sorted(capi.EXPORTED) == syms


def test_struct_sizes_match_header():
    assert F.CANDIDATE.itemsize == 32 and F.RESULT.itemsize == 48 and F.EDGE.itemsize == 48 and F.READ_DESC.itemsize == 24


def test_host_only_entry_points(built_lib):
    for q in (0, 1, 2, 20, 41, 93):
        assert capi.phred_to_prob(q) == O.lib().hco_phred_to_prob(q)
    import math
    for thr in (0.9, 0.95, 0.97, 0.99, 0.995, 0.5):
        x = capi.exp_threshold(thr)
        assert math.exp(x) > thr and not (math.exp(np.nextafter(x, -np.inf)) > thr)
    assert capi.exp_threshold(1.0) > 0.0          # exp(mean) > 1 is impossible for mean <= 0 (POLYTE later iterations)
    assert capi.exp_threshold(-0.5) == -np.inf
    assert b"sm_100a" in built_lib.hc_version()


def test_no_cpu_fallback_without_gpu(built_lib):
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    rs = F.ReadSet.from_lists([(0, "ACGT", "IIII"), (1, "ACGT", "IIII")], [])
    with pytest.raises(capi.HcError) as e:
        capi.Store(rs)
    assert "no CPU fallback" in str(e.value)


def _decode_runs(anchor, start, entries):
    per_anchor = np.repeat(anchor, np.diff(start.astype(np.int64)))
    other = entries["other"] & 0x7fffffff
    anchor_is_2 = (entries["other"] >> 31).astype(bool)
    return np.where(anchor_is_2, other, per_anchor), np.where(anchor_is_2, per_anchor, other)


@pytest.mark.parametrize("order", ["file", "by_read", "by_min_id"])
def test_run_encoding_round_trip(order):
    """hc_candidate -> run-encoded 8-byte records (formats.run_encode) -> the same (ID1, ID2, POS, ORI, ORD), in the same order."""
    rng = np.random.RandomState(5)
    n = 5000
    c = np.zeros(n, dtype=F.CANDIDATE)
    c["idx1"], c["idx2"] = rng.randint(0, 300, size=n), rng.randint(0, 300, size=n)
    c["pos1"], c["pos2"] = rng.randint(0, 1 << 14, size=n), rng.randint(0, 1 << 14, size=n)
    c["ori1"], c["ori2"] = rng.randint(0, 2, size=n), rng.randint(0, 2, size=n)
    c["ord"] = rng.choice([ord("-"), ord("1"), ord("2")], size=n)
    if order == "by_read":
        c = c[np.argsort(c["idx1"], kind="stable")]
    elif order == "by_min_id":
        c = c[np.lexsort((np.maximum(c["idx1"], c["idx2"]), np.minimum(c["idx1"], c["idx2"])))]
    anchor, start, entries = F.run_encode(c)
    assert entries.dtype.itemsize == 8 and start[0] == 0 and start[-1] == n and (np.diff(start.astype(np.int64)) > 0).all()
    assert len(anchor) == len(start) - 1 and (order == "file" or len(anchor) <= 300)
    i1, i2 = _decode_runs(anchor, start, entries)
    assert np.array_equal(i1, c["idx1"]) and np.array_equal(i2, c["idx2"])
    assert np.array_equal(entries["pos"], F.short_candidates(c)["pos"])
    a0, s0, e0 = F.run_encode(c[:0])
    assert len(a0) == 0 and len(e0) == 0 and s0.tolist() == [0]
    big = c.copy()
    big["pos1"][7] = 1 << 14
    with pytest.raises(ValueError):
        F.run_encode(big)


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/hc_b200.h must compile as C99 (no C++ types, no torch types)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc on this box")
    src = tmp_path / "abi.c"
    src.write_text('#include "hc_b200.h"\nint main(void) { return (int)sizeof(hc_candidate) + (int)sizeof(hc_fno_overlap_small) + (int)sizeof(hc_adj_edge); }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert capi.FNO_OVERLAP_SMALL.itemsize == 24 and capi.ADJ_EDGE.itemsize == 16 and capi.SUBREAD_PROBLEM.itemsize == 40
