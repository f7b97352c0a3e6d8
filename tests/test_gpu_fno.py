"""GPU suite: hc_fno1 (CUDA) against the reference's findNextOverlaps() output (golden) and, on
dense random inputs, against the pinned C oracle -- records identical, in processing order."""
import numpy as np
import pytest

from haploconduct_b200 import capi, formats as F
from oracle import oracle as O
from util import fno3_golden_names, fno_golden_names, load_fno3_golden, load_fno_golden, random_fno3_input, random_fno_input

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", fno_golden_names())
def test_fno1_reference_file(built_lib, name):
    fi, ref = load_fno_golden(name)
    ov = capi.fno1(fi)
    assert F.fno_output_file(ov) == ref
    assert ov.tobytes() == O.fno1(fi).tobytes()      # same records, same (processing) order


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_fno1_random_dense_inputs(built_lib, seed):
    fi = random_fno_input(seed, n_vertices=500, n_sr=150, n_edges=60000)
    a, b = capi.fno1(fi), O.fno1(fi)
    assert len(a) == len(b) and len(a) > 1000
    assert a.tobytes() == b.tobytes()


def test_fno1_empty_and_bad_input(built_lib):
    fi = random_fno_input(9, n_edges=10)
    fi.edges = fi.edges[:0]
    assert len(capi.fno1(fi)) == 0
    fi = random_fno_input(9, n_edges=10)
    fi.edges["u"][3] = 10 ** 6
    with pytest.raises(capi.HcError):
        capi.fno1(fi)


@pytest.mark.parametrize("name", fno3_golden_names())
def test_fno3_reference_file(built_lib, name):
    fi, ref = load_fno3_golden(name)
    ov = capi.fno3(fi)
    assert F.fno_lines(ov) == ref                       # discovery order, like the reference's overlaps.txt
    assert ov.tobytes() == O.fno3(fi).tobytes()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fno3_random_dense_inputs(built_lib, seed):
    fi = random_fno3_input(seed, n_originals=30000, n_reads=2000)
    a, b = capi.fno3(fi), O.fno3(fi)
    assert len(a) == len(b) and len(a) > 1000
    assert a.tobytes() == b.tobytes()


# ---- round 2: partitioned first-found-wins table, 24-byte staging / results ------------------------------------------
@pytest.mark.parametrize("name", fno_golden_names())
def test_fno1_small_records_reference_file(built_lib, name):
    fi, ref = load_fno_golden(name)
    ov = capi.fno1(fi, small=True)
    assert F.fno_output_file(ov) == ref
    assert ov.tobytes() == O.fno1(fi).tobytes()


@pytest.mark.parametrize("name", fno3_golden_names())
def test_fno3_small_records_reference_file(built_lib, name):
    fi, ref = load_fno3_golden(name)
    ov = capi.fno3(fi, small=True)
    assert F.fno_lines(ov) == ref
    assert ov.tobytes() == O.fno3(fi).tobytes()


def test_fno_many_partitions_and_both_staging_formats(built_lib, monkeypatch):
    """HC_FNO_PART cuts the first-found-wins table into partitions of a few thousand keys (dozens of passes here, as on
    1e7-edge inputs); HC_FNO_STAGE48 forces the 48-byte staging path that inputs beyond the 24-byte ranges take."""
    fi1 = random_fno_input(11, n_vertices=500, n_sr=150, n_edges=60000)
    fi3 = random_fno3_input(12, n_originals=30000, n_reads=2000)
    ref1, ref3 = O.fno1(fi1).tobytes(), O.fno3(fi3).tobytes()
    for part in ("5000", "700", None):
        for big in (None, "1"):
            if part: monkeypatch.setenv("HC_FNO_PART", part)
            else: monkeypatch.delenv("HC_FNO_PART", raising=False)
            if big: monkeypatch.setenv("HC_FNO_STAGE48", big)
            else: monkeypatch.delenv("HC_FNO_STAGE48", raising=False)
            assert capi.fno1(fi1).tobytes() == ref1
            assert capi.fno3(fi3).tobytes() == ref3
            if not big:
                assert capi.fno1(fi1, small=True).tobytes() == ref1
                assert capi.fno3(fi3, small=True).tobytes() == ref3


def test_fno1_values_beyond_the_small_record(built_lib):
    """Positions >= 2^24 (or negative ones) do not fit the 24-byte record: hc_fno1 falls back to 48-byte staging and
    still equals the oracle, hc_fno1_small reports the range error."""
    fi = random_fno_input(13, n_vertices=300, n_sr=100, n_edges=8000)
    fi.edges["pos1"][::7] += 1 << 25
    fi.edges["len1"][::11] = -5
    assert capi.fno1(fi).tobytes() == O.fno1(fi).tobytes()
    with pytest.raises(capi.HcError):
        capi.fno1(fi, small=True)
