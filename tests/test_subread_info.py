"""SRBuilder::calcSubreadInfo (src/SRBuilder.cpp:536-595): the restatement oracle.calc_subread_info against what the
reference's own function returned (tests/golden/subread_info.npz, oracle/make_golden_subread.py: ref_driver calls the
private member directly) -- CPU -- and hc_subread_info against both -- GPU."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from util import GOLDEN


def _load():
    z = np.load(os.path.join(GOLDEN, "subread_info.npz"))
    return z["problems"], z["pos"], z["vertex"], z["expected"]


def test_restatement_is_pinned():
    P, pos, vertex, exp = _load()
    by_problem = {}
    for row in exp:
        by_problem.setdefault(int(row[0]), {})[int(row[1])] = tuple(int(x) for x in row[2:])
    assert len(by_problem) == len(P)
    for k, (b1, e1, b2, e2, t1, t2) in enumerate(P):
        got = O.calc_subread_info(int(t1), int(t2), pos[b1:e1].tolist(), vertex[b1:e1].tolist(), pos[b2:e2].tolist(), vertex[b2:e2].tolist())
        assert got == by_problem[k], k


@pytest.mark.gpu
def test_device_subread_info_equals_the_reference(built_lib):
    from haploconduct_b200 import capi

    P, pos, vertex, exp = _load()
    probs = np.zeros(len(P), dtype=capi.SUBREAD_PROBLEM)
    for j, f in enumerate(("begin1", "end1", "begin2", "end2", "trim_pos1", "trim_pos2")):
        probs[f] = P[:, j]
    info, first = capi.subread_info(probs, pos, vertex)
    rows = []
    for k, (b1, e1, b2, e2, t1, t2) in enumerate(P):
        own = np.nonzero(first[b1:e1])[0] + b1
        assert not first[b2:e2].any()
        order = own[np.argsort(vertex[own], kind="stable")]
        for i in order:
            rows.append((k, int(vertex[i]), int(info["index1"][i]), int(info["index2"][i]), int(info["startpos1"][i]), int(info["startpos2"][i])))
    assert np.array_equal(np.array(rows, dtype=np.int64), exp)


@pytest.mark.gpu
def test_device_subread_info_bad_input(built_lib):
    from haploconduct_b200 import capi

    probs = np.zeros(2, dtype=capi.SUBREAD_PROBLEM)
    probs["end1"] = [3, 4]
    probs["begin1"] = [0, 2]                      # list 1 of the second problem overlaps the first's
    probs["trim_pos2"] = -1
    with pytest.raises(capi.HcError):
        capi.subread_info(probs, np.zeros(4, np.int32), np.arange(4, dtype=np.uint32))
    info, first = capi.subread_info(probs[:0], np.zeros(0, np.int32), np.zeros(0, np.uint32))
    assert len(info) == 0
