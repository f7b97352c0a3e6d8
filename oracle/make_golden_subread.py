#!/usr/bin/env python
"""Golden vectors for SRBuilder::calcSubreadInfo (src/SRBuilder.cpp:536-595): seeded clique position lists and what the
UNMODIFIED function returns for them (oracle/_ref/ref_driver --subread-info calls the private member directly).  Writes
tests/golden/subread_info.npz.  Run in the build container."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_problems(seed, n):
    rng = np.random.RandomState(seed)
    probs = []
    for _ in range(n):
        k = int(rng.randint(2, 40))
        verts = rng.choice(5000, k, replace=False)
        pos1 = np.sort(rng.randint(0, 400, k)); pos1[0] = 0
        trim1 = int(pos1[min(int(rng.randint(0, 4)), k - 1)]) if rng.random_sample() < 0.8 else 0
        if rng.random_sample() < 0.5:                      # paired-end super-read: list 2 = the same vertices in another order
            order = rng.permutation(k)
            pos2 = np.sort(rng.randint(0, 400, k)); pos2[0] = 0
            trim2 = int(pos2[min(int(rng.randint(0, 4)), k - 1)])
            probs.append((trim1, trim2, pos1.tolist(), verts.tolist(), pos2.tolist(), verts[order].tolist()))
        else:                                              # single-end super-read; some vertices twice (both mates of a paired read)
            v1, p1 = verts.tolist(), pos1.tolist()
            for j in rng.choice(k, int(rng.randint(0, max(k // 3, 1))), replace=False):
                v1.append(int(verts[j])); p1.append(int(p1[-1] + rng.randint(0, 50)))
            probs.append((trim1, -1, p1, v1, [], []))
    return probs


def main():
    probs = make_problems(17, 3000)
    d = tempfile.mkdtemp(prefix="hc_subread_")
    with open(d + "/in.txt", "w") as f:
        for t1, t2, p1, v1, p2, v2 in probs:
            f.write("S %d %d %d %d\n" % (t1, t2, len(p1), len(p2)))
            f.write(" ".join("%d %d" % (p, v) for p, v in zip(p1, v1)) + "\n")
            f.write(" ".join("%d %d" % (p, v) for p, v in zip(p2, v2)) + "\n")
    open(d + "/ov.txt", "w").close()
    subprocess.run([O.REF_DRIVER, "--overlaps", d + "/ov.txt", "--subread-info", d + "/in.txt", d + "/out.txt"], check=True, cwd=d,
                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    lines = open(d + "/out.txt").read().split("\n")
    k = 0
    ref = []
    for t1, t2, p1, v1, p2, v2 in probs:
        assert lines[k].startswith("R\t")
        m = int(lines[k].split("\t")[1]); k += 1
        got = {}
        for _ in range(m):
            t = [int(x) for x in lines[k].split("\t")]; k += 1
            got[t[0]] = tuple(t[1:])
        assert got == O.calc_subread_info(t1, t2, p1, v1, p2, v2), "restatement differs from the reference"
        ref.append(got)
    # flat arrays as hc_subread_info takes them: list 1 of every problem, then its list 2
    pos, vertex, P, exp = [], [], [], []
    for (t1, t2, p1, v1, p2, v2), got in zip(probs, ref):
        b1 = len(pos); pos += p1; vertex += v1
        b2 = len(pos); pos += p2; vertex += v2
        P.append((b1, b1 + len(p1), b2, b2 + len(p2), t1, t2))
        for v in sorted(got):
            exp.append((len(P) - 1, v) + got[v])
    np.savez_compressed(os.path.join(GOLDEN, "subread_info.npz"), problems=np.array(P, dtype=np.int64), pos=np.array(pos, dtype=np.int32),
                        vertex=np.array(vertex, dtype=np.uint32), expected=np.array(exp, dtype=np.int64))
    print("problems=%d entries=%d records=%d restatement_equal=True" % (len(P), len(pos), len(exp)))


if __name__ == "__main__":
    main()
