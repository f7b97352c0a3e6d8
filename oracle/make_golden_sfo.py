#!/usr/bin/env python
"""Golden vectors for the converter scripts/sfo2overlaps.py of the reference: seeded SFO-format inputs and
what the UNMODIFIED script logic writes for them.  The script is Python 2; it is run here from a temporary copy with the three
Python-2-isms rewritten (print statements, xrange, round() half away from zero) under LC_ALL=C (its `sort | uniq` pipeline
compares whole lines when the four numeric keys tie).  Writes tests/golden/sfo_<name>.npz = {sfo text, num_singles, num_pairs,
expected overlaps text}.  Run in the build container (/root/reference)."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SCRIPT = "/root/reference/scripts/sfo2overlaps.py"


def py3_copy(d):
    src = open(REF_SCRIPT).read()
    src = re.sub(r'^(\s*)print (.*)$', r'\1print(\2)', src, flags=re.M)
    src = src.replace("xrange(", "range(")
    src = src.replace("import subprocess\n", "import subprocess\nimport math\n\ndef _round2(x):\n    return math.copysign(math.floor(abs(x) + 0.5), x) if abs(x) % 1 == 0.5 else float(round(x))\n", 1)
    src = src.replace("min(round(100*ovlen/minreadlen), 100)", "min(_round2(100*ovlen/minreadlen), 100)")
    path = os.path.join(d, "sfo2overlaps_py3.py")
    open(path, "w").write(src)
    return path


def make_sfo(seed, ns, npairs, n_lines, read_len=100):
    """Random overlaps between read ends (SFO ids: singles, then /1 ends, then /2 ends), both orientations, all four sign
    combinations of the overhangs, repeated lines, several overlaps per id pair, ends of one pair overlapping each other."""
    rng = np.random.RandomState(seed)
    n_ids = ns + 2 * npairs
    lines = []
    for _ in range(n_lines):
        if npairs and rng.random_sample() < 0.7:       # both ends of two pairs overlap consistently (a real paired overlap)
            a, b = rng.choice(npairs, 2, replace=False) if npairs > 1 else (0, 0)
            ori = "N" if rng.random_sample() < 0.7 else "I"
            for end in (0, 1):
                ia, ib = ns + a + end * npairs, ns + b + (end if ori == "N" else 1 - end) * npairs
                if ia > ib:
                    ia, ib = ib, ia
                oha = int(rng.randint(-60, 60))
                ohb = int(rng.randint(-60, 60))
                ola = int(rng.randint(20, read_len))
                olb = ola + int(rng.randint(-2, 3))
                lines.append("%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d" % (ia, ib, ori, oha, ohb, ola, max(olb, 1), rng.randint(0, 4)))
        else:
            ia, ib = rng.randint(0, n_ids, 2)
            oha, ohb = int(rng.randint(-60, 60)), int(rng.randint(-60, 60))
            ola = int(rng.randint(20, read_len))
            lines.append("%d %d\t%s\t%d\t%d\t%d\t%d\t%d" % (ia, ib, "N" if rng.random_sample() < 0.6 else "I", oha, ohb, ola, ola, rng.randint(0, 4)))
        if rng.random_sample() < 0.1:
            lines.append(lines[rng.randint(0, len(lines))])
    rng.shuffle(lines)
    return "\n".join(lines) + "\n"


def main():
    d = tempfile.mkdtemp(prefix="hc_sfo_")
    script = py3_copy(d)
    env = dict(os.environ, LC_ALL="C")
    for name, seed, ns, npairs, n in (("singles", 1, 300, 0, 4000), ("pairs", 2, 0, 200, 6000), ("mixed", 3, 150, 150, 8000), ("mixed_dense", 4, 20, 25, 5000)):
        text = make_sfo(seed, ns, npairs, n)
        open(os.path.join(d, "in.sfo"), "w").write(text)
        out = subprocess.run([sys.executable, script, "--in", "in.sfo", "--out", "out.txt", "--num_singles", str(ns), "--num_pairs", str(npairs)],
                             cwd=d, env=env, check=True, stdout=subprocess.PIPE, text=True).stdout
        res = open(os.path.join(d, "out.txt")).read()
        np.savez_compressed(os.path.join(GOLDEN, "sfo_" + name + ".npz"), sfo=np.frombuffer(text.encode(), dtype=np.uint8), num_singles=np.int64(ns),
                            num_pairs=np.int64(npairs), overlaps=np.frombuffer(res.encode(), dtype=np.uint8), stdout=np.frombuffer(out.encode(), dtype=np.uint8))
        kinds = [l.split("\t")[11] + l.split("\t")[12] for l in res.split("\n") if l]
        print("%-12s sfo lines=%d overlaps=%d  %s | %s" % (name, text.count("\n"), len(kinds), {k: kinds.count(k) for k in sorted(set(kinds))}, out.replace("\n", "; ")))


if __name__ == "__main__":
    main()
