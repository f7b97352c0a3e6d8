#!/usr/bin/env python
"""Golden vectors for the converter scripts/sam2overlaps.py of the reference (reference-guided mode): seeded
SAM inputs (single-end and interleaved paired-end, soft / hard clips, insertions, deletions, reverse strands, unmapped and
mismatched ends, two reference sequences) and what the UNMODIFIED script logic writes for them.  The script is Python 2; it is
run from a temporary copy with its Python-2-isms rewritten (print statements, xrange, `from time import clock`, round() half
away from zero).  Writes tests/golden/sam_<name>.npz.  Run in the build container (/root/reference)."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SCRIPT = "/root/reference/scripts/sam2overlaps.py"


def py3_copy(d):
    src = open(REF_SCRIPT).read()
    src = re.sub(r'^(\s*)print (.*)$', r'\1print(\2)', src, flags=re.M)
    src = src.replace("xrange(", "range(")
    src = src.replace("from time import clock\n", "import math\n\ndef _round2(x):\n    return math.copysign(math.floor(abs(x) + 0.5), x) if abs(x) % 1 == 0.5 else float(round(x))\n")
    src = src.replace("int(round(ovlen / min(len(seq1), len(seq2)) * 100))", "int(_round2(ovlen / min(len(seq1), len(seq2)) * 100))")
    path = os.path.join(d, "sam2overlaps_py3.py")
    open(path, "w").write(src)
    return path


def rand_cigar(rng, qlen):
    """A CIGAR whose M/I/S lengths add up to qlen (hard clips on top)."""
    kind = rng.randint(0, 8)
    pre = post = ""
    if kind == 1:
        s = int(rng.randint(1, 8)); pre = "%dS" % s; qlen -= s
    elif kind == 2:
        pre = "%dH" % int(rng.randint(1, 8))
    if kind == 3:
        post = "%dH" % int(rng.randint(1, 8))
    if kind in (4, 6) and qlen > 30:
        a = int(rng.randint(5, qlen - 10)); ins = int(rng.randint(1, 4))
        core = "%dM%dI%dM" % (a, ins, qlen - a - ins)
    elif kind in (5, 7) and qlen > 30:
        a = int(rng.randint(5, qlen - 10)); dl = int(rng.randint(1, 5))
        core = "%dM%dD%dM" % (a, dl, qlen - a)
    else:
        core = "%dM" % qlen
    return pre + core + post


def qlen_of(cigar):
    return sum(int(n) for n, t in re.findall(r"(\d+)([MIDSH])", cigar) if t in "MIS")


def sam_line(rng, rid, flag, ref, pos, qlen):
    c = rand_cigar(rng, qlen)
    n = qlen_of(c)
    seq = "".join(rng.choice(list("ACGT"), n))
    return "%s\t%d\t%s\t%d\t60\t%s\t=\t0\t0\t%s\t%s" % (rid, flag, ref, pos, c, seq, "I" * n)


def make_case(seed, n_single, n_pairs, refs):
    rng = np.random.RandomState(seed)
    fasta = "".join(">%s description\n%s\n" % (name, "\n".join("".join(rng.choice(list("ACGT"), min(70, L - o))) for o in range(0, L, 70)))
                    for name, L in refs)
    header = "".join("@SQ\tSN:%s\tLN:%d\n" % (n, L) for n, L in refs) + "@PG\tID:x\n"
    s_lines, p_lines = [], []
    for i in range(n_single):
        name, L = refs[rng.randint(0, len(refs))]
        flag = [0, 16, 4][rng.choice(3, p=[0.55, 0.35, 0.1])]
        s_lines.append(sam_line(rng, "s%d" % i, flag, name, int(rng.randint(1, L + 40)), int(rng.randint(60, 151))))
    for i in range(n_pairs):
        name, L = refs[rng.randint(0, len(refs))]
        p1 = int(rng.randint(1, L))
        p2 = p1 + int(rng.randint(-20, 200))
        r = rng.random_sample()
        if r < 0.45: f1, f2 = 0, 0
        elif r < 0.8: f1, f2, p1, p2 = 16, 16, max(p1, p2), min(p1, p2)
        elif r < 0.9: f1, f2 = 0, 16
        else: f1, f2 = 4, 0
        id2 = "p%d" % i if rng.random_sample() > 0.03 else "q%d" % i
        p_lines.append(sam_line(rng, "p%d" % i, f1, name, max(p1, 1), int(rng.randint(60, 151))))
        p_lines.append(sam_line(rng, id2, f2, name, max(p2, 1), int(rng.randint(60, 151))))
    return fasta, header + "\n".join(s_lines) + "\n", header + "\n".join(p_lines) + "\n"


def main():
    d = tempfile.mkdtemp(prefix="hc_sam_")
    script = py3_copy(d)
    cases = (("singles", 1, 900, 0, [("refA", 1500)], 50, False), ("pairs", 2, 0, 700, [("refA", 1800)], 40, True),
             ("mixed_two_refs", 3, 500, 500, [("refA", 1200), ("refB", 900)], 30, True), ("mixed_min0", 4, 150, 150, [("refA", 700)], 0, False))
    for name, seed, ns, npairs, refs, min_ov, verbose in cases:
        fasta, sam_s, sam_p = make_case(seed, ns, npairs, refs)
        for fn, text in (("ref.fasta", fasta), ("s.sam", sam_s), ("p.sam", sam_p)):
            open(os.path.join(d, fn), "w").write(text)
        cmd = [sys.executable, script, "--ref", "ref.fasta", "--out", "out.txt", "--min_overlap_len", str(min_ov)]
        if ns: cmd += ["--sam_s", "s.sam"]
        if npairs: cmd += ["--sam_p", "p.sam"]
        if verbose: cmd += ["--verbose"]
        out = subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.PIPE, text=True).stdout
        res = open(os.path.join(d, "out.txt")).read()
        np.savez_compressed(os.path.join(GOLDEN, "sam_" + name + ".npz"), fasta=np.frombuffer(fasta.encode(), dtype=np.uint8),
                            sam_s=np.frombuffer(sam_s.encode(), dtype=np.uint8), sam_p=np.frombuffer(sam_p.encode(), dtype=np.uint8),
                            use_s=np.int64(ns > 0), use_p=np.int64(npairs > 0), min_overlap_len=np.int64(min_ov), verbose=np.int64(verbose),
                            overlaps=np.frombuffer(res.encode(), dtype=np.uint8), stdout=np.frombuffer(out.encode(), dtype=np.uint8))
        kinds = [l.split("\t")[11] + l.split("\t")[12] + l.split("\t")[4] for l in res.split("\n") if l]
        print("%-16s overlaps=%d %s" % (name, len(kinds), {k: kinds.count(k) for k in sorted(set(kinds))}))


if __name__ == "__main__":
    main()
