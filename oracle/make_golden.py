"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/ref_driver, built
from /root/reference/src by oracle/Makefile) -- run in the build container only:

    python -m oracle.make_golden

Each fixture holds the reads, the candidates, the parameters and what the reference produced:
per-candidate Edge fields + class (EdgeCalculator::compute_overlap, src/EdgeCalculator.cpp:143-385,
404-413, one thread, input order), the adjacency lists after construct_edges() (:561-666) and the
nonedge_overlaps.txt lines.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from haploconduct_b200 import formats as F, workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def run_case(name: str, rs: F.ReadSet, cands: np.ndarray, ps: dict) -> None:
    d = tempfile.mkdtemp(prefix="hc_golden_")
    F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    F.write_overlaps(d + "/ov.txt", cands, rs.ids)
    kw = dict(singles=d + "/s.fastq" if rs.n_single else None,
              paired1=d + "/p1.fastq" if rs.n_reads > rs.n_single else None,
              paired2=d + "/p2.fastq" if rs.n_reads > rs.n_single else None)
    out = O.run_ref(d, d + "/ov.txt", dump_cands=True, run=True, dump_graph=True, threads=1, **kw, **ps)
    pre = F.prefilter(cands, ps.get("min_overlap_len", 150), ps.get("min_overlap_perc", 0), ps.get("relax_PE_edges", False))
    assert int((pre == 1).sum()) == len(out["cands"]), (int((pre == 1).sum()), len(out["cands"]))
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        ids=rs.ids, descs=rs.descs, bases=rs.bases, quals=rs.quals, n_single=np.int64(rs.n_single), cands=cands,
        ps_keys=np.array(sorted(ps.keys())), ps_vals=np.array([float(ps[k]) for k in sorted(ps.keys())]),
        ref_cands=out["cands"], ref_graph=out["graph"], ref_nonedge=np.array(out.get("nonedge_lines", []), dtype=object).astype(str),
        ref_counts=np.array([out["graph_edges"], out["dup_count"], out["inclusion_count"]], dtype=np.int64),
    )
    # the same run under --ignore_inclusions: OverlapGraph::inclusions as EdgeCalculator.cpp:459-468 sets it
    # (the flag changes nothing else before the graph is handed on, src/ViralQuasispecies.cpp:304)
    inc = O.run_ref(d, d + "/ov.txt", run=True, dump_graph=True, threads=1, ignore_inclusions=1, **kw, **ps)
    assert inc["graph"].tobytes() == out["graph"].tobytes()
    np.savez_compressed(os.path.join(GOLDEN, "insert_" + name + ".npz"), n_vertices=np.int64(rs.n_reads), ref_inclusions=inc["inclusions"])
    cls = np.bincount(out["cands"]["cls"], minlength=3)
    print("%-28s reads=%d cands=%d scored=%d  discard/edge/nonedge=%s graph_edges=%d" %
          (name, rs.n_reads, len(cands), len(out["cands"]), cls.tolist(), out["graph_edges"]))


def save_fno_state(name: str, rs: F.ReadSet, d: str) -> None:
    """What the product's C++ binding (haploconduct_b200/host/hcb_fno.h, hc_fno) starts from: the reads, the state the
    reference's findNextOverlaps* started from (ref_driver --fno-state) and nonedge_overlaps.txt; the expected output
    is ref_lines of <name>.npz."""
    with open(d + "/state.txt", "rb") as f:
        state = f.read()
    nonedge = b""
    if os.path.exists(d + "/nonedge_overlaps.txt"):
        with open(d + "/nonedge_overlaps.txt", "rb") as f:
            nonedge = f.read()
    np.savez_compressed(os.path.join(GOLDEN, "fnostate_" + name + ".npz"), ids=rs.ids, descs=rs.descs, bases=rs.bases, quals=rs.quals,
                        n_single=np.int64(rs.n_single), state=np.frombuffer(state, dtype=np.uint8),
                        nonedge=np.frombuffer(nonedge, dtype=np.uint8))


def run_fno_case(name: str, rs: F.ReadSet, cands: np.ndarray, args: list) -> None:
    """Merge iteration + the reference's own findNextOverlaps(): keeps its inputs (array form) and its overlaps.txt."""
    import subprocess

    d = tempfile.mkdtemp(prefix="hc_golden_fno_")
    F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    F.write_overlaps(d + "/ov.txt", cands, rs.ids)
    cmd = [O.REF_DRIVER, "--overlaps", d + "/ov.txt", "--run", "--merge-fno1", d + "/fno_in.txt", "--fno-state", d + "/state.txt"] + args
    if rs.n_single:
        cmd += ["--singles", d + "/s.fastq"]
    if rs.n_reads > rs.n_single:
        cmd += ["--paired1", d + "/p1.fastq", "--paired2", d + "/p2.fastq"]
    subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    fi = O.parse_fno_dump(d + "/fno_in.txt")
    with open(d + "/overlaps.txt") as f:
        ref = f.read().split("\n")[:-1]
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), visited=fi.visited, label=fi.label, vertex_read=fi.vertex_read,
                        sr_off=fi.sr_off, sr_idx=fi.sr_idx, sr_sub=fi.sr_sub, superread=fi.superread,
                        flags=np.array([fi.resolve_orientations, fi.no_inclusions]), edges=fi.edges, ref_lines=np.array(ref))
    save_fno_state(name, rs, d)
    print("%-28s vertices=%d superreads=%d (paired %d) edges=%d -> %d overlap lines" %
          (name, len(fi.visited), len(fi.superread), int((fi.superread["len2"] > 0).sum()), len(fi.edges), len(ref)))


def run_fno3_case(name: str, rs: F.ReadSet, cands: np.ndarray, args: list) -> None:
    """Clique iteration + the reference's own findNextOverlaps3(): its inputs (array form, reference iteration order) and overlaps.txt."""
    import subprocess

    d = tempfile.mkdtemp(prefix="hc_golden_fno3_")
    F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    F.write_overlaps(d + "/ov.txt", cands, rs.ids)
    cmd = [O.REF_DRIVER, "--overlaps", d + "/ov.txt", "--run", "--merge-fno3", d + "/fno3_in.txt", "--fno-state", d + "/state.txt",
           "--cliques", "1", "--remove_branches", "0"] + args
    if rs.n_single:
        cmd += ["--singles", d + "/s.fastq"]
    if rs.n_reads > rs.n_single:
        cmd += ["--paired1", d + "/p1.fastq", "--paired2", d + "/p2.fastq"]
    subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    fi = O.parse_fno3_dump(d + "/fno3_in.txt")
    with open(d + "/overlaps.txt") as f:
        ref = f.read().split("\n")[:-1]
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), off=fi.off, sr_idx=fi.sr_idx, sr_pos=fi.sr_pos, reads=fi.reads,
                        flags=np.array([fi.no_inclusions]), ref_lines=np.array(ref))
    save_fno_state(name, rs, d)
    print("%-28s originals=%d reads=%d -> %d overlap lines" % (name, len(fi.off) - 1, len(fi.reads), len(ref)))


def run_ingest_case(name: str, rs: F.ReadSet, cands: np.ndarray, seed: int, allow_spaces: bool, ps: dict) -> None:
    """Irregularly spelled overlaps file through the real construct_edges() with thresholds under which every
    candidate that reaches process_overlaps is printed back as a non-edge (edge_threshold 2, merge_contigs -1,
    ov_threshold -1): nonedge_overlaps.txt then holds, re-printed by Overlap::get_overlap_line, first every scored
    line in order (:546-555), then every pre-filtered line in order (:654-660)."""
    d = tempfile.mkdtemp(prefix="hc_golden_ing_")
    F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    text = W.fuzz_overlap_text(cands, rs.ids, seed=seed, allow_spaces=allow_spaces)
    with open(d + "/ov.txt", "wb") as f:
        f.write(text)
    kw = dict(singles=d + "/s.fastq" if rs.n_single else None,
              paired1=d + "/p1.fastq" if rs.n_reads > rs.n_single else None,
              paired2=d + "/p2.fastq" if rs.n_reads > rs.n_single else None)
    out = O.run_ref(d, d + "/ov.txt", run=True, dump_cands=True, threads=1, edge_threshold=2, merge_contigs=-1, ov_threshold=-1,
                    allow_spaced_overlaps=int(allow_spaces), **kw, **ps)
    assert out["graph_edges"] == 0
    lines = out.get("nonedge_lines", [])
    n_f = int(out["len_filtered"])
    assert len(lines) == int(out["scored"]) + n_f, (len(lines), out["scored"], n_f)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"), ids=rs.ids, text=np.frombuffer(text, dtype=np.uint8),
        ps_keys=np.array(sorted(ps.keys())), ps_vals=np.array([float(ps[k]) for k in sorted(ps.keys())]),
        allow_spaces=np.int64(allow_spaces), ref_scored=np.array(lines[:len(lines) - n_f], dtype=object).astype(str),
        ref_filtered=np.array(lines[len(lines) - n_f:], dtype=object).astype(str),
        ref_counts=np.array([out["lines"], out["scored"], n_f, out["self"], out["perc_dropped"], out["bad"]], dtype=np.int64))
    print("%-28s lines=%d scored=%d filtered=%d self=%d perc_dropped=%d skipped=%d" %
          (name, out["lines"], out["scored"], n_f, out["self"], out["perc_dropped"], out["bad"]))


def run_consensus_case(name: str, seed: int, n_problems: int, min_clique_size: int, min_qual: float, **kw) -> None:
    """Pile-ups through the reference's own SRBuilder::consensus (oracle/ref_driver --consensus)."""
    import subprocess
    rs, probs = W.consensus_problems(seed=seed, n_problems=n_problems, **kw)
    d = tempfile.mkdtemp(prefix="hc_golden_cons_")
    with open(d + "/in.txt", "w") as f:
        f.write(W.consensus_problem_text(rs, probs))
    open(d + "/ov.txt", "w").close()
    F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    subprocess.run([O.REF_DRIVER, "--overlaps", d + "/ov.txt", "--singles", d + "/s.fastq", "--paired1", d + "/p1.fastq", "--paired2",
                    d + "/p2.fastq", "--min_clique_size", str(min_clique_size), "--min_qual", repr(min_qual), "--consensus", d + "/in.txt",
                    d + "/out.txt"], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    ref = [l.rstrip("\n").split("\t") for l in open(d + "/out.txt")]
    assert len(ref) == len(probs)
    P, S = F.consensus_arrays(probs)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"), ids=rs.ids, descs=rs.descs, bases=rs.bases, quals=rs.quals, n_single=np.int64(rs.n_single),
        problems=P, seqs=S, params=np.array([min_clique_size, min_qual]), ref_ret=np.array([int(r[1]) for r in ref], dtype=np.int32),
        ref_seq=np.array(["" if r[2] == "0" else r[3] for r in ref], dtype=object).astype(str),
        ref_qual=np.array(["" if r[2] == "0" else r[4] for r in ref], dtype=object).astype(str))
    print("%-28s problems=%d reads=%d non-empty=%d ret<0=%d N columns=%d" % (
        name, len(probs), rs.n_reads, sum(1 for r in ref if r[2] != "0"), sum(1 for r in ref if int(r[1]) < 0), sum(r[3].count("N") for r in ref)))


def fno_cases(full: F.ReadSet) -> None:
    # ---- FindNextOverlaps (FNO1): the reference's overlaps.txt after one merge iteration
    s700 = full.subset(range(0, 700))
    cs = W.seed_candidates(s700, k=24, orientations=((1, 1),))
    run_fno_case("fno1_savage_singles", s700, cs, ["--edge_threshold", "0.97", "--min_overlap_len", "200", "--keep_singletons", "200"])
    for k, (seed, ns, npair, div) in enumerate(((5, 150, 250, (0.0, 0.0, 0.0)), (6, 0, 400, (0.0, 0.0)), (7, 300, 100, (0.0, 0.01)))):
        sx = W.synth_readset(ns, npair, genome_len=1500, n_strains=len(div), divergence=div, seed=seed, n_rate=0.0002, flip_fraction=0.2)
        cx = W.geometry_candidates(sx, 6000, seed=seed + 1, junk_fraction=0.02, min_ov=40)
        run_fno_case("fno1_synth_paired_%d" % k, sx.rs, cx, ["--edge_threshold", "0.9", "--min_overlap_len", "80", "--keep_singletons", "0"]
                     + (["--no_inclusion_overlaps", "1"] if k == 2 else []))
        if k != 1:
            run_fno3_case("fno3_synth_cliques_%d" % k, sx.rs, cx, ["--edge_threshold", "0.9", "--min_overlap_len", "80", "--keep_singletons", "0"]
                          + (["--no_inclusion_overlaps", "1"] if k == 2 else []))
    # config 5: contigs of 1-10 kb (Phred up to 93) tiled over 3 strains, S-S overlaps, the flags of SAVAGE stage b
    # (scripts/pipeline_per_stage.py:214-247: FNO=1, remove_trans=1, optimize=false, ignore_inclusions, keep_singletons =
    # max(min_overlap_len, min_read_len)), one merge iteration + FNO1
    sc = W.synth_readset(500, 0, genome_len=60000, read_len=(1000, 10000), qmax=93, q_lo=30, n_strains=3, divergence=(0.0, 0.01, 0.02),
                         seed=20261019, n_rate=0.0)
    cc = W.geometry_candidates(sc, 30000, seed=11, junk_fraction=0.02, min_ov=100)
    run_fno_case("fno1_contigs_stage_b", sc.rs, cc, ["--edge_threshold", "0.995", "--min_overlap_len", "100", "--keep_singletons", "100",
                                                      "--ignore_inclusions", "1", "--min_read_len", "100"])


def main() -> None:
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(GOLDEN, exist_ok=True)
    if "--only-full" in sys.argv:
        # the two example data sets at their full size: every read, every seed-enumerated candidate of all read types
        R = REF + "/savage/example/input_fas/"
        full = F.load_fastq_set(R + "singles.fastq", R + "paired1.fastq", R + "paired2.fastq")
        c = W.seed_candidates(full, k=24, seed=3)
        run_case("c1_savage_example_full", full, c, dict(edge_threshold=0.97, min_overlap_len=200))
        Rp = REF + "/polyte/example/input/"
        fw = F._read_fastq_records(Rp + "forward.fastq")
        rv = F._read_fastq_records(Rp + "reverse.fastq")
        rs2 = F.ReadSet.from_lists([(i, s.upper(), q) for i, (_, s, q) in enumerate(fw + rv)], [])
        c2 = W.seed_candidates(rs2, k=20, seed=4, orientations=((1, 1), (1, 0)))
        run_case("c2_polyte_example_full", rs2, c2, dict(edge_threshold=0.95, min_overlap_len=126))
        return
    if "--only-fno" in sys.argv:
        R = REF + "/savage/example/input_fas/"
        fno_cases(F.load_fastq_set(R + "singles.fastq", R + "paired1.fastq", R + "paired2.fastq"))
        return
    # C1: savage/example, stage-a parameters (savage.py:384-385, savage/README.md:303: -m 200)
    R = REF + "/savage/example/input_fas/"
    full = F.load_fastq_set(R + "singles.fastq", R + "paired1.fastq", R + "paired2.fastq")
    sub = full.subset(list(range(0, 260)) + list(range(2000, 2120)))
    c = W.seed_candidates(sub, k=24, max_cands=6000, seed=3)
    run_case("c1_savage_example_stage_a", sub, c, dict(edge_threshold=0.97, min_overlap_len=200))
    # C2: polyte/example, all reads as singles (polyte.py:280-288), iteration-1 parameters (polyte.py:599-602)
    Rp = REF + "/polyte/example/input/"
    fw = F._read_fastq_records(Rp + "forward.fastq")[:220]
    rv = F._read_fastq_records(Rp + "reverse.fastq")[:220]
    singles = [(i, s.upper(), q) for i, (_, s, q) in enumerate(fw + rv)]
    rs2 = F.ReadSet.from_lists(singles, [])
    c2 = W.seed_candidates(rs2, k=20, max_cands=6000, seed=4)
    run_case("c2_polyte_example_it1", rs2, c2, dict(edge_threshold=0.95, min_overlap_len=126))
    # synthetic: every TYPE x ORI x ORD case, junk candidates, N and Q=0 bases
    ss = W.synth_readset(160, 160, seed=20261017, n_rate=0.002)
    c3 = W.geometry_candidates(ss, 5000, seed=5)
    run_case("synth_all_types", ss.rs, c3, dict(edge_threshold=0.97, min_overlap_len=60))
    # stage-c like: contigs as singles with wide qualities, min_read_len, merge_contigs > 0
    ss4 = W.synth_readset(120, 0, genome_len=6000, read_len=(400, 1800), qmax=93, q_lo=30, seed=20261019, n_rate=0.0)
    c4 = W.geometry_candidates(ss4, 2500, seed=6, min_ov=60)
    run_case("synth_stage_c_contigs", ss4.rs, c4,
             dict(edge_threshold=0.995, min_overlap_len=100, min_read_len=500, merge_contigs=0.01))
    # non-default ps.mismatch (void overlaps) and relaxed paired-end filter
    ss5 = W.synth_readset(60, 120, seed=99, n_rate=0.001)
    c5 = W.geometry_candidates(ss5, 2500, seed=8)
    run_case("synth_mismatch_void", ss5.rs, c5, dict(edge_threshold=0.9, ov_threshold=0.5, min_overlap_len=120, mismatch=0.01,
                                                      relax_PE_edges=True))

    # ---- candidate ingestion: irregular spellings the parser accepts, through the real text loop
    run_ingest_case("ingest_tabs", ss.rs, c3, 5, False, dict(min_overlap_len=60, min_overlap_perc=30))
    run_ingest_case("ingest_tabs_relaxed", ss5.rs, c5, 6, False, dict(min_overlap_len=150, relax_PE_edges=True))
    run_ingest_case("ingest_spaces", ss.rs, c3, 7, True, dict(min_overlap_len=60, min_overlap_perc=30))

    # ---- super-read consensus: pile-ups through SRBuilder::consensus
    run_consensus_case("consensus_illumina", 3, 160, 3, 0.9)
    run_consensus_case("consensus_wide_qualities", 4, 120, 2, 0.99, qmax=93, read_len=(200, 900))

    fno_cases(full)

if __name__ == "__main__":
    main()
