"""GPU suite: the C++ host mirror of the reference's EdgeCalculator (haploconduct_b200/host/, built
as lib/hc_edgecalc on top of the C ABI) against what the UNMODIFIED reference produced on the same
files (tests/golden): the overlap graph after construct_edges() -- every Edge field of every
adjacency list, in adjacency order -- nonedge_overlaps.txt byte for byte, and the public counters.
With --exact_scores=true the Edge scores are bit-identical; in the default fast mode they are
within 1e-6 relative and the graph (edges, order, integer fields) is still identical."""
import json
import os
import subprocess

import numpy as np
import pytest

from haploconduct_b200 import build as B, formats as F
from oracle import oracle as O
from util import golden_names, load_golden

pytestmark = pytest.mark.gpu
EXE = os.path.join(B.LIBDIR, "hc_edgecalc")


def _run(g, tmp_path, exact, gpu_dedup=False, gpu_parse=False, gpu_fastq=False, dump_sorted=False, ids_table=False):
    d = str(tmp_path)
    F.write_fastq_set(g.rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    F.write_overlaps(d + "/ov.txt", g.cands, g.rs.ids)
    if ids_table:   # headers become names ("@r<id>x some text"), the --IDs table maps the names back to the ids (src/FastqStorage.cpp:60-90)
        for fn in ("s.fastq", "p1.fastq", "p2.fastq"):
            if not os.path.exists(d + "/" + fn):
                continue
            with open(d + "/" + fn) as f:
                lines = f.read().split("\n")
            for k in range(0, len(lines) - 1, 4):
                lines[k] = "@r%sx extra" % lines[k][1:].split()[0]
            with open(d + "/" + fn, "w") as f:
                f.write("\n".join(lines))
        with open(d + "/ids.txt", "w") as f:
            for i in g.rs.ids:
                f.write("%d\t%sr%dx\n" % (int(i), ">" if int(i) % 2 else "", int(i)))
    cmd = [EXE, "--overlaps", d + "/ov.txt", "--dump-graph", d + "/graph.tsv", "--digraph", d + "/digraph.txt",
           "--exact_scores=" + ("true" if exact else "false"), "--gpu_dedup=" + ("true" if gpu_dedup else "false"), "--gpu_parse=" + ("true" if gpu_parse else "false"),
           "--gpu_fastq=" + ("true" if gpu_fastq else "false")]
    if dump_sorted:
        cmd += ["--dump-sorted", d + "/sorted.tsv"]
    if ids_table:
        cmd += ["--IDs", d + "/ids.txt"]
    if g.rs.n_single:
        cmd += ["--singles", d + "/s.fastq"]
    if g.rs.n_reads > g.rs.n_single:
        cmd += ["--paired1", d + "/p1.fastq", "--paired2", d + "/p2.fastq"]
    for k, v in g.ps.items():
        if k in ("min_overlap_len", "min_read_len", "min_overlap_perc"):
            cmd += ["--" + k, str(int(v))]
        elif k == "relax_PE_edges":
            cmd += ["--" + k + "=" + ("true" if v else "false")]
        else:
            cmd += ["--" + k, repr(float(v))]
    out = subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
    summary = json.loads([l for l in out.split("\n") if l.startswith("{")][-1])
    graph = O.parse_graph_dump(d + "/graph.tsv")
    with open(d + "/nonedge_overlaps.txt") as f:
        nonedge = f.read().split("\n")[:-1]
    with open(d + "/digraph.txt") as f:
        digraph = f.read()
    return summary, graph, nonedge, digraph


@pytest.mark.parametrize("name", golden_names())
def test_graph_identical_to_reference_exact_scores(built_lib, tmp_path, name):
    g = load_golden(name)
    summary, graph, nonedge, digraph = _run(g, tmp_path, exact=True)
    assert np.array_equal(graph, g.ref_graph), name     # every field incl. score and mismatch_rate bit for bit, adjacency order
    assert nonedge == g.ref_nonedge
    assert [summary["graph_edges"], summary["dup_count"], summary["inclusion_count"]] == g.ref_counts.tolist()
    want = "".join("%d\t%d\n" % (a, b) for a, b in zip(g.ref_graph["v1"], g.ref_graph["v2"]))
    assert digraph == want


@pytest.mark.parametrize("name", golden_names())
def test_graph_identical_with_device_dedup(built_lib, tmp_path, name):
    """--gpu_dedup=true: the serial insert (:429-545) replaced by hc_dedup_edges + an ordered append."""
    g = load_golden(name)
    summary, graph, nonedge, digraph = _run(g, tmp_path, exact=True, gpu_dedup=True)
    assert np.array_equal(graph, g.ref_graph), name
    assert nonedge == g.ref_nonedge
    assert [summary["graph_edges"], summary["dup_count"], summary["inclusion_count"]] == g.ref_counts.tolist()


@pytest.mark.parametrize("name", golden_names())
def test_graph_identical_with_device_parse_and_dedup(built_lib, tmp_path, name):
    """--gpu_fastq --gpu_parse --gpu_dedup: FASTQ reading (src/FastqStorage.cpp:92-235), the overlaps-file text loop
    (:581-645) and duplicate resolution (:429-545) on the device too: files in, graph out."""
    g = load_golden(name)
    summary, graph, nonedge, digraph = _run(g, tmp_path, exact=True, gpu_dedup=True, gpu_parse=True, gpu_fastq=True)
    assert np.array_equal(graph, g.ref_graph), name
    assert nonedge == g.ref_nonedge
    assert [summary["graph_edges"], summary["dup_count"], summary["inclusion_count"]] == g.ref_counts.tolist()


@pytest.mark.parametrize("gpu_parse", [False, True])
@pytest.mark.parametrize("name,reads", [("ingest_tabs", "synth_all_types"), ("ingest_tabs_relaxed", "synth_mismatch_void"),
                                        ("ingest_spaces", "synth_all_types")])
def test_irregular_overlaps_file(built_lib, tmp_path, name, reads, gpu_parse):
    """An irregularly spelled overlaps file, with thresholds under which every surviving line is printed back:
    nonedge_overlaps.txt must be what the reference's construct_edges() wrote (host parser and device parser)."""
    from util import IngestGolden
    ig, g = IngestGolden(name), load_golden(reads)
    d = str(tmp_path)
    F.write_fastq_set(g.rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    with open(d + "/ov.txt", "wb") as f:
        f.write(ig.text)
    cmd = [EXE, "--overlaps", d + "/ov.txt", "--edge_threshold", "2", "--merge_contigs", "-1", "--ov_threshold", "-1",
           "--singles", d + "/s.fastq", "--paired1", d + "/p1.fastq", "--paired2", d + "/p2.fastq",
           "--gpu_parse=" + ("true" if gpu_parse else "false"), "--allow_spaced_overlaps=" + ("true" if ig.allow_spaces else "false")]
    for k, v in ig.ps.items():
        cmd += ["--" + k + "=true"] if k == "relax_PE_edges" else ["--" + k, str(int(v))]
    out = subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
    assert out.count("incorrect overlap; skipping") == int(ig.ref_counts[5])
    with open(d + "/nonedge_overlaps.txt") as f:
        nonedge = f.read().split("\n")[:-1]
    assert nonedge == ig.ref_scored + ig.ref_filtered


@pytest.mark.parametrize("name", golden_names())
def test_graph_fast_mode(built_lib, tmp_path, name):
    g = load_golden(name)
    summary, graph, nonedge, digraph = _run(g, tmp_path, exact=False)
    ref = g.ref_graph
    assert nonedge == g.ref_nonedge
    assert len(graph) == len(ref)

    def keyed(a):
        d = {}
        for e in a:
            k = (min(int(e["v1"]), int(e["v2"])), max(int(e["v1"]), int(e["v2"])), bool(e["ori1"] == e["ori2"]))
            assert k not in d
            d[k] = e
        return d

    mine, theirs = keyed(graph), keyed(ref)
    assert set(mine) == set(theirs)            # same edges between the same read pairs / relative orientations
    # read pairs with ONE accepted overlap must agree on every field; pairs with several accepted
    # overlaps are resolved by "score >= existing score" (:470) and may pick another representative
    # when their scores agree to 1e-7 -- the documented difference of the fast mode.
    rc = g.ref_cands[g.ref_cands["cls"] == 1]
    multiplicity = {}
    for c in rc:
        k = (min(int(c["v1"]), int(c["v2"])), max(int(c["v1"]), int(c["v2"])), bool(c["ori1"] == c["ori2"]))
        multiplicity[k] = multiplicity.get(k, 0) + 1
    n_checked = 0
    for k, e in theirs.items():
        if multiplicity[k] == 1:
            m = mine[k]
            for f in ("v1", "v2", "pos1", "pos2", "pos3", "pos4", "ori1", "ori2", "ord", "perc", "len1", "len2", "mismatch_rate"):
                assert m[f] == e[f], (k, f)
            assert abs(m["score"] - e["score"]) <= 1e-6 * e["score"]
            n_checked += 1
        else:
            assert abs(mine[k]["score"] - e["score"]) <= 1e-6 * e["score"]
    assert n_checked > 0


SORTED = sorted(f[len("sorted_"):-4] for f in os.listdir(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")) if f.startswith("sorted_"))


@pytest.mark.parametrize("gpu", [False, True])
@pytest.mark.parametrize("name", SORTED)
def test_sort_edges_like_the_reference(built_lib, tmp_path, name, gpu):
    """OverlapGraph::sortEdges() after construct_edges() (src/ViralQuasispecies.cpp:297): adjacency lists and adj_in as the
    unmodified reference leaves them -- with the order taken from hc_build_adjacency (gpu) and with std::sort on the host."""
    g = load_golden(name)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sorted_" + name + ".npz"))
    _run(g, tmp_path, exact=True, gpu_dedup=gpu, gpu_parse=gpu, gpu_fastq=gpu, dump_sorted=True)
    d = str(tmp_path)
    assert O.parse_graph_dump(d + "/sorted.tsv").tobytes() == z["ref_sorted"].tobytes()
    vs, off, src = O.parse_adj_in(d + "/sorted.tsv")
    assert np.array_equal(vs, z["in_vertices"]) and np.array_equal(off, z["in_off"]) and np.array_equal(src, z["in_src"])


@pytest.mark.parametrize("gpu", [False, True])
def test_read_names_through_the_ids_table(built_lib, tmp_path, gpu):
    """--IDs (src/FastqStorage.cpp:60-90): FASTQ headers are names, the table gives the ids the overlaps file uses -- with the
    host FASTQ reader and with the device one (--gpu_fastq), same graph as the fixture with numeric headers."""
    g = load_golden("synth_all_types")
    summary, graph, nonedge, digraph = _run(g, tmp_path, exact=True, gpu_dedup=gpu, gpu_parse=gpu, gpu_fastq=gpu, ids_table=True)
    assert np.array_equal(graph, g.ref_graph)
    assert nonedge == g.ref_nonedge
