// hc_edgecalc -- command-line front end of the host mirror: the EdgeCalculator stage of
// bin/ViralQuasispecies (src/ViralQuasispecies.cpp:233-283) on the B200 path.  Flag names are the
// reference's (src/ViralQuasispecies.cpp:52-98); output files are the reference's
// (nonedge_overlaps.txt in output_dir, digraph.txt on request).  Prints one JSON summary line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/time.h>
#include <thread>
#include <unistd.h>

#include "hcb_host.h"

static double now_s() {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

int main(int argc, char** argv) {
    hcb::ProgramSettings ps;
    std::string dump_graph, digraph, dump_sorted;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        std::string v;
        size_t eq = a.find('=');
        if (eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
        else if (i + 1 < argc) v = argv[++i];
        auto b = [&](const std::string& s) { return s == "true" || s == "1"; };
        if (a == "--singles" || a == "-s") ps.singles_file = v;
        else if (a == "--paired1") ps.paired1_file = v;
        else if (a == "--paired2") ps.paired2_file = v;
        else if (a == "--overlaps") ps.overlaps_file = v;
        else if (a == "--output" || a == "-O") ps.output_dir = v;
        else if (a == "--IDs") ps.id_correspondence = v;
        else if (a == "--max_ov") ps.max_overlaps = strtoul(v.c_str(), NULL, 10);
        else if (a == "--max_reads") ps.max_reads = strtoul(v.c_str(), NULL, 10);
        else if (a == "--threads" || a == "-t") ps.n_threads = atoi(v.c_str());
        else if (a == "--min_overlap_perc") ps.min_overlap_perc = atoi(v.c_str());
        else if (a == "--min_overlap_len") ps.min_overlap_len = atoi(v.c_str());
        else if (a == "--edge_threshold") ps.edge_threshold = atof(v.c_str());
        else if (a == "--ov_threshold") ps.ov_threshold = atof(v.c_str());
        else if (a == "--allow_spaced_overlaps") ps.allow_spaces = b(v);
        else if (a == "--add_duplicates") ps.add_duplicates = b(v);
        else if (a == "--ignore_inclusions") ps.ignore_inclusions = b(v);
        else if (a == "--mismatch") ps.mismatch = atof(v.c_str());
        else if (a == "--merge_contigs") ps.merge_contigs = atof(v.c_str());
        else if (a == "--min_read_len") ps.min_read_len = atoi(v.c_str());
        else if (a == "--relax_PE_edges") ps.relax_PE_edges = b(v);
        else if (a == "--verbose" || a == "-v") ps.verbose = b(v);
        else if (a == "--exact_scores") ps.exact_scores = b(v);
        else if (a == "--gpu_dedup") ps.gpu_dedup = b(v);
        else if (a == "--gpu_parse") ps.gpu_parse = b(v);
        else if (a == "--gpu_fastq") ps.gpu_fastq = b(v);
        else if (a == "--gpus") ps.n_devices = atoi(v.c_str());
        else if (a == "--dump-graph") dump_graph = v;
        else if (a == "--dump-sorted") dump_sorted = v;        // OverlapGraph::sortEdges() (src/ViralQuasispecies.cpp:297), then the same dump + adj_in
        else if (a == "--digraph") digraph = v;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (ps.overlaps_file.empty()) { fprintf(stderr, "No overlaps file provided.\n"); return 1; }
    const double t_start = now_s();
    // --gpu_parse: the overlaps file is read by a thread of its own while the read store is built (the stage is a chain
    // file -> device on both inputs; the reads must be on the device first, the text can be in memory by then)
    hcb::FileBuf ov_text;
    bool ov_ok = false;
    std::thread ov_reader;
    if (ps.gpu_parse && ps.gpu_fastq)
        hcb::FastqStorage::device_ready_hook = [&]() { ov_reader = std::thread([&]() { ov_ok = hcb::read_whole_file(ps.overlaps_file, ov_text, 8); }); };
    double t0 = now_s();
    std::shared_ptr<hcb::FastqStorage> fastq(new hcb::FastqStorage(ps));
    double t_fastq = now_s() - t0;
    std::shared_ptr<hcb::OverlapGraph> graph(new hcb::OverlapGraph(fastq->get_readcount()));
    for (auto& r : fastq->m_read_vec) r.vertex_id = graph->addVertex(r.read_id);   // src/ViralQuasispecies.cpp:258-263
    hcb::EdgeCalculator ec(fastq, graph, ps);
    if (ov_reader.joinable()) {
        ov_reader.join();
        if (ov_ok) { ec.preloaded_overlaps = std::move(ov_text); ec.have_preloaded = true; }
    }
    t0 = now_s();
    ec.construct_edges();
    double t_ce = now_s() - t0;
    const double t_ce_end = now_s();
    if (!dump_graph.empty()) graph->dumpAdjacency(dump_graph);
    if (!digraph.empty()) graph->writeDiGraphToFile(digraph);
    double t_sort = 0;
    if (!dump_sorted.empty()) {
        std::vector<uint32_t> read_len(fastq->mate_len.size() / 2);
        for (size_t i = 0; i < read_len.size(); i++) read_len[i] = fastq->mate_len[2 * i] + fastq->mate_len[2 * i + 1];   // Read::get_len(), src/Read.h:203-212
        const double ts0 = now_s();
        graph->sortEdges(read_len, ps.gpu_dedup ? ps.first_device : -1);
        t_sort = now_s() - ts0;
        graph->dumpAdjacency(dump_sorted, true);
    }
    printf("{\"t_sort_edges_s\": %.6f, \"t_main_s\": %.6f, \"t_graph_files_s\": %.6f, \"reads_single\": %u, \"reads_paired\": %u, \"scored\": %lu, \"t_fastq_s\": %.6f, \"t_construct_edges_s\": %.6f, "
           "\"t_fastq_read_s\": %.3f, \"t_cuda_init_s\": %.3f, \"t_fastq_store_s\": %.3f, \"t_fastq_index_s\": %.3f, \"device_ms\": %.3f, \"parse_device_ms\": %.3f, \"t_ingest_s\": %.3f, \"t_score_s\": %.3f, \"t_edges_s\": %.3f, \"t_write_s\": %.3f, \"graph_edges\": %u, \"dup_count\": %u, \"inclusion_count\": %u}\n",
           t_sort, t_ce_end - t_start, now_s() - t_ce_end, fastq->m_readcount_single, fastq->m_readcount_paired, ec.scored_candidates, t_fastq, t_ce, fastq->t_read_s, fastq->t_cuda_init_s, fastq->t_store_s, fastq->t_index_s, ec.device_ms, ec.parse_device_ms, ec.t_ingest_s, ec.t_score_s, ec.t_edges_s, ec.t_write_s,
           graph->getEdgeCount(), ec.dup_count, ec.inclusion_count);
    // the process ends here: no destructor walk over millions of adjacency entries, no piecewise release of the device
    fflush(stdout);
    fflush(stderr);
    _exit(0);
}
