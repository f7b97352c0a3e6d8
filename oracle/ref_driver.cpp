// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src, see oracle/Makefile) through the EdgeCalculator stage and
// dumps what the parity tests and the CPU baseline need.  It replaces
// src/ViralQuasispecies.cpp (which needs boost::program_options) by filling
// ProgramSettings by hand and replaying src/ViralQuasispecies.cpp:233-281:
// FastqStorage -> OverlapGraph + addVertex per read -> EdgeCalculator.
//
// Modes (combine freely):
//   --dump-cands F   per pre-filtered candidate, in input order, the Edge that
//                    EdgeCalculator::compute_overlap (src/EdgeCalculator.cpp:143)
//                    returns plus the class of src/EdgeCalculator.cpp:404-413.
//   --run            the real EdgeCalculator::construct_edges() (src/EdgeCalculator.cpp:561),
//                    timed as src/ViralQuasispecies.cpp:280-283 times it; writes
//                    nonedge_overlaps.txt into the cwd like the reference does.
//   --dump-graph F   after --run: adjacency lists (adj_out, in order) with all Edge fields.
//   --time-scoring   time only the parallel scoring region (src/EdgeCalculator.cpp:395-423)
//                    on an already parsed batch.
//   --merge-fno1 F   after --run: replay the merge iteration of src/ViralQuasispecies.cpp:297-464
//                    (sortEdges .. cycleRemovalHeuristic, mergeAlongEdges), dump to F everything
//                    SRBuilder::findNextOverlaps (src/FindNextOverlaps.cpp:890-958) reads -- per-vertex
//                    state, super-reads with sub-read indices, and the edge stream in the order the
//                    reference processes it -- then run the reference's findNextOverlaps() itself,
//                    which writes overlaps.txt into the cwd.
//   --fno-state F    with --merge-fno1 / --merge-fno3: also dump to F the STATE SRBuilder::findNextOverlaps /
//                    findNextOverlaps3 start from -- adjacency lists, branching and inclusion edges, vertex labels,
//                    visited, new ids, super-reads with cliques, sub-read indices and original reads -- i.e. the input
//                    of the product's C++ binding (haploconduct_b200/host/hcb_fno.h), which produces the edge stream
//                    itself.
//   --merge-fno3 F   same iteration, but for SRBuilder::findNextOverlaps3 (src/FindNextOverlaps3.cpp:20-173):
//                    dumps the original-read -> super-read lists in the iteration order of the
//                    reference's std::unordered_map, then runs findNextOverlaps3() itself.
//
// compute_overlap / adj_out are private in the reference headers; the dump TU sees them
// through "#define private public" (class layout is unchanged, it links against the
// unmodified objects).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <list>
#include <stack>
#include <unordered_map>
#include <iostream>
#include <fstream>
#include <sstream>
#include <memory>
#include <deque>
#include <functional>
#include <algorithm>
#include <sys/time.h>
#include <omp.h>

#define private public
#include "EdgeCalculator.h"
#include "SRBuilder.h"
#undef private

static double now_s() {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static void usage() {
    std::fprintf(stderr,
        "ref_driver --overlaps F [--singles F] [--paired1 F --paired2 F] [--threads N]\n"
        "  [--edge_threshold X] [--ov_threshold X] [--min_overlap_len N] [--min_overlap_perc N]\n"
        "  [--merge_contigs X] [--mismatch X] [--min_read_len N] [--relax_PE_edges 0|1]\n"
        "  [--ignore_inclusions 0|1] [--max_ov N]\n"
        "  [--dump-cands F] [--run] [--dump-graph F] [--time-scoring] [--reps N]\n");
}

int main(int argc, char** argv) {
    ProgramSettings ps = ProgramSettings();  // zero-init: the struct has no defaults of its own
    // defaults of the option table, src/ViralQuasispecies.cpp:52-98
    ps.max_overlaps = 100000000UL;
    ps.max_reads = 100000000UL;
    ps.n_threads = 1;
    ps.min_clique_size = 4;
    ps.min_qual = 0.9;
    ps.min_overlap_perc = 0;
    ps.min_overlap_len = 150;
    ps.edge_threshold = 0.99;
    ps.ov_threshold = 0.9;
    ps.allow_spaces = false;
    ps.first_it = true;
    ps.add_duplicates = false;
    ps.resolve_orientations = true;
    ps.mismatch = 0;
    ps.optimize = true;
    ps.merge_contigs = 0;
    ps.remove_tips = true;
    ps.max_tip_len = 150;
    ps.store_tips_separately = true;
    ps.base_path = ".";
    ps.careful = true;
    ps.fno = 2;
    ps.output_dir = "";

    std::string dump_cands, dump_graph, dump_sorted, merge_fno1, consensus_in, consensus_out, fno_state, subread_in, subread_out;
    bool fno3 = false, use_cliques = false;
    ps.keep_singletons = 0;
    ps.remove_trans = 1;
    ps.remove_branches = true;
    ps.min_clique_size = 2;
    ps.fno = 1;
    ps.optimize = false;
    ps.original_readcount = 0;
    bool do_run = false, do_time = false;
    int reps = 1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--overlaps") ps.overlaps_file = need("--overlaps");
        else if (a == "--singles") ps.singles_file = need("--singles");
        else if (a == "--paired1") ps.paired1_file = need("--paired1");
        else if (a == "--paired2") ps.paired2_file = need("--paired2");
        else if (a == "--threads") ps.n_threads = std::atoi(need("--threads"));
        else if (a == "--edge_threshold") ps.edge_threshold = std::atof(need("--edge_threshold"));
        else if (a == "--ov_threshold") ps.ov_threshold = std::atof(need("--ov_threshold"));
        else if (a == "--min_overlap_len") ps.min_overlap_len = std::atoi(need("--min_overlap_len"));
        else if (a == "--min_overlap_perc") ps.min_overlap_perc = std::atoi(need("--min_overlap_perc"));
        else if (a == "--merge_contigs") ps.merge_contigs = std::atof(need("--merge_contigs"));
        else if (a == "--mismatch") ps.mismatch = std::atof(need("--mismatch"));
        else if (a == "--min_read_len") ps.min_read_len = std::atoi(need("--min_read_len"));
        else if (a == "--relax_PE_edges") ps.relax_PE_edges = std::atoi(need("--relax_PE_edges")) != 0;
        else if (a == "--ignore_inclusions") ps.ignore_inclusions = std::atoi(need("--ignore_inclusions")) != 0;
        else if (a == "--allow_spaced_overlaps") ps.allow_spaces = std::atoi(need("--allow_spaced_overlaps")) != 0;
        else if (a == "--max_ov") ps.max_overlaps = std::strtoul(need("--max_ov"), NULL, 10);
        else if (a == "--dump-cands") dump_cands = need("--dump-cands");
        else if (a == "--dump-graph") dump_graph = need("--dump-graph");
        else if (a == "--dump-sorted") dump_sorted = need("--dump-sorted");   // the same dump after OverlapGraph::sortEdges(), plus adj_in
        else if (a == "--merge-fno1") merge_fno1 = need("--merge-fno1");
        else if (a == "--merge-fno3") { merge_fno1 = need("--merge-fno3"); fno3 = true; }
        else if (a == "--fno-state") fno_state = need("--fno-state");
        else if (a == "--cliques") use_cliques = std::atoi(need("--cliques")) != 0;
        else if (a == "--min_clique_size") ps.min_clique_size = std::atoi(need("--min_clique_size"));
        else if (a == "--min_qual") ps.min_qual = std::atof(need("--min_qual"));
        else if (a == "--consensus") { consensus_in = need("--consensus"); consensus_out = need("--consensus"); }
        else if (a == "--subread-info") { subread_in = need("--subread-info"); subread_out = need("--subread-info"); }
        else if (a == "--keep_singletons") ps.keep_singletons = std::atoi(need("--keep_singletons"));
        else if (a == "--remove_branches") ps.remove_branches = std::atoi(need("--remove_branches")) != 0;
        else if (a == "--remove_trans") ps.remove_trans = std::atoi(need("--remove_trans"));
        else if (a == "--no_inclusion_overlaps") ps.no_inclusions = std::atoi(need("--no_inclusion_overlaps")) != 0;
        else if (a == "--max_tip_len") ps.max_tip_len = std::atoi(need("--max_tip_len"));
        else if (a == "--run") do_run = true;
        else if (a == "--time-scoring") do_time = true;
        else if (a == "--reps") reps = std::atoi(need("--reps"));
        else { usage(); return 2; }
    }
    if (ps.overlaps_file.empty()) { usage(); return 2; }

    double t0 = now_s();
    std::shared_ptr<FastqStorage> fastq(new FastqStorage(ps));
    double t_fastq = now_s() - t0;
    std::shared_ptr<OverlapGraph> graph(new OverlapGraph(fastq->get_readcount(), fastq, ps));
    for (auto r : fastq->m_read_vec) {
        node_id_t v = graph->addVertex(r->get_read_id());
        r->set_vertex_id(true, v);
    }
    EdgeCalculator ec(fastq, graph, ps);

    // --consensus IN OUT: SRBuilder::consensus (src/SRBuilder.cpp:406-522, with consensus_pos :297-402) on the
    // problems of IN ("P total_len subreads_needed error_correction n", then n lines "pos seq qual"); one line
    // "R ret length seq qual" per problem in OUT ("-" stands for an empty string, length 0).
    if (!consensus_in.empty()) {
        std::shared_ptr<SRBuilder> srb(new SRBuilder(fastq, graph, ps));
        std::ifstream in(consensus_in.c_str());
        FILE* fo = std::fopen(consensus_out.c_str(), "w");
        if (!in.is_open() || !fo) { std::fprintf(stderr, "cannot open the consensus files\n"); return 1; }
        std::string tag;
        double t_cons = 0;
        unsigned long n_prob = 0, n_bases = 0;
        while (in >> tag) {
            int total_len, sub, ec_flag, n;
            in >> total_len >> sub >> ec_flag >> n;
            std::list<int> pos_list;
            std::list<std::string> seq_list, qual_list;
            for (int k = 0; k < n; k++) {
                int pos;
                std::string sq, ql;
                in >> pos >> sq >> ql;
                pos_list.push_back(pos); seq_list.push_back(sq); qual_list.push_back(ql);
                n_bases += sq.size();
            }
            std::string cs, cq;
            const double tc0 = now_s();
            const int ret = srb->consensus(total_len, pos_list, seq_list, qual_list, cs, cq, sub != 0, ec_flag != 0);
            t_cons += now_s() - tc0;
            n_prob++;
            std::fprintf(fo, "R\t%d\t%zu\t%s\t%s\n", ret, cs.size(), cs.empty() ? "-" : cs.c_str(), cq.empty() ? "-" : cq.c_str());
        }
        std::fclose(fo);
        std::printf("{\"consensus_problems\": %lu, \"consensus_bases\": %lu, \"t_consensus_s\": %.6f}\n", n_prob, n_bases, t_cons);
        return 0;
    }

    // --subread-info IN OUT: SRBuilder::calcSubreadInfo (src/SRBuilder.cpp:536-595) on the problems of IN
    // ("S trim_pos1 trim_pos2 n1 n2", then n1 "pos vertex" pairs of list 1 and n2 of list 2); per problem in OUT one line
    // "R m" and m lines "vertex index1 index2 startpos1 startpos2", vertices ascending.
    if (!subread_in.empty()) {
        std::shared_ptr<SRBuilder> srb(new SRBuilder(fastq, graph, ps));
        std::ifstream in(subread_in.c_str());
        FILE* fo = std::fopen(subread_out.c_str(), "w");
        if (!in.is_open() || !fo) { std::fprintf(stderr, "cannot open the subread-info files\n"); return 1; }
        std::string tag;
        while (in >> tag) {
            int t1, t2, n1, n2;
            in >> t1 >> t2 >> n1 >> n2;
            std::list<int> p1, p2;
            std::list<node_id_t> v1, v2;
            for (int k = 0; k < n1; k++) { int p; unsigned long v; in >> p >> v; p1.push_back(p); v1.push_back(v); }
            for (int k = 0; k < n2; k++) { int p; unsigned long v; in >> p >> v; p2.push_back(p); v2.push_back(v); }
            std::unordered_map<node_id_t, SubreadInfo> m = srb->calcSubreadInfo(t1, t2, p1, p2, v1, v2);
            std::vector<node_id_t> keys;
            for (std::unordered_map<node_id_t, SubreadInfo>::iterator it = m.begin(); it != m.end(); ++it) keys.push_back(it->first);
            std::sort(keys.begin(), keys.end());
            std::fprintf(fo, "R\t%zu\n", keys.size());
            for (size_t k = 0; k < keys.size(); k++) {
                const SubreadInfo& si = m[keys[k]];
                std::fprintf(fo, "%lu\t%d\t%d\t%d\t%d\n", keys[k], si.index1, si.index2, si.startpos1, si.startpos2);
            }
        }
        std::fclose(fo);
        return 0;
    }

    // Parse + pre-filter exactly as src/EdgeCalculator.cpp:581-635 (tab-split branch), keeping
    // the 1-based line number of every candidate that reaches process_overlaps.
    std::vector<Overlap> batch;
    std::vector<unsigned long> batch_line;
    unsigned long n_lines = 0, n_self = 0, n_lenfiltered = 0, n_percdropped = 0, n_bad = 0;
    double t_parse = -1;
    if (!dump_cands.empty() || do_time) {
        const double tp0 = now_s();
        std::ifstream in(ps.overlaps_file.c_str());
        if (!in.is_open()) { std::fprintf(stderr, "cannot open %s\n", ps.overlaps_file.c_str()); return 1; }
        std::string line;
        unsigned long i = 0;
        while (getline(in, line) && i < ps.max_overlaps) {
            i++;
            n_lines++;
            boost::trim_if(line, boost::is_any_of("\t "));
            std::vector<std::string> f;
            if (ps.allow_spaces) {
                boost::algorithm::split(f, line, boost::is_any_of("\t "), boost::token_compress_on);
            } else {
                std::stringstream ss(line);
                std::string tmp;
                while (getline(ss, tmp, '\t')) f.push_back(tmp);
            }
            if (f.size() != 13) { n_bad++; continue; }
            Overlap ov(f);
            if (ov.get_id(1) == ov.get_id(2)) { n_self++; continue; }
            bool pass = false, dropped = false;
            if (ov.get_len(1) >= ps.min_overlap_len && ov.get_type(1) == "s" && ov.get_type(2) == "s") {
                if (ov.get_perc() >= ps.min_overlap_perc) pass = true; else dropped = true;
            } else if (ov.get_len(1) >= 0.5 * ps.min_overlap_len && ov.get_len(2) >= 0.5 * ps.min_overlap_len
                       && (ov.get_type(1) == "p" || ov.get_type(2) == "p")) {
                if (ov.get_perc() >= ps.min_overlap_perc) pass = true; else dropped = true;
            } else if (ps.relax_PE_edges && ov.get_len(1) + ov.get_len(2) >= ps.min_overlap_len
                       && (ov.get_type(1) == "p" || ov.get_type(2) == "p")) {
                if (ov.get_perc() >= ps.min_overlap_perc) pass = true; else dropped = true;
            } else {
                n_lenfiltered++;
            }
            if (dropped) n_percdropped++;
            if (pass) { batch.push_back(ov); batch_line.push_back(i); }
        }
        t_parse = now_s() - tp0;   // the single-threaded text loop: getline, trim, split, Overlap ctor, pre-filter
    }

    if (!dump_cands.empty()) {
        FILE* fo = std::fopen(dump_cands.c_str(), "w");
        if (!fo) { std::fprintf(stderr, "cannot write %s\n", dump_cands.c_str()); return 1; }
        std::fprintf(fo, "#line\tclass\tscore\tmm_rate\tpos1\tpos2\tpos3\tpos4\tv1\tv2\tori1\tori2\tord\tperc\tlen1\tlen2\n");
        for (size_t k = 0; k < batch.size(); k++) {
            Edge e = ec.compute_overlap(batch[k]);
            char cls = 'D';
            if (e.get_score() > ps.edge_threshold) cls = 'E';
            else if (e.get_mismatch_rate() != -1 && e.get_mismatch_rate() <= ps.merge_contigs) cls = 'E';
            else if (e.get_score() > ps.ov_threshold && e.get_mismatch_rate() != -1) cls = 'N';
            std::fprintf(fo, "%lu\t%c\t%a\t%a\t%d\t%d\t%d\t%d\t%lu\t%lu\t%d\t%d\t%c\t%d\t%d\t%d\n",
                         batch_line[k], cls, e.get_score(), e.get_mismatch_rate(), e.get_pos(1), e.get_pos(2),
                         e.get_extra_pos(1), e.get_extra_pos(2), e.get_vertex(1), e.get_vertex(2),
                         (int)e.get_ori(1), (int)e.get_ori(2), e.get_ord(), e.get_perc(), e.get_len(1), e.get_len(2));
        }
        std::fclose(fo);
    }

    double t_scoring = -1;
    unsigned long n_edges_t = 0, n_nonedges_t = 0;
    if (do_time) {
        // same loop shape as src/EdgeCalculator.cpp:395-423: static omp-for, per-thread vectors
        // concatenated under critical sections.
        double best = 1e300;
        std::string rep_times;
        for (int r = 0; r < reps; r++) {
            std::vector<Edge> edges;
            std::vector<Overlap> nonedges;
            double ts = now_s();
            unsigned int size = batch.size();
            #pragma omp parallel num_threads(ps.n_threads) shared(batch, edges, nonedges)
            {
                std::vector<Edge> e_t;
                std::vector<Overlap> o_t;
                #pragma omp for
                for (unsigned int i = 0; i < size; i++) {
                    Overlap overlap = batch.at(i);
                    Edge edge = ec.compute_overlap(overlap);
                    if (edge.get_score() > ps.edge_threshold) e_t.push_back(edge);
                    else if (edge.get_mismatch_rate() != -1 && edge.get_mismatch_rate() <= ps.merge_contigs) e_t.push_back(edge);
                    else if (edge.get_score() > ps.ov_threshold && edge.get_mismatch_rate() != -1) o_t.push_back(overlap);
                }
                #pragma omp critical(we)
                { edges.insert(edges.end(), e_t.begin(), e_t.end()); }
                #pragma omp critical(wo)
                { nonedges.insert(nonedges.end(), o_t.begin(), o_t.end()); }
            }
            double dt = now_s() - ts;
            if (dt < best) best = dt;
            char buf[64];
            std::snprintf(buf, sizeof(buf), "%s%.6f", r ? ", " : "", dt);
            rep_times += buf;
            n_edges_t = edges.size();
            n_nonedges_t = nonedges.size();
        }
        t_scoring = best;
        std::printf("{\"rep_times_s\": [%s]}\n", rep_times.c_str());
    }

    double t_construct = -1;
    if (do_run) {
        double ts = now_s();
        ec.construct_edges();
        t_construct = now_s() - ts;
        auto dump = [&](const std::string& path, bool with_in) -> bool {
            FILE* fo = std::fopen(path.c_str(), "w");
            if (!fo) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); return false; }
            std::fprintf(fo, "#v1\tv2\tscore\tmm_rate\tpos1\tpos2\tpos3\tpos4\tori1\tori2\tord\tperc\tlen1\tlen2\n");
            for (size_t v = 0; v < graph->adj_out.size(); v++) {
                for (std::list<Edge>::iterator it = graph->adj_out[v].begin(); it != graph->adj_out[v].end(); ++it) {
                    std::fprintf(fo, "%lu\t%lu\t%a\t%a\t%d\t%d\t%d\t%d\t%d\t%d\t%c\t%d\t%d\t%d\n",
                                 it->get_vertex(1), it->get_vertex(2), it->get_score(), it->get_mismatch_rate(),
                                 it->get_pos(1), it->get_pos(2), it->get_extra_pos(1), it->get_extra_pos(2),
                                 (int)it->get_ori(1), (int)it->get_ori(2), it->get_ord(), it->get_perc(),
                                 it->get_len(1), it->get_len(2));
                }
            }
            // OverlapGraph::inclusions (src/OverlapGraph.h:80) as '#I' lines, set only under ignore_inclusions
            for (size_t v = 0; v < graph->inclusions.size(); v++)
                if (graph->inclusions[v]) std::fprintf(fo, "#I\t%zu\n", v);
            if (with_in) {   // adj_in as sortEdges rebuilds it (src/OverlapGraph.cpp:752-763): '#IN v src src ...'
                for (size_t v = 0; v < graph->adj_in.size(); v++) {
                    if (graph->adj_in[v].empty()) continue;
                    std::fprintf(fo, "#IN\t%zu", v);
                    for (std::list<node_id_t>::iterator it = graph->adj_in[v].begin(); it != graph->adj_in[v].end(); ++it) std::fprintf(fo, "\t%lu", *it);
                    std::fprintf(fo, "\n");
                }
            }
            std::fclose(fo);
            return true;
        };
        if (!dump_graph.empty() && !dump(dump_graph, false)) return 1;
        if (!dump_sorted.empty()) {
            graph->sortEdges();                                   // src/OverlapGraph.cpp:722-764
            if (!dump(dump_sorted, true)) return 1;
        }
    }

    unsigned long fno_lines = 0;
    if (do_run && !merge_fno1.empty() && graph->getEdgeCount() > 0) {
        ps.original_readcount = fastq->get_readcount();
        // src/ViralQuasispecies.cpp:297-367 (merge iteration: error_correction = false, cliques = false)
        graph->sortEdges();
        unsigned int conflict_count;
        graph->vertexLabellingHeuristic(conflict_count);
        graph->checkDuplicateEdges();
        if (ps.ignore_inclusions) graph->removeInclusions();
        graph->removeTransitiveEdges();
        graph->buildOriginalsDict();
        if (ps.remove_tips) graph->removeTips();
        if (ps.remove_branches) graph->removeBranches();
        graph->sortEdges();
        graph->cycleRemovalHeuristic(true);
        if (fno3) ps.fno = 3;
        std::shared_ptr<SRBuilder> srb(new SRBuilder(fastq, graph, ps));
        graph->sortEdges();
        if (use_cliques) {
            // maximal cliques of the undirected overlap graph (the reference shells out to quick-cliques,
            // src/ViralQuasispecies.cpp:400; here a plain Bron-Kerbosch with pivoting writes cliques.txt)
            const size_t Vn = graph->getVertexCount();
            std::vector<std::set<node_id_t>> nb(Vn);
            for (size_t v = 0; v < Vn; v++) for (auto& e : graph->adj_out[v]) { nb[v].insert(e.get_vertex(2)); nb[e.get_vertex(2)].insert(v); }
            std::ofstream cf("cliques.txt");
            std::function<void(std::set<node_id_t>, std::set<node_id_t>, std::set<node_id_t>)> bk =
                [&](std::set<node_id_t> R, std::set<node_id_t> Pn, std::set<node_id_t> X) {
                    if (Pn.empty() && X.empty()) {
                        if (R.size() >= 2) { for (auto v : R) cf << v << " "; cf << "\n"; }
                        return;
                    }
                    node_id_t pivot = Pn.empty() ? *X.begin() : *Pn.begin();
                    std::vector<node_id_t> cand;
                    for (auto v : Pn) if (!nb[pivot].count(v)) cand.push_back(v);
                    for (auto v : cand) {
                        std::set<node_id_t> R2 = R, P2, X2;
                        R2.insert(v);
                        for (auto w : Pn) if (nb[v].count(w)) P2.insert(w);
                        for (auto w : X) if (nb[v].count(w)) X2.insert(w);
                        bk(R2, P2, X2);
                        Pn.erase(v);
                        X.insert(v);
                    }
                };
            std::set<node_id_t> all;
            for (size_t v = 0; v < Vn; v++) if (!nb[v].empty()) all.insert(v);
            bk(std::set<node_id_t>(), all, std::set<node_id_t>());
            cf.close();
            srb->cliquesToSuperreads();                                   // :422
        } else {
            srb->mergeAlongEdges();                                       // :441
        }
        if (!fno_state.empty()) {
            // the state both FindNextOverlaps variants start from (input of hcb::SRBuilder, hcb_fno.h)
            FILE* fs = std::fopen(fno_state.c_str(), "w");
            if (!fs) { std::fprintf(stderr, "cannot write %s\n", fno_state.c_str()); return 1; }
            const size_t Vn = graph->getVertexCount();
            std::fprintf(fs, "P\t%d\t%d\t%d\t%a\t%lu\n", (int)ps.resolve_orientations, (int)ps.no_inclusions, (int)ps.optimize,
                         ps.edge_threshold, (unsigned long)Vn);
            for (size_t v = 0; v < Vn; v++) {
                long nid = -1;
                if (srb->nodes_to_new_IDs.count(v)) nid = (long)srb->nodes_to_new_IDs.at(v);
                const int label = v < graph->vertex_orientations.size() ? (int)graph->getOrientation(v) : 1;
                std::fprintf(fs, "V\t%lu\t%d\t%ld\t%d\n", (unsigned long)v, (int)srb->visited[v], nid, label);
            }
            auto put_edge = [&](const char* tag, long k, const Edge& e) {
                std::fprintf(fs, "%s\t%ld\t%lu\t%lu\t%d\t%d\t%c\t%d\t%d\t%a\t%d\t%d\t%d\n", tag, k, e.get_vertex(1), e.get_vertex(2),
                             e.get_pos(1), e.get_pos(2), e.get_ord(), (int)e.get_ori(1), (int)e.get_ori(2), e.get_score(), e.get_perc(),
                             e.get_len(1), e.get_len(2));
            };
            for (auto& lst : graph->adj_out) for (auto& e : lst) put_edge("A", 0, e);
            for (auto& e : graph->branching_edges) put_edge("B", 0, e);
            for (size_t k = 0; k < graph->inclusion_edges.size(); k++) for (auto& e : graph->inclusion_edges[k]) put_edge("I", (long)k, e);
            auto put_sr = [&](char kind, const Read& r) {
                const unsigned long l1 = r.is_paired() ? r.get_seq(1).size() : r.get_seq(0).size();
                const unsigned long l2 = r.is_paired() ? r.get_seq(2).size() : 0;
                std::fprintf(fs, "S\t%c\t%lu\t%lu\t%lu", kind, r.get_read_id(), l1, l2);
                if (kind != 't') {
                    for (auto node : r.get_sorted_clique(kind == 'p' ? 1 : 0)) {
                        const SubreadInfo si = r.get_subread_info(node);
                        std::fprintf(fs, "\t%lu:%d:%d:%d:%d", (unsigned long)node, si.index1, si.index2, si.startpos1, si.startpos2);
                    }
                }
                std::fprintf(fs, "\nO");
                const std::unordered_map<read_id_t, OriginalIndex> originals = r.get_original_reads();   // the copy the reference iterates
                for (auto it : originals) std::fprintf(fs, "\t%lu:%ld:%ld", it.first, it.second.index1, it.second.index2);
                std::fprintf(fs, "\n");
            };
            for (auto& r : srb->single_SR_vec) put_sr('s', r);
            for (auto& r : srb->paired_SR_vec) put_sr('p', r);
            for (auto& r : srb->trivial_SR_vec) put_sr('t', r);
            std::fclose(fs);
        }
        if (fno3) {
            // what findNextOverlaps3 builds (src/FindNextOverlaps3.cpp:26-76), with the same container
            // types and the same insertion sequence, hence the same iteration order
            FILE* fo = std::fopen(merge_fno1.c_str(), "w");
            if (!fo) { std::fprintf(stderr, "cannot write %s\n", merge_fno1.c_str()); return 1; }
            std::fprintf(fo, "P\t%d\n", (int)ps.no_inclusions);
            std::vector<Read*> srs;
            std::map<Read*, size_t> sr_index;
            for (auto& r : srb->single_SR_vec) { sr_index[&r] = srs.size(); srs.push_back(&r); }
            for (auto& r : srb->paired_SR_vec) { sr_index[&r] = srs.size(); srs.push_back(&r); }
            for (auto& r : srb->trivial_SR_vec) { sr_index[&r] = srs.size(); srs.push_back(&r); }
            for (size_t k = 0; k < srs.size(); k++) {
                Read* r = srs[k];
                unsigned long l1 = r->is_paired() ? r->get_seq(1).size() : r->get_seq(0).size();
                unsigned long l2 = r->is_paired() ? r->get_seq(2).size() : 0;
                std::fprintf(fo, "S\t%lu\t%lu\t%lu\t%lu\n", (unsigned long)k, r->get_read_id(), l1, l2);
            }
            std::unordered_map<read_id_t, node_id_t> original_to_index;
            std::vector<std::vector<Read*>> lists;
            for (Read* rp : srs) {
                std::unordered_map<read_id_t, OriginalIndex> originals = rp->get_original_reads();
                for (auto it : originals) {
                    auto ex = original_to_index.find(it.first);
                    if (ex == original_to_index.end()) {
                        original_to_index.insert(std::make_pair(it.first, (node_id_t)lists.size()));
                        lists.push_back(std::vector<Read*>(1, rp));
                    } else lists[ex->second].push_back(rp);
                }
            }
            std::unordered_map<read_id_t, node_id_t> by_value(original_to_index);   // nodeDictApproach takes it by value (:90)
            for (auto it : by_value) {
                std::fprintf(fo, "O\t%lu", it.first);
                for (Read* rp : lists[it.second]) {
                    OriginalIndex oi = rp->get_original_reads().at(it.first);
                    std::fprintf(fo, "\t%lu:%ld:%ld", (unsigned long)sr_index[rp], oi.index1, oi.index2);
                }
                std::fprintf(fo, "\n");
            }
            std::fclose(fo);
            srb->findNextOverlaps3();                                     // the reference itself -> overlaps.txt
            std::printf("{\"fno3_lines\": %lu}\n", (unsigned long)srb->next_overlaps_count);
            merge_fno1.clear();
        }
        if (!merge_fno1.empty()) {
        // ---- dump what findNextOverlaps reads
        FILE* fo = std::fopen(merge_fno1.c_str(), "w");
        if (!fo) { std::fprintf(stderr, "cannot write %s\n", merge_fno1.c_str()); return 1; }
        const size_t V = graph->getVertexCount();
        std::fprintf(fo, "P\t%d\t%d\t%a\n", (int)ps.resolve_orientations, (int)ps.no_inclusions, ps.edge_threshold);
        std::vector<Read*> srs;
        std::map<Read*, size_t> sr_index;
        for (auto& r : srb->single_SR_vec) { sr_index[&r] = srs.size(); srs.push_back(&r); }
        for (auto& r : srb->paired_SR_vec) { sr_index[&r] = srs.size(); srs.push_back(&r); }
        for (size_t k = 0; k < srs.size(); k++) {
            Read* r = srs[k];
            unsigned long l1 = r->is_paired() ? r->get_seq(1).size() : r->get_seq(0).size();
            unsigned long l2 = r->is_paired() ? r->get_seq(2).size() : 0;
            std::fprintf(fo, "S\t%lu\t%lu\t%lu\t%lu\n", (unsigned long)k, r->get_read_id(), l1, l2);
        }
        // nodes_to_SR exactly as findNextOverlaps builds it (:898-913)
        std::vector<std::vector<size_t>> n2sr(V);
        for (auto& r : srb->single_SR_vec) for (auto node : r.get_sorted_clique(0)) n2sr.at(node).push_back(sr_index[&r]);
        for (auto& r : srb->paired_SR_vec) for (auto node : r.get_sorted_clique(1)) n2sr.at(node).push_back(sr_index[&r]);
        for (size_t v = 0; v < V; v++) {
            Read* rd = fastq->m_read_vec.at(v);
            long nid = -1;
            if (srb->nodes_to_new_IDs.count(v)) nid = (long)srb->nodes_to_new_IDs.at(v);
            int label = v < graph->vertex_orientations.size() ? (int)graph->getOrientation(v) : 1;
            unsigned long l1 = rd->is_paired() ? rd->get_seq(1).size() : rd->get_seq(0).size();
            unsigned long l2 = rd->is_paired() ? rd->get_seq(2).size() : 0;
            std::fprintf(fo, "V\t%lu\t%d\t%ld\t%d\t%lu\t%lu", (unsigned long)v, (int)srb->visited[v], nid, label, l1, l2);
            for (size_t k : n2sr[v]) {
                SubreadInfo si = srs[k]->get_subread_info(v);
                std::fprintf(fo, "\t%lu:%d:%d:%d:%d", (unsigned long)k, si.index1, si.index2, si.startpos1, si.startpos2);
            }
            std::fprintf(fo, "\n");
        }
        auto dump_edge = [&](const Edge& e, char src) {
            std::fprintf(fo, "E\t%c\t%lu\t%lu\t%d\t%d\t%c\t%d\t%d\t%d\t%d\t%d\t%d\n", src, e.get_vertex(1), e.get_vertex(2),
                         e.get_pos(1), e.get_pos(2), e.get_ord(), (int)e.get_ori(1), (int)e.get_ori(2), (int)(e.get_score() == 0),
                         e.get_perc(), e.get_len(1), e.get_len(2));
        };
        // edge stream in processing order: adjacency lists, removed branching/tip edges (:605-631) ...
        for (auto& lst : graph->adj_out) for (auto& e : lst) dump_edge(e, 'a');
        for (auto& e : graph->branching_edges) dump_edge(e, 'b');
        // ... non-edge overlaps not already an edge (:635-697, optimize=false) ...
        if (!ps.optimize) {
            std::ifstream nf((ps.output_dir + "nonedge_overlaps.txt").c_str());
            std::string line;
            while (getline(nf, line)) {
                boost::trim_if(line, boost::is_any_of("\t "));
                std::vector<std::string> f;
                std::stringstream ss(line);
                std::string tmp;
                while (getline(ss, tmp, '\t')) f.push_back(tmp);
                Overlap ov(f);
                Read* r1 = fastq->m_read_vec.at(fastq->m_ID_to_index.at(ov.get_id(1)));
                Read* r2 = fastq->m_read_vec.at(fastq->m_ID_to_index.at(ov.get_id(2)));
                Edge e(0, ov.get_pos(1), ov.get_pos(2), ov.get_ori(1) == "+", ov.get_ori(2) == "+", ov.get_ord(), r1, r2);
                e.set_perc(ov.get_perc());
                e.set_len(ov.get_len(1), ov.get_len(2));
                e.set_vertices(r1->get_vertex_id(true), r2->get_vertex_id(true));
                if (graph->checkEdge(e.get_vertex(1), e.get_vertex(2), true) > 0) continue;
                dump_edge(e, 'n');
            }
        }
        // ... and edges induced through removed inclusion vertices (:816-887)
        for (auto edge_list : graph->inclusion_edges) {
            unsigned int l = edge_list.size();
            for (unsigned int i = 0; i < l; i++) for (unsigned int j = i + 1; j < l; j++) {
                Edge e1 = edge_list.at(i), e2 = edge_list.at(j);
                node_id_t n1, n2; Read *r1, *r2; int pos1; bool o1, o2;
                if (e1.get_vertex(1) == e2.get_vertex(1)) continue;
                else if (e1.get_vertex(1) == e2.get_vertex(2)) { n1 = e2.get_vertex(1); n2 = e1.get_vertex(2); r1 = e2.get_read(1); r2 = e1.get_read(2); pos1 = e2.get_pos(1); o1 = e2.get_ori(1); o2 = e1.get_ori(2); }
                else if (e1.get_vertex(2) == e2.get_vertex(1)) { n1 = e1.get_vertex(1); n2 = e2.get_vertex(2); r1 = e1.get_read(1); r2 = e2.get_read(2); pos1 = e1.get_pos(1); o1 = e1.get_ori(1); o2 = e2.get_ori(2); }
                else continue;
                if (r1->is_paired() || r2->is_paired()) continue;
                int len = std::min(r1->get_len() - pos1, r2->get_len());
                int perc = (int)floor(100 * len / std::min(r1->get_len(), r2->get_len()));
                Edge ne(ps.edge_threshold, pos1, 0, o1, o2, "-", r1, r2);
                ne.set_vertices(n1, n2); ne.set_perc(perc); ne.set_len(len, 0);
                if (graph->checkEdge(n1, n2, true) == -1) dump_edge(ne, 'i');
            }
        }
        std::fclose(fo);
        fno_lines = srb->findNextOverlaps();                             // the reference itself -> overlaps.txt
        }
    }
    std::printf("{\"fno1_lines\": %lu}\n", fno_lines);
    std::printf("{\"reads_single\": %u, \"reads_paired\": %u, \"threads\": %u, \"lines\": %lu, \"scored\": %lu, "
                "\"self\": %lu, \"len_filtered\": %lu, \"perc_dropped\": %lu, \"bad\": %lu, "
                "\"t_fastq_s\": %.6f, \"t_parse_s\": %.6f, \"t_scoring_s\": %.6f, \"scoring_edges\": %lu, \"scoring_nonedges\": %lu, "
                "\"t_construct_edges_s\": %.6f, \"graph_edges\": %u, \"dup_count\": %u, \"inclusion_count\": %u}\n",
                fastq->m_readcount_single, fastq->m_readcount_paired, ps.n_threads, n_lines,
                (unsigned long)batch.size(), n_self, n_lenfiltered, n_percdropped, n_bad, t_fastq, t_parse, t_scoring,
                n_edges_t, n_nonedges_t, t_construct, do_run ? graph->getEdgeCount() : 0u, ec.dup_count,
                ec.inclusion_count);
    return 0;
}
