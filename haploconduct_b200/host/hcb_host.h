// hcb_host.h -- host-side C++ mirror of the reference interfaces that sit on the EdgeCalculator
// path, written against the C ABI of include/hc_b200.h.  Same class names, same argument meaning,
// same error behaviour (message on stderr + exit(1)) as the reference, so a maintainer can swap the
// reference's EdgeCalculator.cpp for hcb_edge_calculator.cpp.
//
//   hcb::ProgramSettings  <- src/Types.h:19-67            (the fields this path reads)
//   hcb::FastqStorage     <- src/FastqStorage.h:58-98     (+ the device replica of the reads)
//   hcb::Overlap          <- src/Overlap.h:20-238
//   hcb::Edge             <- src/Edge.h:18-121
//   hcb::OverlapGraph     <- src/OverlapGraph.h:84-131    (the members process_overlaps touches)
//   hcb::EdgeCalculator   <- src/EdgeCalculator.h:26-63
//
// Nothing here computes a score on the CPU: overlap_score() and construct_edges() call the
// CUDA library through hc_overlap_score() / hc_score_batch().
#ifndef HCB_HOST_H_
#define HCB_HOST_H_

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hc_b200.h"

namespace hcb {

typedef unsigned long read_id_t;   // src/Types.h:93-94
typedef unsigned long node_id_t;

struct ProgramSettings {           // names and defaults of src/Types.h:19-67 / src/ViralQuasispecies.cpp:52-98
    std::string singles_file, paired1_file, paired2_file, overlaps_file, output_dir, id_correspondence;
    unsigned long max_overlaps = 100000000UL;
    unsigned long max_reads = 100000000UL;
    unsigned int n_threads = 1;
    unsigned int min_overlap_perc = 0;
    unsigned int min_overlap_len = 150;
    double edge_threshold = 0.99;
    double ov_threshold = 0.9;
    bool allow_spaces = false;
    bool add_duplicates = false;
    bool resolve_orientations = true;
    bool ignore_inclusions = false;
    double mismatch = 0;
    double merge_contigs = 0;
    unsigned int min_read_len = 0;
    bool relax_PE_edges = false;
    bool verbose = false;
    // additions of this build
    // HC_FLAG_EXACT_EDGE_SCORES + host libm exp(): Edge::score bit-identical to the reference, which the
    // duplicate-edge resolution of process_overlaps (score >= existing score, :470) needs to pick the
    // same representative when two overlaps of one read pair score within 1e-7 of each other.
    bool exact_scores = true;
    // build the read store from the FASTQ text on the GPU (hc_store_create_fastq); m_read_vec then holds ids and
    // read types only, no sequences (the id_correspondence table is not supported on this path)
    bool gpu_fastq = false;
    // parse + pre-filter the overlaps file on the GPU (hc_ingest_overlaps) instead of the text loop of :581-645
    bool gpu_parse = false;
    // resolve duplicate edges on the GPU (hc_dedup_edges) instead of the sequential insert of :429-545
    bool gpu_dedup = false;
    int first_device = 0;
    int n_devices = 1;
};

struct Read {                      // src/Read.h:22-58 (what the path needs)
    bool is_paired = false;
    read_id_t read_id = 0;
    node_id_t vertex_id = 0;
    std::string seq1, seq2, phred1, phred2;
};

// Reads FASTQ exactly like src/FastqStorage.cpp:42-235 (singles upper-cased :123, pairs verbatim
// :196-197, ids through strtoul(.,0) src/Types.h:99-102) and owns the device store.
class FastqStorage {
public:
    explicit FastqStorage(const ProgramSettings& ps);
    ~FastqStorage();
    FastqStorage(const FastqStorage&) = delete;
    FastqStorage& operator=(const FastqStorage&) = delete;

    // --gpu_fastq: called once the CUDA context exists (the page faults of a file being read by another thread and the
    // context creation slow each other down -- both live on the address-space lock -- so a caller that wants to read
    // its next input in the background starts here, not earlier)
    static std::function<void()> device_ready_hook;

    std::vector<Read> m_read_vec;                     // singles first, then pairs (src/FastqStorage.h:88-97)
    std::map<read_id_t, unsigned int> m_ID_to_index;  // src/FastqStorage.h:54
    unsigned int m_readcount_single = 0, m_readcount_paired = 0;
    unsigned int max_read_len = 0;                    // longest mate (decides the candidate record size)
    std::vector<uint64_t> read_ids;                   // read_id of m_read_vec[i], packed (line formatting reads two per candidate)
    std::vector<uint32_t> mate_len;                   // 2 per read: sequence lengths (/1, /2; 0 for the missing mate of a single)
    double t_read_s = 0, t_cuda_init_s = 0, t_store_s = 0, t_index_s = 0;  // wall clock of the constructor's phases (not in the reference)

    unsigned int get_readcount() const { return (unsigned int)m_read_vec.size(); }
    hc_store* device_store() const { return store_; }
    hc_idmap* device_idmap();     // m_ID_to_index on the device (hc_ingest_overlaps), built on first use

private:
    void read_singles(const std::string& path, unsigned long max_reads);
    void read_pairs(const std::string& p1, const std::string& p2, unsigned long max_reads);
    void read_new_ids(const std::string& path);
    std::map<std::string, std::string> new_ids_;
    bool have_new_ids_ = false;
    hc_store* store_ = nullptr;
    hc_idmap* idmap_ = nullptr;
    int first_device_ = 0;
};

struct Overlap {                   // src/Overlap.h:23-35
    read_id_t id1 = 0, id2 = 0;
    unsigned int pos1 = 0, pos2 = 0;
    char ord = '-';
    char ori1 = '+', ori2 = '+';
    unsigned int perc1 = 0, perc2 = 0, len1 = 0, len2 = 0;
    char type1 = 's', type2 = 's';
    long idx1 = -1, idx2 = -1;     // store indices when the device parser resolved them already

    // constructor semantics of src/Overlap.h:39-73; exits like the reference on malformed fields
    static Overlap from_fields(const std::vector<std::string>& f);
    unsigned int get_perc() const { return perc2 > 0 ? (unsigned int)(0.5 * (perc1 + perc2)) : perc1; }   // :203-210
    std::string get_overlap_line() const;                                                                // :234-237
    void append_overlap_line(std::string& buf) const;    // the same line appended to a buffer
};

struct Edge {                      // src/Edge.h:21-37
    double score = 0;
    int pos1 = 0, pos2 = 0, pos3 = 0, pos4 = 0;
    bool ori1 = true, ori2 = true;
    char ord = '-';
    node_id_t vertex1 = 0, vertex2 = 0;
    int overlap_perc = -1, overlap_len = -1, overlap_len1 = -1, overlap_len2 = -1;
    double mismatch_rate = -1;
    void swap_reads();             // src/Edge.h:74-88
};

class OverlapGraph {               // the part of src/OverlapGraph.{h,cpp} that process_overlaps drives
public:
    explicit OverlapGraph(unsigned int V);
    node_id_t addVertex(read_id_t id);                                                   // src/OverlapGraph.cpp:88-92
    void addEdge(const Edge& e);                                                         // :94-101
    // addEdge for es[k] with keep[k] != 0 (all of es if keep is null), in order: the lists are sized first and filled by
    // all host threads; the edges must have distinct (vertex pair, orientation) keys, as the survivors of the insert have
    void addEdges(const std::vector<Edge>& es, const unsigned char* keep);
    double checkEdgeWithOri(node_id_t v, node_id_t w, bool same_ori) const;              // :198-229
    const Edge* getEdgeInfoWithOri(node_id_t v, node_id_t w, bool same_ori) const;       // :285-307
    void removeEdgeWithOri(node_id_t v, node_id_t w, bool same_ori);                     // :150-195
    unsigned int getEdgeCount() const { return edge_count_; }
    unsigned int getVertexCount() const { return (unsigned int)vertex_to_read.size(); }
    // sort every adjacency list by (non-overlap length, vertex2) and rebuild adj_in, :722-764.  read_len[v] = Read::get_len()
    // of vertex v (src/Read.h:203-212).  device >= 0: the order comes from hc_build_adjacency, lists std::sort may order
    // differently (equal keys in a list of more than 16 edges) are settled by std::sort itself; device < 0: std::sort only.
    void sortEdges(const std::vector<uint32_t>& read_len, int device);
    void writeDiGraphToFile(const std::string& path) const;                              // :388-409
    void dumpAdjacency(const std::string& path, bool with_in = false) const;   // every Edge field, adjacency order (test aid)

    std::vector<read_id_t> vertex_to_read;
    std::vector<char> inclusions;                        // src/OverlapGraph.h:80
    std::vector<std::vector<Edge>> adj_out;              // same order as the reference's std::list<Edge>
    std::vector<std::vector<node_id_t>> adj_in;          // filled by sortEdges (:752-763)

private:
    // (min vertex, max vertex, same-orientation flag) -> owner vertex of the unique edge of that key
    // (built when the first lookup / removal asks for it: a stage that only inserts never pays for it)
    mutable std::unordered_map<uint64_t, node_id_t> owner_;
    mutable bool owner_stale_ = false;
    void ensure_owner() const;
    static uint64_t key(node_id_t a, node_id_t b, bool same_ori);
    unsigned int edge_count_ = 0;
};

// A whole file in memory: anonymous mapping (huge pages where the kernel gives them), filled by several threads with pread.
struct FileBuf {
    char* data = nullptr;
    size_t size = 0, mapped = 0;
    FileBuf() {}
    FileBuf(const FileBuf&) = delete;
    FileBuf& operator=(const FileBuf&) = delete;
    FileBuf(FileBuf&& o) noexcept : data(o.data), size(o.size), mapped(o.mapped) { o.data = nullptr; o.size = o.mapped = 0; }
    FileBuf& operator=(FileBuf&& o) noexcept;
    ~FileBuf();
};
// false if the file cannot be opened; threads <= 0: all host threads
bool read_whole_file(const std::string& path, FileBuf& out, int threads = 0);

class EdgeCalculator {             // src/EdgeCalculator.h:26-63
public:
    EdgeCalculator(std::shared_ptr<FastqStorage> fastq, std::shared_ptr<OverlapGraph> graph, const ProgramSettings ps);
    void construct_edges();                                                                       // :561-666
    double overlap_score(std::string seq1, std::string seq2, std::string score1, std::string score2,
                         const unsigned int pos, double& mismatch_rate);                            // :67-139
    double phred_to_prob(const int phred);                                                         // :59-63

    unsigned int self_overlap_count = 0, inclusion_count = 0, dup_count = 0;
    // measurements of the last construct_edges() (not in the reference)
    unsigned long scored_candidates = 0;
    double device_ms = 0, parse_device_ms = 0;
    double t_ingest_s = 0, t_score_s = 0, t_edges_s = 0, t_write_s = 0;   // host wall clock per phase
    // --gpu_parse: the text of the overlaps file, if the caller has read it already (hc_edgecalc reads it while the read
    // store is being built); construct_edges() reads the file itself otherwise
    FileBuf preloaded_overlaps;
    bool have_preloaded = false;

private:
    void process_overlaps(std::vector<Overlap>& batch);                                          // :389-557
    void insert_edge(Edge& e, unsigned int& doubles);                                            // :441-538
    // one line of the text loop (:583-635): 0 skipped / dropped, 1 appended to batch, 2 appended to filtered
    int handle_line(const std::string& line, std::vector<Overlap>& batch, std::vector<Overlap>& filtered);
    void ingest_on_device(std::vector<Overlap>& batch, std::vector<Overlap>& filtered);
    // --gpu_parse: the whole stage on arrays -- the file in one read, one parse call, run-encoded records, small outputs,
    // edges and non-edge lines built by all host threads; false if the records do not fit (then the path above runs)
    bool construct_edges_arrays();
    std::vector<Edge> pending_;   // gpu_dedup: accepted edges of all batches, normalised, in order
    ProgramSettings ps_;
    std::shared_ptr<FastqStorage> fastq_;
    std::shared_ptr<OverlapGraph> graph_;
};

[[noreturn]] void die(const std::string& msg);   // message on stderr + exit(1), the reference's error behaviour

}  // namespace hcb
#endif
