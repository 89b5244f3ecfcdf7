// bioshell_seq.hpp -- C++17 host mirror of the part of bioshell-seq that sits around the
// aligner, written over the C ABI (include/bioshell_align.h).  The reference is compiled Rust;
// its toolchain is not in the build image, so this header is the compiled-language host side:
// same type and function names, argument meaning and error behaviour as the reference (Rust
// panics/Results become C++ exceptions), so call sites and tests read alike.
//
//   Sequence                          bioshell-seq/src/sequence/sequence.rs:8-19
//   count_identical / len_ungapped    sequence.rs:481-490,532-534 ; src/msa/msa.rs:261-269
//   SubstitutionMatrix(List)          src/scoring/substitution_matrix.rs:15-151
//   AlignmentPath / aligned_*         src/alignment/alignment_path.rs:7-204
//   AlignmentStatistics               src/alignment/alignment_statistics.rs:28-81
//   AlignmentReporter + reporters     src/alignment/alignment_reporter.rs:7-64
//   align_all_pairs                   src/alignment/alignment_protocols.rs:83-115
//   SequenceIdentityMatrix            bin/cluster_sequences.rs:77-130
//   align_all_vs_all / align_one_vs_many / PairResults : the new batched entry points
//
// All alignment arithmetic runs on the GPU inside libbioshell_align.so; nothing here aligns.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/bioshell_align.h"

namespace bioshell_seq {

struct BsaError : std::runtime_error {
    int rc;
    BsaError(int rc_, const std::string& msg) : std::runtime_error("libbioshell_align(" + std::to_string(rc_) + "): " + msg), rc(rc_) {}
};

// ---------------------------------------------------------------- sequence
class Sequence {
  public:
    Sequence() = default;
    Sequence(std::string description, std::string seq) : description_(std::move(description)), seq_(std::move(seq)) {}
    static Sequence from_str(const std::string& d, const std::string& s) { return Sequence(d, s); }
    static Sequence from_attrs(std::string d, std::string s) { return Sequence(std::move(d), std::move(s)); }
    const std::string& description() const { return description_; }
    const std::string& as_u8() const { return seq_; }
    size_t len() const { return seq_.size(); }
    std::string to_string(size_t = 0) const { return seq_; }
    // #[derive(PartialEq)] sequence.rs:8 -- description AND residues
    bool operator==(const Sequence& o) const { return description_ == o.description_ && seq_ == o.seq_; }

  private:
    std::string description_, seq_;
};

inline bool is_gap(unsigned char c) { return c == '-' || c == '_'; }

// sequence.rs:481-490 -> msa.rs:261-269
inline size_t count_identical(const Sequence& a, const Sequence& b) {
    if (a.len() != b.len())
        throw std::invalid_argument("AlignedSequencesOfDifferentLengths: expected " + std::to_string(a.len()) +
                                    " found " + std::to_string(b.len()));
    size_t n = 0;
    for (size_t i = 0; i < a.len(); ++i)
        if (a.as_u8()[i] == b.as_u8()[i] && !is_gap((unsigned char)a.as_u8()[i])) ++n;
    return n;
}
// sequence.rs:532-534
inline size_t len_ungapped(const Sequence& s) {
    size_t n = 0;
    for (unsigned char c : s.as_u8()) n += !is_gap(c);
    return n;
}

// ---------------------------------------------------------------- scoring
enum class SubstitutionMatrixList { BLOSUM45, BLOSUM80, PAM250, PAM70, BLOSUM62, PAM120, PAM30 };

class SubstitutionMatrix {
  public:
    int32_t score[441];
    uint8_t aa_indexes[256];
    // substitution_matrix.rs:96-135 (parsed by the library, same rules)
    static SubstitutionMatrix ncbi_matrix_from_buffer(const std::string& text) {
        SubstitutionMatrix m;
        int rc = bsa_parse_ncbi_matrix(text.data(), text.size(), m.score, m.aa_indexes);
        if (rc) throw BsaError(rc, "IncorrectNCBIFormat / CantParseNCBIEntry");
        return m;
    }
    // substitution_matrix.rs:142-150
    static SubstitutionMatrix ncbi_matrix_from_file(const std::string& file_name) {
        FILE* f = std::fopen(file_name.c_str(), "rb");
        if (!f) throw std::runtime_error("FileNotFound: " + file_name);
        std::string text;
        char buf[4096];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
        std::fclose(f);
        return ncbi_matrix_from_buffer(text);
    }
    uint8_t aa_index(uint8_t letter) const {
        if (letter == 255) throw std::out_of_range("index out of bounds: the len is 255 but the index is 255");
        return aa_indexes[letter];
    }
    int32_t score_by_index(uint8_t i, uint8_t j) const { return score[(size_t)i * 21 + j]; }
    int32_t score_by_aa(uint8_t a, uint8_t b) const { return score_by_index(aa_index(a), aa_index(b)); }
};

// ---------------------------------------------------------------- alignment path
// alignment_path.rs:117-139 ; '-' Horizontal (gap in query), '|' Vertical (gap in template), '*' Match
inline std::pair<std::string, std::string> aligned_strings(const std::string& path, const std::string& query,
                                                           const std::string& tmplt, char gap = '-') {
    std::string aq, at;
    size_t qi = 0, ti = 0;
    for (char c : path) {
        if (c == '-') { if (ti >= tmplt.size()) throw std::out_of_range("called `Option::unwrap()` on a `None` value"); aq += gap; at += tmplt[ti++]; }
        else if (c == '|') { if (qi >= query.size()) throw std::out_of_range("called `Option::unwrap()` on a `None` value"); aq += query[qi++]; at += gap; }
        else if (c == '*') { if (qi >= query.size() || ti >= tmplt.size()) throw std::out_of_range("called `Option::unwrap()` on a `None` value"); aq += query[qi++]; at += tmplt[ti++]; }
        else throw std::invalid_argument("Invalid value for AlignmentStep");
    }
    return {aq, at};
}
// alignment_path.rs:199-204
inline std::pair<Sequence, Sequence> aligned_sequences(const std::string& path, const Sequence& query,
                                                       const Sequence& tmplt, char gap = '-') {
    auto p = aligned_strings(path, query.as_u8(), tmplt.as_u8(), gap);
    return {Sequence(query.description(), p.first), Sequence(tmplt.description(), p.second)};
}

// ---------------------------------------------------------------- statistics / reporters
struct AlignmentStatistics {
    std::string query_header, template_header;
    size_t n_identical = 0, query_length = 0, template_length = 0;
    static AlignmentStatistics from_sequences(const Sequence& aq, const Sequence& at) {
        AlignmentStatistics s;
        s.query_header = aq.description();
        s.template_header = at.description();
        s.n_identical = count_identical(aq, at);
        s.query_length = len_ungapped(aq);
        s.template_length = len_ungapped(at);
        return s;
    }
    // alignment_statistics.rs:71-73
    double percent_identity() const {
        return (double)n_identical / (double)std::min(query_length, template_length) * 100.0;
    }
    std::string to_string() const {   // alignment_statistics.rs:76-80
        char buf[512];
        std::snprintf(buf, sizeof buf, "%s %s %6.2f %% %3zu %4zu %4zu", query_header.c_str(), template_header.c_str(),
                      percent_identity(), n_identical, query_length, template_length);
        return buf;
    }
};

struct AlignmentReporter {   // alignment_reporter.rs:7-9
    virtual ~AlignmentReporter() = default;
    virtual void report(const Sequence& aligned_query, const Sequence& aligned_template) = 0;
};

struct MultiReporter : AlignmentReporter {   // alignment_reporter.rs:12-29
    std::vector<std::unique_ptr<AlignmentReporter>> reporters;
    void add_reporter(std::unique_ptr<AlignmentReporter> r) { reporters.push_back(std::move(r)); }
    size_t count_reporters() const { return reporters.size(); }
    void report(const Sequence& q, const Sequence& t) override { for (auto& r : reporters) r->report(q, t); }
};

template <class R>
struct ReportWithSequenceIdentity : AlignmentReporter {   // alignment_reporter.rs:32-64
    double min_seq_id, max_seq_id;
    R reporter;
    ReportWithSequenceIdentity(double lo, double hi, R r) : min_seq_id(lo), max_seq_id(hi), reporter(std::move(r)) {}
    void report(const Sequence& q, const Sequence& t) override {
        double id = AlignmentStatistics::from_sequences(q, t).percent_identity();
        if (id >= min_seq_id && id <= max_seq_id) reporter.report(q, t);
    }
};

class PairResults;

// bin/cluster_sequences.rs:77-130
struct SequenceIdentityMatrix : AlignmentReporter {
    size_t n_sequences;
    std::unordered_map<std::string, size_t> description_to_index;
    std::vector<std::vector<float>> similarity_matrix;
    explicit SequenceIdentityMatrix(const std::vector<Sequence>& seqs) : n_sequences(seqs.size()) {
        for (size_t i = 0; i < seqs.size(); ++i)
            if (!description_to_index.emplace(seqs[i].description(), i).second)
                throw std::invalid_argument("IdenticalSequenceDescriptions: " + seqs[i].description());
        similarity_matrix.assign(n_sequences, std::vector<float>(n_sequences, 0.0f));
    }
    float percent_identity(size_t i, size_t j) const { return similarity_matrix[i][j]; }
    void report(const Sequence& aq, const Sequence& at) override {
        auto st = AlignmentStatistics::from_sequences(aq, at);
        similarity_matrix[description_to_index.at(aq.description())][description_to_index.at(at.description())] =
            (float)st.percent_identity();   // cluster_sequences.rs:128: only [q][t] is written
    }
    inline void fill_from(const PairResults& r);
};

// ---------------------------------------------------------------- the GPU context
class Context {
  public:
    explicit Context(int device = 0) : h_(bsa_create(device)) {
        if (!h_) throw BsaError(BSA_ERR_CUDA, bsa_last_error(nullptr));
    }
    ~Context() { bsa_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    bsa_ctx* raw() { return h_; }
    void ck(int rc) { if (rc) throw BsaError(rc, bsa_last_error(h_)); }
    void set_scoring(const SubstitutionMatrix& m, int32_t go, int32_t ge) { ck(bsa_set_scoring(h_, m.score, m.aa_indexes, go, ge)); }
    void load(int set_id, const std::vector<Sequence>& seqs) {
        std::string res;
        std::vector<uint64_t> off(1, 0);
        for (auto& s : seqs) { res += s.as_u8(); off.push_back(res.size()); }
        ck(bsa_load_sequences(h_, set_id, (const uint8_t*)res.data(), off.data(), (uint32_t)seqs.size()));
    }

  private:
    bsa_ctx* h_;
};

// alignment_protocols.rs:96-97: the trip count of the inner loop before the triangle `break`
inline std::vector<uint32_t> triangle_counts(const std::vector<Sequence>& queries, const std::vector<Sequence>& templates,
                                             bool if_triangle_only) {
    std::vector<uint32_t> c(templates.size(), (uint32_t)queries.size());
    if (!if_triangle_only) return c;
    std::map<std::pair<std::string, std::string>, uint32_t> first;
    for (uint32_t i = 0; i < queries.size(); ++i) first.emplace(std::make_pair(queries[i].description(), queries[i].as_u8()), i);
    for (size_t t = 0; t < templates.size(); ++t) {
        auto it = first.find({templates[t].description(), templates[t].as_u8()});
        if (it != first.end()) c[t] = it->second;
    }
    return c;
}

// Batched results in the reference's report order (template-major).
class PairResults {
  public:
    std::vector<int32_t> scores;
    std::vector<uint32_t> n_identical;
    std::vector<uint32_t> q_counts;
    std::vector<uint64_t> first;            // first[t] = report index of (q=0, t)
    std::vector<size_t> len_q, len_t;       // len_ungapped of the raw sequences
    size_t index(size_t q, size_t t) const {
        if (q >= q_counts[t]) throw std::out_of_range("pair was not aligned");
        return first[t] + q;
    }
    // alignment_statistics.rs:71-73 in f64, then `as f32` (cluster_sequences.rs:128)
    float percent_identity(size_t q, size_t t) const {
        return (float)((double)n_identical[index(q, t)] / (double)std::min(len_q[q], len_t[t]) * 100.0);
    }
};

inline void SequenceIdentityMatrix::fill_from(const PairResults& r) {
    for (size_t t = 0; t < r.q_counts.size(); ++t)
        for (size_t q = 0; q < r.q_counts[t]; ++q) similarity_matrix[q][t] = r.percent_identity(q, t);
}

inline PairResults align_pairs_batched(Context& ctx, const std::vector<Sequence>& queries, const std::vector<Sequence>& templates,
                                       const SubstitutionMatrix& matrix, int32_t gap_open, int32_t gap_extend,
                                       bool if_triangle_only, bool want_identical = true) {
    if (queries.empty() || templates.empty())   // alignment_protocols.rs:86-87: max().unwrap() panics
        throw std::invalid_argument("called `Option::unwrap()` on a `None` value (empty sequence set)");
    ctx.set_scoring(matrix, gap_open, gap_extend);
    ctx.load(0, queries);
    ctx.load(1, templates);
    PairResults r;
    r.q_counts = triangle_counts(queries, templates, if_triangle_only);
    r.first.assign(templates.size() + 1, 0);
    for (size_t t = 0; t < templates.size(); ++t) r.first[t + 1] = r.first[t] + r.q_counts[t];
    for (auto& s : queries) r.len_q.push_back(len_ungapped(s));
    for (auto& s : templates) r.len_t.push_back(len_ungapped(s));
    r.scores.assign(r.first.back(), 0);
    r.n_identical.assign(r.first.back(), 0);
    uint64_t n = 0;
    ctx.ck(bsa_align_all_pairs(ctx.raw(), 0, 1, r.q_counts.data(), 0, (uint32_t)templates.size(),
                               BSA_WANT_SCORE | (want_identical ? BSA_WANT_IDENTICAL : 0u), r.scores.data(),
                               r.n_identical.data(), &n));
    return r;
}
inline PairResults align_all_vs_all(Context& ctx, const std::vector<Sequence>& seqs, const SubstitutionMatrix& m, int32_t go, int32_t ge) {
    return align_pairs_batched(ctx, seqs, seqs, m, go, ge, true);
}
inline PairResults align_one_vs_many(Context& ctx, const std::vector<Sequence>& queries, const std::vector<Sequence>& db,
                                     const SubstitutionMatrix& m, int32_t go, int32_t ge, bool want_identical = false) {
    return align_pairs_batched(ctx, queries, db, m, go, ge, false, want_identical);
}

// alignment_protocols.rs:83-115 with the reference's argument list (+ the context): every pair
// is replayed into reporter.report(aligned_query, aligned_template) in template-major order.
inline size_t align_all_pairs(Context& ctx, const std::vector<Sequence>& queries, const std::vector<Sequence>& templates,
                              const SubstitutionMatrix& matrix, int32_t gap_open, int32_t gap_extend, bool if_triangle_only,
                              AlignmentReporter& reporter) {
    if (queries.empty() || templates.empty())
        throw std::invalid_argument("called `Option::unwrap()` on a `None` value (empty sequence set)");
    ctx.set_scoring(matrix, gap_open, gap_extend);
    ctx.load(0, queries);
    ctx.load(1, templates);
    auto counts = triangle_counts(queries, templates, if_triangle_only);
    size_t reported = 0;
    for (size_t t = 0; t < templates.size(); ++t) {
        const uint32_t cnt = counts[t];
        if (!cnt) continue;
        std::vector<uint32_t> qi(cnt), ti(cnt, (uint32_t)t);
        size_t cap = 0;
        for (uint32_t q = 0; q < cnt; ++q) { qi[q] = q; cap += queries[q].len() + templates[t].len(); }
        std::vector<uint8_t> buf(cap + 1);
        std::vector<uint64_t> off(cnt + 1);
        ctx.ck(bsa_align_pairs_paths(ctx.raw(), 0, 1, qi.data(), ti.data(), cnt, nullptr, nullptr, buf.data(), off.data()));
        for (uint32_t q = 0; q < cnt; ++q) {
            std::string path((const char*)buf.data() + off[q], (size_t)(off[q + 1] - off[q]));
            auto al = aligned_sequences(path, queries[q], templates[t], '-');
            reporter.report(al.first, al.second);
            ++reported;
        }
    }
    return reported;
}

}  // namespace bioshell_seq
