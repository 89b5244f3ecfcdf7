// test_host_mirror.cpp -- the reference's own aligner test (bioshell-seq/tests/test_aligners.rs:13-58)
// and doc-tests, transcribed against the C++ host mirror; everything aligns on the GPU through
// the C ABI.  Built by __graft_entry__.build(), run by tests/test_gpu_cpp_host.py (needs a B200).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "bioshell_seq.hpp"

using namespace bioshell_seq;

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
    } while (0)

struct GlobalAlignmentTestCase { const char *query, *tmplt, *aligned_query, *aligned_template, *alignment; int score; };
static const GlobalAlignmentTestCase GLOBAL_CASES[3] = {
    {"A", "AW", "A-", "AW", "*-", -6},
    {"AR", "ARK", "AR-", "ARK", "**-", -1},
    {"MAVRLLKTHL", "MKNITCYL", "MAVRLLKTHL", "M--KNITCYL", "*||*******", -2},
};

struct Collect : AlignmentReporter {
    std::vector<std::pair<Sequence, Sequence>> pairs;
    void report(const Sequence& q, const Sequence& t) override { pairs.emplace_back(q, t); }
};

static std::string align_one(Context& ctx, const SubstitutionMatrix& m, const std::string& q, const std::string& t, int* score) {
    ctx.set_scoring(m, -10, -2);
    ctx.load(0, {Sequence("query", q)});
    ctx.load(1, {Sequence("template", t)});
    uint32_t qi = 0, ti = 0;
    std::vector<uint8_t> buf(q.size() + t.size() + 1);
    uint64_t off[2];
    int32_t s;
    ctx.ck(bsa_align_pairs_paths(ctx.raw(), 0, 1, &qi, &ti, 1, &s, nullptr, buf.data(), off));
    *score = s;
    return std::string((const char*)buf.data(), (size_t)off[1]);
}

static int run(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <BLOSUM62 NCBI file>\n", argv[0]); return 2; }
    SubstitutionMatrix blosum62 = SubstitutionMatrix::ncbi_matrix_from_file(argv[1]);
    CHECK(blosum62.score_by_aa('C', 'C') == 9 && blosum62.score_by_aa('W', 'W') == 11);     // scoring/mod.rs:46-47
    CHECK(blosum62.score_by_aa('X', 'X') == -1 && blosum62.score_by_aa('A', 'W') == -3);
    Context ctx(0);

    // test_global_aligner, both orientations
    for (const auto& c : GLOBAL_CASES) {
        int score;
        std::string path = align_one(ctx, blosum62, c.query, c.tmplt, &score);
        auto al = aligned_strings(path, c.query, c.tmplt, '-');
        CHECK(score == c.score);
        CHECK(al.first == c.aligned_query && al.second == c.aligned_template);
        CHECK(path == c.alignment);
        path = align_one(ctx, blosum62, c.tmplt, c.query, &score);
        al = aligned_strings(path, c.tmplt, c.query, '-');
        CHECK(score == c.score);
        CHECK(al.second == c.aligned_query && al.first == c.aligned_template);
    }
    // alignment_path.rs:155-159 ; alignment_statistics.rs:16-26
    auto ex = aligned_strings("**-**", "ALIV", "ALRIV", '-');
    CHECK(ex.first == "AL-IV" && ex.second == "ALRIV");
    auto st = AlignmentStatistics::from_sequences(Sequence("query", "EIIIDSYNQFSDR----SYQFMTPSLFVR"),
                                                  Sequence("templ", "ETVKEAYDLYPDRRYFGSFQFLYPSLFLR"));
    CHECK(st.query_length == 25 && st.template_length == 29 && st.n_identical == 12);
    CHECK(st.to_string() == "query templ  48.00 %  12   25   29");

    // align_all_pairs: t-major replay == batched results == SequenceIdentityMatrix
    std::vector<Sequence> seqs = {Sequence("1clf:A", "AYKIADSCVSCGACASECPVNAISQGDSIFVIDADTCIDCGNCANVCPVGAPVQE"),
                                  Sequence("1dur:A", "AYVINDSCIACGACKPECPVNCIQEGSIYAIDADSCIDCGSCASVCPVGAPNPED"),
                                  Sequence("1fca:A", "AYVINEACISCGACEPECPVDAISQGGSRYVIDADTCIDCGACAGVCPVDAPVQA"),
                                  Sequence("short", "ACDC")};
    Collect col;
    size_t n = align_all_pairs(ctx, seqs, seqs, blosum62, -10, -1, true, col);
    CHECK(n == 6 && col.pairs.size() == 6);
    const int order[6][2] = {{0, 1}, {0, 2}, {1, 2}, {0, 3}, {1, 3}, {2, 3}};   // alignment_protocols.rs:94-102
    PairResults res = align_all_vs_all(ctx, seqs, blosum62, -10, -1);
    SequenceIdentityMatrix replayed(seqs), batched(seqs);
    for (size_t k = 0; k < 6; ++k) {
        CHECK(col.pairs[k].first.description() == seqs[order[k][0]].description());
        CHECK(col.pairs[k].second.description() == seqs[order[k][1]].description());
        CHECK(col.pairs[k].first.len() == col.pairs[k].second.len());
        CHECK(count_identical(col.pairs[k].first, col.pairs[k].second) == res.n_identical[k]);
        replayed.report(col.pairs[k].first, col.pairs[k].second);
    }
    batched.fill_from(res);
    CHECK(replayed.similarity_matrix == batched.similarity_matrix);
    CHECK(batched.similarity_matrix[1][0] == 0.0f);    // the reference never writes [t][q]

    // error behaviour
    bool threw = false;
    try { ctx.set_scoring(blosum62, -1, -2); } catch (const BsaError& e) { threw = e.rc == BSA_ERR_UNSUPPORTED_GAPS; }
    CHECK(threw);
    threw = false;
    try { align_all_vs_all(ctx, {}, blosum62, -10, -1); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    std::printf("host mirror ok: 3 reference KATs x 2 orientations, replay order, identity matrix\n");
    return 0;
}

int main(int argc, char** argv) {
    try {
        return run(argc, argv);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 3;
    }
}
