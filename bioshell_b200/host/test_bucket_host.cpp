// test_bucket_host.cpp -- drives the compiled bucket-clustering driver (bioshell_bucket.hpp) so that
// tests/test_bucket_clustering.py can compare it with the restatement (oracle/pybucket.py) and with the
// golden clusterings of the reference's own test sequences:
//   usage : test_bucket_host <BLOSUM62 file> <sequences file: one sequence per line> <id_level> <n_threads>
//   output: one line per cluster, member indices separated by blanks; then "stats above below aligned"
// Without an sm_100 device it exits with 3 (there is no CPU fallback).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "bioshell_bucket.hpp"

using namespace bioshell_seq;

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: test_bucket_host <matrix> <sequences> <id_level> <n_threads>\n"); return 2; }
    std::ifstream mf(argv[1]);
    std::stringstream mtext;
    mtext << mf.rdbuf();
    std::ifstream sf(argv[2]);
    std::vector<Sequence> seqs;
    std::string line;
    while (std::getline(sf, line))
        if (!line.empty()) seqs.push_back(Sequence::from_str("seq" + std::to_string(seqs.size()), line));
    const float id_level = std::stof(argv[3]);
    const size_t n_threads = (size_t)std::stoul(argv[4]);
    try {
        const SubstitutionMatrix blosum62 = SubstitutionMatrix::ncbi_matrix_from_buffer(mtext.str());
        Context ctx(0);
        BucketClustering bc(ctx, seqs, id_level, blosum62);
        const std::vector<Cluster> clusters = bc.run_n(n_threads);
        for (const Cluster& c : clusters) {
            for (size_t k = 0; k < c.members.size(); ++k) std::printf(k ? " %zu" : "%zu", c.members[k]);
            std::printf("\n");
        }
        std::printf("stats %zu %zu %zu\n", bc.stats.above_threshold, bc.stats.below_threshold, bc.stats.aligned);
    } catch (const BsaError& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return e.rc == BSA_ERR_CUDA ? 3 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
