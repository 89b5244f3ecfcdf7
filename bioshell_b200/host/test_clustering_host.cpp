// test_clustering_host.cpp -- drives the C++ clustering mirror (bioshell_clustering.hpp) from a text
// file so that tests/test_cpp_host.py can compare it with the oracle (oracle/pyhclust.py):
//   input : n  linkage  cutoff  outlier_cutoff \n  (n-1) x "mat_i mat_j dist" \n  n x n distances
//   output: leaf order, clusters at the cutoff (sorted by size, stable) with their medoids, outliers,
//           leaf order after balance_clustering_tree
// With --gpu the merge log in the file is ignored and the tree comes from bsa_hclust on the device.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "bioshell_clustering.hpp"

using namespace bioshell_clustering;

static void print_ids(const char* tag, const std::vector<size_t>& ids) {
    std::printf("%s", tag);
    for (size_t i : ids) std::printf(" %zu", i);
    std::printf("\n");
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: test_clustering_host <input> [--gpu]\n"); return 2; }
    const bool gpu = argc > 2 && !std::strcmp(argv[2], "--gpu");
    std::ifstream in(argv[1]);
    size_t n;
    int linkage;
    float cutoff, outlier_cutoff;
    in >> n >> linkage >> cutoff >> outlier_cutoff;
    std::vector<uint32_t> mi(n > 1 ? n - 1 : 0), mj(mi.size());
    std::vector<float> md(mi.size()), dist(n * n);
    for (size_t s = 0; s < mi.size(); ++s) in >> mi[s] >> mj[s] >> md[s];
    for (float& d : dist) in >> d;
    if (!in) { std::fprintf(stderr, "bad input file\n"); return 2; }
    const DistanceFn distance = [&](size_t i, size_t j) { return dist[i * n + j]; };
    try {
        Tree root;
        if (gpu) {
            Context ctx(0);
            root = hierarchical_clustering(ctx, n, dist, (Linkage)linkage);
        } else {
            root = tree_from_merge_log(n, mi, mj, md);
        }
        print_ids("order", retrieve_data_id(*root));
        auto clusters = retrieve_clusters(*root, cutoff);
        std::stable_sort(clusters.begin(), clusters.end(), [](const ClusteringTreeNode* a, const ClusteringTreeNode* b) {
            return a->value.cluster_size < b->value.cluster_size;      // cluster_sequences.rs:205
        });
        for (const ClusteringTreeNode* c : clusters) {
            std::printf("cluster %zu medoid %zu :", c->value.cluster_size, medoid_by_min_max(*c, distance));
            print_ids("", retrieve_data_id(*c));
        }
        print_ids("outliers", retrieve_outliers(n, distance, outlier_cutoff));
        balance_clustering_tree(*root, distance);
        print_ids("balanced", retrieve_data_id(*root));
    } catch (const bioshell_seq::BsaError& e) {
        std::fprintf(stderr, "BsaError: %s\n", e.what());
        return 3;
    }
    std::printf("clustering host ok\n");
    return 0;
}
