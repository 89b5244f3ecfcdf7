// bioshell_clustering.hpp -- C++17 host mirror of bioshell-clustering's hierarchical module over the
// C ABI (`bsa_hclust`): the consumer of the identity matrix (bin/cluster_sequences.rs:189-230).
// Same names and argument meaning as the reference; the O(n^3) part -- closest-pair scans and
// matrix updates -- runs on the GPU, the tree (O(n)) is rebuilt here from the merge log.
//
//   HierarchicalCluster / ClusteringTreeNode   hierarchical/hierarchical.rs:12-20, bioshell-datastructures/src/tree.rs:41-118
//   hierarchical_clustering                    hierarchical/hierarchical.rs:22-80
//   Linkage (single_link ... wards_method)     hierarchical/strategies/mod.rs:25-92
//   balance_clustering_tree                    hierarchical/hierarchical.rs:86-100,242-287
//   medoid_by_min_max                          hierarchical/hierarchical.rs:106-134
//   retrieve_clusters / _data_id / _data       hierarchical/hierarchical.rs:139-196
//   retrieve_outliers                          hierarchical/hierarchical.rs:198-217
//
// Every traversal uses an explicit stack (single linkage gives chain-like trees, 10^5 levels deep)
// and balance_clustering_tree is linear: the outermost leaf ids are cached per node and a subtree
// mirror (`rotate`, tree.rs:106-118) is a pending flag pushed down once at the end -- the same
// decisions and the same final tree as the reference's eager form.
#pragma once
#include <cfloat>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

#include "bioshell_seq.hpp"

namespace bioshell_clustering {

using bioshell_seq::Context;

struct HierarchicalCluster {   // hierarchical.rs:12-17
    size_t cluster_size;
    float merging_distance;
};

struct ClusteringTreeNode {    // BinaryTreeNode<HierarchicalCluster>, tree.rs:41-118
    size_t id = 0;
    HierarchicalCluster value{1, 0.0f};
    std::unique_ptr<ClusteringTreeNode> left, right;
    bool is_leaf() const { return !left && !right; }
    bool has_left() const { return (bool)left; }
    bool has_right() const { return (bool)right; }
    ~ClusteringTreeNode() {    // iterative: a chain of 10^5 unique_ptrs must not recurse
        std::vector<std::unique_ptr<ClusteringTreeNode>> stack;
        if (left) stack.push_back(std::move(left));
        if (right) stack.push_back(std::move(right));
        while (!stack.empty()) {
            auto nd = std::move(stack.back());
            stack.pop_back();
            if (nd->left) stack.push_back(std::move(nd->left));
            if (nd->right) stack.push_back(std::move(nd->right));
        }
    }
};
using Tree = std::unique_ptr<ClusteringTreeNode>;

// strategies/mod.rs:25-92, in the order of bsa_hclust's linkage codes
enum class Linkage : int { single_link = 0, complete_link = 1, average_link = 2, median_link = 3, centroid_link = 4, wards_method = 5 };

using DistanceFn = std::function<float(size_t, size_t)>;

// Replays the bookkeeping of hierarchical.rs:34-77 (the `clusters` map keyed by matrix index,
// merged node stored at i, last cluster moved to j) on a merge log.
inline Tree tree_from_merge_log(size_t n_data, const std::vector<uint32_t>& mat_i, const std::vector<uint32_t>& mat_j,
                                const std::vector<float>& merge_dist) {
    if (n_data == 0) throw std::invalid_argument("no data to cluster");
    std::unordered_map<size_t, Tree> clusters;
    for (size_t i = 0; i < n_data; ++i) {
        auto c = std::make_unique<ClusteringTreeNode>();
        c->id = i;
        clusters.emplace(i, std::move(c));
    }
    size_t order = n_data, current = n_data;
    for (size_t s = 0; s + 1 < n_data; ++s) {
        const size_t i = mat_i.at(s), j = mat_j.at(s);
        Tree ci = std::move(clusters.at(i)), cj = std::move(clusters.at(j));
        clusters.erase(i);
        clusters.erase(j);
        auto c = std::make_unique<ClusteringTreeNode>();
        c->value = {ci->value.cluster_size + cj->value.cluster_size, merge_dist.at(s)};
        c->left = std::move(ci);
        c->right = std::move(cj);
        c->id = current;
        clusters[i] = std::move(c);
        const size_t last = order - 1;
        if (j < last) {
            clusters[j] = std::move(clusters.at(last));
            clusters.erase(last);
        }
        --order;
        ++current;
    }
    return std::move(clusters.at(0));
}

// hierarchical.rs:22-80.  `dist` is n x n row-major; only dist[i * n + j] with i > j is read, as
// HierarchicalClusteringMatrix::new evaluates its closure (clustering_matrix.rs:14-19).
inline Tree hierarchical_clustering(Context& ctx, size_t n_data, const std::vector<float>& dist, Linkage strategy) {
    if (dist.size() != n_data * n_data) throw std::invalid_argument("distance matrix must be n_data x n_data");
    const size_t k = n_data > 1 ? n_data - 1 : 1;
    std::vector<uint32_t> mi(k), mj(k);
    std::vector<float> md(k);
    ctx.ck(bsa_hclust(ctx.raw(), (uint32_t)n_data, dist.data(), (int)strategy, 0, mi.data(), mj.data(), md.data()));
    return tree_from_merge_log(n_data, mi, mj, md);
}
inline Tree hierarchical_clustering(Context& ctx, size_t n_data, const DistanceFn& distance, Linkage strategy) {
    std::vector<float> m(n_data * n_data, 0.0f);
    for (size_t i = 1; i < n_data; ++i)
        for (size_t j = 0; j < i; ++j) m[i * n_data + j] = distance(i, j);
    return hierarchical_clustering(ctx, n_data, m, strategy);
}

// hierarchical.rs:178-184: leaf ids in depth-first pre-order
inline std::vector<size_t> retrieve_data_id(const ClusteringTreeNode& cluster) {
    std::vector<size_t> out;
    std::vector<const ClusteringTreeNode*> stack{&cluster};
    while (!stack.empty()) {
        const ClusteringTreeNode* nd = stack.back();
        stack.pop_back();
        if (nd->is_leaf()) out.push_back(nd->id);
        if (nd->right) stack.push_back(nd->right.get());
        if (nd->left) stack.push_back(nd->left.get());
    }
    return out;
}

// hierarchical.rs:190-196
template <class T>
std::vector<T> retrieve_data(const ClusteringTreeNode& cluster, const std::vector<T>& all_data) {
    std::vector<T> out;
    for (size_t i : retrieve_data_id(cluster)) out.push_back(all_data.at(i));
    return out;
}

// hierarchical.rs:139-172: the subtrees whose merging distance is within max_distance (pre-order, left first)
inline std::vector<const ClusteringTreeNode*> retrieve_clusters(const ClusteringTreeNode& root, float max_distance) {
    std::vector<const ClusteringTreeNode*> clusters;
    if (root.value.merging_distance <= max_distance) return {&root};
    std::vector<std::pair<const ClusteringTreeNode*, bool>> stack{{&root, true}};
    while (!stack.empty()) {
        auto [node, expand] = stack.back();
        stack.pop_back();
        if (!expand || node->is_leaf()) {
            clusters.push_back(node);
            continue;
        }
        if (node->value.merging_distance > max_distance) {
            for (const ClusteringTreeNode* ch : {node->right.get(), node->left.get()})   // right first: left is visited first
                if (ch) stack.push_back({ch, ch->value.merging_distance > max_distance});
        }
    }
    return clusters;
}

namespace detail {
inline std::vector<ClusteringTreeNode*> postorder(ClusteringTreeNode& root) {
    std::vector<ClusteringTreeNode*> out;
    std::vector<std::pair<ClusteringTreeNode*, bool>> stack{{&root, false}};
    while (!stack.empty()) {
        auto [nd, done] = stack.back();
        stack.pop_back();
        if (done) {
            out.push_back(nd);
            continue;
        }
        stack.push_back({nd, true});
        if (nd->right) stack.push_back({nd->right.get(), false});
        if (nd->left) stack.push_back({nd->left.get(), false});
    }
    return out;
}
}  // namespace detail

// hierarchical.rs:86-100 with the rotation rule of :242-287
inline void balance_clustering_tree(ClusteringTreeNode& root, const DistanceFn& distance) {
    std::unordered_map<const ClusteringTreeNode*, size_t> lm, rm;
    std::unordered_map<const ClusteringTreeNode*, bool> flip;
    for (ClusteringTreeNode* nd : detail::postorder(root)) {
        if (nd->is_leaf()) {
            lm[nd] = rm[nd] = nd->id;
            continue;
        }
        ClusteringTreeNode *left = nd->left.get(), *right = nd->right.get();
        bool a = false, b = false;
        if (right->is_leaf() && left->is_leaf()) {
        } else if (right->is_leaf()) {
            a = distance(right->id, lm[left]) < distance(right->id, rm[left]);
        } else if (left->is_leaf()) {
            b = distance(left->id, lm[right]) > distance(left->id, rm[right]);
        } else {
            const size_t rr = rm[right], rl = lm[right], lr = rm[left], ll = lm[left];
            const float d[4] = {distance(lr, rl), distance(ll, rl), distance(lr, rr), distance(ll, rr)};
            int k = 0;
            for (int t = 1; t < 4; ++t)
                if (d[t] < d[k]) k = t;      // Iterator::min_by returns the first minimum
            a = (k & 1) != 0;
            b = (k & 2) != 0;
        }
        if (a) {
            flip[left] = !flip[left];
            std::swap(lm[left], rm[left]);
        }
        if (b) {
            flip[right] = !flip[right];
            std::swap(lm[right], rm[right]);
        }
        lm[nd] = lm[left];
        rm[nd] = rm[right];
    }
    std::vector<std::pair<ClusteringTreeNode*, bool>> stack{{&root, false}};   // push the pending mirrors down
    while (!stack.empty()) {
        auto [nd, mirrored] = stack.back();
        stack.pop_back();
        auto it = flip.find(nd);
        if (it != flip.end() && it->second) mirrored = !mirrored;
        if (mirrored) std::swap(nd->left, nd->right);
        if (nd->left) stack.push_back({nd->left.get(), mirrored});
        if (nd->right) stack.push_back({nd->right.get(), mirrored});
    }
}

// hierarchical.rs:106-134: the member whose largest distance to another member is smallest (first minimum)
inline size_t medoid_by_min_max(const ClusteringTreeNode& cluster, const DistanceFn& distance_fn) {
    const std::vector<size_t> members = retrieve_data_id(cluster);
    if (members.size() == 1) return members[0];
    float best = FLT_MAX;
    size_t best_index = 0;
    for (size_t i = 0; i < members.size(); ++i) {
        float mx = -FLT_MAX;
        for (size_t j = 0; j < members.size(); ++j) {
            if (i == j) continue;
            const float d = distance_fn(members[i], members[j]);
            if (d > mx) mx = d;
        }
        if (mx < best) {
            best = mx;
            best_index = i;
        }
    }
    return members[best_index];
}

// hierarchical.rs:198-217: elements whose nearest neighbour is farther than cutoff
inline std::vector<size_t> retrieve_outliers(size_t n_data, const DistanceFn& distance_fn, float cutoff) {
    std::vector<size_t> out;
    if (n_data < 2) return out;
    for (size_t i = 0; i < n_data; ++i) {
        float mn = FLT_MAX;
        for (size_t j = 0; j < n_data; ++j) {
            if (i == j) continue;
            const float d = distance_fn(i, j);
            if (d < mn) mn = d;
        }
        if (mn > cutoff) out.push_back(i);
    }
    return out;
}

}  // namespace bioshell_clustering
