"""Host mirror of bioshell-clustering's hierarchical module over the C ABI (`bsa_hclust`) and
the `cluster_sequences` driver built on it -- SURVEY.md 8(f) ranks 1 and 2.

  hierarchical_clustering      bioshell-clustering/src/hierarchical/hierarchical.rs:22-80
  ClusteringTreeNode           .../hierarchical.rs:12-20, bioshell-datastructures/src/tree.rs:41-118
  single_link ... wards_method .../strategies/mod.rs:25-92
  balance_clustering_tree      .../hierarchical.rs:86-100,242-287
  medoid_by_min_max            .../hierarchical.rs:106-134
  retrieve_clusters/_data(_id) .../hierarchical.rs:139-193
  retrieve_outliers            .../hierarchical.rs:198-217
  cluster_sequences            bin/cluster_sequences.rs:133-261

The O(n^3) part -- the closest-pair scans and matrix updates of every merge -- runs on the GPU
(csrc/hclust_kernels.cuh); the tree (O(n)) is rebuilt here from the merge log.  Linkage rules are
passed BY NAME (the objects below) because the arithmetic happens on the device.
"""
import ctypes as C
import sys

import numpy as np

from .alignment import rust_fixed

from . import _lib
from .alignment import SequenceIdentityMatrix, align_all_vs_all, default_context
from .sequence_id import LabelStyle, sequence_label

sys.setrecursionlimit(max(sys.getrecursionlimit(), 1000000))


class _Linkage:
    def __init__(self, name, code):
        self.name, self.code = name, code

    def __repr__(self):
        return "<linkage %s>" % self.name


single_link = _Linkage("single_link", 0)
complete_link = _Linkage("complete_link", 1)
average_link = _Linkage("average_link", 2)
median_link = _Linkage("median_link", 3)
centroid_link = _Linkage("centroid_link", 4)
wards_method = _Linkage("wards_method", 5)


class HierarchicalCluster:
    """hierarchical.rs:12-17"""
    __slots__ = ("cluster_size", "merging_distance")

    def __init__(self, cluster_size, merging_distance):
        self.cluster_size, self.merging_distance = cluster_size, merging_distance


class ClusteringTreeNode:
    """BinaryTreeNode<HierarchicalCluster> (tree.rs:41-118)."""
    __slots__ = ("id", "value", "_left", "_right")

    def __init__(self, value):
        self.id, self.value, self._left, self._right = 0, value, None, None

    def set_left(self, node):
        self._left = node
        return self

    def set_right(self, node):
        self._right = node
        return self

    def has_left(self):
        return self._left is not None

    def has_right(self):
        return self._right is not None

    def is_leaf(self):
        return self._left is None and self._right is None

    def left(self):
        return self._left

    def right(self):
        return self._right

    def rotate(self):
        """tree.rs:106-118: swap the children in the whole subtree (iterative here)."""
        stack = [self]
        while stack:
            nd = stack.pop()
            nd._left, nd._right = nd._right, nd._left
            if nd._left is not None:
                stack.append(nd._left)
            if nd._right is not None:
                stack.append(nd._right)


def _as_matrix(n_data, distance_func):
    """Evaluate the distance closure the way HierarchicalClusteringMatrix::new does: only for
    i > j (clustering_matrix.rs:14-19).  A numpy n x n array is used as is ([i][j], i > j)."""
    if callable(distance_func):
        m = np.zeros((n_data, n_data), np.float32)
        for i in range(1, n_data):
            for j in range(i):
                m[i, j] = distance_func(i, j)
        return m
    m = np.ascontiguousarray(distance_func, np.float32)
    if m.shape != (n_data, n_data):
        raise ValueError("distance matrix must be n_data x n_data")
    return m


def hclust_merge_log(n_data, distance_func, clustering_strategy, ctx=None):
    """The GPU part: (mat_i, mat_j, merging_distance) per merge step."""
    ctx = ctx or default_context()
    m = _as_matrix(n_data, distance_func)
    k = max(n_data - 1, 1)
    mi, mj, md = np.zeros(k, np.uint32), np.zeros(k, np.uint32), np.zeros(k, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx._ck(ctx._L.bsa_hclust(ctx._h, n_data, p(m), clustering_strategy.code, 0, p(mi), p(mj), p(md)))
    return mi[:n_data - 1], mj[:n_data - 1], md[:n_data - 1]


def tree_from_merge_log(n_data, mat_i, mat_j, merge_dist):
    """Replays the bookkeeping of hierarchical.rs:34-77 (the `clusters` map keyed by matrix
    index, merged node stored at i, last cluster moved to j) on a merge log."""
    clusters = {}
    for i in range(n_data):                                  # :34-39
        c = ClusteringTreeNode(HierarchicalCluster(1, np.float32(0.0)))
        c.id = i
        clusters[i] = c
    order, current = n_data, n_data
    for s in range(n_data - 1):                              # :42-77
        i, j = int(mat_i[s]), int(mat_j[s])
        ci, cj = clusters.pop(i), clusters.pop(j)
        c = ClusteringTreeNode(HierarchicalCluster(ci.value.cluster_size + cj.value.cluster_size,
                                                   np.float32(merge_dist[s])))
        c.set_left(ci).set_right(cj)
        c.id = current
        clusters[i] = c
        last = order - 1
        if j < last:
            clusters[j] = clusters.pop(last)
        order -= 1
        current += 1
    return clusters.pop(0)


def hierarchical_clustering(n_data, distance_func, clustering_strategy, ctx=None):
    """hierarchical.rs:22-80.  `distance_func` is a closure (i, j) -> f32 or an n x n array; the
    merges run on the GPU, the tree is rebuilt from their log."""
    mi, mj, md = hclust_merge_log(n_data, distance_func, clustering_strategy, ctx)
    return tree_from_merge_log(n_data, mi, mj, md)


def retrieve_data_id(cluster):
    """hierarchical.rs:178-184: leaf ids in depth-first pre-order."""
    out, stack = [], [cluster]
    while stack:
        nd = stack.pop()
        if nd.is_leaf():
            out.append(nd.id)
        if nd._right is not None:
            stack.append(nd._right)
        if nd._left is not None:
            stack.append(nd._left)
    return out


def retrieve_data(cluster, all_data):
    """hierarchical.rs:190-196"""
    return [all_data[i] for i in retrieve_data_id(cluster)]


def retrieve_clusters(clustering_root, max_distance):
    """hierarchical.rs:139-172 (pre-order, left before right; explicit stack instead of recursion)."""
    if clustering_root.value.merging_distance <= max_distance:
        return [clustering_root]
    clusters = []
    stack = [(clustering_root, True)]            # (node, expand?)
    while stack:
        node, expand = stack.pop()
        if not expand or node.is_leaf():
            clusters.append(node)
            continue
        if node.value.merging_distance > max_distance:
            for ch in (node._right, node._left):             # pushed right first -> left is visited first
                if ch is not None:
                    stack.append((ch, ch.value.merging_distance > max_distance))
    return clusters


def _if_rotate(left, right, lm, rm, distance):
    """hierarchical.rs:242-287 on cached leftmost / rightmost leaf ids (lm / rm, by node)."""
    if right.is_leaf() and left.is_leaf():
        return False, False
    if right.is_leaf():
        return bool(distance(right.id, lm[left]) < distance(right.id, rm[left])), False
    if left.is_leaf():
        return False, bool(distance(left.id, lm[right]) > distance(left.id, rm[right]))
    rr, rl, lr, ll = rm[right], lm[right], rm[left], lm[left]
    d = [distance(lr, rl), distance(ll, rl), distance(lr, rr), distance(ll, rr)]
    k = min(range(4), key=lambda t: (d[t], t))    # Iterator::min_by returns the first minimum
    return ((False, False), (True, False), (False, True), (True, True))[k]


def balance_clustering_tree(root, distance):
    """hierarchical.rs:86-100: post-order, mirror a child's subtree (`rotate`, tree.rs:106-118) to
    bring similar leaves together.  Same decisions and same final tree as the reference, in O(n):
    the reference walks to the outermost leaves at every node and mirrors subtrees eagerly (both
    O(subtree), quadratic on the chain-like trees single linkage produces); here the outermost
    leaf ids are cached per node and a mirror is a pending flag pushed down once at the end."""
    lm, rm, flip = {}, {}, {}
    for nd in _postorder(root):                  # children before parents, left subtree first
        if nd.is_leaf():
            lm[nd] = rm[nd] = nd.id
            continue
        left, right = nd._left, nd._right        # nobody above has mirrored nd yet (post-order)
        a, b = _if_rotate(left, right, lm, rm, distance)
        if a:
            flip[left] = not flip.get(left, False)
            lm[left], rm[left] = rm[left], lm[left]
        if b:
            flip[right] = not flip.get(right, False)
            lm[right], rm[right] = rm[right], lm[right]
        lm[nd], rm[nd] = lm[left], rm[right]
    stack = [(root, False)]                      # push the pending mirrors down
    while stack:
        nd, mirrored = stack.pop()
        mirrored ^= flip.get(nd, False)
        if mirrored:
            nd._left, nd._right = nd._right, nd._left
        if nd._left is not None:
            stack.append((nd._left, mirrored))
        if nd._right is not None:
            stack.append((nd._right, mirrored))


def _postorder(root):
    out, stack = [], [(root, False)]
    while stack:
        nd, done = stack.pop()
        if done:
            out.append(nd)
            continue
        stack.append((nd, True))
        if nd._right is not None:
            stack.append((nd._right, False))
        if nd._left is not None:
            stack.append((nd._left, False))
    return out


def medoid_by_min_max(cluster, distance_fn):
    """hierarchical.rs:106-134"""
    members = retrieve_data_id(cluster)
    if len(members) == 1:
        return members[0]
    fmax = np.finfo(np.float32).max
    best, best_index = fmax, 0
    for i in range(len(members)):
        mx = -fmax
        for j in range(len(members)):
            if i != j:
                d = distance_fn(members[i], members[j])
                if d > mx:
                    mx = d
        if mx < best:
            best, best_index = mx, i
    return members[best_index]


def medoid_by_min_max_matrix(cluster, dist):
    """`medoid_by_min_max` when the distances are an n x n array (`distance_fn(i, j) = dist[i, j]`):
    the same first-minimum of the row maxima, as array operations instead of m^2 closure calls."""
    members = retrieve_data_id(cluster)
    if len(members) == 1:
        return members[0]
    fmax = np.finfo(np.float32).max
    idx = np.asarray(members)
    sub = np.array(dist[np.ix_(idx, idx)], np.float32)
    sub[np.isnan(sub)] = -fmax                   # `d > mx` is false for a NaN: it never raises the maximum
    np.fill_diagonal(sub, -fmax)                 # i != j
    mx = sub.max(axis=1)
    best = int(np.argmin(mx))                    # first minimum, like the strict `mx < best`
    return members[best] if mx[best] < fmax else members[0]


def retrieve_outliers(n_data, distance_fn, cutoff):
    """hierarchical.rs:198-217"""
    out = []
    if n_data < 2:
        return out
    fmax = np.finfo(np.float32).max
    for i in range(n_data):
        mn = fmax
        for j in range(n_data):
            if i != j:
                d = distance_fn(i, j)
                if d < mn:
                    mn = d
        if mn > cutoff:
            out.append(i)
    return out


# ---------------------------------------------------------------------------
# bin/cluster_sequences.rs
# ---------------------------------------------------------------------------
def format_fasta(seq, width=0):
    """`impl Display for Sequence` (bioshell-seq/src/sequence/display_sequence.rs:26-29)."""
    s = seq.to_string(0)
    if width:
        s = "\n".join(s[k:k + width] for k in range(0, len(s), width))
    return "> %s\n%s\n" % (seq.description(), s)


def cluster_sequences(sequences, linkage, gap_open=-10, gap_extend=-2, identity_cutoff=None, medoids=False,
                      detect_outliers=None, prefix="", sequence_width=0, distance_matrix=None, fasta=None,
                      name_width=32, reference_compat=True, ctx=None, write_files=True):
    """bin/cluster_sequences.rs:133-261 on the batched GPU path.

    reference_compat=True keeps the reference's identity matrix exactly as its reporter fills it
    (only [q][t], q < t is ever written, cluster_sequences.rs:122-130), so the clustering -- which
    reads distance(i, j) for i > j -- sees 100.0 everywhere, as the reference binary does at this
    snapshot (SURVEY.md 3.1 note).  reference_compat=False mirrors the matrix first.
    Returns a dict with the tree, the clusters (lists of sequence indices) and the sequence order.
    """
    ctx = ctx or default_context()
    res = align_all_vs_all(sequences, "BLOSUM62", gap_open, gap_extend, ctx=ctx)        # :173-177
    reporter = SequenceIdentityMatrix(sequences, name_width)
    reporter.fill_from(res)
    ident = reporter.similarity_matrix
    if not reference_compat:
        ident = np.maximum(ident, ident.T)
    dist = (np.float32(100.0) - ident).astype(np.float32)                                # :181,191
    n = len(sequences)
    out = {"identity": ident, "outliers": None, "clusters": None, "medoids": None}
    distance_fn = lambda i, j: dist[i, j]
    if detect_outliers is not None:                                                      # :180-187
        d = dist.copy()
        np.fill_diagonal(d, np.finfo(np.float32).max)
        out["outliers"] = [int(i) for i in np.nonzero(d.min(axis=1) > (np.float32(100.0) - np.float32(detect_outliers)))[0]] \
            if n >= 2 else []
        return out
    if linkage not in (single_link, complete_link, average_link):                       # :193-198
        raise ValueError("Exactly one of --single-link, --complete-link, or --average-link must be specified.")
    clustering = hierarchical_clustering(n, dist, linkage, ctx)
    out["tree"] = clustering
    if identity_cutoff is not None:                                                      # :203-226
        clusters = retrieve_clusters(clustering, (np.float32(100.0) - np.float32(identity_cutoff)))
        clusters.sort(key=lambda c: c.value.cluster_size)                                # stable, like sort_by
        out["clusters"] = [retrieve_data_id(c) for c in clusters]
        if write_files:
            for i, c in enumerate(clusters):
                with open("%scluster_%d-%d.fasta" % (prefix, i, c.value.cluster_size), "w") as fh:
                    for sid in retrieve_data_id(c):
                        fh.write(format_fasta(sequences[sid], sequence_width) + "\n")
        if medoids:
            out["medoids"] = [medoid_by_min_max_matrix(c, dist) for c in clusters]
            if write_files:
                for i, c in enumerate(clusters):
                    with open("%scenter_%d-%d.fasta" % (prefix, i, c.value.cluster_size), "w") as fh:
                        fh.write(format_fasta(sequences[out["medoids"][i]], sequence_width) + "\n")
    balance_clustering_tree(clustering, distance_fn)                                     # :228
    seq_order = retrieve_data(clustering, list(range(n)))
    out["order"] = seq_order
    if distance_matrix and write_files:                                                  # :232-248
        with open(distance_matrix, "w") as fh:
            style = LabelStyle.FullId(True, name_width)                                  # :231
            labels = {i: sequence_label(sequences[i].description(), style) for i in seq_order}
            for k, i in enumerate(seq_order):
                for l, j in enumerate(seq_order):
                    fh.write("%s\t%s\t%s\t%d\t%d\n" % (labels[i], labels[j], rust_fixed(ident[i, j], 6, 3), k, l))
                fh.write("\n")
    if fasta and write_files:                                                            # :251-256
        with open(fasta, "w") as fh:
            for i in seq_order:
                fh.write(format_fasta(sequences[i]) + "\n")
    return out
