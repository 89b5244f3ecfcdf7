"""Independent pure-Python restatement of bioshell-clustering's hierarchical clustering and its
tree helpers.  TEST INFRASTRUCTURE (small cases only).

  HierarchicalClusteringMatrix  bioshell-clustering/src/hierarchical/clustering_matrix.rs:1-75
  hierarchical_clustering       .../hierarchical.rs:22-80
  balance_clustering_tree       .../hierarchical.rs:86-100, if_rotate :242-287
  medoid_by_min_max             .../hierarchical.rs:106-134
  retrieve_clusters             .../hierarchical.rs:139-172
  retrieve_data_id              .../hierarchical.rs:178-184
  retrieve_outliers             .../hierarchical.rs:198-217
  BinaryTreeNode::rotate        bioshell-datastructures/src/tree.rs:106-118
"""
import numpy as np

F32MAX = np.finfo(np.float32).max
f32 = np.float32


def rule_fn(name):
    def single(si, sj, sk, dij, dik, djk):
        return min(dik, djk)

    def complete(si, sj, sk, dij, dik, djk):
        return max(dik, djk)

    def average(si, sj, sk, dij, dik, djk):
        d = f32(1.0) / f32(si + sj)
        return f32(f32(f32(d * f32(si)) * dik) + f32(f32(d * f32(sj)) * djk))

    def median(si, sj, sk, dij, dik, djk):
        return f32(f32(f32(f32(0.5) * dik) + f32(f32(0.5) * djk)) - f32(f32(0.25) * dij))

    def centroid(si, sj, sk, dij, dik, djk):
        d = f32(1.0) / f32(si + sj)
        a = f32(f32(f32(d * f32(si)) * dik) + f32(f32(d * f32(sj)) * djk))
        b = f32(f32(f32(f32(f32(si) * f32(sj)) * d) * d) * dij)
        return f32(a - b)

    def ward(si, sj, sk, dij, dik, djk):
        d = f32(1.0) / f32(si + sj + sk)
        a = f32(f32(f32(d * f32(si + sk)) * dik) + f32(f32(d * f32(sj + sk)) * djk))
        return f32(a - f32(f32(f32(sk) * d) * dij))

    return dict(single=single, complete=complete, average=average, median=median, centroid=centroid,
                ward=ward)[name]


class Node:
    def __init__(self, nid, size, dist, left=None, right=None):
        self.id, self.cluster_size, self.merging_distance, self.left, self.right = nid, size, dist, left, right

    def is_leaf(self):
        return self.left is None and self.right is None


def hierarchical_clustering(n, distance, rule):
    """distance(i, j) is called for i > j only; returns (root, merge log)."""
    rule = rule_fn(rule) if isinstance(rule, str) else rule
    m = [[f32(0.0)] * n for _ in range(n)]
    for i in range(1, n):
        for j in range(i):
            m[i][j] = m[j][i] = f32(distance(i, j))
    sizes = [1] * n
    clusters = {i: Node(i, 1, f32(0.0)) for i in range(n)}
    order, cur, log = n, n, []
    while len(clusters) > 1:
        best, bi, bj = F32MAX, 0, 0
        for j in range(1, order):
            for i in range(j):
                if m[i][j] < best:
                    best, bi, bj = m[i][j], i, j
        i, j = bi, bj
        ci, cj = clusters.pop(i), clusters.pop(j)
        dij = m[i][j]
        log.append((i, j, ci.id, cj.id, float(dij)))
        clusters[i] = Node(cur, ci.cluster_size + cj.cluster_size, dij, ci, cj)
        si, sj = sizes[i], sizes[j]
        res = [rule(si, sj, sizes[j], m[i][j], m[i][k], m[j][k]) for k in range(order)]
        for k in range(order):
            m[i][k] = m[k][i] = res[k]
        m[i][i] = f32(0.0)
        sizes[i] = si + sj
        last = order - 1
        if j < last:
            clusters[j] = clusters.pop(last)
            order -= 1
            m[j], m[order] = m[order], m[j]
            for r in range(order):
                m[r][j] = m[r][order]
        else:
            order -= 1
        cur += 1
    return clusters[0], log


def retrieve_data_id(node):
    out = []

    def rec(nd):
        if nd.is_leaf():
            out.append(nd.id)
        if nd.left is not None:
            rec(nd.left)
        if nd.right is not None:
            rec(nd.right)
    rec(node)
    return out


def rotate(node):
    if node.left is not None:
        rotate(node.left)
    if node.right is not None:
        rotate(node.right)
    node.left, node.right = node.right, node.left


def _leftmost(c):
    return _leftmost(c.left) if c.left is not None else c.id


def _rightmost(c):
    return _rightmost(c.right) if c.right is not None else c.id


def _if_rotate(c, distance):
    if c.is_leaf():
        return False, False
    left, right = c.left, c.right
    if right.is_leaf() and left.is_leaf():
        return False, False
    if right.is_leaf():
        return (distance(right.id, _leftmost(left)) < distance(right.id, _rightmost(left))), False
    if left.is_leaf():
        return False, (distance(left.id, _leftmost(right)) > distance(left.id, _rightmost(right)))
    rr, rl, lr, ll = _rightmost(right), _leftmost(right), _rightmost(left), _leftmost(left)
    d = [distance(lr, rl), distance(ll, rl), distance(lr, rr), distance(ll, rr)]
    k = min(range(4), key=lambda t: (d[t], t))     # min_by keeps the first minimum
    return {0: (False, False), 1: (True, False), 2: (False, True), 3: (True, True)}[k]


def balance_clustering_tree(root, distance):
    def rec(nd):
        if nd.left is not None:
            rec(nd.left)
        if nd.right is not None:
            rec(nd.right)
        a, b = _if_rotate(nd, distance)
        if a:
            rotate(nd.left)
        if b:
            rotate(nd.right)
    rec(root)


def retrieve_clusters(root, max_distance):
    out = []

    def rec(nd):
        if nd.is_leaf():
            out.append(nd)
        elif nd.merging_distance > max_distance:
            for ch in (nd.left, nd.right):
                if ch is not None:
                    if ch.merging_distance <= max_distance:
                        out.append(ch)
                    else:
                        rec(ch)
    if root.merging_distance <= max_distance:
        out.append(root)
    else:
        rec(root)
    return out


def medoid_by_min_max(cluster, distance):
    members = retrieve_data_id(cluster)
    if len(members) == 1:
        return members[0]
    best, best_i = F32MAX, 0
    for i in range(len(members)):
        mx = -F32MAX
        for j in range(len(members)):
            if i != j:
                d = distance(members[i], members[j])
                if d > mx:
                    mx = d
        if mx < best:
            best, best_i = mx, i
    return members[best_i]


def retrieve_outliers(n, distance, cutoff):
    out = []
    if n < 2:
        return out
    for i in range(n):
        mn = F32MAX
        for j in range(n):
            if i != j:
                d = distance(i, j)
                if d < mn:
                    mn = d
        if mn > cutoff:
            out.append(i)
    return out
