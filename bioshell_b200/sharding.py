"""Multi-GPU sharding of the pair set: one process per GPU, each rank takes a
contiguous range of TEMPLATES (all its queries), balanced by cells.  Results are position
addressed (t-major), so the ranks' outputs simply concatenate -- no collective on the data
path.  `plan_shards` mirrors bsa_plan_shards (csrc/bsa_api.cu) so that every rank derives
the same bounds from the same lengths without talking to anyone."""
import numpy as np


def plan_shards(len_q, len_t, q_counts, n_shards):
    """bounds[0]=0 <= ... <= bounds[n_shards]=|T|; shard r = templates [bounds[r], bounds[r+1])."""
    len_q = np.asarray(len_q, np.float64)
    len_t = np.asarray(len_t, np.float64)
    nT = len(len_t)
    qoff = np.concatenate([[0.0], np.cumsum(len_q)])
    cnt = np.full(nT, len(len_q), np.int64) if q_counts is None else np.minimum(np.asarray(q_counts, np.int64), len(len_q))
    work = len_t * qoff[cnt]
    pre = np.concatenate([[0.0], np.cumsum(work)])
    bounds = np.zeros(n_shards + 1, np.int64)
    for r in range(1, n_shards):
        want = pre[-1] * r / n_shards
        b = int(np.searchsorted(pre, want, side="left"))
        bounds[r] = min(max(b, bounds[r - 1]), nT)
    bounds[n_shards] = nT
    return bounds


def shard_result_offsets(q_counts, bounds):
    """Index of each shard's first result in the whole t-major result array."""
    first = np.concatenate([[0], np.cumsum(np.asarray(q_counts, np.int64))])
    return first[np.asarray(bounds, np.int64)]
