"""Literal pure-Python restatement of bioshell-seq's bucket clustering (TEST INFRASTRUCTURE):
bioshell-seq/src/sequence/bucket_clustering/{bucket_clustering.rs:141-309, kmers.rs:17-121}.
One representative at a time, one alignment at a time (through the C oracle), exactly in the
reference's order.  Parity status: the reference's two tests only run the function
(tests/test_bucket_clustering.rs:36-62) and hold no expected output, so the pin is
tests/golden/bucket_kats.json: clusterings of the reference's own test sequences derived from a
table of (pinned-oracle) identities and a table of k-mer verdicts computed independently of this
file (tools/gen_bucket_golden.py), both stored so that the expected clusters can be checked by hand."""
import numpy as np

from . import c_oracle

ORDER = "ARNDCQEGHILKMFPSTWYVXacgtuacgt-_Z*"


def letter_index():
    t = {}
    for i, ch in enumerate(ORDER):
        t[ord(ch)] = i
    t[ord("B")] = t[ord("N")]
    t[ord("Z")] = t[ord("Q")]
    return t


LETTER = letter_index()


def generate_kmers(seq, k):
    if k == 0 or k > 6 or len(seq) < k:
        return []
    mask = (1 << (5 * k)) - 1
    code, out = 0, []
    for i, b in enumerate(seq):
        if b not in LETTER:
            raise ValueError("InvalidOneLetterCode")
        x = LETTER[b]
        assert x <= 31
        code = ((code << 5) | x) & mask
        if i + 1 >= k:
            out.append(code)
    return sorted(set(out))


def count_intersection_sorted(a, b):
    i = j = c = 0
    while i < len(a) and j < len(b):
        if a[i] < b[j]:
            i += 1
        elif a[i] > b[j]:
            j += 1
        else:
            c += 1; i += 1; j += 1
    return c


def kmer_identity_bounds(different, k, min_len):
    if min_len == 0:
        return np.float32(0), np.float32(0)
    wrap = lambda v: v % (1 << 64)
    upper = np.float32(wrap(min_len - (different // k + 1))) / np.float32(min_len)
    lower = np.float32(wrap(min_len - (different + k - 1))) / np.float32(min_len)
    return max(lower, np.float32(0)), min(upper, np.float32(1))


def suggest_word_length(x):
    x = np.float32(x)
    for lim, k in ((0.95, 6), (0.90, 5), (0.85, 5), (0.80, 4), (0.75, 4), (0.70, 3), (0.60, 3), (0.50, 2)):
        if x >= np.float32(lim):
            return k
    return 1


def run(seqs, id_level, score, aa, n_threads=1):
    """seqs: list of bytes.  Returns the clusters as lists of member indices."""
    id_level = np.float32(id_level)
    k = suggest_word_length(id_level)
    order = sorted(range(len(seqs)), key=lambda i: -len(seqs[i]))          # stable
    kmers = [generate_kmers(s, k) for s in seqs]

    def identity(rep, cand):
        shared = count_intersection_sorted(kmers[cand], kmers[rep])
        different = max(len(kmers[cand]) - shared, 0)
        shorter = min(len(seqs[cand]), len(seqs[rep]))
        lo, up = kmer_identity_bounds(different, k, shorter)
        if lo >= id_level:
            return "above"
        if up < id_level:
            return "below"
        r = c_oracle.align_pair(seqs[rep], seqs[cand], score, aa, -11, -1)
        return np.float32(r["n_identical"]) / np.float32(shorter)

    def merge(c1, c2):
        for b in c2:
            assigned = False
            for a in c1:
                v = identity(a[0], b[0])
                if isinstance(v, str):
                    if v == "above":
                        a[1].extend(b[1]); assigned = True
                        break
                    continue
                if v >= id_level:
                    a[1].extend(b[1]); assigned = True
                    break
            if not assigned:
                c1.append(b)
        return c1

    singles = [[i, [i]] for i in order]
    if n_threads == 1:
        return [c[1] for c in merge([], singles)]
    n_threads = min(max(n_threads, 1), len(singles))
    chunk = -(-len(singles) // n_threads)
    cl = [merge([], [[c[0], list(c[1])] for c in singles[b:b + chunk]]) for b in range(0, len(singles), chunk)]
    while len(cl) > 1:
        cl = [merge(cl[b], cl[b + 1]) if b + 1 < len(cl) else cl[b] for b in range(0, len(cl), 2)]
    return [c[1] for c in cl.pop()]
