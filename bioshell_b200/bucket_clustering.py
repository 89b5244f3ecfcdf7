"""Host mirror of bioshell-seq's k-mer bucket clustering (CD-HIT-like greedy incremental
clustering) with its aligner calls batched onto the GPU -- SURVEY.md 8(f) rank 3.

  bucket_clustering / bucket_clustering_n   bioshell-seq/src/sequence/bucket_clustering/bucket_clustering.rs:33-67
  BucketClustering::{new, run, run_n, merge, sequence_identity}   .../bucket_clustering.rs:141-309
  generate_kmers, count_intersection_sorted, kmer_identity_bounds, suggest_word_length
                                            .../kmers.rs:17-121
  standard_letter_to_index                  bioshell-seq/src/chemical/residue_types.rs:537-580

The reference walks the representatives one by one and aligns (GlobalAligner, BLOSUM62, -11/-1)
only when the k-mer bounds are inconclusive, stopping at the first hit.  Here the k-mer verdicts of
one candidate against all current representatives are taken first (host, cheap); the inconclusive
ones that precede the first certain hit are aligned in ONE GPU batch (in blocks, stopping at the
first block that contains a hit) and then scanned in the reference's order, so the resulting
clustering is identical; only the `aligned` statistic can be larger than the reference's.
A batch is one candidate (the template, columns) against a list of representatives (queries,
rows): the representatives are gathered into a set of their own on the device
(bsa_gather_sequences) and the batch runs on the forward score + identity kernels
(bsa_align_all_pairs) -- no direction store, no traceback.  The compiled twin of this driver is
bioshell_b200/host/bioshell_bucket.hpp.
"""
import numpy as np

from .alignment import default_context
from .sequence import Sequence, pack

_INVALID = 255
_ORDER = "ARNDCQEGHILKMFPSTWYVXacgtuacgt-_Z*"       # StandardResidueType::TYPES code1, residue_types.rs:499-533


def _letter_table():
    t = np.full(256, _INVALID, np.uint8)
    for i, ch in enumerate(_ORDER):                 # later entries overwrite earlier ones (:540-542)
        t[ord(ch)] = i
    t[ord("B")] = t[ord("N")]                       # :544
    t[ord("Z")] = t[ord("Q")]                       # :546
    return t


STANDARD_LETTER_TO_INDEX = _letter_table()


def standard_letter_to_index(letter):
    """residue_types.rs:572-580"""
    b = letter if isinstance(letter, int) else ord(letter)
    idx = int(STANDARD_LETTER_TO_INDEX[b])
    if idx == _INVALID:
        raise ValueError("InvalidOneLetterCode: %r" % chr(b))
    return idx


def generate_kmers(seq, k):
    """kmers.rs:17-46: sorted, de-duplicated 5-bit-per-symbol k-mers as u32."""
    seq = np.frombuffer(bytes(seq), np.uint8)
    if k == 0 or k > 6 or len(seq) < k:
        return np.zeros(0, np.uint32)
    x = STANDARD_LETTER_TO_INDEX[seq]
    if np.any(x == _INVALID):
        bad = seq[np.nonzero(x == _INVALID)[0][0]]
        raise ValueError("InvalidOneLetterCode: %r" % chr(int(bad)))
    if np.any(x > 31):
        raise AssertionError("symbol value %d exceeds maximum allowed value 31" % int(x.max()))
    code = np.zeros(len(seq) - k + 1, np.uint64)
    for d in range(k):                              # code = ((code << 5) | x) & mask, vectorised
        code = (code << np.uint64(5)) | x[d:len(seq) - k + 1 + d].astype(np.uint64)
    return np.unique(code.astype(np.uint32))


def count_intersection_sorted(a, b):
    """kmers.rs:55-78"""
    return int(np.intersect1d(a, b, assume_unique=True).size)


def _usize_sub(a, b):
    """`a - b` on usize in a release build (wrapping)."""
    return (a - b) % (1 << 64)


def kmer_identity_bounds(different_kmers, kmer_len, min_seq_len):
    """kmers.rs:88-100 (f32 results)."""
    if min_seq_len == 0:
        return np.float32(0.0), np.float32(0.0)
    min_mut = different_kmers // kmer_len + 1
    upper = np.float32(_usize_sub(min_seq_len, min_mut)) / np.float32(min_seq_len)
    max_mut = different_kmers + kmer_len - 1
    lower = np.float32(_usize_sub(min_seq_len, max_mut)) / np.float32(min_seq_len)
    return np.float32(max(lower, np.float32(0.0))), np.float32(min(upper, np.float32(1.0)))


def suggest_word_length(identity_level):
    """kmers.rs:110-121"""
    for lim, k in ((0.95, 6), (0.90, 5), (0.85, 5), (0.80, 4), (0.75, 4), (0.70, 3), (0.60, 3), (0.50, 2)):
        if identity_level >= np.float32(lim):
            return k
    return 1


class Cluster:
    """bucket_clustering.rs:77-91"""
    __slots__ = ("representative", "members")

    def __init__(self, representative, members=None):
        self.representative = representative
        self.members = [representative] if members is None else members

    def clone(self):
        return Cluster(self.representative, list(self.members))


class BucketClustering:
    """bucket_clustering.rs:69-75,141-309"""
    BLOCK = 64          # inconclusive representatives aligned per GPU batch

    def __init__(self, sequences, id_level, ctx=None):
        self.id_level = np.float32(id_level)
        self.word_size = suggest_word_length(self.id_level)
        self.sequences = list(sequences)
        lens = np.array([s.len() for s in self.sequences], np.int64)
        self.sequence_order = [int(i) for i in np.argsort(-lens, kind="stable")]        # :148-149 (stable sort)
        self.kmer_sets = [generate_kmers(s.as_u8(), self.word_size) for s in self.sequences]
        self.stats = dict(above_threshold=0, below_threshold=0, aligned=0)
        self._ctx = ctx or default_context()
        res, off = pack(self.sequences)
        self._ctx.load_sequences(self.SET_ALL, res, off)
        self._lens = lens

    # sets of the context this driver uses: all sequences, the gathered representatives, the candidate
    SET_ALL, SET_REPS, SET_CAND = 6, 7, 5

    def _identical(self, reps, cand):
        """n_identical of GlobalAligner(BLOSUM62, -11, -1) for query = each representative (rows) against
        template = the candidate (columns), :296-300, on the forward kernels."""
        ctx = self._ctx
        ctx.set_scoring("BLOSUM62", -11, -1)        # :223,300 -- the context may have been used by others since
        ctx.gather_sequences(self.SET_ALL, self.SET_REPS, reps)
        ctx.gather_sequences(self.SET_ALL, self.SET_CAND, [cand])
        _, nid = ctx.align_all_pairs(self.SET_REPS, self.SET_CAND, want_score=False)
        return nid

    def run(self):
        """bucket_clustering.rs:162-169"""
        return self.merge([], [Cluster(i) for i in self.sequence_order])

    def run_n(self, n_threads):
        """bucket_clustering.rs:171-204: the same chunk / pairwise-merge tree, evaluated in order."""
        if n_threads == 1:
            return self.run()
        singles = [Cluster(i) for i in self.sequence_order]
        n_threads = min(max(n_threads, 1), len(singles))
        chunk = -(-len(singles) // n_threads)
        clusterings = [self.merge([], singles[b:b + chunk]) for b in range(0, len(singles), chunk)]
        while len(clusterings) > 1:
            nxt = []
            for b in range(0, len(clusterings), 2):
                pair = clusterings[b:b + 2]
                nxt.append(self.merge([c.clone() for c in pair[0]], [c.clone() for c in pair[1]])
                           if len(pair) == 2 else pair[0])
            clusterings = nxt
        return clusterings.pop()

    def _verdict(self, rep, cand):
        """The k-mer part of sequence_identity (:272-292): +1 above, -1 below, 0 inconclusive."""
        shared = count_intersection_sorted(self.kmer_sets[cand], self.kmer_sets[rep])
        different = max(len(self.kmer_sets[cand]) - shared, 0)
        shorter = int(min(self._lens[cand], self._lens[rep]))
        lower, upper = kmer_identity_bounds(different, self.word_size, shorter)
        if lower >= self.id_level:
            return 1
        if upper < self.id_level:
            return -1
        return 0

    def merge(self, clusters1, clusters2):
        """bucket_clustering.rs:209-270"""
        for b in clusters2:
            cand = b.representative
            assigned = False
            pos = 0
            while pos < len(clusters1) and not assigned:
                # k-mer verdicts up to the first certain hit; inconclusive ones are collected
                pending = []
                hit = None
                while pos < len(clusters1) and len(pending) < self.BLOCK:
                    v = self._verdict(clusters1[pos].representative, cand)
                    if v == 1:
                        hit = pos
                        pos += 1
                        break
                    if v == 0:
                        pending.append(pos)
                    else:
                        self.stats["below_threshold"] += 1
                    pos += 1
                if pending:
                    reps = [clusters1[p].representative for p in pending]
                    # query = representative (rows), template = candidate (columns): :296-300
                    nid = self._identical(reps, cand)
                    self.stats["aligned"] += len(reps)
                    for p, n_identical in zip(pending, nid):
                        shorter = np.float32(min(self._lens[cand], self._lens[clusters1[p].representative]))
                        if np.float32(n_identical) / shorter >= self.id_level:         # :306-307, :245
                            clusters1[p].members.extend(b.members)
                            assigned = True
                            break
                if not assigned and hit is not None:
                    self.stats["above_threshold"] += 1
                    clusters1[hit].members.extend(b.members)
                    assigned = True
            if not assigned:
                clusters1.append(b)
        return clusters1


def bucket_clustering(sequences, id_level, ctx=None):
    """bucket_clustering.rs:33-41 -> list of clusters, each a list of Sequence."""
    bc = BucketClustering(sequences, id_level, ctx)
    return [[sequences[i] for i in c.members] for c in bc.run()]


def bucket_clustering_n(sequences, id_level, n_threads, ctx=None):
    """bucket_clustering.rs:60-67"""
    bc = BucketClustering(sequences, id_level, ctx)
    return [[sequences[i] for i in c.members] for c in bc.run_n(n_threads)]
