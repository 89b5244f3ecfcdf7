"""Host mirror of bioshell-seq's alignment module over the C ABI.

Same names and argument meaning as the reference so call sites (and tests) read
alike:
  align_all_pairs          bioshell-seq/src/alignment/alignment_protocols.rs:83-115
  AlignmentReporter        bioshell-seq/src/alignment/alignment_reporter.rs:7-9
  AlignmentStatistics      bioshell-seq/src/alignment/alignment_statistics.rs:28-81
  AlignmentStep / AlignmentPath  bioshell-seq/src/alignment/alignment_path.rs:7-115
  aligned_strings/_sequences  bioshell-seq/src/alignment/alignment_path.rs:161-204
  SequenceIdentityMatrix   bin/cluster_sequences.rs:77-130
plus the new batched entry points the north star asks for (`align_all_vs_all`,
`align_one_vs_many`) which never materialise alignment strings.

All alignment arithmetic happens in libbioshell_align.so on the GPU.  Nothing here
falls back to a CPU aligner.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from .scoring import SubstitutionMatrix
from .sequence import Sequence, count_identical, len_ungapped, pack, ungapped_lengths


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One `bsa_ctx`, used by one host thread at a time.  `device` is a CUDA device index
    (bsa_create) or a list of them / "all" (bsa_create_multi: one process, every GPU behind the
    same calls, results streamed tile-wise into the caller's host buffers)."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        if isinstance(device, (list, tuple)) or device == "all":
            ids = [] if device == "all" else [int(d) for d in device]
            arr = (C.c_int * max(len(ids), 1))(*ids)
            self._h = self._L.bsa_create_multi(arr, len(ids))
            self.device = ids[0] if ids else 0
        else:
            self._h = self._L.bsa_create(int(device))
            self.device = int(device)
        if not self._h:
            raise _lib.BsaError(-4, self._L.bsa_last_error(None).decode())
        self.n_devices = self._L.bsa_context_devices(self._h)
        self._sets = {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.bsa_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc:
            raise _lib.BsaError(rc, self._L.bsa_last_error(self._h).decode())

    # ---- scoring / data ----
    def set_scoring(self, matrix, gap_open, gap_extend):
        if isinstance(matrix, str):
            matrix = SubstitutionMatrix.load(matrix)
        self._matrix = matrix
        self._ck(self._L.bsa_set_scoring(self._h, _p(matrix.score), _p(matrix.aa_indexes),
                                         int(gap_open), int(gap_extend)))

    def load_sequences(self, set_id, residues, offsets):
        res = np.ascontiguousarray(residues, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        self._ck(self._L.bsa_load_sequences(self._h, int(set_id), _p(res) if res.size else None,
                                            _p(off), len(off) - 1))
        self._sets[int(set_id)] = (len(off) - 1, np.diff(off.astype(np.int64)))

    def gather_sequences(self, src_set, dst_set, idx):
        """bsa_gather_sequences: sequence idx[i] of `src_set` becomes sequence i of `dst_set`, on the device."""
        ix = np.ascontiguousarray(idx, np.uint32)
        self._ck(self._L.bsa_gather_sequences(self._h, int(src_set), int(dst_set), _p(ix), len(ix)))
        self._sets[int(dst_set)] = (len(ix), self._sets[int(src_set)][1][ix.astype(np.int64)])

    def n_sequences(self, set_id):
        return self._sets[int(set_id)][0]

    # ---- alignment ----
    def align_all_pairs(self, q_set, t_set, q_counts=None, t_begin=0, t_end=None, want_score=True,
                        want_identical=True, scores=None, n_identical=None, device_out=False):
        """bsa_align_all_pairs.  Returns (scores int32[k], n_identical uint32[k]) in t-major
        report order.  `scores`/`n_identical` may be preallocated numpy arrays (e.g. views of
        pinned memory) or, with device_out=True, raw device pointers (ints)."""
        nT = self._sets[int(t_set)][0]
        nQ = self._sets[int(q_set)][0]
        if t_end is None:
            t_end = nT
        qc = None
        if q_counts is not None:
            qc = np.ascontiguousarray(q_counts, np.uint32)
            if len(qc) != nT:
                raise ValueError("q_counts needs one entry per template")
            n_res = int(qc[t_begin:t_end].astype(np.int64).sum())
        else:
            n_res = (t_end - t_begin) * nQ
        flags = (_lib.WANT_SCORE if want_score else 0) | (_lib.WANT_IDENTICAL if want_identical else 0)
        if device_out:
            flags |= _lib.OUT_DEVICE
            sp = C.c_void_p(scores) if (want_score and scores) else None
            ip = C.c_void_p(n_identical) if (want_identical and n_identical) else None
        else:
            if want_score and scores is None:
                scores = np.empty(n_res, np.int32)
            if want_identical and n_identical is None:
                n_identical = np.empty(n_res, np.uint32)
            sp = _p(scores) if want_score else None
            ip = _p(n_identical) if want_identical else None
        nr = C.c_uint64()
        self._ck(self._L.bsa_align_all_pairs(self._h, int(q_set), int(t_set), _p(qc), int(t_begin),
                                             int(t_end), flags, sp, ip, C.byref(nr)))
        assert nr.value == n_res
        return scores, n_identical

    def all_vs_all(self, set_id, want_score=True, want_identical=True):
        n = self._sets[int(set_id)][0]
        return self.align_all_pairs(set_id, set_id, np.arange(n, dtype=np.uint32),
                                    want_score=want_score, want_identical=want_identical)

    def one_vs_many(self, q_set, db_set, want_score=True, want_identical=False):
        return self.align_all_pairs(q_set, db_set, None, want_score=want_score,
                                    want_identical=want_identical)

    def align_pairs_paths(self, q_set, t_set, q_idx, t_idx, want_paths=True):
        """bsa_align_pairs_paths -> (scores, n_identical, [path bytes per pair])."""
        qi = np.ascontiguousarray(q_idx, np.uint32)
        ti = np.ascontiguousarray(t_idx, np.uint32)
        n = len(qi)
        lq = self._sets[int(q_set)][1][qi.astype(np.int64)] if n else np.zeros(0, np.int64)
        lt = self._sets[int(t_set)][1][ti.astype(np.int64)] if n else np.zeros(0, np.int64)
        scores = np.zeros(n, np.int32)
        nid = np.zeros(n, np.uint32)
        poff = np.zeros(n + 1, np.uint64)
        buf = np.zeros(int((lq + lt).sum()) + 1, np.uint8) if want_paths else None
        self._ck(self._L.bsa_align_pairs_paths(self._h, int(q_set), int(t_set), _p(qi), _p(ti), n,
                                               _p(scores), _p(nid), _p(buf), _p(poff)))
        paths = None
        if want_paths:
            raw = buf.tobytes()
            paths = [raw[int(poff[i]):int(poff[i + 1])] for i in range(n)]
        return scores, nid, paths

    def local_align_pairs(self, q_set, t_set, q_idx, t_idx, want_paths=True):
        """bsa_local_align_pairs -> dict(score, end_q, end_t, start_q, start_t, paths)."""
        qi = np.ascontiguousarray(q_idx, np.uint32)
        ti = np.ascontiguousarray(t_idx, np.uint32)
        n = len(qi)
        lq = self._sets[int(q_set)][1][qi.astype(np.int64)] if n else np.zeros(0, np.int64)
        lt = self._sets[int(t_set)][1][ti.astype(np.int64)] if n else np.zeros(0, np.int64)
        out = {k: np.zeros(n, np.uint32) for k in ("end_q", "end_t", "start_q", "start_t")}
        out["score"] = np.zeros(n, np.int32)
        poff = np.zeros(n + 1, np.uint64)
        buf = np.zeros(int((lq + lt).sum()) + 1, np.uint8) if want_paths else None
        self._ck(self._L.bsa_local_align_pairs(self._h, int(q_set), int(t_set), _p(qi), _p(ti), n, _p(out["score"]),
                                               _p(out["end_q"]), _p(out["end_t"]), _p(out["start_q"]),
                                               _p(out["start_t"]), _p(buf), _p(poff)))
        if want_paths:
            raw = buf.tobytes()
            out["paths"] = [raw[int(poff[i]):int(poff[i + 1])] for i in range(n)]
        return out

    def plan_shards(self, q_set, t_set, q_counts, n_shards):
        qc = None if q_counts is None else np.ascontiguousarray(q_counts, np.uint32)
        b = np.zeros(n_shards + 1, np.uint32)
        self._ck(self._L.bsa_plan_shards(self._h, int(q_set), int(t_set), _p(qc), int(n_shards), _p(b)))
        return b

    def stats(self):
        st = _lib.Stats()
        self._ck(self._L.bsa_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def measure_int_peak(self, which=0):
        ops, mhz = C.c_double(), C.c_double()
        self._ck(self._L.bsa_measure_int_peak(self._h, int(which), C.byref(ops), C.byref(mhz)))
        return ops.value, mhz.value


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---------------------------------------------------------------------------
# alignment_path.rs
# ---------------------------------------------------------------------------
class AlignmentStep(enum.Enum):
    """alignment_path.rs:7-32: `Horizontal` consumes a template residue (gap in the query, printed
    '-'), `Vertical` a query residue (gap in the template, '|'), `Match` both ('*')."""
    Horizontal = "-"
    Vertical = "|"
    Match = "*"

    def __str__(self):                                    # alignment_path.rs:34-47
        return self.value

    @classmethod
    def try_from(cls, value):                             # alignment_path.rs:50-62
        c = chr(value) if isinstance(value, int) else value
        try:
            return cls(c)
        except ValueError:
            raise ValueError("Invalid value for AlignmentStep")


class AlignmentPath(str):
    """alignment_path.rs:75-115.  A `str` of step glyphs -- what the GPU traceback writes and what
    `AlignmentPath::to_string()` prints -- with the reference's constructors and iterator."""

    @classmethod
    def try_from(cls, s):                                 # alignment_path.rs:88-104
        s = s.decode("latin-1") if isinstance(s, (bytes, bytearray)) else s
        for c in s:
            AlignmentStep.try_from(c)
        return cls(s)

    @classmethod
    def from_attrs(cls, path):                            # alignment_path.rs:81
        return cls("".join(str(AlignmentStep(st) if not isinstance(st, AlignmentStep) else st) for st in path))

    def iter(self):                                       # alignment_path.rs:84
        return (AlignmentStep(c) for c in self)

    def to_string(self):
        return str(self)


def aligned_symbols(path, query, template, gap_symbol=ord("-")):
    """alignment_path.rs:117-139: '-' takes a template symbol, '|' a query symbol, '*' both."""
    p = np.frombuffer(path.encode() if isinstance(path, str) else bytes(path), np.uint8)
    q = np.frombuffer(bytes(query), np.uint8)
    t = np.frombuffer(bytes(template), np.uint8)
    takes_q = p != ord("-")
    takes_t = p != ord("|")
    if takes_q.sum() > len(q) or takes_t.sum() > len(t):
        raise IndexError("called `Option::unwrap()` on a `None` value")   # iterator exhausted
    aq = np.full(len(p), gap_symbol, np.uint8)
    at = np.full(len(p), gap_symbol, np.uint8)
    aq[takes_q] = q[:int(takes_q.sum())]
    at[takes_t] = t[:int(takes_t.sum())]
    return aq.tobytes(), at.tobytes()


def aligned_strings(path, query, template, gap_symbol="-"):
    """alignment_path.rs:161-167"""
    aq, at = aligned_symbols(path, query.encode(), template.encode(), ord(gap_symbol))
    return aq.decode("latin-1"), at.decode("latin-1")


def aligned_sequences(path, query, template, gap_symbol="-"):
    """alignment_path.rs:199-204: results inherit the descriptions."""
    aq, at = aligned_symbols(path, query.as_u8(), template.as_u8(), ord(gap_symbol))
    return Sequence(query.description(), aq), Sequence(template.description(), at)


# ---------------------------------------------------------------------------
# alignment_statistics.rs / alignment_reporter.rs / cluster_sequences.rs
# ---------------------------------------------------------------------------
def rust_fixed(x, width, prec):
    """`{:width.prec}` of a Rust float: non-finite values print as NaN / inf / -inf (Python: nan)."""
    x = float(x)
    if x != x:
        return "NaN".rjust(width)
    if x in (float("inf"), float("-inf")):
        return ("inf" if x > 0 else "-inf").rjust(width)
    return "%*.*f" % (width, prec, x)


class AlignmentStatistics:
    """alignment_statistics.rs:28-81.  `label_style` is a `sequence_id.LabelStyle`
    (sequence_label.rs:33-63); None keeps the raw descriptions (the batched path never builds
    labels per pair -- the reference runs ~17 regex scans per reported pair here)."""

    def __init__(self, query_header, template_header, n_identical, query_length, template_length):
        self.query_header, self.template_header = query_header, template_header
        self.n_identical, self.query_length, self.template_length = n_identical, query_length, template_length

    @classmethod
    def from_sequences(cls, aligned_query, aligned_template, label_style=None):
        qh, th = aligned_query.description(), aligned_template.description()
        if label_style is not None:
            from .sequence_id import sequence_label
            qh, th = sequence_label(qh, label_style), sequence_label(th, label_style)
        return cls(qh, th, count_identical(aligned_query, aligned_template),
                   len_ungapped(aligned_query), len_ungapped(aligned_template))

    @classmethod
    def from_strings(cls, query_name, query_sequence, template_name, template_sequence, label_style=None):
        return cls.from_sequences(Sequence(query_name, query_sequence),
                                  Sequence(template_name, template_sequence), label_style)

    def percent_identity(self):
        """alignment_statistics.rs:71-73 (f64; 0/0 -> NaN as in Rust)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            return float(np.float64(self.n_identical) /
                         np.float64(min(self.query_length, self.template_length)) * 100.0)

    def __str__(self):
        """alignment_statistics.rs:76-80"""
        return "%s %s %s %% %3d %4d %4d" % (self.query_header, self.template_header,
                                            rust_fixed(self.percent_identity(), 6, 2), self.n_identical,
                                            self.query_length, self.template_length)


class AlignmentReporter:
    """alignment_reporter.rs:7-9"""

    def report(self, aligned_query, aligned_template):
        raise NotImplementedError


class MultiReporter(AlignmentReporter):
    """alignment_reporter.rs:12-29"""

    def __init__(self):
        self.reporters = []

    def add_reporter(self, reporter):
        self.reporters.append(reporter)

    def count_reporters(self):
        return len(self.reporters)

    def report(self, aligned_query, aligned_template):
        for r in self.reporters:
            r.report(aligned_query, aligned_template)


class CollectReporter(AlignmentReporter):
    """Keeps every reported pair (test helper; no reference counterpart)."""

    def __init__(self):
        self.pairs = []

    def report(self, aligned_query, aligned_template):
        self.pairs.append((aligned_query, aligned_template))


class SequenceIdentityMatrix(AlignmentReporter):
    """bin/cluster_sequences.rs:77-130: `similarity_matrix[q][t] = percent_identity as f32`;
    only the entries the reporter is called for are written (upper triangle for the
    triangle protocol -- the reference never mirrors them, SURVEY.md 3.1 note)."""

    def __init__(self, sequences, name_width=0):
        self.description_to_index = {}
        for i, s in enumerate(sequences):
            if s.description() in self.description_to_index:
                raise ValueError("IdenticalSequenceDescriptions: %s" % s.description())
            self.description_to_index[s.description()] = i
        self.n_sequences = len(sequences)
        self.similarity_matrix = np.zeros((self.n_sequences, self.n_sequences), np.float32)

    def percent_identity(self, i, j):
        return self.similarity_matrix[i, j]

    def report(self, aligned_query, aligned_template):
        st = AlignmentStatistics.from_sequences(aligned_query, aligned_template)
        qi = self.description_to_index[aligned_query.description()]
        ti = self.description_to_index[aligned_template.description()]
        self.similarity_matrix[qi, ti] = np.float32(st.percent_identity())

    def fill_from(self, results):
        """Batched equivalent of replaying every pair through `report`: same matrix, no strings."""
        q, t = results.pair_indices()
        self.similarity_matrix[q, t] = results.percent_identity()


# ---------------------------------------------------------------------------
# the protocol
# ---------------------------------------------------------------------------
def triangle_counts(queries, templates, if_triangle_only):
    """How many queries the inner loop of alignment_protocols.rs:96-97 visits for each
    template before `if if_triangle_only && template == query { break }`."""
    nq = len(queries)
    if not if_triangle_only:
        return np.full(len(templates), nq, np.uint32)
    first = {}
    for i, q in enumerate(queries):
        first.setdefault((q.description(), q.as_u8()), i)
    return np.array([first.get((t.description(), t.as_u8()), nq) for t in templates], np.uint32)


class PairResults:
    """Scores / identical counts of a batched run, in the reference's t-major report order."""

    def __init__(self, scores, n_identical, q_counts, len_q_ungapped, len_t_ungapped):
        self.scores, self.n_identical = scores, n_identical
        self.q_counts = np.asarray(q_counts, np.int64)
        self.first = np.concatenate([[0], np.cumsum(self.q_counts)])
        self._lq, self._lt = len_q_ungapped, len_t_ungapped

    def __len__(self):
        return int(self.first[-1])

    def pair_indices(self):
        t = np.repeat(np.arange(len(self.q_counts)), self.q_counts)
        q = np.arange(len(t)) - self.first[t]
        return q, t

    def index(self, q, t):
        if q >= self.q_counts[t]:
            raise KeyError("pair (%d,%d) was not aligned" % (q, t))
        return int(self.first[t] + q)

    def percent_identity(self):
        """alignment_statistics.rs:71-73 in f64, then `as f32` (cluster_sequences.rs:128)."""
        if self.n_identical is None:
            raise ValueError("n_identical was not requested for this run (want_identical=False)")
        q, t = self.pair_indices()
        mn = np.minimum(self._lq[q], self._lt[t]).astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (self.n_identical.astype(np.float64) / mn * 100.0).astype(np.float32)


def _prepare(ctx, queries, templates, matrix, gap_open, gap_extend):
    ctx = ctx or default_context()
    if len(queries) == 0 or len(templates) == 0:
        # alignment_protocols.rs:86-87: `.max().unwrap()` on an empty iterator panics
        raise ValueError("called `Option::unwrap()` on a `None` value (empty sequence set)")
    ctx.set_scoring(matrix, gap_open, gap_extend)
    qres, qoff = pack(queries)
    same = templates is queries
    ctx.load_sequences(0, qres, qoff)
    if same:
        tres, toff = qres, qoff
        t_set = 0
    else:
        tres, toff = pack(templates)
        ctx.load_sequences(1, tres, toff)
        t_set = 1
    return ctx, t_set, ungapped_lengths(qres, qoff), ungapped_lengths(tres, toff)


def align_all_vs_all(sequences, matrix, gap_open, gap_extend, ctx=None):
    """New batched entry point: strict upper triangle of one set, scores + identical counts."""
    return align_pairs_batched(sequences, sequences, matrix, gap_open, gap_extend, True, ctx)


def align_one_vs_many(queries, database, matrix, gap_open, gap_extend, ctx=None, want_identical=False):
    return align_pairs_batched(queries, database, matrix, gap_open, gap_extend, False, ctx,
                               want_identical=want_identical)


def align_pairs_batched(queries, templates, matrix, gap_open, gap_extend, if_triangle_only, ctx=None,
                        want_identical=True):
    ctx, t_set, lq, lt = _prepare(ctx, queries, templates, matrix, gap_open, gap_extend)
    counts = triangle_counts(queries, templates, if_triangle_only)
    scores, nid = ctx.align_all_pairs(0, t_set, counts, want_identical=want_identical)
    return PairResults(scores, nid, counts, lq, lt)      # nid stays None when it was not requested


def align_all_pairs(queries, templates, matrix, gap_open, gap_extend, if_triangle_only, reporter,
                    ctx=None, chunk_pairs=4096):
    """alignment_protocols.rs:83-115 with the reference's signature: every aligned pair is
    replayed into `reporter.report(aligned_query, aligned_template)` in the reference's
    template-major order.  Alignment paths come from the GPU traceback kernel."""
    ctx, t_set, _, _ = _prepare(ctx, queries, templates, matrix, gap_open, gap_extend)
    counts = triangle_counts(queries, templates, if_triangle_only).astype(np.int64)
    t_all = np.repeat(np.arange(len(templates)), counts)
    first = np.concatenate([[0], np.cumsum(counts)])
    q_all = np.arange(len(t_all)) - first[t_all]
    for b in range(0, len(t_all), chunk_pairs):
        qs, ts = q_all[b:b + chunk_pairs], t_all[b:b + chunk_pairs]
        _, _, paths = ctx.align_pairs_paths(0, t_set, qs, ts)
        for q, t, path in zip(qs, ts, paths):
            aq, at = aligned_sequences(path, queries[int(q)], templates[int(t)], "-")
            reporter.report(aq, at)
    return len(t_all)


class LocalAlignment:
    """Single-pair convenience with the reference's method names (local.rs:18-284)."""

    def __init__(self, max_seq_length, ctx=None):
        self.max_length = max_seq_length + 1
        self._ctx = ctx or default_context()
        self._r = None

    def align(self, query, template, matrix, gap_open, gap_extend):
        q = query.encode() if isinstance(query, str) else bytes(query)
        t = template.encode() if isinstance(template, str) else bytes(template)
        if len(q) >= self.max_length or len(t) >= self.max_length:
            raise IndexError("index out of bounds: sequence longer than the aligner capacity")
        self._ctx.set_scoring(matrix, gap_open, gap_extend)
        res, off = pack([q, t])
        self._ctx.load_sequences(7, res, off)
        self._r = self._ctx.local_align_pairs(7, 7, [0], [1])
        return int(self._r["score"][0])

    def backtrace(self):
        """-> (path, query_start, template_start)  (local.rs:213-273)"""
        return AlignmentPath(self._r["paths"][0].decode()), int(self._r["start_q"][0]), int(self._r["start_t"][0])

    def recent_score(self):
        return int(self._r["score"][0])

    def recent_end_point(self):
        return int(self._r["end_q"][0]), int(self._r["end_t"][0])


class GlobalAligner:
    """Single-pair convenience with the reference's method names (global.rs:17-203); each
    `align` is a one-pair batch on the GPU."""

    def __init__(self, max_seq_length, ctx=None):
        self.max_length = max_seq_length + 1
        self._ctx = ctx or default_context()
        self._score = 0
        self._path = b""

    def align(self, query, template, matrix, gap_open, gap_extend):
        q = query.encode() if isinstance(query, str) else bytes(query)
        t = template.encode() if isinstance(template, str) else bytes(template)
        if len(q) >= self.max_length or len(t) >= self.max_length:
            raise IndexError("index out of bounds: sequence longer than the aligner capacity")
        self._ctx.set_scoring(matrix, gap_open, gap_extend)
        res, off = pack([q, t])
        self._ctx.load_sequences(7, res, off)
        s, _, p = self._ctx.align_pairs_paths(7, 7, [0], [1])
        self._score, self._path = int(s[0]), p[0]
        return self._score

    def backtrace(self):
        return AlignmentPath(self._path.decode())

    def recent_score(self):
        return self._score
