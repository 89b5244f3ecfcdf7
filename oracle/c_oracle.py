"""ctypes binding of the C oracle (oracle/bioshell_oracle.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libbioshell_oracle.so")
_lib = None

ERR = {0: "ok", -1: "IncorrectNCBIFormat", -2: "CantParseNCBIEntry", -3: "reference would panic",
       -4: "alloc", -5: "sequence longer than aligner capacity"}


class OracleError(RuntimeError):
    def __init__(self, rc):
        super().__init__("oracle rc=%d (%s)" % (rc, ERR.get(rc, "?")))
        self.rc = rc


class _SeqSet(C.Structure):
    _fields_ = [("res", C.c_void_p), ("res_off", C.c_void_p), ("desc", C.c_void_p),
                ("desc_off", C.c_void_p), ("n", C.c_uint32)]


def build(force=False):
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("bioshell_oracle.c", "hclust_oracle.c", "local_oracle.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def use_native():
    """bench.py only: rebuild the oracle with -O3 -march=native for THIS host's CPU (BASELINE.md section 3
    promises that for the timed CPU baseline) into oracle/_ref/libbioshell_oracle_native.so and bind
    that from now on.  The portable build stays what the tests use (it travels to the GPU box, whose
    CPU may differ from the build container's).  Falls back to the portable build if gcc fails."""
    global _SO, _lib
    native = os.path.join(_HERE, "_ref", "libbioshell_oracle_native.so")
    if _SO == native:
        return True
    try:
        os.makedirs(os.path.dirname(native), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fPIC", "-std=c11", "-shared", "-o", native] +
                              [os.path.join(_HERE, f) for f in ("bioshell_oracle.c", "hclust_oracle.c", "local_oracle.c")] +
                              ["-lpthread"], stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        return False
    _SO, _lib = native, None
    return True


def lib():
    global _lib
    if _lib is None:
        if not _SO.endswith("_native.so"):
            build()
        L = C.CDLL(_SO)
        L.orc_parse_ncbi.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_parse_ncbi.restype = C.c_int
        L.orc_align_pair.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                     C.c_void_p, C.c_int32, C.c_int32, C.c_size_t,
                                     C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int64),
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_align_pair.restype = C.c_int
        L.orc_all_pairs_layout.argtypes = [C.POINTER(_SeqSet), C.POINTER(_SeqSet), C.c_int,
                                           C.c_void_p]
        L.orc_all_pairs_layout.restype = C.c_uint64
        L.orc_align_all_pairs.argtypes = [C.POINTER(_SeqSet), C.POINTER(_SeqSet), C.c_void_p,
                                          C.c_void_p, C.c_int32, C.c_int32, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p] + [C.c_void_p] * 7 + \
                                         [C.POINTER(C.c_double)]
        L.orc_align_all_pairs.restype = C.c_int
        L.orc_align_pair_list.argtypes = [C.POINTER(_SeqSet), C.POINTER(_SeqSet), C.c_void_p,
                                          C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_align_pair_list.restype = C.c_int
        L.orc_percent_identity.argtypes = [C.c_uint64] * 3
        L.orc_percent_identity.restype = C.c_double
        L.orc_percent_identity_f32.argtypes = [C.c_uint64] * 3
        L.orc_percent_identity_f32.restype = C.c_float
        L.orc_expand_and_count.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p,
                                           C.c_void_p, C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_expand_and_count.restype = C.c_int
        L.orc_hclust.argtypes = [C.c_uint32, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.orc_hclust.restype = C.c_int64
        L.orc_local_align.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_int32)] + [C.POINTER(C.c_uint64)] * 4 + [C.c_void_p]
        L.orc_local_align.restype = C.c_int64
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def parse_ncbi(text):
    """-> (score[441] int32, aa_index[256] uint8) per substitution_matrix.rs:96-135."""
    if isinstance(text, str):
        text = text.encode()
    score = np.zeros(441, np.int32)
    idx = np.zeros(256, np.uint8)
    rc = lib().orc_parse_ncbi(text, len(text), _ptr(score), _ptr(idx))
    if rc:
        raise OracleError(rc)
    return score, idx


def align_pair(q, t, score, aa_index, go, ge, lmax=None):
    """One pair on raw bytes.  Returns dict(score, path, aligned_q, aligned_t, n_identical,
    len_q, len_t, identity)."""
    q = bytes(q)
    t = bytes(t)
    n, m = len(q), len(t)
    if lmax is None:
        lmax = max(n, m)
    qa = np.frombuffer(q, np.uint8) if n else np.zeros(1, np.uint8)
    ta = np.frombuffer(t, np.uint8) if m else np.zeros(1, np.uint8)
    path = np.zeros(n + m + 1, np.uint8)
    aq = np.zeros(n + m + 1, np.uint8)
    at = np.zeros(n + m + 1, np.uint8)
    sc = C.c_int32()
    pl = C.c_int64()
    nid, lq, lt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib().orc_align_pair(_ptr(qa), n, _ptr(ta), m, _ptr(score), _ptr(aa_index), go, ge,
                              lmax, C.byref(sc), _ptr(path), C.byref(pl), _ptr(aq), _ptr(at),
                              C.byref(nid), C.byref(lq), C.byref(lt))
    if rc:
        raise OracleError(rc)
    L = pl.value
    return dict(score=sc.value, path=path[:L].tobytes().decode(), aligned_q=aq[:L].tobytes(),
                aligned_t=at[:L].tobytes(), n_identical=nid.value, len_q=lq.value, len_t=lt.value,
                identity=lib().orc_percent_identity(nid.value, lq.value, lt.value)
                if min(lq.value, lt.value) else float("nan"))


class SeqSet:
    """Packed Vec<Sequence>: list of (description, residue-bytes)."""

    def __init__(self, seqs, descs=None):
        seqs = [bytes(s) for s in seqs]
        if descs is None:
            descs = [("seq%d" % i).encode() for i in range(len(seqs))]
        descs = [d.encode() if isinstance(d, str) else bytes(d) for d in descs]
        self.n = len(seqs)
        self.res_off = np.zeros(self.n + 1, np.uint64)
        self.res_off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        self.desc_off = np.zeros(self.n + 1, np.uint64)
        self.desc_off[1:] = np.cumsum([len(d) for d in descs], dtype=np.uint64)
        self.res = np.frombuffer(b"".join(seqs) + b"\0", np.uint8).copy()
        self.desc = np.frombuffer(b"".join(descs) + b"\0", np.uint8).copy()
        self.c = _SeqSet(self.res.ctypes.data, self.res_off.ctypes.data, self.desc.ctypes.data,
                         self.desc_off.ctypes.data, self.n)

    @classmethod
    def from_packed(cls, res, off, descs=None):
        self = cls.__new__(cls)
        self.n = len(off) - 1
        self.res = np.ascontiguousarray(res, np.uint8)
        self.res_off = np.ascontiguousarray(off, np.uint64)
        if descs is None:
            # unique fixed-width descriptions "syn|%07d" 
            d = np.char.add("syn|", np.char.zfill(np.arange(self.n).astype(str), 7))
            descs = "".join(d.tolist()).encode()
            self.desc = np.frombuffer(descs + b"\0", np.uint8).copy()
            self.desc_off = (np.arange(self.n + 1, dtype=np.uint64) * np.uint64(11))
        else:
            descs = [x.encode() if isinstance(x, str) else bytes(x) for x in descs]
            self.desc = np.frombuffer(b"".join(descs) + b"\0", np.uint8).copy()
            self.desc_off = np.zeros(self.n + 1, np.uint64)
            self.desc_off[1:] = np.cumsum([len(x) for x in descs], dtype=np.uint64)
        self.c = _SeqSet(self.res.ctypes.data, self.res_off.ctypes.data, self.desc.ctypes.data,
                         self.desc_off.ctypes.data, self.n)
        return self


def align_all_pairs(Q, T, score, aa_index, go, ge, triangle, n_threads=1, clear_mode=1,
                    with_backtrace=True):
    """alignment_protocols.rs:83-115 + the SequenceIdentityMatrix reporter.  Outputs are in
    report (t-major) order."""
    L = lib()
    first = np.zeros(T.n + 1, np.uint64)
    npairs = int(L.orc_all_pairs_layout(C.byref(Q.c), C.byref(T.c), int(triangle), _ptr(first)))
    out = dict(first=first, n_pairs=npairs,
               score=np.zeros(npairs, np.int32), n_identical=np.zeros(npairs, np.uint32),
               len_q=np.zeros(npairs, np.uint32), len_t=np.zeros(npairs, np.uint32),
               identity=np.zeros(npairs, np.float32), q=np.zeros(npairs, np.uint32),
               t=np.zeros(npairs, np.uint32))
    cells = C.c_double()
    rc = L.orc_align_all_pairs(C.byref(Q.c), C.byref(T.c), _ptr(score), _ptr(aa_index), go, ge,
                               int(triangle), clear_mode, int(with_backtrace), n_threads,
                               _ptr(first), _ptr(out["score"]), _ptr(out["n_identical"]),
                               _ptr(out["len_q"]), _ptr(out["len_t"]), _ptr(out["identity"]),
                               _ptr(out["q"]), _ptr(out["t"]), C.byref(cells))
    if rc:
        raise OracleError(rc)
    out["cells"] = cells.value
    return out


def align_pair_list(Q, T, score, aa_index, go, ge, pq, pt, lmax, n_threads=1, clear_mode=1,
                    with_backtrace=True):
    pq = np.ascontiguousarray(pq, np.uint32)
    pt = np.ascontiguousarray(pt, np.uint32)
    n = len(pq)
    sc = np.zeros(n, np.int32)
    nid = np.zeros(n, np.uint32)
    cells = C.c_double()
    rc = lib().orc_align_pair_list(C.byref(Q.c), C.byref(T.c), _ptr(score), _ptr(aa_index), go,
                                   ge, int(lmax), clear_mode, int(with_backtrace), n_threads,
                                   _ptr(pq), _ptr(pt), n, _ptr(sc), _ptr(nid), C.byref(cells))
    if rc:
        raise OracleError(rc)
    return dict(score=sc, n_identical=nid, cells=cells.value)


LINKAGE = {"single": 0, "complete": 1, "average": 2, "median": 3, "centroid": 4, "ward": 5}


def hclust(dist, rule):
    """hierarchical_clustering (bioshell-clustering/src/hierarchical/hierarchical.rs:22-80) on a full
    n x n f32 matrix of which only [i][j], i > j is read.  Returns the merge log."""
    d = np.ascontiguousarray(dist, np.float32)
    n = d.shape[0]
    out = dict(mat_i=np.zeros(max(n - 1, 1), np.uint32), mat_j=np.zeros(max(n - 1, 1), np.uint32),
               id_i=np.zeros(max(n - 1, 1), np.uint32), id_j=np.zeros(max(n - 1, 1), np.uint32),
               dist=np.zeros(max(n - 1, 1), np.float32))
    rc = lib().orc_hclust(n, _ptr(d), LINKAGE[rule] if isinstance(rule, str) else int(rule), _ptr(out["mat_i"]),
                          _ptr(out["mat_j"]), _ptr(out["id_i"]), _ptr(out["id_j"]), _ptr(out["dist"]))
    if rc < 0:
        raise OracleError(int(rc))
    return {k: v[:rc] for k, v in out.items()}


def local_align(q, t, score, aa_index, go, ge):
    """LocalAlignment::align + backtrace (bioshell-seq/src/alignment/local.rs:83-273) on raw bytes."""
    q, t = bytes(q), bytes(t)
    n, m = len(q), len(t)
    qi = aa_index[np.frombuffer(q, np.uint8)] if n else np.zeros(1, np.uint8)
    ti = aa_index[np.frombuffer(t, np.uint8)] if m else np.zeros(1, np.uint8)
    qi, ti = np.ascontiguousarray(qi, np.uint8), np.ascontiguousarray(ti, np.uint8)
    path = np.zeros(n + m + 1, np.uint8)
    sc = C.c_int32()
    eq, et, sq, st = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib().orc_local_align(_ptr(qi), n, _ptr(ti), m, _ptr(score), go, ge, C.byref(sc), C.byref(eq), C.byref(et),
                               C.byref(sq), C.byref(st), _ptr(path))
    if rc < 0:
        raise OracleError(int(rc))
    return dict(score=sc.value, path=path[:rc].tobytes().decode(), end_q=eq.value, end_t=et.value,
                start_q=sq.value, start_t=st.value)
