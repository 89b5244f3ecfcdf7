"""ctypes binding of libbioshell_align.so (include/bioshell_align.h).

The library is the product: there is no Python or CPU implementation of the
alignment behind it.  If the shared object is missing this module raises at import
of the first symbol; if no sm_100 GPU is visible, ``Context()`` raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSA_LIB_PATH: A/B experiments against another build of the same library (tools/); the product is the in-tree one
LIB_PATH = os.environ.get("BSA_LIB_PATH") or os.path.join(_HERE, "libbioshell_align.so")

OK = 0
ERRORS = {-1: "BAD_ARG", -2: "UNSUPPORTED_GAPS", -3: "RANGE", -4: "CUDA", -5: "OOM", -6: "FORMAT",
          -7: "ALPHABET", -8: "EMPTY"}
WANT_SCORE, WANT_IDENTICAL, OUT_DEVICE, IN_DEVICE = 1, 2, 4, 8

# every symbol include/bioshell_align.h declares
SYMBOLS = ["bsa_device_count", "bsa_create", "bsa_create_multi", "bsa_context_devices", "bsa_destroy", "bsa_last_error", "bsa_gather_sequences",
           "bsa_parse_ncbi_matrix", "bsa_set_scoring", "bsa_load_sequences", "bsa_align_all_pairs",
           "bsa_all_vs_all", "bsa_one_vs_many", "bsa_plan_shards", "bsa_align_pairs_paths",
           "bsa_host_alloc_pinned", "bsa_host_free_pinned", "bsa_get_stats", "bsa_measure_int_peak",
           "bsa_hclust", "bsa_local_align_pairs"]


class BsaError(RuntimeError):
    def __init__(self, rc, msg=""):
        super().__init__("libbioshell_align: %s (%d)%s" % (ERRORS.get(rc, "?"), rc, ": " + msg if msg else ""))
        self.rc = rc


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("cells", C.c_uint64), ("padded_cells", C.c_uint64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_uint32),
                ("items", C.c_uint32), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("fallback_pairs", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib = None


def lib():
    """Load the shared library (once).  Raises OSError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no fallback implementation." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    L.bsa_device_count.restype = C.c_int
    L.bsa_create.argtypes = [C.c_int]
    L.bsa_create.restype = vp
    L.bsa_create_multi.argtypes = [vp, C.c_int]
    L.bsa_create_multi.restype = vp
    L.bsa_context_devices.argtypes = [vp]
    L.bsa_context_devices.restype = C.c_int
    L.bsa_destroy.argtypes = [vp]
    L.bsa_destroy.restype = None
    L.bsa_last_error.argtypes = [vp]
    L.bsa_last_error.restype = C.c_char_p
    L.bsa_parse_ncbi_matrix.argtypes = [C.c_char_p, C.c_size_t, vp, vp]
    L.bsa_set_scoring.argtypes = [vp, vp, vp, i32, i32]
    L.bsa_load_sequences.argtypes = [vp, C.c_int, vp, vp, u32]
    L.bsa_gather_sequences.argtypes = [vp, C.c_int, C.c_int, vp, u32]
    L.bsa_align_all_pairs.argtypes = [vp, C.c_int, C.c_int, vp, u32, u32, u32, vp, vp, C.POINTER(u64)]
    L.bsa_all_vs_all.argtypes = [vp, C.c_int, u32, vp, vp]
    L.bsa_one_vs_many.argtypes = [vp, C.c_int, C.c_int, u32, vp, vp]
    L.bsa_plan_shards.argtypes = [vp, C.c_int, C.c_int, vp, u32, vp]
    L.bsa_align_pairs_paths.argtypes = [vp, C.c_int, C.c_int, vp, vp, u64, vp, vp, vp, vp]
    L.bsa_hclust.argtypes = [vp, u32, vp, C.c_int, u32, vp, vp, vp]
    L.bsa_local_align_pairs.argtypes = [vp, C.c_int, C.c_int, vp, vp, u64] + [vp] * 7
    L.bsa_host_alloc_pinned.argtypes = [C.c_size_t]
    L.bsa_host_alloc_pinned.restype = vp
    L.bsa_host_free_pinned.argtypes = [vp]
    L.bsa_host_free_pinned.restype = None
    L.bsa_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.bsa_measure_int_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in SYMBOLS:
        getattr(L, name)
    _lib = L
    return L
