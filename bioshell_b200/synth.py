"""Deterministic synthetic protein sets -- ctypes binding of csrc/synth.c.
Benchmark/test input only; see the C file for the generator's definition."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbsa_synth.so")
_lib = None

# BASELINE.json configs (seeds from SURVEY.md 8d)
CONFIGS = {
    "cfg1": dict(n=1000, dist=0, lo=50, hi=500, seed=1001),
    "cfg2": dict(n=10000, dist=1, lo=30, hi=4000, mu=5.45, sigma=0.65, seed=1002),
    "cfg3": dict(n=100000, dist=1, lo=30, hi=4000, mu=5.45, sigma=0.65, seed=1003),
    "cfg4q": dict(n=1000, dist=1, lo=30, hi=4000, mu=5.45, sigma=0.65, seed=1004),
    "cfg4db": dict(n=1000000, dist=1, lo=30, hi=4000, mu=5.45, sigma=0.65, seed=1005),
    # cfg5 is a PAIR set (see pair_set): 16 pairs, sequence 2p = query, 2p+1 = template of pair p,
    # lengths U{5000..35000}, even pairs homologous, pair 0 forced to 34,350 x 35,000 residues
    "cfg5": dict(pairs=16, lo=5000, hi=35000, fixed_q=34350, fixed_t=35000, seed=1006),
}


def build():
    src = os.path.join(_HERE, "csrc", "synth.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.bsa_synth_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                            C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        _lib.bsa_synth_generate.restype = C.c_uint64
        _lib.bsa_synth_pair_set.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.bsa_synth_pair_set.restype = C.c_uint64
    return _lib


def generate(n, seed, dist=1, lo=30, hi=4000, mu=5.45, sigma=0.65, homolog_fraction=0.25):
    """-> (residues uint8[total], offsets uint64[n+1])"""
    L = _load()
    off = np.zeros(n + 1, np.uint64)
    total = L.bsa_synth_generate(seed, n, dist, lo, hi, mu, sigma, homolog_fraction, None,
                                 off.ctypes.data_as(C.c_void_p))
    res = np.zeros(max(int(total), 1), np.uint8)
    L.bsa_synth_generate(seed, n, dist, lo, hi, mu, sigma, homolog_fraction,
                         res.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p))
    return res[:int(total)], off


def pair_set(pairs, seed, lo=5000, hi=35000, fixed_q=0, fixed_t=0):
    """BASELINE configs[4]: `pairs` (query, template) pairs as 2*pairs sequences -> (residues, offsets).
    Pair p is (sequence 2p, sequence 2p+1); even pairs are homologous, pair 0 has the fixed lengths."""
    L = _load()
    off = np.zeros(2 * pairs + 1, np.uint64)
    total = L.bsa_synth_pair_set(seed, pairs, lo, hi, fixed_q, fixed_t, None, off.ctypes.data_as(C.c_void_p))
    res = np.zeros(max(int(total), 1), np.uint8)
    L.bsa_synth_pair_set(seed, pairs, lo, hi, fixed_q, fixed_t, res.ctypes.data_as(C.c_void_p),
                         off.ctypes.data_as(C.c_void_p))
    return res[:int(total)], off


def config(name, n=None):
    c = dict(CONFIGS[name])
    if "pairs" in c:
        if n is not None:
            c["pairs"] = n
        return pair_set(**c)
    if n is not None:
        c["n"] = n
    return generate(**c)


def descriptions(n):
    return ["syn|%07d" % i for i in range(n)]
