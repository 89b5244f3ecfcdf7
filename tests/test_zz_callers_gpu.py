"""GPU end-to-end tests of the callers added late in round 1 (the needleman_wunsh command line, the
C++ clustering mirror).  Their CPU twins (tests/test_reporters.py, tests/test_cpp_host.py) run the
same code with the device call replaced by the oracle; these two run the real thing.  The file
sorts last on purpose: they were written after the round's GPU budget was spent."""
import io
import subprocess

import numpy as np
import pytest

from test_cpp_host import CEXE, _build_if_missing, _clustering_case
from test_reporters import NW_ARGS, _expected_text, _run_cli


@pytest.mark.gpu
def test_needleman_wunsh_cli_on_gpu(ctx, tmp_path, oracle_matrices):
    got = io.StringIO()
    _run_cli(tmp_path, NW_ARGS, got)
    assert got.getvalue() == _expected_text(oracle_matrices)
    # query set against a template set: every pair, self pairs included (if_triangle_only = false)
    t = tmp_path / "t.fasta"
    t.write_text(">t1\nMAVRLLKTHL\n>t2\nMKNITCYL\n")
    got = io.StringIO()
    _run_cli(tmp_path, ["-t", str(t), "--identity"], got)
    assert len(got.getvalue().strip().split("\n")) == 8
    assert np.all([ln.count("%") == 1 for ln in got.getvalue().strip().split("\n")])


@pytest.mark.gpu
def test_cpp_clustering_mirror_on_gpu(tmp_path):
    _build_if_missing()
    for rule in ("single", "complete", "average"):
        path, want = _clustering_case(tmp_path, 70, 5, rule, 7)
        r = subprocess.run([CEXE, str(path), "--gpu"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert [ln.rstrip() for ln in r.stdout.strip().split("\n")] == [w.rstrip() for w in want]


@pytest.mark.gpu
def test_wavefront_short_queries_against_long_templates(ctx, oracle_matrices):
    """K3 with queries far shorter than one hand-off batch (1 .. 100 rows) against templates of
    4,100-6,000 columns: the boundary batches of the wavefront kernel at their edge sizes."""
    from bioshell_b200 import synth
    from oracle import c_oracle
    rng = np.random.default_rng(77)
    long_res, long_off = synth.generate(3, seed=4242, dist=0, lo=4100, hi=6000)
    raw_long = long_res.tobytes()
    shorts = [bytes(rng.choice(list(b"ARNDCQEGHILKMFPSTWYV"), n).astype(np.uint8)) for n in (1, 2, 15, 16, 17, 31, 32, 33, 47, 64, 65, 100)]
    seqs = [raw_long[int(long_off[i]):int(long_off[i + 1])] for i in range(3)] + shorts
    res = np.frombuffer(b"".join(seqs), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    qi = np.array([3 + k for k in range(len(shorts))] * 2)
    ti = np.array([0] * len(shorts) + [2] * len(shorts))
    s, nid, paths = ctx.align_pairs_paths(0, 0, qi, ti)
    M = oracle_matrices["BLOSUM62"]
    for k in range(len(qi)):
        one = c_oracle.align_pair(seqs[qi[k]], seqs[ti[k]], M[0], M[1], -10, -1)
        assert (one["score"], one["n_identical"], one["path"]) == (s[k], nid[k], paths[k].decode()), (qi[k], ti[k])
