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
