"""The C++ host mirror (bioshell_b200/host/bioshell_seq.hpp): the reference's own aligner test
transcribed to C++ runs against the GPU through the C ABI; without a GPU the same program must
fail loudly (no CPU fallback)."""
import os
import subprocess

import pytest

from bioshell_b200 import _lib
from bioshell_b200.scoring import ncbi_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bioshell_b200", "host", "test_host_mirror")


def _run(tmp_path):
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    f = tmp_path / "BLOSUM62"
    f.write_text(ncbi_text("BLOSUM62"))
    return subprocess.run([EXE, str(f)], capture_output=True, text=True, timeout=300)


@pytest.mark.gpu
def test_cpp_host_mirror_reference_kats_on_gpu(tmp_path):
    r = _run(tmp_path)
    assert r.returncode == 0, r.stderr
    assert "host mirror ok" in r.stdout


def test_cpp_host_mirror_fails_loudly_without_gpu(tmp_path):
    if _lib.lib().bsa_device_count() > 0:
        pytest.skip("a GPU is present")
    r = _run(tmp_path)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
