"""Substitution matrices -- host mirror of bioshell-seq/src/scoring/.

``SubstitutionMatrixList`` / ``SubstitutionMatrix`` follow
bioshell-seq/src/scoring/substitution_matrix.rs:15-151; the NCBI text is parsed by the
C ABI (``bsa_parse_ncbi_matrix``, same rules as ``ncbi_matrix_from_buffer``, :96-135).
The seven shipped matrices are NCBI data held in data/matrices.json
(tools/gen_matrices.py) and re-rendered to NCBI text on load.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "matrices.json")
_tables = None


class SubstitutionMatrixList:
    """Names of the shipped matrices (substitution_matrix.rs:15-23)."""
    BLOSUM45 = "BLOSUM45"
    BLOSUM80 = "BLOSUM80"
    PAM250 = "PAM250"
    PAM70 = "PAM70"
    BLOSUM62 = "BLOSUM62"
    PAM120 = "PAM120"
    PAM30 = "PAM30"
    ALL = ("BLOSUM45", "BLOSUM62", "BLOSUM80", "PAM30", "PAM70", "PAM120", "PAM250")


def _load_tables():
    global _tables
    if _tables is None:
        with open(_DATA) as fh:
            _tables = json.load(fh)
    return _tables


def ncbi_text(name):
    """Render a shipped matrix as NCBI-format text (comment lines, header row starting
    with a blank, one row per letter)."""
    t = _load_tables()[name]
    letters, rows = t["letters"], t["rows"]
    width = max(len(str(v)) for r in rows for v in r) + 1
    out = ["#  %s substitution matrix (NCBI format, https://ftp.ncbi.nih.gov/blast/matrices/)" % name,
           "#  rendered from bioshell_b200/data/matrices.json"]
    out.append(" " + "".join(c.rjust(width) for c in letters))
    for c, r in zip(letters, rows):
        out.append(c + "".join(str(v).rjust(width) for v in r) + " ")
    return "\n".join(out) + "\n"


class SubstitutionMatrix:
    """21x21 i32 table + byte->index LUT (substitution_matrix.rs:30-34)."""

    def __init__(self, score, aa_index):
        self.score = np.ascontiguousarray(score, np.int32).reshape(441)
        self.aa_indexes = np.ascontiguousarray(aa_index, np.uint8).reshape(256)

    @classmethod
    def load(cls, matrix_name):
        """substitution_matrix.rs:57-68"""
        return cls.ncbi_matrix_from_buffer(ncbi_text(matrix_name))

    @classmethod
    def ncbi_matrix_from_buffer(cls, text):
        """substitution_matrix.rs:96-135 (through the C ABI)."""
        if isinstance(text, str):
            text = text.encode()
        score = np.zeros(441, np.int32)
        idx = np.zeros(256, np.uint8)
        rc = _lib.lib().bsa_parse_ncbi_matrix(text, len(text), score.ctypes.data_as(C.c_void_p),
                                              idx.ctypes.data_as(C.c_void_p))
        if rc:
            raise _lib.BsaError(rc, "IncorrectNCBIFormat / CantParseNCBIEntry")
        return cls(score, idx)

    @classmethod
    def ncbi_matrix_from_file(cls, file_name):
        """substitution_matrix.rs:142-150"""
        with open(file_name, "rb") as fh:
            return cls.ncbi_matrix_from_buffer(fh.read())

    def aa_index(self, aa_letter):
        """substitution_matrix.rs:75 (byte 255 is out of the reference's [u8;255])."""
        b = aa_letter if isinstance(aa_letter, int) else ord(aa_letter)
        if b == 255:
            raise IndexError("index out of bounds: the len is 255 but the index is 255")
        return int(self.aa_indexes[b])

    def score_by_index(self, i, j):
        """substitution_matrix.rs:81-83"""
        return int(self.score[i * 21 + j])

    def score_by_aa(self, a, b):
        """substitution_matrix.rs:89-91"""
        return self.score_by_index(self.aa_index(a), self.aa_index(b))
