"""Minimal host mirror of bioshell-seq's `Sequence` and the identity helpers the
reporters use (bioshell-seq/src/sequence/sequence.rs:8-19,481-490,532-534;
src/msa/msa.rs:261-269)."""
import numpy as np

_GAPS = (ord("-"), ord("_"))


class Sequence:
    """`Sequence { description: String, seq: Vec<u8> }` (sequence.rs:13-19); equality
    compares description AND bytes (`#[derive(PartialEq)]`, sequence.rs:8)."""
    __slots__ = ("_description", "_seq")

    def __init__(self, description, seq):
        self._description = description
        # `Sequence::new` / `from_str` drop blanks from a string (sequence.rs:34-39,69-71,190-192);
        # bytes are taken as they are, like `from_attrs` (sequence.rs:54-56)
        if isinstance(seq, str):
            seq = seq.replace(" ", "")
            try:
                seq = seq.encode("latin-1")
            except UnicodeEncodeError:                       # `c as u8` keeps the low byte
                seq = bytes(ord(c) & 0xFF for c in seq)
        self._seq = bytes(seq)

    @classmethod
    def from_str(cls, description, seq):
        return cls(description, seq)

    from_attrs = from_str

    def description(self):
        return self._description

    def as_u8(self):
        return self._seq

    def len(self):
        return len(self._seq)

    __len__ = len

    def to_string(self, _width=0):
        return self._seq.decode("latin-1")

    def __eq__(self, other):
        return isinstance(other, Sequence) and self._description == other._description and \
            self._seq == other._seq

    def __hash__(self):
        return hash((self._description, self._seq))

    def __repr__(self):
        return "Sequence(%r, %r)" % (self._description, self._seq)


def count_identical(si, sj):
    """sequence.rs:481-490 -> msa.rs:261-269: equal raw bytes, gap symbols excluded."""
    a = si.as_u8() if isinstance(si, Sequence) else (si.encode() if isinstance(si, str) else bytes(si))
    b = sj.as_u8() if isinstance(sj, Sequence) else (sj.encode() if isinstance(sj, str) else bytes(sj))
    if len(a) != len(b):
        raise ValueError("AlignedSequencesOfDifferentLengths: expected %d found %d" % (len(a), len(b)))
    x = np.frombuffer(a, np.uint8)
    y = np.frombuffer(b, np.uint8)
    return int(np.count_nonzero((x == y) & (x != _GAPS[0]) & (x != _GAPS[1])))


def len_ungapped(s):
    """sequence.rs:532-534"""
    a = s.as_u8() if isinstance(s, Sequence) else (s.encode() if isinstance(s, str) else bytes(s))
    x = np.frombuffer(a, np.uint8)
    return int(np.count_nonzero((x != _GAPS[0]) & (x != _GAPS[1])))


def pack(seqs):
    """list of Sequence/bytes -> (residues uint8[total], offsets uint64[n+1])."""
    raw = [s.as_u8() if isinstance(s, Sequence) else (s.encode() if isinstance(s, str) else bytes(s))
           for s in seqs]
    off = np.zeros(len(raw) + 1, np.uint64)
    if raw:
        off[1:] = np.cumsum([len(r) for r in raw], dtype=np.uint64)
    res = np.frombuffer(b"".join(raw) + b"\0", np.uint8)[:-1].copy() if raw else np.zeros(0, np.uint8)
    return res, off


def ungapped_lengths(res, off):
    """len_ungapped of every sequence of a packed set (vectorised)."""
    res = np.asarray(res, np.uint8)
    off = np.asarray(off, np.uint64).astype(np.int64)
    keep = ((res != _GAPS[0]) & (res != _GAPS[1])).astype(np.int64)
    c = np.concatenate([[0], np.cumsum(keep)])
    return (c[off[1:]] - c[off[:-1]]).astype(np.int64)
