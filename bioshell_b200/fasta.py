"""FASTA input for the drivers of the alignment path: `FastaIterator` with the reference's
parsing modes and `load_sequences` (bioshell-seq/src/sequence/parse_fasta.rs:35-303,
bioshell-seq/src/sequence/mod.rs:44-54).  Host text handling only.

Behaviour kept from the reference, quirks included:
  * a record is emitted only when residues were collected: a header followed directly by
    another header (or by the end of the file) is dropped (parse_fasta.rs:186-217);
  * the header is `line_with_newline[1..].trim()`, taken from the UNTRIMMED line, so an indented
    "   > name" keeps its '>' in the description (parse_fasta.rs:208,212);
  * a line starting with '#' is an error (parse_fasta.rs:200-205);
  * `Sequence::new` removes blanks from the residues (sequence.rs:34-39,190-192);
  * the clean modes drop everything between '(' and ')' (state survives line breaks), keep only
    the allowed letters and upper-case them (parse_fasta.rs:262-289).
"""
import io

from .sequence import Sequence

PROTEIN_LETTERS = b"ACDEFGHIKLMNPQRSTVWYBJOUXZ-_"                                  # parse_fasta.rs:293
PROTEIN_LETTERS_STOP = b"ACDEFGHIKLMNPQRSTVWYBJOUXZ-_*"                            # :296
PROTEIN_LETTERS_STOP_SMALL = b"ACDEFGHIKLMNPQRSTVWYBJOUXZacdefghiklmnopqrtsvwx-_*"  # :299 (as written there)
NUCLEIC_LETTERS = b"cgmtu-_"                                                       # :291 (as written there)

RAW = "raw"
CLEAN_PROTEIN = "clean-protein"
CLEAN_PROTEIN_STOP = "clean-protein-stop"
CLEAN_PROTEIN_STOP_SMALL = "clean-protein-stop-small"
CLEAN_NUCLEIC = "clean-nucleic"

_ALLOWED = {CLEAN_PROTEIN: PROTEIN_LETTERS, CLEAN_PROTEIN_STOP: PROTEIN_LETTERS_STOP,
            CLEAN_PROTEIN_STOP_SMALL: PROTEIN_LETTERS_STOP_SMALL, CLEAN_NUCLEIC: NUCLEIC_LETTERS}


class InvalidFastaFormat(ValueError):
    """`SequenceError::InvalidFastaFormat { line, description }`"""

    def __init__(self, line, description):
        ValueError.__init__(self, "%s: %r" % (description, line))
        self.line = line
        self.description = description


class _AllowedCharsOnly:
    """parse_fasta.rs:262-289"""

    def __init__(self, allowed):
        self.inside_parentheses = False
        self.allowed = frozenset(allowed)

    def parse_line(self, line, out):
        for b in line.encode("utf-8"):
            if b == 0x28:
                self.inside_parentheses = True
            elif b == 0x29:
                self.inside_parentheses = False
            elif self.inside_parentheses:
                continue
            elif b in self.allowed:
                out.append(chr(b).upper())


class FastaIterator:
    """Iterates `Sequence`s of a text stream (a file object, or any iterable of lines).  `mode` is
    one of the module's mode names or a callable `f(line: str, out: list[str])` (the reference's
    `FastaParsingMode::Custom`)."""

    def __init__(self, stream, mode=RAW):
        if isinstance(stream, (str, bytes)):
            stream = io.StringIO(stream.decode("utf-8") if isinstance(stream, bytes) else stream)
        self._lines = iter(stream)
        self._header = ""
        self._seq = []
        self._done = False
        if callable(mode):
            self._push = mode
        elif mode == RAW:
            self._push = lambda line, out: out.append(line)
        elif mode in _ALLOWED:
            self._push = _AllowedCharsOnly(_ALLOWED[mode]).parse_line
        else:
            raise ValueError("unknown FASTA parsing mode %r" % (mode,))

    def __iter__(self):
        return self

    def _take(self):
        ret = Sequence(self._header, "".join(self._seq))
        self._seq = []
        return ret

    def __next__(self):
        if self._done:
            raise StopIteration
        for buffer in self._lines:
            if isinstance(buffer, bytes):
                buffer = buffer.decode("utf-8")
            line = buffer.strip()
            if line.startswith("#"):
                raise InvalidFastaFormat(line, "Fasta line must not start with '#' character")
            if line.startswith(">"):
                header = buffer[1:].strip()
                if self._seq_len() > 0:
                    ret = self._take()
                    self._header = header
                    return ret
                self._header = header
            elif line:
                self._push(line, self._seq)
        self._done = True
        if self._seq_len() > 0:
            return self._take()
        raise StopIteration

    def _seq_len(self):
        return sum(len(s) for s in self._seq)


def load_sequences(seq_or_fname, seq_name=""):
    """bioshell-seq/src/sequence/mod.rs:44-54: an argument containing a '.' is a FASTA file name
    (read in Raw mode), anything else is the sequence itself."""
    if "." in seq_or_fname:
        with open(seq_or_fname, "rb") as fh:      # binary: lines end at '\n' only, like read_line
            return list(FastaIterator(fh, RAW))
    return [Sequence.from_str(seq_name, seq_or_fname)]
