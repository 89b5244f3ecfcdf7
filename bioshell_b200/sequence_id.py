"""Host mirror of the reference's sequence labels: the text side of `cluster_sequences`'
distance-matrix output (bin/cluster_sequences.rs:232-248 labels every row with
`sequence_label(description, FullId{sort:true, n:name_width})`).

Follows bioshell-seq/src/sequence/sequence_id.rs (SeqId :9-44, priority order :58-78,
Display :104-128, parse_sequence_id :212-254, expand_taxids :256-274, SeqIdList :277-345,
the pattern table :375-444) and sequence_label.rs:3-63.  Plain host code: nothing here
touches the GPU path.  The patterns are the reference's own, written for Python's `re`
(POSIX classes spelled out, `$` as `\\Z` because a Rust `$` never matches before a newline).
"""
import re

# variant name -> (sort priority (sequence_id.rs:58-78), display format (sequence_id.rs:104-128))
_KINDS = {
    "CypId": (0, "%s"), "PDB": (1, "pdb|%s"), "SwissProt": (2, "sp|%s"), "UniProtKB": (3, "UniProt|%s"),
    "TrEmbl": (4, "tr|%s"), "UniProtEntry": (5, "%s"), "UniParc": (6, "%s"), "UniRef": (7, "%s"),
    "RefSeq": (8, "ref|%s"), "GenBank": (9, "gb|%s"), "Ensembl": (10, "Ensembl|%s"), "DDBJ": (11, "dbj|%s"),
    "NCBIGI": (12, "gi|%s"), "KEGG": (13, "%s"), "Default": (14, "%s"), "TaxId": (15, "taxid=%s"),
    "Organism": (16, "[organism=%s]"),
}


class SeqId:
    """One recognised identifier: `SeqId::<kind>(value)`; equality is kind AND value, ordering is
    the kind's priority only (sequence_id.rs:46-56)."""
    __slots__ = ("kind", "_value")

    def __init__(self, kind, value):
        if kind not in _KINDS:
            raise ValueError("unknown SeqId kind %r" % (kind,))
        self.kind = kind
        self._value = value

    def value(self):
        return self._value

    def order_priority(self):
        return _KINDS[self.kind][0]

    def __eq__(self, other):
        return isinstance(other, SeqId) and self.kind == other.kind and self._value == other._value

    def __hash__(self):
        return hash((self.kind, self._value))

    def __lt__(self, other):
        return self.order_priority() < other.order_priority()

    def __str__(self):
        if self.kind == "PDB" and self._value.startswith("pdb_"):      # sequence_id.rs:107
            return self._value
        return _KINDS[self.kind][1] % self._value

    def __repr__(self):
        return "SeqId.%s(%r)" % (self.kind, self._value)


def sanitize_filename(name):
    """bioshell-core/src/io/utils.rs:544-556"""
    out = []
    for c in name:
        if c in '/\\:*?"<>|':
            out.append("_")
        elif ord(c) < 32 or 127 <= ord(c) < 160:      # char::is_control
            continue
        else:
            out.append(c)
    return "".join(out)


class SeqIdList(list):
    """`SeqIdList(Vec<SeqId>)` (sequence_id.rs:277-345)."""

    def sort(self):                       # stable, by priority only (Vec::sort on Ord)
        list.sort(self, key=SeqId.order_priority)

    def to_string(self):
        out = []
        for i, sid in enumerate(self):
            if i > 0:
                out.append(" " if sid.kind == "Organism" else "|")        # :337-339
            out.append(str(sid))
        return "".join(out)

    __str__ = to_string

    def file_name(self):
        """sequence_id.rs:282-295"""
        if not self:
            return "sequence_ids"
        name = self.to_string().replace("|", "_").replace("]", "").replace(" ", "_").replace("[organism=", "")
        return sanitize_filename(name.strip("_|"))


_ALPHA = "A-Za-z"
_ALNUM = "A-Za-z0-9"
_GB_TAIL = r"(?:\Z|[^\w]|_)"

# sequence_id.rs:375-444, in the reference's order (a match blanks its span for the later patterns)
_PATTERNS = [(re.compile(p), k) for p, k in [
    (r"(?i:\[?taxid=(\d+))", "TaxId"),
    (r"(?i:\[?TaxID=(\d+))", "TaxId"),
    (r"OX=(\d+)", "TaxId"),
    (r"(?i:(?:\b|\|)taxid\|(\d+))", "TaxId"),
    (r"(?:^|[|>]|\b)([a-z][a-z0-9]{2,4}:[A-Za-z0-9_.-]+)(?:\Z|[|>]|\b)", "KEGG"),
    (r"(?:^|pdb|\s+|\|)([0-9][A-Za-z0-9]{3}(?::[_]?[A-Za-z0-9]{1,3})?)(?:\Z|[ |])", "PDB"),
    (r"\b(pdb_[A-Za-z0-9]{8}(?::[_]?[A-Za-z0-9]{1,3})?)(?:\Z|[ |])", "PDB"),
    (r"(?:\b|\|\>|_)((?:AC|NC|NG|NT|NW|NZ|NM|NR|XM|XR|AP|NP|YP|XP|WP)_[0-9]+\.\d+)\b", "RefSeq"),
    (r"(?:\b|\|\>)sp[|.]([A-Z0-9]{6}|[A-Z0-9]{10})(?:-\d+)?[|.]", "SwissProt"),
    (r"(?:\b|\|)tr[|.]([A-Z0-9]{6}|[A-Z0-9]{10})(?:-\d+)?[|.]", "TrEmbl"),
    (r"(?:\b|\|)([A-Z0-9]{3,}_[A-Z0-9]{3,5})\b", "UniProtEntry"),
    (r"\b([OPQ][0-9][A-Z0-9]{3}[0-9]|[A-NR-Z][0-9](?:[A-Z][A-Z0-9]{2}[0-9]){1,2})(?:-\d+)?\b", "UniProtKB"),
    (r"\b(UniRef\d{2,3}_[A-Z0-9]+)\b", "UniRef"),
    (r"\b(UPI[0-9A-F]{10})\b", "UniParc"),
    (r"\bGI:(\d+)\b", "NCBIGI"),
    (r"\bgi\|(\d+)\b", "NCBIGI"),
    (r"\b(ENS[TPGR][0-9]{11})\b", "Ensembl"),
    (r"(?:\b|\|)dbj\|([A-Z]{3}[0-9]{5}(?:\.\d+)?)\b", "DDBJ"),
    (r"(?:\b|\||_)gb\|([A-Z]{1,3}[0-9]{4,8}(?:\.\d+)?)\b", "GenBank"),
    (r"(?:\b|\||_)([A-Z][0-9]{5}(?:\.\d+)?)" + _GB_TAIL, "GenBank"),
    (r"(?:\b|\||_)([A-Z]{2}[0-9]{6}(?:\.\d+)?)" + _GB_TAIL, "GenBank"),
    (r"(?:\b|\||_)([A-Z]{2}[0-9]{8}(?:\.\d+)?)" + _GB_TAIL, "GenBank"),
    (r"(?:\b|\||_)([A-Z]{3}[0-9]{5}(?:\.\d+)?)" + _GB_TAIL, "GenBank"),
    (r"(?:\b|\||_)([A-Z]{3}[0-9]{7}(?:\.\d+)?)" + _GB_TAIL, "GenBank"),
    (r"(?:\b|\||_)([A-Z]{4}[0-9]{8,10}(?:\.\d+)?)" + _GB_TAIL, "GenBank"),
    (r"\[organism=([%s][%s. ]*)\]" % (_ALPHA, _ALNUM), "Organism"),
    (r"\[([%s][%s. ]*)\]" % (_ALPHA, _ALNUM), "Organism"),
    (r"\bOS=\s*([^|=\r\n]*[^|=\s])\s*(?:\||\Z)", "Organism"),
    (r"(?:^|[^\w]|_)(CYP[0-9]+[A-Z]{1,3}[0-9]+[a-z]?(?:v[0-9]{1,2})?(?:P(?:[0-9]+|[NC])?)?X?(?:_[A-Z]{1,4})?)" + _GB_TAIL,
     "CypId"),
    (r"(?:^|[^\w]|_)(Cyp[0-9]+[a-z]{1,3}[0-9]+[a-z]?(?:v[0-9]{1,2})?(?:P(?:[0-9]+|[NC])?)?X?(?:_[A-Z]{1,4})?)" + _GB_TAIL,
     "CypId"),
]]

_TAXID_LIST_RE = re.compile(r"(?i)\btaxid(?:=|\|)(\d+(?:,\d+)*)(?:\|)?[ \t]*")


def expand_taxids(text):
    """sequence_id.rs:256-274: `taxid=1,2` -> `taxid|1| taxid|2| `"""
    if "taxid" not in text:
        return text
    return _TAXID_LIST_RE.sub(lambda m: " ".join("taxid|%s|" % i for i in m.group(1).split(",")) + " ",
                              text).rstrip()


def parse_sequence_id(description):
    """sequence_id.rs:212-254: every pattern scans the description in table order; a match is
    recorded with its start offset and blanked, so later patterns cannot reuse it; the
    identifiers come back in order of appearance; none found -> Default(first word)."""
    buf = list(expand_taxids(description))
    found = []
    for pattern, kind in _PATTERNS:
        view = "".join(buf)
        for m in pattern.finditer(view):
            if m.group(1) is None:
                continue
            found.append((m.start(), SeqId(kind, m.group(1))))
            buf[m.start():m.end()] = " " * (m.end() - m.start())
    found.sort(key=lambda sv: sv[0])
    ids = SeqIdList(v for _, v in found)
    if not ids:
        words = description.split()
        ids.append(SeqId("Default", words[0] if words else ""))
    return ids


class LabelStyle:
    """sequence_label.rs:3-12"""
    __slots__ = ("style", "sort", "n")

    def __init__(self, style, sort=False, n=0):
        self.style, self.sort, self.n = style, sort, n

    @classmethod
    def FirstId(cls, sort, n):
        return cls("FirstId", sort, n)

    @classmethod
    def FullId(cls, sort, n):
        return cls("FullId", sort, n)

    @classmethod
    def Description(cls, n):
        return cls("Description", False, n)


def sequence_label(description, style):
    """sequence_label.rs:33-63 (lengths count bytes there; descriptions here are ASCII in practice,
    a multi-byte cut would panic in the reference)"""
    if style.style == "Description":
        return description[:min(len(description), style.n)]
    ids = parse_sequence_id(description)
    if style.sort:
        ids.sort()
    if style.style == "FirstId":
        out = str(ids[0])
        if style.n == 0:
            return out
        return out[:style.n] if len(out) > style.n else out.rjust(style.n)
    desc = ids.to_string()
    return desc if style.n == 0 else desc[:min(len(desc), style.n)]
