"""The text side of the path -- FASTA input, sequence ids and labels, the cluster_sequences
command line -- against the reference's own tests (tests/golden/ref_text_kats.json cites them)."""
import json
import os

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import cli, fasta
from bioshell_b200.clustering import format_fasta
from bioshell_b200.sequence_id import LabelStyle, SeqId, SeqIdList, expand_taxids, parse_sequence_id, sequence_label

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tk():
    with open(os.path.join(ROOT, "tests", "golden", "ref_text_kats.json")) as fh:
        return json.load(fh)


# ----------------------------------------------------------------------------- sequence ids
def test_first_id_detection(tk):
    for text, kind, value in tk["seq_id_first"]["cases"]:
        ids = parse_sequence_id(text)
        assert ids[0] == SeqId(kind, value), text
        assert str(ids[0]) == str(SeqId(kind, value))


def test_id_lists(tk):
    for c in tk["seq_id_lists"]["cases"]:
        ids = parse_sequence_id(c["description"])
        if c["sort"]:
            ids.sort()
        if "kinds" in c:
            assert [i.kind for i in ids] == c["kinds"], c["description"]
        if "to_string" in c:
            assert ids.to_string() == c["to_string"]
        if "first" in c:
            assert ids[0] == SeqId(*c["first"])
        if "ids" in c:
            assert list(ids) == [SeqId(k, v) for k, v in c["ids"]]


def test_frog_virus_header_follows_the_code_not_the_stale_test():
    # see "frog_virus_note" in the golden file: the three ids the reference test names, in its order
    ids = parse_sequence_id(">sp.Q6GZX3.002L_FRG3G Uncharacterized protein OS=Frog virus 3 (isolate Goorha) OX=654924")
    ids.sort()
    assert list(ids[:3]) == [SeqId("SwissProt", "Q6GZX3"), SeqId("UniProtEntry", "002L_FRG3G"), SeqId("TaxId", "654924")]
    assert ids.to_string().startswith("sp|Q6GZX3|002L_FRG3G|taxid=654924")


def test_list_display_and_stable_priority_sort(tk):
    for c in tk["seq_id_list_display"]["cases"]:
        ids = SeqIdList(SeqId(k, v) for k, v in c["ids"])
        ids.sort()
        assert ids.to_string() == c["sorted_to_string"]
    # same priority keeps the order of appearance (Vec::sort is stable)
    ids = SeqIdList([SeqId("PDB", "2aza"), SeqId("CypId", "CYP1A1"), SeqId("PDB", "2gb1")])
    ids.sort()
    assert [i.value() for i in ids] == ["CYP1A1", "2aza", "2gb1"]


def test_file_names(tk):
    for text, expected in tk["file_name"]["cases"]:
        assert parse_sequence_id(text).file_name() == expected, text
    assert SeqIdList().file_name() == "sequence_ids"


def test_pdb_and_species_patterns(tk):
    for text, is_match in tk["pdb_id"]["cases"]:
        ids = parse_sequence_id(text)
        assert len(ids) == 1 and ids[0].kind == ("PDB" if is_match else "Default"), text
    for text, is_match in tk["species"]["cases"]:
        assert parse_sequence_id(text)[0].kind == ("Organism" if is_match else "Default"), text


def test_multiple_matches_and_taxid_expansion(tk):
    for text, n in tk["multiple_matches"]["cases"]:
        assert len(parse_sequence_id(text)) == n, text
    assert expand_taxids("x taxid=1,22|y") == "x taxid|1| taxid|22| y"
    assert expand_taxids("no ids here") == "no ids here"
    assert expand_taxids("TaxID=7") == "TaxID=7"          # the guard is a case-sensitive `contains("taxid")`


def test_default_id_is_the_first_word():
    assert parse_sequence_id("syn|0000042") == [SeqId("Default", "syn|0000042")]
    assert parse_sequence_id("hello world") == [SeqId("Default", "hello")]
    assert parse_sequence_id("") == [SeqId("Default", "")]


def test_sequence_label_styles(tk):
    k = tk["sequence_label"]
    d = k["description"]
    assert sequence_label(d, LabelStyle.FirstId(False, 15)) == k["first_id_unsorted_15"]
    assert sequence_label(d, LabelStyle.FullId(True, 0)) == k["full_id_sorted_0"]
    assert sequence_label(d, LabelStyle.FullId(True, 10)) == k["full_id_sorted_0"][:10]
    assert sequence_label(d, LabelStyle.FirstId(False, 5)) == "sp|A0"
    assert sequence_label(d, LabelStyle.FirstId(False, 0)) == "sp|A0A009IHW8"
    assert sequence_label(d, LabelStyle.Description(7)) == d[:7]
    assert sequence_label(d, LabelStyle.Description(0)) == ""      # `description[0..min(len, 0)]` as written


def test_statistics_with_label_style(tk):
    k = tk["statistics_with_label"]
    st = bs.AlignmentStatistics.from_strings("query", k["query"], "template", k["template"],
                                             LabelStyle.Description(k["description_n"]))
    assert str(st) == k["printed"]


# ----------------------------------------------------------------------------- FASTA
def test_fasta_collect(tk):
    k = tk["fasta_collect"]
    recs = list(fasta.FastaIterator(k["text"], k["mode"]))
    assert len(recs) == k["n_records"] and recs[0].len() == k["len_0"]
    assert recs[1].to_string(0) == k["seq_1"]
    assert recs[0].description() == "2gb1"
    # the header comes from the UNTRIMMED line minus its first byte (parse_fasta.rs:208,212)
    assert recs[1].description() == "> 2azaA"
    raw = list(fasta.FastaIterator(k["text"], fasta.RAW))
    assert [r.to_string(0) for r in raw] == [r.to_string(0) for r in recs]   # blanks go in Sequence::new


def test_fasta_clean_modes(tk):
    k = tk["fasta_clean"]
    recs = list(fasta.FastaIterator(k["text"], k["mode"]))
    assert [r.to_string(0) for r in recs] == k["expected"]
    stop = list(fasta.FastaIterator(k["text"], fasta.CLEAN_PROTEIN_STOP))
    assert stop[0].to_string(0).endswith("PRLIA*LNIL")
    small = list(fasta.FastaIterator(">x\nacd(zz\nzz)EFb\n", fasta.CLEAN_PROTEIN_STOP_SMALL))
    assert small[0].to_string(0) == "ACDEF"          # parenthesis state survives the line break; 'b' is not allowed
    custom = list(fasta.FastaIterator(">x\nabc\n", lambda line, out: out.append(line[::-1])))
    assert custom[0].to_string(0) == "cba"


def test_fasta_quirks_and_errors():
    recs = list(fasta.FastaIterator(">a\n>b\nMK\n\n>c\n", fasta.RAW))
    assert [(r.description(), r.to_string(0)) for r in recs] == [("b", "MK")]     # empty records vanish
    recs = list(fasta.FastaIterator("MKV\n>b\nAA", fasta.RAW))
    assert [(r.description(), r.to_string(0)) for r in recs] == [("", "MKV"), ("b", "AA")]
    with pytest.raises(fasta.InvalidFastaFormat):
        list(fasta.FastaIterator(">a\nMK\n# comment\n", fasta.RAW))
    assert list(fasta.FastaIterator("", fasta.RAW)) == []
    recs = list(fasta.FastaIterator(b">a b  \r\nMK\r\nV V\r\n", fasta.RAW))
    assert (recs[0].description(), recs[0].to_string(0)) == ("a b", "MKVV")


def test_load_sequences_and_display(tmp_path):
    p = tmp_path / "in.fasta"
    p.write_text("> one\nMKVLA\nGG\n>two words\nAAAA\n")
    seqs = bs.load_sequences(str(p))
    assert [(s.description(), s.to_string(0)) for s in seqs] == [("one", "MKVLAGG"), ("two words", "AAAA")]
    assert bs.load_sequences("MK VL", "name") == [bs.Sequence("name", "MKVL")]      # no '.' -> the sequence itself
    assert format_fasta(seqs[0]) == "> one\nMKVLAGG\n"
    assert format_fasta(seqs[0], 3) == "> one\nMKV\nLAG\nG\n"
    # round trip through the writer the drivers use
    q = tmp_path / "out.fasta"
    q.write_text("".join(format_fasta(s, 4) + "\n" for s in seqs))
    assert bs.load_sequences(str(q)) == seqs


def test_sequence_constructors_follow_the_reference():
    assert bs.Sequence("d", "A B C").as_u8() == b"ABC"          # Sequence::new / from_str drop blanks
    assert bs.Sequence("d", b"A B").as_u8() == b"A B"           # from_attrs takes the bytes as they are


# ----------------------------------------------------------------------------- command line
def test_cli_arguments_match_the_reference_binary():
    a = cli.build_parser().parse_args(["in.fasta"])
    assert (a.open, a.extend, a.n_threads, a.prefix, a.name_width, a.sequence_width) == (-10, -2, 1, "", 20, 80)
    assert a.detect_outliers is None and a.identity_cutoff is None and a.bucket_clustering is None
    assert not (a.single_link or a.complete_link or a.average_link or a.medoids or a.verbose)
    a = cli.build_parser().parse_args(["in.fasta", "-o", "-11", "-e", "-1", "--complete-link", "-c", "40", "-m",
                                       "--prefix", "job_", "--fasta", "o.fasta", "--distance-matrix", "d.tsv",
                                       "-w", "12", "--sequence-width", "0", "-b", "0.8", "--n-threads", "4",
                                       "--detect-outliers", "30"])
    assert (a.open, a.extend, a.complete_link, a.identity_cutoff, a.medoids) == (-11, -1, True, 40.0, True)
    assert (a.prefix, a.fasta, a.distance_matrix, a.name_width, a.sequence_width) == ("job_", "o.fasta", "d.tsv", 12, 0)
    assert (a.bucket_clustering, a.n_threads, a.detect_outliers) == (0.8, 4, 30.0)


def test_cli_fails_loudly_without_a_device(tmp_path):
    from bioshell_b200 import _lib
    if _lib.lib().bsa_device_count() > 0:
        pytest.skip("a GPU is present")
    p = tmp_path / "in.fasta"
    p.write_text(">a\nMKVLA\n>b\nMKVLG\n")
    with pytest.raises(bs.BsaError) as e:
        cli.main([str(p), "--single-link", "-c", "40", "--prefix", str(tmp_path) + "/"])
    assert "no CPU fallback" in str(e.value)
    assert not list(tmp_path.glob("cluster_*"))


@pytest.mark.gpu
def test_cli_end_to_end_on_gpu(ctx, tmp_path, oracle_matrices):
    """FASTA file in, cluster / medoid / ordered FASTA and the labelled distance matrix out, against
    the oracle pipeline (oracle aligner -> identity matrix -> oracle clustering)."""
    from bioshell_b200 import synth
    from oracle import c_oracle, pyhclust
    f32 = np.float32
    n = 40
    res, off = synth.generate(n, seed=91, dist=0, lo=30, hi=110, homolog_fraction=0.6)
    raw = res.tobytes()
    names = ["sp|P%05d|SYN%d_HUMAN synthetic protein %d [taxid=9606]" % (10000 + i, i, i) for i in range(n)]
    seqs = [bs.Sequence(names[i], raw[int(off[i]):int(off[i + 1])]) for i in range(n)]
    infile = tmp_path / "in.fasta"
    infile.write_text("".join(format_fasta(s, 60) + "\n" for s in seqs))
    prefix = str(tmp_path) + "/job_"
    rc = cli.main([str(infile), "--average-link", "-c", "40", "-m", "--symmetric", "--prefix", prefix,
                   "--distance-matrix", str(tmp_path / "dm.tsv"), "--fasta", str(tmp_path / "ordered.fasta"),
                   "-w", "24", "--sequence-width", "50"])
    assert rc == 0
    M = oracle_matrices["BLOSUM62"]
    S = c_oracle.SeqSet.from_packed(res, off)
    ref = c_oracle.align_all_pairs(S, S, M[0], M[1], -10, -2, True)
    ident = np.zeros((n, n), f32)
    ident[ref["q"], ref["t"]] = ref["identity"]
    ident = np.maximum(ident, ident.T)
    dist = (f32(100.0) - ident).astype(f32)
    root, _ = pyhclust.hierarchical_clustering(n, lambda i, j: dist[i, j], "average")
    cls = pyhclust.retrieve_clusters(root, f32(60.0))
    cls.sort(key=lambda c: c.cluster_size)
    for i, c in enumerate(cls):
        members = pyhclust.retrieve_data_id(c)
        got = bs.load_sequences("%scluster_%d-%d.fasta" % (prefix, i, c.cluster_size))
        assert got == [seqs[m] for m in members]
        med = bs.load_sequences("%scenter_%d-%d.fasta" % (prefix, i, c.cluster_size))
        assert med == [seqs[pyhclust.medoid_by_min_max(c, lambda a, b: dist[a, b])]]
    pyhclust.balance_clustering_tree(root, lambda i, j: dist[i, j])
    order = pyhclust.retrieve_data_id(root)
    assert bs.load_sequences(str(tmp_path / "ordered.fasta")) == [seqs[i] for i in order]
    rows = [ln.split("\t") for ln in (tmp_path / "dm.tsv").read_text().split("\n") if ln]
    assert len(rows) == n * n
    style = LabelStyle.FullId(True, 24)
    for r, (k, l) in zip(rows[:3 * n], [(k, l) for k in range(3) for l in range(n)]):
        assert r[0] == sequence_label(names[order[k]], style) and r[1] == sequence_label(names[order[l]], style)
        assert r[2] == "%6.3f" % ident[order[k], order[l]] and (int(r[3]), int(r[4])) == (k, l)
    assert rows[0][0] == ("sp|P%05d|SYN%d_HUMAN|taxid=9606" % (10000 + order[0], order[0]))[:24]
