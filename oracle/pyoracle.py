"""Independent pure-Python restatement of the BioShell global-alignment path.

TEST INFRASTRUCTURE.  Written separately from oracle/bioshell_oracle.c (full 2-D
tables, dict-based traces) so the two readings of the reference can be compared
on random inputs.  Follows the language-neutral spec of SURVEY.md Appendix A:
  fill       bioshell-seq/src/alignment/global.rs:57-145
  walk       bioshell-seq/src/alignment/global.rs:146-201
  matrix     bioshell-seq/src/scoring/substitution_matrix.rs:96-135
  expansion  bioshell-seq/src/alignment/alignment_path.rs:117-139
  identity   bioshell-seq/src/msa/msa.rs:261-269, src/sequence/sequence.rs:532-534,
             src/alignment/alignment_statistics.rs:71-73
Pure-Python loops: use for small cases only.
"""

N = 21


def parse_ncbi(text):
    """-> (score: list[441], aa_index: list[256])."""
    if isinstance(text, bytes):
        text = text.decode("latin-1")
    score = [0] * (N * N)
    aa = [0] * 256
    i = 0
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith("#") or line.startswith(" "):
            continue
        v = line.split()
        if len(v) < 23:
            raise ValueError("IncorrectNCBIFormat: %r" % line)
        aa[ord(v[0][0])] = i
        for j in range(1, 21):
            val = int(v[j])
            score[i * N + j - 1] = val
            score[(j - 1) * N + i] = val
        vx = int(v[len(v) - 2])
        score[i * N + 20] = vx
        score[20 * N + i] = vx
        i += 1
        if i == 20:
            break
    aa[ord("X")] = 20
    score[20 * N + 20] = -1
    return score, aa


def fill(q, t, score, aa, go, ge, lmax):
    """Full tables H,E,F plus arrows/e_trace/f_trace (dicts default 0)."""
    n, m = len(q), len(t)
    qi = [aa[b] for b in q]
    ti = [aa[b] for b in t]
    imp = go * (lmax + 1)
    H = [[0] * (m + 1) for _ in range(n + 1)]
    E = [[0] * (m + 1) for _ in range(n + 1)]
    F = [[0] * (m + 1) for _ in range(n + 1)]
    arrows, et, ft = {}, {}, {}
    for j in range(1, m + 1):
        H[0][j] = E[0][j] = go + (j - 1) * ge
        F[0][j] = imp
        arrows[(0, j)] = 1
        et[(0, j)] = 1
    for i in range(1, n + 1):
        H[i][0] = F[i][0] = go + (i - 1) * ge
        E[i][0] = imp
        arrows[(i, 0)] = 4
        ft[(i, 0)] = 1
        for j in range(1, m + 1):
            ee, eh, ef = E[i][j - 1] + ge, H[i][j - 1] + go, F[i][j - 1] + go
            if ee >= eh and ee >= ef:
                E[i][j], et[(i, j)] = ee, 1
            else:
                E[i][j], et[(i, j)] = max(eh, ef), 0
            ff, fh, fe = F[i - 1][j] + ge, H[i - 1][j] + go, E[i - 1][j] + go
            if ff >= fh and ff >= fe:
                F[i][j], ft[(i, j)] = ff, 1
            else:
                F[i][j], ft[(i, j)] = max(fh, fe), 0
            d = H[i - 1][j - 1] + score[qi[i - 1] * N + ti[j - 1]]
            h = max(d, E[i][j], F[i][j])
            H[i][j] = h
            arrows[(i, j)] = (1 if h == E[i][j] else 0) + (2 if h == d else 0) + (4 if h == F[i][j] else 0)
    return H, arrows, et, ft


def walk(n, m, arrows, et, ft):
    i, j, st, out = n, m, "H", []
    while i > 0 or j > 0:
        if st == "H":
            a = arrows.get((i, j), 0)
            if a & 2:
                if i == 0 or j == 0:
                    raise RuntimeError("panic: underflow")
                out.append("*"); i -= 1; j -= 1
            elif a & 1:
                st = "E"
            elif a & 4:
                st = "F"
            else:
                raise RuntimeError("panic: invalid H traceback state")
        elif st == "E":
            if j == 0:
                raise RuntimeError("panic: underflow")
            out.append("-"); j -= 1
            st = "E" if et.get((i, j + 1), 0) == 1 else "H"
        else:
            if i == 0:
                raise RuntimeError("panic: underflow")
            out.append("|"); i -= 1
            st = "F" if ft.get((i + 1, j), 0) == 1 else "H"
    return "".join(reversed(out))


def expand(path, q, t, gap=ord("-")):
    aq, at, qi, ti = bytearray(), bytearray(), 0, 0
    for c in path:
        if c == "-":
            aq.append(gap); at.append(t[ti]); ti += 1
        elif c == "|":
            aq.append(q[qi]); at.append(gap); qi += 1
        else:
            aq.append(q[qi]); at.append(t[ti]); qi += 1; ti += 1
    return bytes(aq), bytes(at)


GAPS = (ord("-"), ord("_"))


def count_identical(a, b):
    return sum(1 for x, y in zip(a, b) if x == y and x not in GAPS)


def len_ungapped(a):
    return sum(1 for x in a if x not in GAPS)


def align_pair(q, t, score, aa, go, ge, lmax=None):
    q, t = bytes(q), bytes(t)
    if lmax is None:
        lmax = max(len(q), len(t))
    H, arrows, et, ft = fill(q, t, score, aa, go, ge, lmax)
    path = walk(len(q), len(t), arrows, et, ft)
    aq, at = expand(path, q, t)
    nid, lq, lt = count_identical(aq, at), len_ungapped(aq), len_ungapped(at)
    ident = (nid / min(lq, lt) * 100.0) if min(lq, lt) else float("nan")
    return dict(score=H[len(q)][len(t)], path=path, aligned_q=aq, aligned_t=at, n_identical=nid,
                len_q=lq, len_t=lt, identity=ident)


def all_pairs_order(Q, T, triangle):
    """Report order of alignment_protocols.rs:94-102; Q/T are lists of (desc, bytes)."""
    out = []
    for t in range(len(T)):
        for q in range(len(Q)):
            if triangle and T[t] == Q[q]:
                break
            out.append((q, t))
    return out


def local_align(q, t, score, aa, go, ge):
    """Independent restatement of LocalAlignment (bioshell-seq/src/alignment/local.rs:83-273):
    full tables, the H selection written as the reference's three comparisons."""
    q, t = bytes(q), bytes(t)
    n, m = len(q), len(t)
    H = [[0] * (m + 1) for _ in range(n + 1)]
    E = [[0] * (m + 1) for _ in range(n + 1)]
    F = [[0] * (m + 1) for _ in range(n + 1)]
    arrows, et, ft = {}, {}, {}
    recent, best = 0, (0, 0, "H")
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            ee, eo = E[i][j - 1] + ge, max(H[i][j - 1] + go, F[i][j - 1] + go)
            if ee > 0 and ee >= eo:
                E[i][j], et[(i, j)] = ee, 1
            elif eo > 0:
                E[i][j], et[(i, j)] = eo, 2
            else:
                E[i][j], et[(i, j)] = 0, 0
            ff, fo = F[i - 1][j] + ge, max(H[i - 1][j] + go, E[i - 1][j] + go)
            if ff > 0 and ff >= fo:
                F[i][j], ft[(i, j)] = ff, 1
            elif fo > 0:
                F[i][j], ft[(i, j)] = fo, 2
            else:
                F[i][j], ft[(i, j)] = 0, 0
            d = H[i - 1][j - 1] + score[aa[q[i - 1]] * N + aa[t[j - 1]]]
            h, fl = 0, 0
            for val, bit in ((d, 2), (E[i][j], 1), (F[i][j], 4)):
                if val > h:
                    h, fl = val, bit
                elif val == h and h > 0:
                    fl |= bit
            H[i][j], arrows[(i, j)] = h, fl
            for val, st in ((H[i][j], "H"), (E[i][j], "E"), (F[i][j], "F")):
                if val > recent:
                    recent, best = val, (i, j, st)
    i, j, st = best
    out = []
    while True:
        if st == "H":
            a = arrows.get((i, j), 0)
            if a == 0:
                break
            if a & 2:
                out.append("*"); i -= 1; j -= 1
            elif a & 1:
                st = "E"
            else:
                st = "F"
        elif st == "E":
            tt = et.get((i, j), 0)
            if tt == 0:
                break
            out.append("-"); j -= 1
            st = "E" if tt == 1 else "H"
        else:
            tt = ft.get((i, j), 0)
            if tt == 0:
                break
            out.append("|"); i -= 1
            st = "F" if tt == 1 else "H"
    return dict(score=recent, path="".join(reversed(out)), end_q=best[0], end_t=best[1], start_q=i, start_t=j)
