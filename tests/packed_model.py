"""Scalar Python model of the kernels' packed-lane arithmetic (gotoh_kernels.cuh):
v = score << (cs+2) | prio << cs | count, one signed max per selection.  Used by the
CPU tests to check the ALGORITHM (tie rules, forward-carried identity, dropped sentinel,
direction nibbles) against the oracle without a GPU.  Test infrastructure only."""


def packed_align(q, t, score, aa, go, ge, cs, want_dirs=False):
    """q rows (query), t columns (template); raw bytes.  Returns (score, n_identical, dirs)
    where dirs[(i,j)] is the 4-bit direction code the DIRS kernel stores."""
    n, m = len(q), len(t)
    S, U = 1 << (cs + 2), 1 << cs
    GE, GO, MASK, PH, PV = ge * S, go * S, ~(3 * U), 2 * U, 1 * U
    gaps = (ord("-"), ord("_"))

    def T(a, b):
        ident = 1 if (cs > 0 and a == b and a not in gaps) else 0
        return score[aa[a] * 21 + aa[b]] * S + 3 * U + ident

    Hc = [(go + j * ge) * S for j in range(m)]          # top border H[0][j+1]
    Fr = [h + GO for h in Hc]                           # eager F for row 1 (opened)
    dirs = {}
    hb = go * S
    for i in range(n):
        hin, er = hb, hb + GO                           # left border, eager E (opened)
        hb += GE
        hd = 0 if i == 0 else (go + (i - 1) * ge) * S   # H[i][0] of the previous row
        for c in range(m):
            e = er | PH
            f = Fr[c] | PV
            d = hd + T(q[i], t[c])
            h = max(d, e, f)
            if want_dirs:
                dirs[(i + 1, c + 1)] = (h & 3) | (((er | Fr[c]) & 3) << 2)
            hc = h & MASK
            hg = hc + GO
            er = max(e + GE, hg)
            Fr[c] = max(f + GE, hg)
            hd, Hc[c] = Hc[c], hc
    if m == 0:
        v = 0 if n == 0 else (go + (n - 1) * ge) * S
    else:
        v = Hc[m - 1]
    assert -(1 << 31) <= v < (1 << 31)
    return v >> (cs + 2), v & (U - 1), dirs


def walk_dirs(n, m, dirs):
    """traceback_kernel's walk over the 4-bit codes (cs = 0 layout)."""
    out, i, j, st = [], n, m, 0
    while i > 0 and j > 0:
        nib = dirs[(i, j)]
        if st == 0:
            hd = nib & 3
            if hd == 3:
                out.append("*"); i -= 1; j -= 1
            elif hd == 2:
                st = 1
            elif hd == 1:
                st = 2
            else:
                raise RuntimeError("invalid direction")
        elif st == 1:
            out.append("-"); j -= 1
            st = 1 if nib & 8 else 0
        else:
            out.append("|"); i -= 1
            st = 2 if nib & 4 else 0
    out.extend("-" * j)
    out.extend("|" * i)
    return "".join(reversed(out))


def tagged_align(q, t, score, aa, go, ge, cs, xb=5, K=4, R=16):
    """Model of the TAG cell (gotoh_kernels.cuh, cell_row_tag): the E/F values keep their H-max
    priority for good (opened from H with prio 2 / 1 already in place) and the extend-beats-open
    tie rule (global.rs:109,122) comes from a STREAK field x below the priority: an extension
    adds 1 to x, an opening has x = 0.  x only matters in the comparison that follows the
    increment, so it is simply cleared when E enters the next lane (every K columns) and every R
    rows for F -- it never reaches 2**xb.   v = score << (cs+xb+2) | prio << (cs+xb) | x << cs | count
    4 ALU-pipe instructions per cell (VIMNMX3, LOP3, 2 VIADDMNMX) + 3 IMAD."""
    n, m = len(q), len(t)
    U, X1 = 1 << cs, 1 << cs
    P1 = 1 << (cs + xb)
    S = 1 << (cs + xb + 2)
    XMASK = ((1 << xb) - 1) << cs
    MASK = ~(3 * P1 | XMASK)
    GEX = ge * S + X1
    GOE, GOF = go * S + 2 * P1, go * S + P1
    gaps = (ord("-"), ord("_"))

    def T(a, b):
        ident = 1 if (cs > 0 and a == b and a not in gaps) else 0
        return score[aa[a] * 21 + aa[b]] * S + 3 * P1 + ident

    Hc = [(go + j * ge) * S for j in range(m)]
    Fr = [h + GOF for h in Hc]
    hb = go * S
    for i in range(n):
        if i % R == 0:
            Fr = [f & ~XMASK for f in Fr]
        hin, er = hb, hb + GOE
        hb += ge * S
        hd = 0 if i == 0 else (go + (i - 1) * ge) * S
        for c in range(m):
            if c % K == 0:
                er &= ~XMASK                      # lane boundary: after the shuffle
            d = hd + T(q[i], t[c])
            h = max(d, er, Fr[c])
            hc = h & MASK
            er = max(er + GEX, hc + GOE)
            Fr[c] = max(Fr[c] + GEX, hc + GOF)
            assert (er & XMASK) >> cs <= K and (Fr[c] & XMASK) >> cs <= R
            assert (er >> (cs + xb)) & 3 == 2 and (Fr[c] >> (cs + xb)) & 3 == 1
            hd, Hc[c] = Hc[c], hc
    if m == 0:
        v = 0 if n == 0 else (go + (n - 1) * ge) * S
    else:
        v = Hc[m - 1]
    assert -(1 << 31) <= v < (1 << 31)
    return v >> (cs + xb + 2), v & (U - 1)


def frame_align(q, t, score, aa, go, ge, cs, xb=5, K=4, R=16, pad=False, phase=0, etag=False):
    """Model of the FRAME cell (gotoh_kernels.cuh, TAG mode since round 2): the TAG cell in a moving
    frame.  Every stored DP value of cell (i, j) carries score - (i + j) * ge, so that BOTH gap
    extensions cost nothing in the frame:
        D* = H*[i-1][j-1] + (s - 2 ge)         E*[j+1] = max(E*[j], H* + go - ge)      F*[i+1] = max(F*[i], H* + go - ge)
    and the borders become the constant go - ge.  Extend-beats-open (global.rs:109,122):
      * E keeps the streak field: the extension adds 1 to x (and nothing to the score), openings have x = 0;
        x is cleared when E enters the next lane (every K columns).
      * F needs no addition at all: the OPENING candidate of row i carries x = R-1 - (i mod R), so an
        older opening outranks a newer one on ties; every R rows the stored F values get the top bit of x
        (an OR), which outranks every candidate of the next R rows.
    6 instructions per cell: IMAD (diagonal), VIMNMX3, LOP3, IMAD (E opening), 2 VIADDMNMX.
    pad=True models the EVEN-ALIGNED stream of the two-row kernels (stream_block_tag2): a query of odd
    length is preceded by one PAD row whose profile row adds 0 on the diagonal of the frame (go - ge in
    DP column 1), which reproduces the top border exactly -- H* = go - ge, the stored F keep their top tag
    bit -- so the end-of-sequence flag always sits on the second row of a double step; the only special
    case is the first real row's diagonal input in column 1, H*[0][0] = 0, which lane 0 takes when the
    row above is the PAD row.  `phase` is where the query starts inside the kernel's blocks of R rows.
    etag=True models BSA_ETAG: E uses the scheme of F along the columns of a lane -- the OPENING of column c
    carries the tag K-1-c, the extension adds nothing (E = max(H + GOE_c, E), one VIADDMNMX with a warp-uniform
    constant), and an E that crosses a lane boundary gets the full tag field, older than every opening of the
    next lane: 5 instructions per cell (IMAD, VIMNMX3, LOP3, 2 VIADDMNMX)."""
    n, m = len(q), len(t)
    assert R & (R - 1) == 0 and R <= 1 << (xb - 1)
    PAD = -1
    rows = list(q)
    if pad and n % 2 == 1:
        rows = [PAD] + rows
    U, X1 = 1 << cs, 1 << cs
    P1 = 1 << (cs + xb)
    S = 1 << (cs + xb + 2)
    XMASK = ((1 << xb) - 1) << cs
    MASK = ~(3 * P1 | XMASK)
    XTOP = (1 << (xb - 1)) * X1
    GOE = (go - ge) * S + 2 * P1
    GOF = (go - ge) * S + P1
    HB = (go - ge) * S                      # H*[0][j], j >= 1, and H*[i][0], i >= 1
    gaps = (ord("-"), ord("_"))

    def T(a, b, c):
        if a == PAD:
            return (HB if c == 0 else 0) + 3 * P1
        ident = 1 if (cs > 0 and a == b and a not in gaps) else 0
        return (score[aa[a] * 21 + aa[b]] - 2 * ge) * S + 3 * P1 + ident

    Hc = [HB] * m
    Fr = [HB + GOF + XTOP] * m             # eager F of row 1: opened from the top border, older than any candidate
    lo, hi = 0, 0
    for i in range(len(rows)):
        g = phase + i
        if g % R == 0:
            Fr = [f | XTOP for f in Fr]
        cF = GOF + (R - 1 - g % R) * X1
        er = HB + GOE
        hd = 0 if (i == 0 or rows[i - 1] == PAD) else HB
        for c in range(m):
            if c % K == 0:
                er = (er | XMASK) if etag else (er & ~XMASK)      # lane boundary: after the shuffle
            d = hd + T(rows[i], t[c], c)
            h = max(d, er, Fr[c])
            hc = h & MASK
            if etag:
                assert K - 1 < (1 << xb) - 1
                er = max(hc + GOE + (K - 1 - c % K) * X1, er)
            else:
                er = max(er + X1, hc + GOE)
            Fr[c] = max(hc + cF, Fr[c])
            assert etag or (er & XMASK) >> cs <= K
            assert (er >> (cs + xb)) & 3 == 2 and (Fr[c] >> (cs + xb)) & 3 == 1
            lo, hi = min(lo, d, er, Fr[c]), max(hi, d, er, Fr[c])
            hd, Hc[c] = Hc[c], hc
    assert -(1 << 31) <= lo and hi < (1 << 31)
    if m == 0:
        return (0 if n == 0 else go + (n - 1) * ge), 0
    if n == 0:
        return go + (m - 1) * ge, 0
    v = Hc[m - 1]
    return (v >> (cs + xb + 2)) + (n + m) * ge, v & (U - 1)


def frame16_align(q, t, score, aa, go, ge, pad=True, mpad=0):
    """Model of the 16-bit score-only cell in the moving frame (gotoh_kernels.cuh, stream_block16_fa): one
    half of a packed register, biased by 0x8000, every operation modulo 2**16 as the U16x2 instructions do:
        t = max(hd + T, e)   h = max(t, f)   e = max(h + (go - ge), e)   f = max(h + (go - ge), f)
    borders go - ge / 2 (go - ge), a PAD row (profile 0; go - ge in DP column 1) ahead of odd-length queries,
    `mpad` extra padded columns with profile 0 to the right (the shorter template of a pair).
    Returns the score, or None when a value leaves the 16-bit range (the host's range check must refuse it)."""
    n, m = len(q), len(t)
    if m == 0:
        return 0 if n == 0 else go + (n - 1) * ge
    if n == 0:
        return go + (m - 1) * ge
    B, M16 = 0x8000, 0xffff
    PAD = -1
    rows = ([PAD] if pad and n % 2 else []) + list(q)
    ok = [True]

    def wadd(a, b):                      # 16-bit wrapping add of a two's-complement addend to a biased value
        r = a + b
        if not 0 <= r <= M16:
            ok[0] = False
        return r & M16

    def T(a, c):
        if c >= m:
            return 0
        if a == PAD:
            return (go - ge) if c == 0 else 0
        return score[aa[a] * 21 + aa[t[c]]] - 2 * ge

    HB = B + go - ge
    GOF = go - ge
    FB = wadd(HB, GOF)
    W = m + mpad
    H = [HB] * W
    F = [FB] * W
    for i, a in enumerate(rows):
        hd = B if (i == 0 or rows[i - 1] == PAD) else HB      # H*[.][0] of the row above; H*[0][0] = 0
        e = FB
        for c in range(W):
            tt = max(wadd(hd, T(a, c)), e)
            h = max(tt, F[c])
            e = max(wadd(h, GOF), e)
            F[c] = max(wadd(h, GOF), F[c])
            hd, H[c] = H[c], h
    if not ok[0]:
        return None
    return H[m - 1] - B + (n + m) * ge


def wave_frame_align(q, t, score, aa, go, ge, pre=4):
    """Scalar model of the K3 wavefront cell (wave_kernels.cuh, wave_block_frame + traceback_kernel):
    v = (score - (i + j) ge) << 3 | prio << 1 | m.  The rows are padded at the TOP to a multiple of 4 with
    rows of a pad residue (profile entry 6 = score 0, priority 3) -- `pre` more such rows model the steps a
    lane spends before its first row -- columns are padded on the right to a multiple of 256.  Returns
    (score, path, words) where words[(t, lane_block)] is the 32-bit direction word of padded row t."""
    n, m = len(q), len(t)
    pad = (4 - n % 4) % 4
    mp = (m + 255) // 256 * 256
    HB = (go - ge) * 8
    GOE, GOF = HB + 4, HB + 2
    EB, FB = HB + GOE, HB + GOF

    def T(a, c):
        if a is None or c >= m:
            return 6
        return (score[aa[a] * 21 + aa[t[c]]] - 2 * ge) * 8 + 6

    H = [HB] * mp
    F = [FB] * mp
    words = {}
    hin_prev = HB                      # H of the left border one row up (the diagonal of column 1)
    for tt in range(-pre, n + pad):
        a = q[tt - pad] if tt >= pad else None
        # left border entry of this row (block 0's ring): H*[i][0], E*[i][1]; H*[0][0] = 0 sits on the row
        # right above row 1 (the last padding row, or -- no padding -- the initial diagonal)
        hin, er = HB, EB
        if tt == pad - 1:
            hin = 0
        hd = hin_prev
        if tt == 0 and pad == 0 and pre == 0:
            hd = 0
        if pad == 0 and tt == 0 and pre > 0:
            hd = 0                     # the kernel: hdiag of (block 0, lane 0) starts at 0 when there is no padding
        dA = dF = 0
        for c in range(mp):
            f = F[c]
            d = hd + T(a, c)
            h = max(d, er, f)
            dA = ((dA >> 3) | ((((h & ~1) | (er & 1)) & 7) << 29)) & 0xffffffff
            f1 = f | 1
            dF = (dF * 2 + f1 - f) & 0xffffffff
            hc = h & ~7
            er = max(hc + GOE, er | 1)
            F[c] = max(hc + GOF, f1)
            hd, H[c] = H[c], hc
            if c % 8 == 7:
                if tt >= 0:
                    words[(tt, c // 8)] = (dA & 0xffffff00) | (dF & 0xff)
                dA = dF = 0
        hin_prev = hin
        for v in H + F:
            assert -(1 << 31) <= v < (1 << 31)
    if m == 0:
        return (0 if n == 0 else go + (n - 1) * ge), "|" * n, words
    if n == 0:
        return go + (m - 1) * ge, "-" * m, words
    sc = (H[m - 1] >> 3) + (n + m) * ge

    def look(i, j):
        w = words[(i - 1 + pad, (j - 1) // 8)]
        c = (j - 1) % 8
        a3 = (w >> (8 + 3 * c)) & 7
        return a3 >> 1, a3 & 1, ((w >> (7 - c)) & 1) ^ 1

    # the run-based walk of traceback_kernel: 32 cells ahead in the current direction at a time
    out, i, j, st = [], n, m, 0
    while i > 0 and j > 0:
        cells = []
        for k in range(32):
            ii, jj = (i - k if st != 1 else i), (j - k if st != 2 else j)
            ok = ii >= 1 and jj >= 1 and (jj - 1) // 256 == (j - 1) // 256
            cells.append(look(ii, jj) if ok else None)
        flags = [c is not None and (c[0] == 3 if st == 0 else c[st] == 1) for c in cells]
        t1 = flags.index(False) if False in flags else 32
        seen = t1 < 32 and cells[t1] is not None
        if st == 0:
            out += ["*"] * t1
            i -= t1; j -= t1
            if seen:
                hd = cells[t1][0]
                assert hd in (1, 2)
                st = 1 if hd == 2 else 2
        else:
            run = t1 + (1 if seen else 0)
            out += ["-" if st == 1 else "|"] * run
            if st == 1:
                j -= run
            else:
                i -= run
            if seen:
                st = 0
    out += ["-"] * j + ["|"] * i
    return sc, "".join(reversed(out)), words


def wave_ring_schedule(X, WB=32, PF=8, U=2, span=31):
    """Scalar model of the boundary hand-off of the K3 wavefront consumer (stream_block with WRING,
    gotoh_kernels.cuh): which boundary entry lane 0 takes at every step and how many entries must
    have been published when a batch is fetched.  Returns (used, fetches): used[S] = entry index
    taken at step S (S < X), fetches = [(step, first entry, last entry + 1, needed published)]."""
    nsteps = (X + span + (U - 1)) // U * U
    fetches = []
    ring_cur = {l: l for l in range(WB) if l < X}          # initial batch, waits for min(WB, X)
    fetches.append((-1, 0, min(WB, X), min(WB, X)))
    ring_nxt = {}
    used = {}
    for s in range(0, nsteps, U):
        if (s & (WB - 1)) == PF:
            base = (s & ~(WB - 1)) + WB
            if base < X:
                need = min(base + WB, X)
                ring_nxt = {l: base + l for l in range(WB) if base + l < X}
                fetches.append((s, base, need, need))
        for u in range(U):
            S = s + u
            if S < X:
                used[S] = ring_cur.get(S & (WB - 1))
            if (S & (WB - 1)) == WB - 1:
                ring_cur = ring_nxt
    return used, fetches
