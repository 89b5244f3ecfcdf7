/*
 * bioshell_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A literal, single-translation-unit C restatement of the one hot path of
 * dgront/BioShell v4 that this repository accelerates: all-vs-all global
 * alignment (Needleman-Wunsch with Gotoh affine gaps) as run by
 * bioshell-seq's `align_all_pairs`.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs may load this library; the
 * product path (bioshell_b200/csrc, libbioshell_align.so) never links or calls
 * it and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference is 100% Rust and cannot be built in
 * this image (no rustc/cargo, unvendored crates), so the pin is
 * the reference's own known-answer tests, restated in tests/golden/ref_kats.json
 * and checked by tests/test_oracle_kats.py:
 *   bioshell-seq/tests/test_aligners.rs:13-58            (3 global cases x 2 orientations)
 *   bioshell-seq/tests/test_substitution_matrix.rs:3-9   (BLOSUM80 values)
 *   bioshell-seq/src/scoring/mod.rs:46-67                (BLOSUM62 doc-tests)
 *   bioshell-seq/src/scoring/similarity_score.rs:33-65   (encode + lookup doc-tests)
 *   bioshell-seq/src/alignment/alignment_path.rs:155-197 (path expansion)
 *   bioshell-seq/src/alignment/alignment_statistics.rs:16-26
 *   bioshell-seq/src/sequence/sequence.rs:476-479,526-530, src/msa/msa.rs:243-247
 * plus an independent second restatement (oracle/pyoracle.py) that is compared
 * with this file on random inputs (tests/test_oracle_cross.py).
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference checkout).  The odd parts of the reference are kept on purpose:
 * 3-term E/F recurrences, the capacity-dependent "impossible" sentinel,
 * unknown byte -> index 0, forced X/X = -1, mirrored matrix writes, identity
 * on raw bytes, t-major pair order, description+bytes equality for the
 * triangle break, f64 -> f32 identity.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NCOL 21 /* substitution_matrix.rs:25 */

#define ORC_OK 0
#define ORC_ERR_FORMAT (-1)    /* IncorrectNCBIFormat, substitution_matrix.rs:108 */
#define ORC_ERR_PARSE (-2)     /* CantParseNCBIEntry, :115,123 */
#define ORC_ERR_PANIC (-3)     /* a Rust panic in the reference (index/underflow/invalid state) */
#define ORC_ERR_ALLOC (-4)
#define ORC_ERR_CAPACITY (-5)  /* sequence longer than the aligner capacity (Vec index panic) */

/* ------------------------------------------------------------------ */
/* substitution_matrix.rs:96-135  ncbi_matrix_from_buffer             */
/* ------------------------------------------------------------------ */

/* Rust `str::parse::<i32>()`: optional sign, >=1 ASCII digit, nothing else. */
static int parse_i32(const char *s, size_t n, int32_t *out) {
    size_t k = 0;
    int neg = 0;
    int64_t v = 0;
    if (n == 0) return -1;
    if (s[0] == '+' || s[0] == '-') { neg = (s[0] == '-'); k = 1; }
    if (k == n) return -1;
    for (; k < n; ++k) {
        if (s[k] < '0' || s[k] > '9') return -1;
        v = v * 10 + (s[k] - '0');
        if (v > 2147483648LL) return -1;
    }
    if (neg) v = -v;
    if (v > 2147483647LL || v < -2147483648LL) return -1;
    *out = (int32_t)v;
    return 0;
}

static int is_ws(char c) { /* char::is_whitespace restricted to ASCII */
    return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v' || c == '\n';
}

/*
 * score441: 21x21 row-major i32 (substitution_matrix.rs:31); aa_index256: the
 * reference holds [u8;255] (byte 255 panics on lookup, :32,75); we keep 256
 * entries and leave entry 255 = 0, callers that need the panic check byte 255.
 */
int orc_parse_ncbi(const char *text, size_t len, int32_t *score441, uint8_t *aa_index256) {
    memset(score441, 0, sizeof(int32_t) * ORC_NCOL * ORC_NCOL); /* :38 */
    memset(aa_index256, 0, 256);                                   /* :38 */
    size_t pos = 0, i = 0;
    while (pos < len) { /* reader.lines() :100 */
        size_t e = pos;
        while (e < len && text[e] != '\n') ++e;
        const char *line = text + pos;
        size_t ll = e - pos;
        pos = e + 1;
        if (ll > 0 && line[ll - 1] == '\r') --ll;            /* lines() strips \r\n */
        if (ll > 0 && (line[0] == '#' || line[0] == ' ')) continue; /* :105 */
        /* split_whitespace :106 */
        const char *tok[64];
        size_t tl[64];
        size_t nt = 0, k = 0;
        while (k < ll) {
            while (k < ll && is_ws(line[k])) ++k;
            if (k >= ll) break;
            size_t s = k;
            while (k < ll && !is_ws(line[k])) ++k;
            if (nt < 64) { tok[nt] = line + s; tl[nt] = k - s; }
            ++nt;
        }
        if (nt > 64) nt = 64;
        if (nt < 23) return ORC_ERR_FORMAT;                   /* :108 */
        uint8_t ch = (uint8_t)tok[0][0];                      /* :109 */
        if (ch == 255) return ORC_ERR_PANIC;                  /* [u8;255] index */
        aa_index256[ch] = (uint8_t)i;                         /* :110 */
        for (size_t j = 1; j < 21; ++j) {                     /* :112-119 */
            int32_t v;
            if (parse_i32(tok[j], tl[j], &v)) return ORC_ERR_PARSE;
            score441[i * ORC_NCOL + (j - 1)] = v;             /* :117 */
            score441[(j - 1) * ORC_NCOL + i] = v;             /* :118 mirrored write */
        }
        int32_t vx;                                           /* :121-126, the X column */
        if (parse_i32(tok[nt - 2], tl[nt - 2], &vx)) return ORC_ERR_PARSE;
        score441[i * ORC_NCOL + 20] = vx;
        score441[20 * ORC_NCOL + i] = vx;
        if (++i == 20) break;                                 /* :127-128 */
    }
    aa_index256['X'] = 20;                                    /* :130 */
    score441[20 * ORC_NCOL + 20] = -1;                        /* :132 */
    return ORC_OK;
}

/* similarity_score.rs:125-134 -- bytes -> matrix indices through the LUT */
void orc_encode(const uint8_t *raw, size_t n, const uint8_t *aa_index256, uint8_t *out) {
    for (size_t k = 0; k < n; ++k) out[k] = aa_index256[raw[k]];
}

/* ------------------------------------------------------------------ */
/* global.rs:17-54  GlobalAligner::new                                 */
/* ------------------------------------------------------------------ */
typedef struct orc_aligner {
    size_t max_length;  /* = max_seq_length + 1, global.rs:37 (row stride of the planes) */
    size_t sentinel_length; /* the (Lmax+1) of global.rs:72; == max_length unless the test
                               harness models a bigger reference aligner with small planes */
    size_t qlen, tlen;
    int32_t recent_score;
    int32_t *H, *E, *F, *Hp, *Ep, *Fp;          /* :43-48 */
    uint8_t *arrows, *e_trace, *f_trace;        /* :49-51, (Lmax+1)^2 bytes each, row-major */
} orc_aligner;

void orc_aligner_free(orc_aligner *a) {
    if (!a) return;
    free(a->H); free(a->E); free(a->F); free(a->Hp); free(a->Ep); free(a->Fp);
    free(a->arrows); free(a->e_trace); free(a->f_trace);
    free(a);
}

orc_aligner *orc_aligner_new(size_t max_seq_length) {
    orc_aligner *a = (orc_aligner *)calloc(1, sizeof(orc_aligner));
    if (!a) return NULL;
    size_t L = max_seq_length + 1;
    a->max_length = L;
    a->sentinel_length = L;
    a->H = (int32_t *)calloc(L, 4); a->E = (int32_t *)calloc(L, 4); a->F = (int32_t *)calloc(L, 4);
    a->Hp = (int32_t *)calloc(L, 4); a->Ep = (int32_t *)calloc(L, 4); a->Fp = (int32_t *)calloc(L, 4);
    a->arrows = (uint8_t *)calloc(L * L, 1);
    a->e_trace = (uint8_t *)calloc(L * L, 1);
    a->f_trace = (uint8_t *)calloc(L * L, 1);
    if (!a->H || !a->E || !a->F || !a->Hp || !a->Ep || !a->Fp || !a->arrows || !a->e_trace ||
        !a->f_trace) {
        orc_aligner_free(a);
        return NULL;
    }
    return a;
}

/*
 * global.rs:57-145  GlobalAligner::align.  q_idx/t_idx are matrix indices
 * (orc_encode).  clear_mode 0 = faithful: clear both trace planes over the
 * whole (Lmax+1)^2 capacity as the reference does (:69-70); 1 = "DP-only":
 * clear just the (n+1)x(m+1) extent this pair touches (same results; used so
 * the baseline can report the clear cost separately).
 */
int orc_aligner_align(orc_aligner *a, const uint8_t *q_idx, size_t n, const uint8_t *t_idx,
                      size_t m, const int32_t *score441, int32_t gap_open, int32_t gap_extend,
                      int clear_mode, int32_t *out_score) {
    if (n >= a->max_length || m >= a->max_length) return ORC_ERR_CAPACITY;
    const size_t L = a->max_length;
    int32_t *H = a->H, *Hp = a->Hp, *E = a->E, *Ep = a->Ep, *F = a->F, *Fp = a->Fp;
    a->qlen = n;   /* :66 */
    a->tlen = m;   /* :67 */
    if (clear_mode == 0) {                      /* :69-70 */
        memset(a->e_trace, 0, L * L);
        memset(a->f_trace, 0, L * L);
    } else {
        for (size_t i = 0; i <= n; ++i) {
            memset(a->e_trace + i * L, 0, m + 1);
            memset(a->f_trace + i * L, 0, m + 1);
        }
    }
    const int32_t impossible = (int32_t)((int64_t)gap_open * (int64_t)a->sentinel_length); /* :72 */

    int32_t tmp = gap_open;                     /* :75 */
    H[0] = 0; E[0] = 0; F[0] = 0;               /* :76-79 */
    for (size_t j = 1; j <= m; ++j) {           /* :81-88 */
        F[j] = impossible;
        H[j] = tmp;
        E[j] = tmp;
        a->arrows[j] = 1;
        a->e_trace[j] = 1;
        tmp += gap_extend;
    }
    int32_t gap_started_in_q = gap_open;        /* :90 */
    for (size_t i = 1; i <= n; ++i) {           /* :91 */
        int32_t *s;
        s = H; H = Hp; Hp = s;                  /* :92-94 */
        s = E; E = Ep; Ep = s;
        s = F; F = Fp; Fp = s;
        uint8_t *arr = a->arrows + i * L, *et = a->e_trace + i * L, *ft = a->f_trace + i * L;
        arr[0] = 4;                             /* :96 */
        ft[0] = 1;                              /* :97 */
        H[0] = gap_started_in_q;                /* :99-101 */
        F[0] = gap_started_in_q;
        E[0] = impossible;
        const int32_t *srow = score441 + (size_t)q_idx[i - 1] * ORC_NCOL;
        for (size_t j = 1; j <= m; ++j) {       /* :103-139 */
            int32_t e_e = E[j - 1] + gap_extend;
            int32_t e_h = H[j - 1] + gap_open;
            int32_t e_f = F[j - 1] + gap_open;
            if (e_e >= e_h && e_e >= e_f) { E[j] = e_e; et[j] = 1; }          /* :109-111 */
            else { E[j] = e_h > e_f ? e_h : e_f; et[j] = 0; }                  /* :112-115 */
            int32_t f_f = Fp[j] + gap_extend;
            int32_t f_h = Hp[j] + gap_open;
            int32_t f_e = Ep[j] + gap_open;
            if (f_f >= f_h && f_f >= f_e) { F[j] = f_f; ft[j] = 1; }          /* :122-124 */
            else { F[j] = f_h > f_e ? f_h : f_e; ft[j] = 0; }                  /* :125-128 */
            int32_t h_diag = Hp[j - 1] + srow[t_idx[j - 1]];                   /* :131 */
            int32_t h = h_diag > E[j] ? h_diag : E[j];                         /* :132 */
            if (F[j] > h) h = F[j];
            H[j] = h;
            uint8_t fl = 0;                                                    /* :134-138 */
            if (h == E[j]) fl += 1;
            if (h == h_diag) fl += 2;
            if (h == F[j]) fl += 4;
            arr[j] = fl;
        }
        gap_started_in_q += gap_extend;         /* :140 */
    }
    a->H = H; a->Hp = Hp; a->E = E; a->Ep = Ep; a->F = F; a->Fp = Fp;
    a->recent_score = H[m];                     /* :143 */
    if (out_score) *out_score = H[m];
    return ORC_OK;
}

/*
 * global.rs:146-201  GlobalAligner::backtrace.  Writes the path glyphs
 * ('*' Match, '-' Horizontal, '|' Vertical; alignment_path.rs:40-47) in
 * forward order into path (capacity >= n+m), returns the length or <0.
 */
int64_t orc_aligner_backtrace(const orc_aligner *a, uint8_t *path) {
    const size_t L = a->max_length;
    size_t i = a->qlen, j = a->tlen;
    int state = 0; /* 0=H 1=E 2=F */
    size_t len = 0;
    const size_t cap = a->qlen + a->tlen;
    while (i > 0 || j > 0) {                    /* :156 */
        if (state == 0) {
            uint8_t ar = a->arrows[i * L + j];  /* :159 */
            if (ar & 2) {                       /* :161-165 */
                if (i == 0 || j == 0) return ORC_ERR_PANIC; /* usize underflow */
                if (len >= cap) return ORC_ERR_PANIC;
                path[len++] = '*'; --i; --j;
            } else if (ar & 1) state = 1;       /* :166-167 */
            else if (ar & 4) state = 2;         /* :168-169 */
            else return ORC_ERR_PANIC;          /* :171 */
        } else if (state == 1) {                /* :175-184 */
            if (j == 0) return ORC_ERR_PANIC;   /* j -= 1 underflow */
            if (len >= cap) return ORC_ERR_PANIC;
            path[len++] = '-';
            --j;
            state = (a->e_trace[i * L + (j + 1)] == 1) ? 1 : 0;
        } else {                                /* :186-195 */
            if (i == 0) return ORC_ERR_PANIC;
            if (len >= cap) return ORC_ERR_PANIC;
            path[len++] = '|';
            --i;
            state = (a->f_trace[(i + 1) * L + j] == 1) ? 2 : 0;
        }
    }
    for (size_t k = 0; k < len / 2; ++k) {      /* :199 reverse */
        uint8_t c = path[k]; path[k] = path[len - 1 - k]; path[len - 1 - k] = c;
    }
    return (int64_t)len;
}

/*
 * alignment_path.rs:117-139 aligned_symbols (raw bytes, gap '-'), followed by
 * msa.rs:261-269 sum_identical and sequence.rs:532-534 len_ungapped on the two
 * expanded strings.  aq/at may be NULL when only the statistics are wanted.
 */
int orc_expand_and_count(const uint8_t *path, size_t plen, const uint8_t *q_raw, size_t n,
                         const uint8_t *t_raw, size_t m, uint8_t gap, uint8_t *aq, uint8_t *at,
                         uint64_t *n_identical, uint64_t *len_q, uint64_t *len_t) {
    size_t qi = 0, ti = 0;
    uint64_t nid = 0, lq = 0, lt = 0;
    for (size_t k = 0; k < plen; ++k) {
        uint8_t a, b;
        if (path[k] == '-') {                 /* Horizontal: gap in the query */
            if (ti >= m) return ORC_ERR_PANIC;
            a = gap; b = t_raw[ti++];
        } else if (path[k] == '|') {          /* Vertical: gap in the template */
            if (qi >= n) return ORC_ERR_PANIC;
            a = q_raw[qi++]; b = gap;
        } else if (path[k] == '*') {
            if (qi >= n || ti >= m) return ORC_ERR_PANIC;
            a = q_raw[qi++]; b = t_raw[ti++];
        } else return ORC_ERR_FORMAT;
        if (aq) aq[k] = a;
        if (at) at[k] = b;
        if (a == b && a != '-' && a != '_') ++nid;      /* msa.rs:264 */
        if (a != '-' && a != '_') ++lq;                 /* sequence.rs:533 */
        if (b != '-' && b != '_') ++lt;
    }
    if (n_identical) *n_identical = nid;
    if (len_q) *len_q = lq;
    if (len_t) *len_t = lt;
    return ORC_OK;
}

/* alignment_statistics.rs:71-73 ; `as f32` of bin/cluster_sequences.rs:128 */
double orc_percent_identity(uint64_t n_identical, uint64_t len_q, uint64_t len_t) {
    uint64_t mn = len_q < len_t ? len_q : len_t;
    return (double)n_identical / (double)mn * 100.0;
}
float orc_percent_identity_f32(uint64_t n_identical, uint64_t len_q, uint64_t len_t) {
    return (float)orc_percent_identity(n_identical, len_q, len_t);
}

/*
 * One pair, start to finish, on raw bytes: encode (similarity_score.rs:125-134),
 * align, backtrace, expand, count.  lmax is the aligner capacity the reference
 * would have used (alignment_protocols.rs:86-88); it only enters through the
 * sentinel.  path/aq/at need capacity n+m (any may be NULL).
 */
int orc_align_pair(const uint8_t *q_raw, size_t n, const uint8_t *t_raw, size_t m,
                   const int32_t *score441, const uint8_t *aa_index256, int32_t gap_open,
                   int32_t gap_extend, size_t lmax, int32_t *out_score, uint8_t *path,
                   int64_t *path_len, uint8_t *aq, uint8_t *at, uint64_t *n_identical,
                   uint64_t *len_q, uint64_t *len_t) {
    size_t cap = n > m ? n : m;
    if (lmax < cap) return ORC_ERR_CAPACITY;
    /* planes sized for this pair; sentinel as if the aligner had capacity lmax */
    orc_aligner *a = orc_aligner_new(cap);
    if (!a) return ORC_ERR_ALLOC;
    a->sentinel_length = lmax + 1;
    uint8_t *qi = (uint8_t *)malloc(n + 1), *ti = (uint8_t *)malloc(m + 1);
    uint8_t *p = path ? path : (uint8_t *)malloc(n + m + 1);
    int rc = ORC_OK;
    if (!qi || !ti || !p) { rc = ORC_ERR_ALLOC; goto done; }
    orc_encode(q_raw, n, aa_index256, qi);
    orc_encode(t_raw, m, aa_index256, ti);
    rc = orc_aligner_align(a, qi, n, ti, m, score441, gap_open, gap_extend, 1, out_score);
    if (rc) goto done;
    {
        int64_t pl = orc_aligner_backtrace(a, p);
        if (pl < 0) { rc = (int)pl; goto done; }
        if (path_len) *path_len = pl;
        rc = orc_expand_and_count(p, (size_t)pl, q_raw, n, t_raw, m, '-', aq, at, n_identical,
                                  len_q, len_t);
    }
done:
    free(qi); free(ti);
    if (!path) free(p);
    orc_aligner_free(a);
    return rc;
}

/* ------------------------------------------------------------------ */
/* alignment_protocols.rs:83-115  align_all_pairs                      */
/* ------------------------------------------------------------------ */

/* sequence.rs:8 `#[derive(PartialEq)]`: description AND bytes must match. */
static int seq_eq(const uint8_t *ra, uint64_t na, const uint8_t *da, uint64_t dna,
                  const uint8_t *rb, uint64_t nb, const uint8_t *db, uint64_t dnb) {
    return na == nb && dna == dnb && memcmp(da, db, dna) == 0 && memcmp(ra, rb, na) == 0;
}

typedef struct orc_seqset {           /* a Vec<Sequence>: packed residues + packed descriptions */
    const uint8_t *res;  const uint64_t *res_off;   /* n+1 offsets */
    const uint8_t *desc; const uint64_t *desc_off;  /* n+1 offsets */
    uint32_t n;
} orc_seqset;

/*
 * How many queries the inner loop of alignment_protocols.rs:96-97 visits for
 * template t before the triangle `break` (nq when not triangle / no match).
 */
static uint32_t inner_count(const orc_seqset *Q, const orc_seqset *T, uint32_t t, int triangle) {
    if (!triangle) return Q->n;
    for (uint32_t q = 0; q < Q->n; ++q) {
        if (seq_eq(T->res + T->res_off[t], T->res_off[t + 1] - T->res_off[t],
                   T->desc + T->desc_off[t], T->desc_off[t + 1] - T->desc_off[t],
                   Q->res + Q->res_off[q], Q->res_off[q + 1] - Q->res_off[q],
                   Q->desc + Q->desc_off[q], Q->desc_off[q + 1] - Q->desc_off[q]))
            return q;
    }
    return Q->n;
}

/* Fills first[t] (t = 0..nt) with the report index of pair (q=0, t); returns the pair count. */
uint64_t orc_all_pairs_layout(const orc_seqset *Q, const orc_seqset *T, int triangle,
                              uint64_t *first /* nt+1 */) {
    uint64_t k = 0;
    for (uint32_t t = 0; t < T->n; ++t) {
        first[t] = k;
        k += inner_count(Q, T, t, triangle);
    }
    first[T->n] = k;
    return k;
}

typedef struct {
    const orc_seqset *Q, *T;
    const int32_t *score441;
    const uint8_t *aa_index256;
    int32_t go, ge;
    size_t lmax;
    int clear_mode, with_backtrace;
    const uint64_t *first;
    int32_t *scores; uint32_t *n_identical; uint32_t *len_q; uint32_t *len_t; float *identity;
    uint32_t *out_q, *out_t;
    uint32_t next_t;            /* dynamic dealing of templates over threads */
    pthread_mutex_t mu;
    int rc;
    double cells;               /* alignment_protocols.rs:104 */
} ap_job;

static void *ap_worker(void *arg) {
    ap_job *J = (ap_job *)arg;
    orc_aligner *a = orc_aligner_new(J->lmax);              /* alignment_protocols.rs:88 */
    uint8_t *ti = (uint8_t *)malloc(J->lmax + 1), *qi = (uint8_t *)malloc(J->lmax + 1);
    uint8_t *path = (uint8_t *)malloc(2 * J->lmax + 2);
    double cells = 0.0;
    int rc = (a && ti && qi && path) ? ORC_OK : ORC_ERR_ALLOC;
    while (rc == ORC_OK) {
        pthread_mutex_lock(&J->mu);
        uint32_t t = J->next_t < J->T->n ? J->next_t++ : UINT32_MAX;
        pthread_mutex_unlock(&J->mu);
        if (t == UINT32_MAX) break;
        const uint8_t *tr = J->T->res + J->T->res_off[t];
        size_t m = J->T->res_off[t + 1] - J->T->res_off[t];
        orc_encode(tr, m, J->aa_index256, ti);              /* :95 */
        uint64_t cnt = J->first[t + 1] - J->first[t];
        for (uint32_t q = 0; q < cnt; ++q) {                /* :96-97 */
            const uint8_t *qr = J->Q->res + J->Q->res_off[q];
            size_t n = J->Q->res_off[q + 1] - J->Q->res_off[q];
            orc_encode(qr, n, J->aa_index256, qi);          /* :98 */
            int32_t sc;
            rc = orc_aligner_align(a, qi, n, ti, m, J->score441, J->go, J->ge, J->clear_mode,
                                   &sc);                    /* :99 */
            if (rc) break;
            uint64_t k = J->first[t] + q, nid = 0, lq = 0, lt = 0;
            if (J->with_backtrace) {
                int64_t pl = orc_aligner_backtrace(a, path); /* :100 */
                if (pl < 0) { rc = (int)pl; break; }
                rc = orc_expand_and_count(path, (size_t)pl, qr, n, tr, m, '-', NULL, NULL, &nid,
                                          &lq, &lt);        /* :101-102 + reporter */
                if (rc) break;
            }
            if (J->scores) J->scores[k] = sc;
            if (J->n_identical) J->n_identical[k] = (uint32_t)nid;
            if (J->len_q) J->len_q[k] = (uint32_t)lq;
            if (J->len_t) J->len_t[k] = (uint32_t)lt;
            if (J->identity) J->identity[k] = orc_percent_identity_f32(nid, lq, lt);
            if (J->out_q) J->out_q[k] = q;
            if (J->out_t) J->out_t[k] = t;
            cells += (double)n * (double)m;                 /* :104 */
        }
    }
    pthread_mutex_lock(&J->mu);
    if (rc && !J->rc) J->rc = rc;
    J->cells += cells;
    pthread_mutex_unlock(&J->mu);
    free(ti); free(qi); free(path);
    orc_aligner_free(a);
    return NULL;
}

/*
 * align_all_pairs with the SequenceIdentityMatrix reporter folded in
 * (bin/cluster_sequences.rs:122-130): outputs are indexed by report order
 * (t-major; see orc_all_pairs_layout).  n_threads = 1 reproduces the
 * reference's single-threaded loop; n_threads > 1 deals templates over
 * threads with one aligner each (as bucket_clustering.rs:222 does) -- the
 * stand-in for the north-star's "rayon" baseline.  Any output may be NULL.
 * with_backtrace = 0 skips traceback/identity (score-only timing).
 */
int orc_align_all_pairs(const orc_seqset *Q, const orc_seqset *T, const int32_t *score441,
                        const uint8_t *aa_index256, int32_t gap_open, int32_t gap_extend,
                        int triangle, int clear_mode, int with_backtrace, int n_threads,
                        const uint64_t *first, int32_t *scores, uint32_t *n_identical,
                        uint32_t *len_q, uint32_t *len_t, float *identity, uint32_t *out_q,
                        uint32_t *out_t, double *cells) {
    (void)triangle; /* the triangle rule is already folded into `first` (orc_all_pairs_layout) */
    if (Q->n == 0 || T->n == 0) return ORC_ERR_PANIC;     /* max().unwrap() :86-87 */
    size_t lmax = 0;
    for (uint32_t i = 0; i < Q->n; ++i)
        if (Q->res_off[i + 1] - Q->res_off[i] > lmax) lmax = Q->res_off[i + 1] - Q->res_off[i];
    for (uint32_t i = 0; i < T->n; ++i)
        if (T->res_off[i + 1] - T->res_off[i] > lmax) lmax = T->res_off[i + 1] - T->res_off[i];
    ap_job J;
    memset(&J, 0, sizeof(J));
    J.Q = Q; J.T = T; J.score441 = score441; J.aa_index256 = aa_index256;
    J.go = gap_open; J.ge = gap_extend; J.lmax = lmax; J.clear_mode = clear_mode;
    J.with_backtrace = with_backtrace; J.first = first;
    J.scores = scores; J.n_identical = n_identical; J.len_q = len_q; J.len_t = len_t;
    J.identity = identity; J.out_q = out_q; J.out_t = out_t;
    pthread_mutex_init(&J.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    if (n_threads == 1) {
        ap_worker(&J);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
        int started = 0;
        for (int i = 0; i < n_threads; ++i)
            if (pthread_create(&th[i], NULL, ap_worker, &J) == 0) ++started; else break;
        if (started == 0) ap_worker(&J);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&J.mu);
    if (cells) *cells = J.cells;
    return J.rc;
}

/*
 * A bounded sample of the same loop for the CPU baseline: aligns the listed
 * (q,t) pairs only, with one aligner of capacity lmax per thread (so the
 * faithful clear cost is that of the full data set, not of the sample).
 */
typedef struct {
    const orc_seqset *Q, *T;
    const int32_t *score441; const uint8_t *aa_index256;
    int32_t go, ge; size_t lmax; int clear_mode, with_backtrace;
    const uint32_t *pq, *pt; uint64_t n_pairs, next; uint64_t grain;
    int32_t *scores; uint32_t *n_identical;
    pthread_mutex_t mu; int rc; double cells;
} sp_job;

static void *sp_worker(void *arg) {
    sp_job *J = (sp_job *)arg;
    orc_aligner *a = orc_aligner_new(J->lmax);
    uint8_t *ti = (uint8_t *)malloc(J->lmax + 1), *qi = (uint8_t *)malloc(J->lmax + 1);
    uint8_t *path = (uint8_t *)malloc(2 * J->lmax + 2);
    double cells = 0.0;
    int rc = (a && ti && qi && path) ? ORC_OK : ORC_ERR_ALLOC;
    while (rc == ORC_OK) {
        pthread_mutex_lock(&J->mu);
        uint64_t b = J->next; J->next += J->grain;
        pthread_mutex_unlock(&J->mu);
        if (b >= J->n_pairs) break;
        uint64_t e = b + J->grain < J->n_pairs ? b + J->grain : J->n_pairs;
        for (uint64_t k = b; k < e && rc == ORC_OK; ++k) {
            uint32_t q = J->pq[k], t = J->pt[k];
            const uint8_t *qr = J->Q->res + J->Q->res_off[q], *tr = J->T->res + J->T->res_off[t];
            size_t n = J->Q->res_off[q + 1] - J->Q->res_off[q];
            size_t m = J->T->res_off[t + 1] - J->T->res_off[t];
            orc_encode(tr, m, J->aa_index256, ti);
            orc_encode(qr, n, J->aa_index256, qi);
            int32_t sc; uint64_t nid = 0;
            rc = orc_aligner_align(a, qi, n, ti, m, J->score441, J->go, J->ge, J->clear_mode, &sc);
            if (rc) break;
            if (J->with_backtrace) {
                int64_t pl = orc_aligner_backtrace(a, path);
                if (pl < 0) { rc = (int)pl; break; }
                rc = orc_expand_and_count(path, (size_t)pl, qr, n, tr, m, '-', NULL, NULL, &nid,
                                          NULL, NULL);
            }
            if (J->scores) J->scores[k] = sc;
            if (J->n_identical) J->n_identical[k] = (uint32_t)nid;
            cells += (double)n * (double)m;
        }
    }
    pthread_mutex_lock(&J->mu);
    if (rc && !J->rc) J->rc = rc;
    J->cells += cells;
    pthread_mutex_unlock(&J->mu);
    free(ti); free(qi); free(path);
    orc_aligner_free(a);
    return NULL;
}

int orc_align_pair_list(const orc_seqset *Q, const orc_seqset *T, const int32_t *score441,
                        const uint8_t *aa_index256, int32_t gap_open, int32_t gap_extend,
                        uint64_t lmax, int clear_mode, int with_backtrace, int n_threads,
                        const uint32_t *pq, const uint32_t *pt, uint64_t n_pairs,
                        int32_t *scores, uint32_t *n_identical, double *cells) {
    sp_job J;
    memset(&J, 0, sizeof(J));
    J.Q = Q; J.T = T; J.score441 = score441; J.aa_index256 = aa_index256;
    J.go = gap_open; J.ge = gap_extend; J.lmax = (size_t)lmax; J.clear_mode = clear_mode;
    J.with_backtrace = with_backtrace; J.pq = pq; J.pt = pt; J.n_pairs = n_pairs;
    J.scores = scores; J.n_identical = n_identical;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    J.grain = n_pairs / ((uint64_t)n_threads * 64) + 1;
    pthread_mutex_init(&J.mu, NULL);
    if (n_threads == 1) {
        sp_worker(&J);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
        int started = 0;
        for (int i = 0; i < n_threads; ++i)
            if (pthread_create(&th[i], NULL, sp_worker, &J) == 0) ++started; else break;
        if (started == 0) sp_worker(&J);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&J.mu);
    if (cells) *cells = J.cells;
    return J.rc;
}
