/*
 * local_oracle.c -- CPU ORACLE for SURVEY.md 8(f) rank 4: local alignment (Smith-Waterman with
 * Gotoh matrices) as bioshell-seq's LocalAlignment does it.  TEST INFRASTRUCTURE.
 *
 * Literal restatement of bioshell-seq/src/alignment/local.rs:83-273 (align :83-207,
 * backtrace :213-273): STOP semantics at <= 0, three trace planes, first-strict-maximum best
 * cell in row-major order, H/E/F candidates for the best cell.
 * Parity status: PINNED by bioshell-seq/tests/test_aligners.rs:73-149 (3 cases x 2
 * orientations; tests/golden/ref_kats.json).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NCOL 21

/* q_idx/t_idx: matrix indices (orc_encode).  path gets the glyphs in forward order (capacity
 * n+m).  Returns path length, or <0 (-3: the reference would panic). */
int64_t orc_local_align(const uint8_t *q_idx, size_t n, const uint8_t *t_idx, size_t m,
                        const int32_t *score441, int32_t gap_open, int32_t gap_extend,
                        int32_t *out_score, uint64_t *end_q, uint64_t *end_t, uint64_t *start_q,
                        uint64_t *start_t, uint8_t *path) {
    const size_t L = (n > m ? n : m) + 1;
    int32_t *H = calloc(L, 4), *E = calloc(L, 4), *F = calloc(L, 4);
    int32_t *Hp = calloc(L, 4), *Ep = calloc(L, 4), *Fp = calloc(L, 4);
    uint8_t *arrows = calloc((n + 1) * (m + 1), 1), *et = calloc((n + 1) * (m + 1), 1),
            *ft = calloc((n + 1) * (m + 1), 1);
    if (!H || !E || !F || !Hp || !Ep || !Fp || !arrows || !et || !ft) return -4;
    const size_t W = m + 1;
    int32_t recent = 0;                       /* local.rs:95-98 */
    size_t best_i = 0, best_j = 0;
    int best_state = 0;
    for (size_t i = 1; i <= n; ++i) {         /* :111 */
        int32_t *s;
        s = H; H = Hp; Hp = s;
        s = E; E = Ep; Ep = s;
        s = F; F = Fp; Fp = s;
        H[0] = 0; E[0] = 0; F[0] = 0;         /* :116-118 */
        const int32_t *srow = score441 + (size_t)q_idx[i - 1] * ORC_NCOL;
        for (size_t j = 1; j <= m; ++j) {     /* :120 */
            int32_t e_e = E[j - 1] + gap_extend, e_h = H[j - 1] + gap_open, e_f = F[j - 1] + gap_open;
            int32_t e_open = e_h > e_f ? e_h : e_f;
            if (e_e > 0 && e_e >= e_open) { E[j] = e_e; et[i * W + j] = 1; }       /* :127-129 */
            else if (e_open > 0) { E[j] = e_open; et[i * W + j] = 2; }             /* :130-132 */
            else { E[j] = 0; et[i * W + j] = 0; }                                  /* :133-136 */
            int32_t f_f = Fp[j] + gap_extend, f_h = Hp[j] + gap_open, f_e = Ep[j] + gap_open;
            int32_t f_open = f_h > f_e ? f_h : f_e;
            if (f_f > 0 && f_f >= f_open) { F[j] = f_f; ft[i * W + j] = 1; }
            else if (f_open > 0) { F[j] = f_open; ft[i * W + j] = 2; }
            else { F[j] = 0; ft[i * W + j] = 0; }
            int32_t h_diag = Hp[j - 1] + srow[t_idx[j - 1]];                       /* :156 */
            int32_t h_val = 0;
            uint8_t fl = 0;
            if (h_diag > h_val) { h_val = h_diag; fl = 2; }                        /* :161-166 */
            else if (h_diag == h_val && h_val > 0) fl |= 2;
            if (E[j] > h_val) { h_val = E[j]; fl = 1; }                            /* :168-173 */
            else if (E[j] == h_val && h_val > 0) fl |= 1;
            if (F[j] > h_val) { h_val = F[j]; fl = 4; }                            /* :175-180 */
            else if (F[j] == h_val && h_val > 0) fl |= 4;
            H[j] = h_val;
            arrows[i * W + j] = fl;
            if (H[j] > recent) { recent = H[j]; best_i = i; best_j = j; best_state = 0; }   /* :185-202 */
            if (E[j] > recent) { recent = E[j]; best_i = i; best_j = j; best_state = 1; }
            if (F[j] > recent) { recent = F[j]; best_i = i; best_j = j; best_state = 2; }
        }
    }
    /* backtrace :213-273 */
    size_t i = best_i, j = best_j, len = 0;
    int state = best_state;
    int64_t rc = 0;
    for (;;) {
        if (state == 0) {
            uint8_t a = arrows[i * W + j];
            if (a == 0) break;
            if (a & 2) { if (i == 0 || j == 0) { rc = -3; break; } path[len++] = '*'; --i; --j; }
            else if (a & 1) state = 1;
            else if (a & 4) state = 2;
            else { rc = -3; break; }
        } else if (state == 1) {
            uint8_t t = et[i * W + j];
            if (t == 0) break;
            if (j == 0) { rc = -3; break; }
            path[len++] = '-'; --j;
            state = t == 1 ? 1 : 0;
        } else {
            uint8_t t = ft[i * W + j];
            if (t == 0) break;
            if (i == 0) { rc = -3; break; }
            path[len++] = '|'; --i;
            state = t == 1 ? 2 : 0;
        }
    }
    for (size_t k = 0; k < len / 2; ++k) { uint8_t c = path[k]; path[k] = path[len - 1 - k]; path[len - 1 - k] = c; }
    if (out_score) *out_score = recent;
    if (end_q) *end_q = best_i;
    if (end_t) *end_t = best_j;
    if (start_q) *start_q = i;
    if (start_t) *start_t = j;
    free(H); free(E); free(F); free(Hp); free(Ep); free(Fp); free(arrows); free(et); free(ft);
    return rc < 0 ? rc : (int64_t)len;
}
