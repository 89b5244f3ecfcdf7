/*
 * synth.c -- deterministic synthetic protein sets.  Host-only
 * utility for tests and bench.py: there is no network, so the benchmark inputs are
 * generated.  Residues follow the Swiss-Prot background over the 20 standard amino
 * acids; 25 % of the sequences are "homologs" (a mutated copy of an earlier
 * sequence: substitutions at rate U[0.05,0.6], indels at 2 %/site with geometric
 * lengths of mean 3) so that gaps, ties and high identities occur.
 *
 * RNG: splitmix64, u = (x >> 11) * 2^-53, one stream per sequence
 * (state = seed * 0x9E3779B97F4A7C15 + (index + 1) * 0xD1B54A32D192ED03).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const char AA[21] = "ARNDCQEGHILKMFPSTWYV";
/* Swiss-Prot composition in percent, same order as AA */
static const double FREQ[20] = {8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96,
                                9.66, 5.84, 2.42, 3.86, 4.70, 6.56, 5.34, 1.08, 2.92, 6.87};

static uint8_t draw_residue(uint64_t *s, const double *cum) {
    double u = u01(s);
    int k = 0;
    while (k < 19 && u >= cum[k]) ++k;
    return (uint8_t)AA[k];
}

static uint32_t geometric(uint64_t *s, double mean) {
    /* P(len = k) = (1-p)^(k-1) p, p = 1/mean */
    double p = 1.0 / mean, u = u01(s);
    uint32_t k = 1 + (uint32_t)floor(log(1.0 - u) / log(1.0 - p));
    return k < 1 ? 1 : (k > 50 ? 50 : k);
}

/*
 * dist 0: length ~ U{lo..hi};  dist 1: length = clip(round(exp(N(mu, sigma))), lo, hi).
 * Two-call protocol: pass res == NULL to get the total residue count (offsets are
 * filled, n+1 entries); then call again with the buffer.  Both calls are deterministic.
 */
uint64_t bsa_synth_generate(uint64_t seed, uint32_t n, int dist, uint32_t lo, uint32_t hi, double mu,
                            double sigma, double homolog_fraction, uint8_t *res, uint64_t *off) {
    double cum[20], tot = 0.0;
    for (int i = 0; i < 20; ++i) tot += FREQ[i];
    double acc = 0.0;
    for (int i = 0; i < 20; ++i) { acc += FREQ[i] / tot; cum[i] = acc; }
    /* homologs copy earlier sequences, so a scratch copy of everything is kept */
    uint64_t cap = (uint64_t)n * (hi < 64 ? 64 : hi > 600 ? 600 : hi) + 1024, used = 0;
    uint8_t *buf = (uint8_t *)malloc(cap);
    uint64_t *o = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n + 1));
    uint8_t *tmp = (uint8_t *)malloc((size_t)hi + 64);
    if (!buf || !o || !tmp) { free(buf); free(o); free(tmp); return 0; }
    o[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t st = seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)i + 1) * 0xD1B54A32D192ED03ULL;
        uint32_t len;
        if (dist == 0) {
            len = lo + (uint32_t)(u01(&st) * (double)(hi - lo + 1));
            if (len > hi) len = hi;
        } else {
            double u1 = u01(&st), u2 = u01(&st);
            if (u1 < 1e-300) u1 = 1e-300;
            double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            double L = floor(exp(mu + sigma * z) + 0.5);
            len = L < lo ? lo : (L > hi ? hi : (uint32_t)L);
        }
        uint32_t outl = 0;
        int homolog = (i > 0) && (u01(&st) < homolog_fraction);
        if (homolog) {
            uint32_t p = (uint32_t)(u01(&st) * (double)i);
            if (p >= i) p = i - 1;
            const uint8_t *par = buf + o[p];
            uint32_t pl = (uint32_t)(o[p + 1] - o[p]);
            double rate = 0.05 + 0.55 * u01(&st);
            uint32_t k = 0;
            while (k < pl && outl < hi) {
                double u = u01(&st);
                if (u < 0.01) {
                    k += geometric(&st, 3.0);                       /* deletion */
                } else if (u < 0.02) {
                    uint32_t g = geometric(&st, 3.0);               /* insertion */
                    for (uint32_t x = 0; x < g && outl < hi; ++x) tmp[outl++] = draw_residue(&st, cum);
                } else {
                    tmp[outl++] = (u01(&st) < rate) ? draw_residue(&st, cum) : par[k];
                    ++k;
                }
            }
            while (outl < lo) tmp[outl++] = draw_residue(&st, cum);
        } else {
            for (uint32_t x = 0; x < len; ++x) tmp[outl++] = draw_residue(&st, cum);
        }
        if (used + outl > cap) {
            cap = (used + outl) * 2;
            uint8_t *nb = (uint8_t *)realloc(buf, cap);
            if (!nb) { free(buf); free(o); free(tmp); return 0; }
            buf = nb;
        }
        memcpy(buf + used, tmp, outl);
        used += outl;
        o[i + 1] = used;
    }
    if (off) memcpy(off, o, sizeof(uint64_t) * ((size_t)n + 1));
    if (res) memcpy(res, buf, used);
    free(buf); free(o); free(tmp);
    return used;
}

/*
 * BASELINE configs[4]: n_pairs titin-scale pairs.  Sequence 2p is the
 * query of pair p, sequence 2p+1 its template; lengths ~ U{lo..hi}.  Even pairs are HOMOLOG
 * pairs (the query is a mutated copy of the template: substitutions at rate U[0.05,0.6],
 * indels at 2 %/site, geometric lengths of mean 3), odd pairs are unrelated.  Pair 0 is forced
 * to fixed_q x fixed_t residues (34,350 x 35,000: titin against a 35,000-residue template)
 * when fixed_t != 0: the mutated copy is cut, or extended with random residues, to fixed_q.
 * Same two-call protocol as bsa_synth_generate.
 */
uint64_t bsa_synth_pair_set(uint64_t seed, uint32_t n_pairs, uint32_t lo, uint32_t hi, uint32_t fixed_q,
                            uint32_t fixed_t, uint8_t *res, uint64_t *off) {
    double cum[20], tot = 0.0;
    for (int i = 0; i < 20; ++i) tot += FREQ[i];
    double acc = 0.0;
    for (int i = 0; i < 20; ++i) { acc += FREQ[i] / tot; cum[i] = acc; }
    uint32_t cap = (hi > fixed_t ? hi : fixed_t);
    if (fixed_q > cap) cap = fixed_q;
    uint8_t *tq = (uint8_t *)malloc((size_t)cap + 64), *tt = (uint8_t *)malloc((size_t)cap + 64);
    if (!tq || !tt) { free(tq); free(tt); return 0; }
    uint64_t used = 0;
    if (off) off[0] = 0;
    for (uint32_t p = 0; p < n_pairs; ++p) {
        uint64_t st = seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)p + 1) * 0xD1B54A32D192ED03ULL;
        uint32_t lt = lo + (uint32_t)(u01(&st) * (double)(hi - lo + 1));
        uint32_t lq = lo + (uint32_t)(u01(&st) * (double)(hi - lo + 1));
        if (lt > hi) lt = hi;
        if (lq > hi) lq = hi;
        const int fixed = (p == 0 && fixed_t != 0);
        if (fixed) { lt = fixed_t; lq = fixed_q; }
        for (uint32_t x = 0; x < lt; ++x) tt[x] = draw_residue(&st, cum);
        uint32_t outl = 0;
        if ((p & 1u) == 0) {
            const uint32_t lim = fixed ? lq : cap;
            double rate = 0.05 + 0.55 * u01(&st);
            uint32_t k = 0;
            while (k < lt && outl < lim) {
                double u = u01(&st);
                if (u < 0.01) {
                    k += geometric(&st, 3.0);
                } else if (u < 0.02) {
                    uint32_t g = geometric(&st, 3.0);
                    for (uint32_t x = 0; x < g && outl < lim; ++x) tq[outl++] = draw_residue(&st, cum);
                } else {
                    tq[outl++] = (u01(&st) < rate) ? draw_residue(&st, cum) : tt[k];
                    ++k;
                }
            }
            while (outl < (fixed ? lq : lo)) tq[outl++] = draw_residue(&st, cum);
        } else {
            for (uint32_t x = 0; x < lq; ++x) tq[outl++] = draw_residue(&st, cum);
        }
        if (res) { memcpy(res + used, tq, outl); memcpy(res + used + outl, tt, lt); }
        if (off) { off[2 * p + 1] = used + outl; off[2 * p + 2] = used + outl + lt; }
        used += (uint64_t)outl + lt;
    }
    free(tq); free(tt);
    return used;
}
