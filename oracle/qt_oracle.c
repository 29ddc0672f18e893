/*
 * qt_oracle.c — CPU ORACLE (test infrastructure only; see qt_oracle.h).
 *
 * n-generic restatement of the reference's CPU negacyclic-polymul functions.  The
 * reference (benlwk/ntt-gpu-qTESLA, NTT.cu) fixes n=1024 ("level < 10", NTT.cu:1064),
 * BATCH=2 and P=8404993 at compile time; here log2(n), B and q are run-time values and
 * every reduction is a plain `% q` like the reference's CPU twins, so all results are
 * canonical in [0,q).
 */
#include "qt_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* parameter sets                                                                        */
/* ------------------------------------------------------------------------------------ */

static uint32_t powmod(uint32_t b, uint64_t e, uint32_t q) {
    uint64_t r = 1, x = b % q;
    while (e) {
        if (e & 1) r = r * x % q;
        x = x * x % q;
        e >>= 1;
    }
    return (uint32_t)r;
}

static const struct { uint32_t n, q, psi; } k_sets[QTO_NUM_SETS] = {
    {512, 4205569u, 0},      /* psi derived: g^((q-1)/2n), smallest working g (=17) */
    {1024, 8404993u, 2083362u}, /* psi pinned: Phi[1] of constants.h:11 (main.cu:26 comment) */
    {1024, 343576577u, 0},   /* g = 3 */
    {2048, 856145921u, 0},   /* g = 3 */
};

int qto_get_params(int set, qto_params* p) {
    if (set < 0 || set >= QTO_NUM_SETS || !p) return -1;
    uint32_t n = k_sets[set].n, q = k_sets[set].q, psi = k_sets[set].psi;
    uint32_t logn = 0;
    while ((1u << logn) < n) logn++;
    if (!psi) {
        for (uint32_t g = 2;; g++) {
            psi = powmod(g, (q - 1) / (2 * n), q);
            if (powmod(psi, n, q) == q - 1) break; /* order exactly 2n */
        }
    }
    p->set = set;
    p->n = n;
    p->logn = logn;
    p->q = q;
    p->psi = psi;
    p->psi_inv = powmod(psi, q - 2, q);
    p->omega = (uint32_t)((uint64_t)psi * psi % q);
    p->omega_inv = powmod(p->omega, q - 2, q);
    p->n_inv = powmod(n, q - 2, q);
    /* -q^-1 mod 2^32 by Newton iteration */
    uint32_t inv = q; /* q*q == 1 mod 8 */
    for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
    p->qinv_neg = 0u - inv;
    p->barrett_mu48 = (uint32_t)((1ull << 48) / q);
    return 0;
}

uint32_t qto_bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

int qto_tables(int set, uint32_t* bitrev, uint32_t* Phi, uint32_t* invPhi, uint32_t* tf0,
               uint32_t* ti0) {
    qto_params p;
    if (qto_get_params(set, &p)) return -1;
    uint64_t ps = 1, ips = p.n_inv, w = 1, iw = 1;
    for (uint32_t i = 0; i < p.n; i++) {
        if (bitrev) bitrev[i] = qto_bitrev(i, p.logn);
        if (Phi) Phi[i] = (uint32_t)ps;
        if (invPhi) invPhi[i] = (uint32_t)ips;
        if (tf0) tf0[i] = (uint32_t)w;
        if (ti0) ti0[i] = (uint32_t)iw;
        ps = ps * p.psi % p.q;
        ips = ips * p.psi_inv % p.q;
        w = w * p.omega % p.q;
        iw = iw * p.omega_inv % p.q;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* scalar arithmetic (NTT.cu:33-47)                                                      */
/* ------------------------------------------------------------------------------------ */

uint32_t qto_addmod(uint32_t a, uint32_t b, uint32_t q) {
    uint64_t s = (uint64_t)a + b;
    return (uint32_t)(s >= q ? s - q : s);
}
uint32_t qto_submod(uint32_t a, uint32_t b, uint32_t q) {
    uint64_t d = (uint64_t)a + (a < b ? q : 0) - b;
    return (uint32_t)(d >= q ? d - q : d);
}
uint32_t qto_mulmod(uint32_t a, uint32_t b, uint32_t q) { return (uint32_t)((uint64_t)a * b % q); }

/* ------------------------------------------------------------------------------------ */
/* transforms                                                                            */
/* ------------------------------------------------------------------------------------ */

/* Gentleman-Sande / DIF, natural -> bit-reversed (NTT.cu:1063-1083, 1246-1265) */
void qto_gs_dif(int set, uint32_t* a, size_t B, const uint32_t* tw) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    const uint32_t n = P.n, q = P.q;
    for (size_t b = 0; b < B; b++) {
        uint32_t* p = a + b * n;
        for (uint32_t lvl = 0; lvl < P.logn; lvl++) {
            uint32_t m = n >> lvl, half = m >> 1, stride = 1u << lvl;
            for (uint32_t k = 0; k < n; k += m)
                for (uint32_t j = 0; j < half; j++) {
                    uint32_t u = p[k + j], v = p[k + j + half];
                    p[k + j] = qto_addmod(u, v, q);
                    p[k + j + half] = qto_mulmod(qto_submod(u, v, q), tw[(j * stride) % n], q);
                }
        }
    }
}

/* Cooley-Tukey / DIT, bit-reversed -> natural (NTT.cu:1206-1221, 1478-1493) */
void qto_ct_dit(int set, uint32_t* a, size_t B, const uint32_t* tw) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    const uint32_t n = P.n, q = P.q;
    for (size_t b = 0; b < B; b++) {
        uint32_t* p = a + b * n;
        uint32_t k = n / 2;
        for (uint32_t l = 1; l < n; l *= 2, k >>= 1)
            for (uint32_t s = 0; s < n; s += 2 * l)
                for (uint32_t j = 0; j < l; j++) {
                    uint32_t t = qto_mulmod(p[j + l + s], tw[j * k], q);
                    uint32_t u = p[j + s];
                    p[j + l + s] = qto_submod(u, t, q);
                    p[j + s] = qto_addmod(u, t, q);
                }
    }
}

/* Stockham autosort, natural -> natural (NTT.cu:1170-1191, 1343-1365); result left in a */
void qto_stockham(int set, uint32_t* a, size_t B, const uint32_t* tw, uint32_t* scratch) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    const uint32_t n = P.n, q = P.q;
    for (size_t b = 0; b < B; b++) {
        uint32_t* in = a + b * n;
        uint32_t* out = scratch;
        for (uint32_t lvl = 0; lvl < P.logn; lvl++) {
            uint32_t stride = 1u << lvl, groups = n / (2 * stride);
            for (uint32_t j = 0; j < stride; j++)
                for (uint32_t s = 0; s < groups; s++) {
                    uint32_t u = in[s * stride + j], v = in[s * stride + j + n / 2];
                    out[2 * s * stride + j] = qto_addmod(u, v, q);
                    out[(2 * s + 1) * stride + j] =
                        qto_mulmod(qto_submod(u, v, q), tw[(s * stride) % n], q);
                }
            uint32_t* t = in; /* the reference swaps contents; swapping roles is equivalent */
            in = out;
            out = t;
        }
        if (in != a + b * n) memcpy(a + b * n, in, n * sizeof(uint32_t));
    }
}

void qto_bitrev_copy(int set, const uint32_t* in, uint32_t* out, size_t B) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    for (size_t b = 0; b < B; b++)
        for (uint32_t j = 0; j < P.n; j++) out[b * P.n + j] = in[b * P.n + qto_bitrev(j, P.logn)];
}

void qto_scale(int set, uint32_t* a, size_t B, const uint32_t* tbl) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    for (size_t b = 0; b < B; b++)
        for (uint32_t i = 0; i < P.n; i++) a[b * P.n + i] = qto_mulmod(a[b * P.n + i], tbl[i], P.q);
}

typedef struct {
    qto_params P;
    uint32_t *bitrev, *Phi, *invPhi, *tf0, *ti0;
} tables_t;

static int tables_make(int set, tables_t* T) {
    if (qto_get_params(set, &T->P)) return -1;
    uint32_t n = T->P.n;
    uint32_t* blk = (uint32_t*)malloc(5u * n * sizeof(uint32_t));
    if (!blk) return -1;
    T->bitrev = blk;
    T->Phi = blk + n;
    T->invPhi = blk + 2 * n;
    T->tf0 = blk + 3 * n;
    T->ti0 = blk + 4 * n;
    return qto_tables(set, T->bitrev, T->Phi, T->invPhi, T->tf0, T->ti0);
}
static void tables_free(tables_t* T) { free(T->bitrev); }

void qto_ntt_forward(int set, uint32_t* a, size_t B) {
    tables_t T;
    if (tables_make(set, &T)) return;
    qto_scale(set, a, B, T.Phi);      /* NTT.cu:1866-1870 */
    qto_gs_dif(set, a, B, T.tf0);     /* NTT.cu:1875 */
    tables_free(&T);
}

void qto_ntt_inverse(int set, uint32_t* a, size_t B) {
    tables_t T;
    if (tables_make(set, &T)) return;
    qto_ct_dit(set, a, B, T.ti0);     /* NTT.cu:1845 */
    qto_scale(set, a, B, T.invPhi);   /* NTT.cu:1846-1849 */
    tables_free(&T);
}

void qto_ntt_forward_natural(int set, uint32_t* a, size_t B) {
    tables_t T;
    if (tables_make(set, &T)) return;
    uint32_t* s = (uint32_t*)malloc(T.P.n * sizeof(uint32_t));
    qto_scale(set, a, B, T.Phi);      /* NTT.cu:1166-1167 */
    qto_stockham(set, a, B, T.tf0, s);
    free(s);
    tables_free(&T);
}

void qto_ntt_inverse_natural(int set, uint32_t* a, size_t B) {
    tables_t T;
    if (tables_make(set, &T)) return;
    uint32_t* s = (uint32_t*)malloc(T.P.n * sizeof(uint32_t));
    qto_stockham(set, a, B, T.ti0, s);
    qto_scale(set, a, B, T.invPhi);   /* NTT.cu:1367-1370 */
    free(s);
    tables_free(&T);
}

void qto_pointwise(int set, const uint32_t* a, const uint32_t* b, uint32_t* c, size_t B) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    for (size_t i = 0; i < B * P.n; i++) c[i] = qto_mulmod(a[i], b[i], P.q);
}

static void polymul_one_chunk(const tables_t* T, const uint32_t* x, const uint32_t* y, uint32_t* z,
                              size_t B, int variant, uint32_t* w /* 3*B*n + n words */) {
    const int set = T->P.set;
    const uint32_t n = T->P.n;
    const size_t N = B * n;
    uint32_t *X = w, *Y = w + N, *Z = w + 2 * N, *s = w + 3 * N;
    memcpy(X, x, N * sizeof(uint32_t));
    memcpy(Y, y, N * sizeof(uint32_t));
    qto_scale(set, X, B, T->Phi);
    qto_scale(set, Y, B, T->Phi);
    switch (variant) {
    case QTO_VARIANT_GS_CT: /* NTT.cu:1826-1849 */
        qto_gs_dif(set, X, B, T->tf0);
        qto_gs_dif(set, Y, B, T->tf0);
        qto_pointwise(set, X, Y, z, B);
        qto_ct_dit(set, z, B, T->ti0);
        break;
    case QTO_VARIANT_GS_GS: /* NTT.cu:1875-1888 */
        qto_gs_dif(set, X, B, T->tf0);
        qto_gs_dif(set, Y, B, T->tf0);
        qto_bitrev_copy(set, X, Z, B);
        qto_bitrev_copy(set, Y, X, B);
        qto_pointwise(set, Z, X, Y, B);
        qto_gs_dif(set, Y, B, T->ti0);
        qto_bitrev_copy(set, Y, z, B);
        break;
    case QTO_VARIANT_CT_CT: /* NTT.cu:1922-1933 */
        qto_bitrev_copy(set, X, Z, B);
        qto_ct_dit(set, Z, B, T->tf0);
        qto_bitrev_copy(set, Y, X, B);
        qto_ct_dit(set, X, B, T->tf0);
        qto_pointwise(set, Z, X, Y, B);
        qto_bitrev_copy(set, Y, z, B);
        qto_ct_dit(set, z, B, T->ti0);
        break;
    default: /* QTO_VARIANT_STOCKHAM, NTT.cu:1965-1974 */
        qto_stockham(set, X, B, T->tf0, s);
        qto_stockham(set, Y, B, T->tf0, s);
        qto_pointwise(set, X, Y, z, B);
        qto_stockham(set, z, B, T->ti0, s);
        break;
    }
    qto_scale(set, z, B, T->invPhi);
}

int qto_polymul(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int variant) {
    tables_t T;
    if (variant < 0 || variant > QTO_VARIANT_STOCKHAM || tables_make(set, &T)) return -1;
    const uint32_t n = T.P.n;
    const size_t chunk = 64;
    uint32_t* w = (uint32_t*)malloc((3 * chunk * n + n) * sizeof(uint32_t));
    for (size_t b = 0; b < B; b += chunk) {
        size_t c = B - b < chunk ? B - b : chunk;
        polymul_one_chunk(&T, x + b * n, y + b * n, z + b * n, c, variant, w);
    }
    free(w);
    tables_free(&T);
    return 0;
}

int qto_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int qto_polymul_omp(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B,
                    int threads) {
    tables_t T;
    if (tables_make(set, &T)) return -1;
    const uint32_t n = T.P.n;
    const size_t chunk = 16;
    const long nchunks = (long)((B + chunk - 1) / chunk);
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel num_threads(threads)
#endif
    {
        uint32_t* w = (uint32_t*)malloc((3 * chunk * n + n) * sizeof(uint32_t));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long c = 0; c < nchunks; c++) {
            size_t b = (size_t)c * chunk;
            size_t cnt = B - b < chunk ? B - b : chunk;
            polymul_one_chunk(&T, x + b * n, y + b * n, z + b * n, cnt, QTO_VARIANT_GS_CT, w);
        }
        free(w);
    }
    tables_free(&T);
    return used;
}

/* O(n^2) negacyclic schoolbook mod q; index pattern of naive (NTT.cu:151-164):
 * z[i] = sum_{j<=i} x[j]*y[i-j] - sum_{j>i} x[j]*y[n+i-j] */
void qto_schoolbook(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    qto_params P;
    if (qto_get_params(set, &P)) return;
    const uint32_t n = P.n, q = P.q;
    for (size_t b = 0; b < B; b++) {
        const uint32_t *xx = x + b * n, *yy = y + b * n;
        for (uint32_t i = 0; i < n; i++) {
            uint64_t pos = 0, neg = 0; /* each term < 2^60; reduce every step to stay exact */
            for (uint32_t j = 0; j <= i; j++) pos = (pos + (uint64_t)xx[j] * yy[i - j]) % q;
            for (uint32_t j = i + 1; j < n; j++) neg = (neg + (uint64_t)xx[j] * yy[n + i - j]) % q;
            z[b * n + i] = qto_submod((uint32_t)pos, (uint32_t)neg, q);
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Nussbaumer over Z/(2^32-1)  (ring macros NTT.cu:102-134)                              */
/* ------------------------------------------------------------------------------------ */

static inline uint32_t r_add(uint32_t a, uint32_t b) { /* modadd, NTT.cu:102-106 */
    uint32_t t = a + b;
    return t + (t < a);
}
static inline uint32_t r_sub(uint32_t a, uint32_t b) { /* modsub, NTT.cu:108 */
    return (a - b) - (uint32_t)(b > a);
}
static inline uint32_t r_fold(uint64_t T) { return r_add((uint32_t)T, (uint32_t)(T >> 32)); }
static inline uint32_t r_mul(uint32_t a, uint32_t b) { return r_fold((uint64_t)a * b); } /* NTT.cu:110-114 */
static inline uint32_t r_muladd(uint32_t c, uint32_t a, uint32_t b) { /* modmuladd, NTT.cu:117-121 */
    return r_fold((uint64_t)a * b + c);
}
static inline uint32_t r_norm(uint32_t a) { return a + (uint32_t)(a == 0xFFFFFFFFu); } /* NTT.cu:124 */
static inline uint32_t r_half(uint32_t a) { /* moddiv2 = normalize; div2, NTT.cu:123,131 */
    a = r_norm(a);
    return (uint32_t)(((uint64_t)a + (uint64_t)(uint32_t)(0u - (a & 1u))) >> 1);
}
static inline uint32_t r_neg(uint32_t a) { return r_norm(0xFFFFFFFFu - a); } /* NTT.cu:132 */

static void ring_schoolbook_one(uint32_t* z, const uint32_t* x, const uint32_t* y, uint32_t n) {
    /* naive, NTT.cu:147-165: A over j<=i, B over j>i, z = A - B, operations in this order */
    for (uint32_t i = 0; i < n; i++) {
        uint32_t A = r_mul(x[0], y[i]), Bn = 0;
        for (uint32_t j = 1; j <= i; j++) A = r_muladd(A, x[j], y[i - j]);
        for (uint32_t j = i + 1; j < n; j++) Bn = r_muladd(Bn, x[j], y[n + i - j]);
        z[i] = r_sub(A, Bn);
    }
}

void qto_ring_schoolbook(uint32_t n, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    for (size_t b = 0; b < B; b++) ring_schoolbook_one(z + b * n, x + b * n, y + b * n, n);
}

static int nuss_split(uint32_t n, uint32_t* m, uint32_t* r) {
    switch (n) {
    case 512: *m = 16; *r = 32; return 0;
    case 1024: *m = 32; *r = 32; return 0; /* the reference's 32x32, NTT.cu:185-193 */
    case 2048: *m = 32; *r = 64; return 0;
    default: return -1;
    }
}

/* rotation exponent of stage j, group i: brv_{logm-j}(i) << j, scaled by r/m so that
 * w^(r/m) is the 2m-th root (the reference has r == m, NTT.cu:198-203) */
static inline uint32_t nuss_rot(uint32_t i, uint32_t j, uint32_t logm, uint32_t r, uint32_t m) {
    return (qto_bitrev(i, logm - j) << j) * (r / m);
}

int qto_nussbaumer(uint32_t n, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    uint32_t m, r;
    if (nuss_split(n, &m, &r)) return -1;
    uint32_t logm = 0;
    while ((1u << logm) < m) logm++;
    const uint32_t rows = 2 * m;
    uint32_t* X = (uint32_t*)malloc((3u * rows * r + r) * sizeof(uint32_t));
    uint32_t *Y = X + rows * r, *Z = Y + rows * r, *T = Z + rows * r;
    for (size_t b = 0; b < B; b++) {
        const uint32_t *xx = x + b * n, *yy = y + b * n;
        uint32_t* zz = z + b * n;
        /* rows X_i[j] = x[m*j+i]; rows m..2m-1 are copies (NTT.cu:185-193) */
        for (uint32_t i = 0; i < m; i++)
            for (uint32_t j = 0; j < r; j++) {
                X[i * r + j] = X[(i + m) * r + j] = xx[m * j + i];
                Y[i * r + j] = Y[(i + m) * r + j] = yy[m * j + i];
            }
        /* forward: rotate-and-add butterflies, no multiplications (NTT.cu:195-235) */
        for (int j = (int)logm - 1; j >= 0; j--)
            for (uint32_t i = 0; i < (1u << (logm - j)); i++) {
                uint32_t sr = nuss_rot(i, (uint32_t)j, logm, r, m);
                for (uint32_t t = 0; t < (1u << j); t++) {
                    uint32_t I = (i << (j + 1)) + t, L = I + (1u << j);
                    for (int op = 0; op < 2; op++) {
                        uint32_t* V = op ? Y : X;
                        for (uint32_t a = sr; a < r; a++) T[a] = V[L * r + a - sr];
                        for (uint32_t a = 0; a < sr; a++) T[a] = r_neg(V[L * r + r + a - sr]);
                        for (uint32_t a = 0; a < r; a++) {
                            V[L * r + a] = r_sub(V[I * r + a], T[a]);
                            V[I * r + a] = r_add(V[I * r + a], T[a]);
                        }
                    }
                }
            }
        /* 2m negacyclic products of length r (NTT.cu:237-239) */
        for (uint32_t i = 0; i < rows; i++) ring_schoolbook_one(Z + i * r, X + i * r, Y + i * r, r);
        /* inverse: logm+1 stages with halving (NTT.cu:241-269) */
        for (uint32_t j = 0; j <= logm; j++)
            for (uint32_t i = 0; i < (1u << (logm - j)); i++) {
                uint32_t sr = nuss_rot(i, j, logm, r, m);
                for (uint32_t t = 0; t < (1u << j); t++) {
                    uint32_t A = (i << (j + 1)) + t, Bq = A + (1u << j);
                    for (uint32_t a = 0; a < r; a++) {
                        T[a] = r_half(r_sub(Z[A * r + a], Z[Bq * r + a]));
                        Z[A * r + a] = r_half(r_add(Z[A * r + a], Z[Bq * r + a]));
                    }
                    for (uint32_t a = 0; a + sr < r; a++) Z[Bq * r + a] = T[a + sr];
                    for (uint32_t a = r - sr; a < r; a++) Z[Bq * r + a] = r_neg(T[a - (r - sr)]);
                }
            }
        /* recombination with u^m = w (NTT.cu:271-276) */
        for (uint32_t i = 0; i < m; i++) {
            zz[i] = r_sub(Z[i * r], Z[(m + i) * r + r - 1]);
            for (uint32_t j = 1; j < r; j++) zz[m * j + i] = r_add(Z[i * r + j], Z[(m + i) * r + j - 1]);
        }
    }
    free(X);
    return 0;
}

/* Same index maps over Z_q: add/sub/halve mod q, schoolbook products mod q, canonical. */
int qto_nussbaumer_modq(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    qto_params P;
    uint32_t m, r;
    if (qto_get_params(set, &P) || nuss_split(P.n, &m, &r)) return -1;
    const uint32_t n = P.n, q = P.q;
    const uint32_t half = (q + 1) / 2; /* 2^-1 mod q */
    uint32_t logm = 0;
    while ((1u << logm) < m) logm++;
    const uint32_t rows = 2 * m;
    uint32_t* X = (uint32_t*)malloc((3u * rows * r + r) * sizeof(uint32_t));
    uint32_t *Y = X + rows * r, *Z = Y + rows * r, *T = Z + rows * r;
    for (size_t b = 0; b < B; b++) {
        const uint32_t *xx = x + b * n, *yy = y + b * n;
        uint32_t* zz = z + b * n;
        for (uint32_t i = 0; i < m; i++)
            for (uint32_t j = 0; j < r; j++) {
                X[i * r + j] = X[(i + m) * r + j] = xx[m * j + i];
                Y[i * r + j] = Y[(i + m) * r + j] = yy[m * j + i];
            }
        for (int j = (int)logm - 1; j >= 0; j--)
            for (uint32_t i = 0; i < (1u << (logm - j)); i++) {
                uint32_t sr = nuss_rot(i, (uint32_t)j, logm, r, m);
                for (uint32_t t = 0; t < (1u << j); t++) {
                    uint32_t I = (i << (j + 1)) + t, L = I + (1u << j);
                    for (int op = 0; op < 2; op++) {
                        uint32_t* V = op ? Y : X;
                        for (uint32_t a = sr; a < r; a++) T[a] = V[L * r + a - sr];
                        for (uint32_t a = 0; a < sr; a++) T[a] = qto_submod(0, V[L * r + r + a - sr], q);
                        for (uint32_t a = 0; a < r; a++) {
                            V[L * r + a] = qto_submod(V[I * r + a], T[a], q);
                            V[I * r + a] = qto_addmod(V[I * r + a], T[a], q);
                        }
                    }
                }
            }
        for (uint32_t i = 0; i < rows; i++)
            for (uint32_t k = 0; k < r; k++) {
                uint64_t pos = 0, neg = 0;
                for (uint32_t j = 0; j <= k; j++) pos = (pos + (uint64_t)X[i * r + j] * Y[i * r + k - j]) % q;
                for (uint32_t j = k + 1; j < r; j++) neg = (neg + (uint64_t)X[i * r + j] * Y[i * r + r + k - j]) % q;
                Z[i * r + k] = qto_submod((uint32_t)pos, (uint32_t)neg, q);
            }
        for (uint32_t j = 0; j <= logm; j++)
            for (uint32_t i = 0; i < (1u << (logm - j)); i++) {
                uint32_t sr = nuss_rot(i, j, logm, r, m);
                for (uint32_t t = 0; t < (1u << j); t++) {
                    uint32_t A = (i << (j + 1)) + t, Bq = A + (1u << j);
                    for (uint32_t a = 0; a < r; a++) {
                        T[a] = qto_mulmod(qto_submod(Z[A * r + a], Z[Bq * r + a], q), half, q);
                        Z[A * r + a] = qto_mulmod(qto_addmod(Z[A * r + a], Z[Bq * r + a], q), half, q);
                    }
                    for (uint32_t a = 0; a + sr < r; a++) Z[Bq * r + a] = T[a + sr];
                    for (uint32_t a = r - sr; a < r; a++) Z[Bq * r + a] = qto_submod(0, T[a - (r - sr)], q);
                }
            }
        for (uint32_t i = 0; i < m; i++) {
            zz[i] = qto_submod(Z[i * r], Z[(m + i) * r + r - 1], q);
            for (uint32_t j = 1; j < r; j++)
                zz[m * j + i] = qto_addmod(Z[i * r + j], Z[(m + i) * r + j - 1], q);
        }
    }
    free(X);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* input streams                                                                         */
/* ------------------------------------------------------------------------------------ */

uint64_t qto_fill_xorshift_pair(uint64_t s, uint32_t q, uint32_t* x, uint32_t* y, size_t count) {
    for (size_t i = 0; i < count; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        x[i] = (uint32_t)(s >> 11) % q;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        y[i] = (uint32_t)(s >> 11) % q;
    }
    return s;
}

uint64_t qto_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void qto_fill_splitmix(uint64_t seed, uint64_t first, uint32_t q, uint32_t* a, size_t count) {
    for (size_t i = 0; i < count; i++) a[i] = (uint32_t)(qto_splitmix64(seed + first + i) % q);
}
