/*
 * qt_cpu_fast.c — CPU BASELINE in the style of the qTESLA submission's C code (poly_ntt / poly_mul):
 * Montgomery reduction with PARAM_QINV, merged psi twiddles in bit-reversed order, lazy ranges.
 *
 * THIS IS TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE (same rule as qt_oracle.h): only tests/ and
 * bench.py's cpu_baseline / --impl reference legs load it.
 *
 * Provenance and parity status.  BASELINE.json names "the qTESLA C poly_ntt/poly_mul" as the CPU path to time
 * beside the GPU.  That source (the NIST submission package, qTESLA round-2 "Reference_implementation/
 * qTesla-*" poly.c, no pinned version — it is not a dependency of the reference repo) is NOT under
 * /root/reference and not on this machine.  Its only traces in the reference are PARAM_QINV
 * (/root/reference/main.cuh:15 — equal to -q^-1 mod 2^32 for q = 8404993) and the commented-out Montgomery
 * `reduce` inside barrett_red (/root/reference/NTT.cu:390-396: q2 = (ip*PARAM_QINV) & 0xFFFFFFFF; q2 *= P;
 * ip += q2; res = ip >> 32).  So this file is a RESTATEMENT of the published algorithm ("restatement, qTESLA
 * source unavailable"; parity with the qTESLA C code itself is UNPINNED), and it is pinned instead to
 *   - the reference's own CPU path: forward output == Phi-scale + radix2NTTGS (NTT.cu:1866-1876) bit for bit,
 *     product == the reference's CPU compositions (oracle port qto_polymul, itself pinned to oracle/_ref and
 *     the golden vectors) for qTESLA-III, and
 *   - the O(n^2) schoolbook for the three sets the reference has no artefact for
 * (tests/test_oracle.py::test_fast_cpu_*).
 *
 * Published algorithm restated (qTESLA poly.c: ntt(), nttinv(), reduce(), poly_pointwise(), poly_mul()):
 *   reduce(a)      : Montgomery, a * 2^-32 mod q via q^-1 mod 2^32                (NTT.cu:390-396 sketch)
 *   ntt(a, zeta)   : Cooley-Tukey, len = n/2 .. 1, one twiddle zeta[k++] = psi^brv(k) * 2^32 per block,
 *                    t = reduce(zeta * a[j+len]); a[j+len] = a[j] - t; a[j] = a[j] + t  — natural -> bit-reversed,
 *                    no separate psi scaling pass
 *   nttinv(a)      : Gentleman-Sande with the inverse twiddles, len = 1 .. n/2, then * n^-1
 *   poly_mul(x, y) : ntt(x), ntt(y), pointwise reduce(x*y), nttinv — 2^32 factors folded into the last scale
 * Ranges: signed 32-bit residues; q < 2^25 runs all log2(n) forward levels without any reduction
 * ((1 + log2 n) q < 2^31) and reduces once in the middle of the inverse; the 29/30-bit moduli renormalise
 * after every level, as the submission's barr_reduce does for the provable sets.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "qt_oracle.h"

#define QT_MAXN 2048

typedef struct {
    int ready;
    uint32_t n, logn, q, qinv; /* qinv = q^-1 mod 2^32 (= -PARAM_QINV) */
    int lazy;                  /* q < 2^25 */
    int32_t zeta[QT_MAXN];     /* psi^brv(k) * 2^32 mod q, centred */
    int32_t zeta_inv[QT_MAXN]; /* psi^-brv(k) * 2^32, the value nttinv's block k uses */
    int32_t f;                 /* n^-1 * 2^64 mod q: last scale, cancels the 2^-32 of pointwise and of itself */
    int32_t r1;                /* 2^32 mod q: reduce(v * r1) = v mod q, a plain reduction sweep */
} fast_tab;

static fast_tab g_tab[QTO_NUM_SETS];

static uint32_t pw(uint32_t b, uint64_t e, uint32_t q) {
    uint64_t r = 1, x = b % q;
    for (; e; e >>= 1, x = x * x % q)
        if (e & 1) r = r * x % q;
    return (uint32_t)r;
}
static int32_t centre(uint64_t v, uint32_t q) { return v > q / 2 ? (int32_t)((int64_t)v - q) : (int32_t)v; }

static const fast_tab* tab(int set) {
    fast_tab* T = &g_tab[set];
    if (T->ready) return T;
#pragma omp critical(qt_fast_tab)
    if (!T->ready) {
        qto_params p;
        qto_get_params(set, &p);
        T->n = p.n; T->logn = p.logn; T->q = p.q;
        T->qinv = 0u - p.qinv_neg;
        T->lazy = p.q < (1u << 25);
        const uint64_t R = (1ull << 32) % p.q;
        for (uint32_t k = 0; k < p.n; k++) {
            const uint32_t e = qto_bitrev(k, p.logn);
            T->zeta[k] = centre((uint64_t)pw(p.psi, e, p.q) * R % p.q, p.q);
            T->zeta_inv[k] = centre((uint64_t)pw(p.psi_inv, e, p.q) * R % p.q, p.q);
        }
        T->r1 = centre(R, p.q);
        T->f = centre((uint64_t)p.n_inv * (R * R % p.q) % p.q, p.q);
        __sync_synchronize();
        T->ready = 1;
    }
    return T;
}

/* reduce(): a * 2^-32 mod q, |result| < q for |a| < 2^31 q */
static inline int32_t mont(int64_t a, uint32_t q, uint32_t qinv) {
    const int32_t t = (int32_t)((uint32_t)a * qinv);
    return (int32_t)((a - (int64_t)t * (int32_t)q) >> 32);
}
/* (-2q, 2q) -> (-q, q) without branches */
static inline int32_t renorm(int32_t v, int32_t q) {
    v += q & (v >> 31);         /* (-2q, 0) -> (-q, q); no intermediate leaves 32 bits even for the 30-bit q */
    v -= q & ~((v - q) >> 31);  /* [q, 2q) -> [0, q) */
    return v;
}

#if defined(__x86_64__) && defined(__GNUC__) && !defined(QT_NO_CLONES)
#define QT_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define QT_CLONES
#endif

QT_CLONES
static void fast_ntt(const fast_tab* T, int32_t* a) {
    const uint32_t n = T->n, q = T->q, qinv = T->qinv;
    uint32_t k = 1;
    for (uint32_t len = n >> 1; len >= 1; len >>= 1) {
        for (uint32_t s = 0; s < n; s += 2 * len) {
            const int64_t z = T->zeta[k++];
            int32_t* lo = a + s;
            int32_t* hi = a + s + len;
            for (uint32_t j = 0; j < len; j++) {
                const int32_t t = mont(z * hi[j], q, qinv);
                hi[j] = lo[j] - t;
                lo[j] = lo[j] + t;
            }
        }
        if (!T->lazy)
            for (uint32_t i = 0; i < n; i++) a[i] = renorm(a[i], (int32_t)q);
    }
}

QT_CLONES
static void fast_nttinv(const fast_tab* T, int32_t* a) {
    const uint32_t n = T->n, q = T->q, qinv = T->qinv, logn = T->logn;
    uint32_t lvl = 0;
    for (uint32_t len = 1; len < n; len <<= 1, lvl++) {
        /* block b of this level undoes forward block k = n/(2 len) + b */
        uint32_t k = n / (2 * len);
        for (uint32_t s = 0; s < n; s += 2 * len) {
            const int64_t z = T->zeta_inv[k++];
            int32_t* lo = a + s;
            int32_t* hi = a + s + len;
            for (uint32_t j = 0; j < len; j++) {
                const int32_t u = lo[j], v = hi[j];
                lo[j] = u + v;
                hi[j] = mont(z * (u - v), q, qinv);
            }
        }
        if (!T->lazy) {
            for (uint32_t i = 0; i < n; i++) a[i] = renorm(a[i], (int32_t)q);
        } else if (lvl == logn / 2) { /* sums double per level: one sweep keeps 2^levels q below 2^31 */
            for (uint32_t i = 0; i < n; i++) a[i] = mont((int64_t)a[i] * T->r1, q, qinv);
        }
    }
    for (uint32_t i = 0; i < n; i++) {
        int32_t v = mont((int64_t)a[i] * T->f, q, qinv);
        v += (int32_t)q & (v >> 31);
        a[i] = v;
    }
}

QT_CLONES
static void fast_pointwise(const fast_tab* T, int32_t* c, const int32_t* a, const int32_t* b) {
    const uint32_t n = T->n, q = T->q, qinv = T->qinv;
    if (T->lazy) /* |a|,|b| <= (1+logn) q: the product fits the reduction's domain only after one operand is reduced */
        for (uint32_t i = 0; i < n; i++) c[i] = mont((int64_t)mont((int64_t)a[i] * T->r1, q, qinv) * b[i], q, qinv);
    else
        for (uint32_t i = 0; i < n; i++) c[i] = mont((int64_t)a[i] * b[i], q, qinv);
}

/* forward transform alone, canonical output: equals qto_ntt_forward (Phi scale + radix2NTTGS) */
void qto_fast_ntt_forward(int set, uint32_t* a, size_t B) {
    const fast_tab* T = tab(set);
    const uint32_t n = T->n, q = T->q;
    for (size_t b = 0; b < B; b++) {
        int32_t* p = (int32_t*)(a + b * n);
        fast_ntt(T, p);
        for (uint32_t i = 0; i < n; i++) {
            int32_t v = mont((int64_t)p[i] * T->r1, q, T->qinv); /* any |v| < 2^31 -> (-q, q) */
            v += (int32_t)q & (v >> 31);
            p[i] = v;
        }
    }
}

/* poly_mul over a batch, OpenMP over polynomials; returns the number of threads used */
int qto_fast_polymul(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int threads) {
    if (set < 0 || set >= QTO_NUM_SETS) return -1;
    const fast_tab* T = tab(set);
    const uint32_t n = T->n;
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (long b = 0; b < (long)B; b++) {
        int32_t X[QT_MAXN], Y[QT_MAXN];
        memcpy(X, x + (size_t)b * n, n * sizeof(int32_t));
        memcpy(Y, y + (size_t)b * n, n * sizeof(int32_t));
        fast_ntt(T, X);
        fast_ntt(T, Y);
        fast_pointwise(T, X, X, Y);
        fast_nttinv(T, X);
        memcpy(z + (size_t)b * n, X, n * sizeof(int32_t));
    }
    return used;
}
