/*
 * qt_oracle.h — CPU ORACLE for the batched qTESLA negacyclic polynomial multiplication.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (ntt-gpu-qtesla_b200/csrc + include/qtesla_b200.h) never links, loads or calls it.
 *
 * It is an n-generic plain-C restatement of the CPU functions of the reference
 * (benlwk/ntt-gpu-qTESLA, NTT.cu), which hard-code n = 1024, ten levels and BATCH = 2.
 * Every function cites the reference file:line it follows.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - qTESLA-III (n=1024, q=8404993): PINNED.  Tables are checked against the literals
 *     of constants.h (sha256) and every transform/product against the reference's own
 *     CPU functions compiled from /root/reference (oracle/_ref) and the committed golden
 *     vectors in tests/golden/.
 *   - Nussbaumer ring Z/(2^32-1), n=1024: PINNED against nussbaumer_fft (NTT.cu:167-277).
 *   - qTESLA-I / p-I / p-III: the reference holds no table, code or vector for them
 *     ("parity unpinned" w.r.t. reference artefacts); they are pinned by mathematics only:
 *     the O(n^2) negacyclic schoolbook mod q, whose result does not depend on psi.
 *
 * Layout everywhere: batch-major uint32_t a[B*n], coefficient i of polynomial b at
 * a[b*n+i] (NTT.cu:975, 1159, 1576).
 */
#ifndef QT_ORACLE_H
#define QT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* parameter-set ids (shared with include/qtesla_b200.h) */
enum {
    QTO_SET_I = 0,     /* qTESLA-I     n=512  q=4205569   */
    QTO_SET_III = 1,   /* qTESLA-III   n=1024 q=8404993   (the reference's set, main.cuh:13-21) */
    QTO_SET_P_I = 2,   /* qTESLA-p-I   n=1024 q=343576577 */
    QTO_SET_P_III = 3, /* qTESLA-p-III n=2048 q=856145921 */
    QTO_NUM_SETS = 4
};

/* negacyclic polymul compositions of the reference's CPU drivers */
enum {
    QTO_VARIANT_GS_CT = 0,    /* test_NTT_GS_CT_BATCH   NTT.cu:1820-1857 (with the %P fix of 1868) */
    QTO_VARIANT_GS_GS = 1,    /* test_NTT_nega_GS       NTT.cu:1860-1906 */
    QTO_VARIANT_CT_CT = 2,    /* test_NTT_nega_CT       NTT.cu:1908-1953 */
    QTO_VARIANT_STOCKHAM = 3  /* test_NTT_Stockham_nega NTT.cu:1955-1984 */
};

typedef struct {
    int set;
    uint32_t n, logn, q;
    uint32_t psi, psi_inv;     /* primitive 2n-th root of unity and its inverse */
    uint32_t omega, omega_inv; /* psi^2 and its inverse (fg0 / ig0 of main.cu:25-27) */
    uint32_t n_inv;            /* n^-1 mod q (Ni of main.cu:26) */
    uint32_t qinv_neg;         /* -q^-1 mod 2^32 (PARAM_QINV of main.cuh:15) */
    uint32_t barrett_mu48;     /* floor(2^48 / q) (MIU of main.cuh:20) */
} qto_params;

int qto_get_params(int set, qto_params* out); /* 0 ok, -1 bad set */

/* The five distinct tables of constants.h:3-35, each n words:
 * bitrev[i]=brv_logn(i), Phi[i]=psi^i, invPhi[i]=n^-1*psi^-i, tf0[i]=omega^i, ti0[i]=omega^-i.
 * Any output pointer may be NULL. */
int qto_tables(int set, uint32_t* bitrev, uint32_t* Phi, uint32_t* invPhi, uint32_t* tf0,
               uint32_t* ti0);

/* --- scalar helpers (NTT.cu:33-47, 61-79, 341-361) --- */
uint32_t qto_addmod(uint32_t a, uint32_t b, uint32_t q);
uint32_t qto_submod(uint32_t a, uint32_t b, uint32_t q);
uint32_t qto_mulmod(uint32_t a, uint32_t b, uint32_t q);
uint32_t qto_bitrev(uint32_t x, uint32_t bits);

/* --- single transforms, in place on a[B*n]; tw = tf0 (forward) or ti0 (inverse) --- */
void qto_gs_dif(int set, uint32_t* a, size_t B, const uint32_t* tw);   /* radix2NTTGS/radix2INTTGS NTT.cu:1058-1084,1241-1266: natural -> bit-reversed */
void qto_ct_dit(int set, uint32_t* a, size_t B, const uint32_t* tw);   /* radix2NTT/radix2INTT NTT.cu:1201-1222,1473-1494: bit-reversed -> natural */
void qto_stockham(int set, uint32_t* a, size_t B, const uint32_t* tw, uint32_t* scratch); /* radix2NTTStock/radix2INTTStock butterfly loops NTT.cu:1170-1191,1343-1365: natural -> natural; result in a */
void qto_bitrev_copy(int set, const uint32_t* in, uint32_t* out, size_t B); /* bit_reverse_copy(_tbl) NTT.cu:81-100 */
void qto_scale(int set, uint32_t* a, size_t B, const uint32_t* tbl);    /* a[b*n+i] = a*tbl[i] % q, NTT.cu:1866-1870, 1896-1899 */

/* --- the drop-in entry points' semantics --- */
/* forward: Phi-scale + GS/DIF  => NTT domain, bit-reversed order: pos i = x(psi^(2*brv(i)+1)) */
void qto_ntt_forward(int set, uint32_t* a, size_t B);
/* inverse of the above: CT/DIT with ti0 + invPhi scale (n^-1 inside invPhi) */
void qto_ntt_inverse(int set, uint32_t* a, size_t B);
/* natural-order NTT-domain variants (what the Stockham pipeline leaves, NTT.cu:2040-2049) */
void qto_ntt_forward_natural(int set, uint32_t* a, size_t B);
void qto_ntt_inverse_natural(int set, uint32_t* a, size_t B);
void qto_pointwise(int set, const uint32_t* a, const uint32_t* b, uint32_t* c, size_t B); /* NTT.cu:1155-1160, 1884-1885 */
/* whole product z = x*y mod (X^n+1, q); x,y are not modified */
int qto_polymul(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int variant);
/* O(n^2) schoolbook negacyclic product mod q (index pattern of naive, NTT.cu:151-164) */
void qto_schoolbook(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B);

/* --- Nussbaumer, ring Z/(2^32-1), operation order of nussbaumer_fft NTT.cu:167-277 --- */
/* n = m*r with (m,r) = (16,32) n=512, (32,32) n=1024, (32,64) n=2048 */
int qto_nussbaumer(uint32_t n, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B);
/* ring schoolbook of length n (naive, NTT.cu:147-165) */
void qto_ring_schoolbook(uint32_t n, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B);
/* Nussbaumer structure over Z_q (same index maps, arithmetic mod q, canonical output) —
 * oracle of the kernel's Z_q mode; must equal qto_schoolbook */
int qto_nussbaumer_modq(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B);

/* --- input streams --- */
/* golden PRNG of SURVEY.md 8c-3: xorshift64, x[i]=out%q then y[i]=out%q; returns new state */
uint64_t qto_fill_xorshift_pair(uint64_t state, uint32_t q, uint32_t* x, uint32_t* y, size_t count);
/* bench stream of SURVEY.md 8d: a[i] = splitmix64(seed + first + i) % q */
void qto_fill_splitmix(uint64_t seed, uint64_t first, uint32_t q, uint32_t* a, size_t count);
uint64_t qto_splitmix64(uint64_t x);

/* --- CPU baseline legs (bench.py cpu_baseline / --impl reference, kind "port") --- */
/* qto_polymul(GS_CT) over the batch with OpenMP; returns threads used */
int qto_polymul_omp(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B,
                    int threads);
int qto_max_threads(void);
/* qTESLA-style CPU path (qt_cpu_fast.c: Montgomery reduce with PARAM_QINV, merged twiddles, lazy ranges;
 * "restatement, qTESLA source unavailable"), OpenMP over polynomials; returns threads used.  Results are
 * canonical and equal to qto_polymul / qto_ntt_forward. */
int qto_fast_polymul(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int threads);
void qto_fast_ntt_forward(int set, uint32_t* a, size_t B);

#ifdef __cplusplus
}
#endif
#endif
