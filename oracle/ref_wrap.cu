/*
 * ref_wrap.cu — builds the UNMODIFIED reference (benlwk/ntt-gpu-qTESLA) CPU functions into
 * oracle/_ref/libqtref.so.  TEST INFRASTRUCTURE ONLY (see qt_oracle.h).
 *
 * The reference translation unit is included where it lies (-I/root/reference); no reference
 * source is copied into this repository.  Including NTT.cu (rather than linking it) is what
 * gives access to the `static` nussbaumer_fft / naive (NTT.cu:147,167).
 *
 * The reference is compiled for BATCH=2, NTTSIZE=1024, P=8404993 (main.cuh:7-21); the wrappers
 * below loop over the caller's batch two polynomials at a time.
 */
#include "NTT.cu" /* /root/reference/NTT.cu, pulls main.cuh + constants.h */

#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

static uint32_t g_tf0[NTTSIZE], g_ti0[NTTSIZE];
static bool g_tw_ready = false;

/* twiddle precompute exactly as main.cu:119-129 does it (fg0 = 2893, main.cu:26), O(n) here */
static void ref_twiddles() {
    if (g_tw_ready) return;
    const uint64_t fg0 = 2893;
    g_tf0[0] = g_ti0[0] = 1;
    for (int i = 1; i < NTTSIZE; i++) g_tf0[i] = (uint32_t)((uint64_t)g_tf0[i - 1] * fg0 % (uint64_t)P);
    for (int i = 1; i < NTTSIZE; i++) g_ti0[i] = g_tf0[NTTSIZE - i];
    g_tw_ready = true;
}

extern "C" {

int qtref_n(void) { return NTTSIZE; }
unsigned qtref_q(void) { return P; }
int qtref_batch(void) { return BATCH; }
unsigned qtref_qinv(void) { return (unsigned)PARAM_QINV; }
unsigned qtref_miu(void) { uint32_t m = MIU return m; } /* macro carries its own ';' (main.cuh:20) */

/* host tables of constants.h: 0 bitrev_tbl, 1 Phi, 2 invPhi; 3/4 tf0/ti0 as main.cu builds them */
int qtref_table(int which, uint32_t* out) {
    ref_twiddles();
    const uint32_t* src = which == 0 ? bitrev_tbl : which == 1 ? Phi : which == 2 ? invPhi
                        : which == 3 ? g_tf0 : which == 4 ? g_ti0 : nullptr;
    if (!src) return -1;
    memcpy(out, src, NTTSIZE * sizeof(uint32_t));
    return 0;
}

static void load2(uint32_t* dst, const uint32_t* src, size_t b, size_t B) {
    for (int k = 0; k < BATCH; k++) {
        if (b + k < B) memcpy(dst + k * NTTSIZE, src + (b + k) * NTTSIZE, NTTSIZE * sizeof(uint32_t));
        else memset(dst + k * NTTSIZE, 0, NTTSIZE * sizeof(uint32_t));
    }
}
static void store2(uint32_t* dst, const uint32_t* src, size_t b, size_t B) {
    for (int k = 0; k < BATCH && b + k < B; k++)
        memcpy(dst + (b + k) * NTTSIZE, src + k * NTTSIZE, NTTSIZE * sizeof(uint32_t));
}
static void scale2(uint32_t* a, const uint32_t* tbl) { /* NTT.cu:1866-1870 / 1896-1899 */
    for (int i = 0; i < BATCH * NTTSIZE; i++) a[i] = (uint64_t)a[i] * (uint64_t)tbl[i % NTTSIZE] % P;
}

/* forward: Phi scale + radix2NTTGS  (NTT.cu:1866-1876) */
void qtref_forward(uint32_t* a, size_t B) {
    ref_twiddles();
    uint32_t t[BATCH * NTTSIZE];
    for (size_t b = 0; b < B; b += BATCH) { load2(t, a, b, B); scale2(t, Phi); radix2NTTGS(t, g_tf0); store2(a, t, b, B); }
}
/* inverse: radix2INTT + invPhi scale (NTT.cu:1845-1849) */
void qtref_inverse(uint32_t* a, size_t B) {
    ref_twiddles();
    uint32_t t[BATCH * NTTSIZE];
    for (size_t b = 0; b < B; b += BATCH) { load2(t, a, b, B); radix2INTT(t, g_ti0, 8396785u); scale2(t, invPhi); store2(a, t, b, B); }
}

/* whole product with the reference's CPU function compositions.
 * variant 0 GS->CT (NTT.cu:1820-1857 with the %P fix of 1868), 1 GS/GS (1860-1906),
 * 2 CT/CT (1908-1953), 3 Stockham (1955-1984). Returns threads used. */
int qtref_polymul(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int variant, int threads) {
    ref_twiddles();
    const long pairs = (long)((B + BATCH - 1) / BATCH);
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (long pi = 0; pi < pairs; pi++) {
        const size_t b = (size_t)pi * BATCH;
        uint32_t X[BATCH * NTTSIZE], Y[BATCH * NTTSIZE], Z[BATCH * NTTSIZE], W[BATCH * NTTSIZE], V[BATCH * NTTSIZE];
        load2(X, x, b, B);
        load2(Y, y, b, B);
        if (variant == 3) {
            radix2NTTStock(X, g_tf0, W); /* applies Phi itself (NTT.cu:1166-1167) */
            radix2NTTStock(Y, g_tf0, W);
            for (int i = 0; i < BATCH * NTTSIZE; i++) Z[i] = ((uint64_t)X[i] * (uint64_t)Y[i]) % P;
            radix2INTTStock(Z, g_ti0, 8396785u, W, invPhi); /* applies invPhi itself (NTT.cu:1367-1370) */
        } else {
            scale2(X, Phi);
            scale2(Y, Phi);
            if (variant == 0) {
                radix2NTTGS(X, g_tf0); radix2NTTGS(Y, g_tf0);
                for (int i = 0; i < BATCH * NTTSIZE; i++) Z[i] = ((uint64_t)X[i] * (uint64_t)Y[i]) % P;
                radix2INTT(Z, g_ti0, 8396785u);
            } else if (variant == 1) {
                radix2NTTGS(X, g_tf0); radix2NTTGS(Y, g_tf0);
                bit_reverse_copy(X, W); bit_reverse_copy(Y, V);
                for (int i = 0; i < BATCH * NTTSIZE; i++) X[i] = ((uint64_t)W[i] * (uint64_t)V[i]) % P;
                radix2INTTGS(X, g_ti0, 8396785u);
                bit_reverse_copy(X, Z);
            } else {
                bit_reverse_copy(X, W); bit_reverse_copy(Y, V);
                radix2NTT(W, g_tf0); radix2NTT(V, g_tf0);
                for (int i = 0; i < BATCH * NTTSIZE; i++) X[i] = ((uint64_t)W[i] * (uint64_t)V[i]) % P;
                bit_reverse_copy_tbl(X, Z);
                radix2INTT(Z, g_ti0, 8396785u);
            }
            scale2(Z, invPhi);
        }
        store2(z, Z, b, B);
    }
    return used;
}

/* nussbaumer_fft (NTT.cu:167-277) per polynomial; the reference leaks its 193 mallocs per call */
void qtref_nussbaumer(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    for (size_t b = 0; b < B; b++) nussbaumer_fft(z + b * NTTSIZE, x + b * NTTSIZE, y + b * NTTSIZE);
}
void qtref_naive(const uint32_t* x, const uint32_t* y, uint32_t* z, unsigned n) { naive(z, x, y, n); }
unsigned qtref_barrett_cpu(unsigned long long v) { return barrett_red_cpu(v); }
int qtref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- the reference's own GPU kernels as a same-box baseline --------------------------------
 * Launches the UNMODIFIED __global__ kernels of NTT.cu in the order of test_NTT_CT_GS_nega_gpu
 * (NTT.cu:2388-2425: bit_reverse_copy_tbl_Phi_gpu x2, 10 radix2NTT_gpu0/1 levels per operand, pointwise_mult,
 * 10 GS_radix2INTT_gpu0/2 levels, bit_reverse_copy_tbl_invPhi_gpu = 34 launches), with the grid taken from the
 * caller's batch instead of the BATCH macro (the kernels index by blockIdx.x, NTT.cu:957).  Of the five drivers
 * this is the one whose forward path applies the psi scale, i.e. computes the negacyclic product.
 * x, y, z are HOST arrays of B*1024 words.  Times with CUDA events:
 *   *kernel_ms = the 34 launches, operands device-resident (per repetition, averaged over reps)
 *   *total_ms  = synchronous cudaMemcpy of x, y + the launches + cudaMemcpy of z, the reference's own
 *                convention (NTT.cu:2383-2428), with the caller's buffers as they are (pinned or pageable)
 * Returns 0, or a cudaError_t. */
static void ref_ct_gs_launches(uint32_t* d_x, uint32_t* d_X, uint32_t* d_y, uint32_t* d_Y, uint32_t* d_Z, unsigned B) {
    uint32_t* unused = nullptr; /* the reference passes uninitialised d_tf0/d_ti0; its kernels read __constant__ tables */
    bit_reverse_copy_tbl_Phi_gpu<<<B, NTTSIZE>>>(d_x, d_X);
    bit_reverse_copy_tbl_Phi_gpu<<<B, NTTSIZE>>>(d_y, d_Y);
    for (int op = 0; op < 2; op++) {
        uint32_t* d = op ? d_Y : d_X;
        for (unsigned lvl = 1; lvl <= 5; lvl++) radix2NTT_gpu0<<<B, NTTSIZE >> lvl>>>(d, unused, 1u << lvl, lvl);
        for (unsigned lvl = 6; lvl <= 10; lvl++) radix2NTT_gpu1<<<B, 1u << (lvl - 1)>>>(d, unused, 1u << (lvl - 1), lvl);
    }
    pointwise_mult<<<B, NTTSIZE>>>(d_X, d_Y, d_Z);
    for (unsigned lvl = 0; lvl <= 4; lvl++) GS_radix2INTT_gpu0<<<B, 512u >> lvl>>>(d_Z, unused, lvl);
    for (unsigned lvl = 5; lvl <= 9; lvl++) GS_radix2INTT_gpu2<<<B, 32u << (lvl - 5)>>>(d_Z, unused, lvl);
    bit_reverse_copy_tbl_invPhi_gpu<<<B, NTTSIZE>>>(d_Z, d_x);
}

int qtref_gpu_ct_gs(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int reps, float* kernel_ms, float* total_ms) {
    const size_t bytes = B * NTTSIZE * sizeof(uint32_t);
    uint32_t* d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 5 && e == cudaSuccess; i++) e = cudaMalloc((void**)&d[i], bytes);
    cudaEvent_t e0, e1;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        /* reference convention: copies inside the timed region */
        float acc = 0.f;
        for (int r = 0; r < reps + 1 && e == cudaSuccess; r++) { /* first pass = warm-up */
            cudaEventRecord(e0);
            cudaMemcpy(d[0], x, bytes, cudaMemcpyHostToDevice);
            cudaMemcpy(d[2], y, bytes, cudaMemcpyHostToDevice);
            ref_ct_gs_launches(d[0], d[1], d[2], d[3], d[4], (unsigned)B);
            cudaMemcpy(z, d[0], bytes, cudaMemcpyDeviceToHost);
            cudaEventRecord(e1);
            e = cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r) acc += ms;
        }
        if (total_ms) *total_ms = acc / reps;
        /* kernel-only: operands resident (d_x is overwritten by the result, so it is refreshed outside the timer) */
        acc = 0.f;
        for (int r = 0; r < reps && e == cudaSuccess; r++) {
            cudaMemcpy(d[0], x, bytes, cudaMemcpyHostToDevice);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            ref_ct_gs_launches(d[0], d[1], d[2], d[3], d[4], (unsigned)B);
            cudaEventRecord(e1);
            e = cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            acc += ms;
        }
        if (kernel_ms) *kernel_ms = acc / reps;
        if (e == cudaSuccess) e = cudaGetLastError();
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    for (int i = 0; i < 5; i++)
        if (d[i]) cudaFree(d[i]);
    return (int)e;
}

} /* extern "C" */
