/*
 * qtesla_b200_reference_api.h — source-compatible replacements for the reference's harness-level
 * operators (benlwk/ntt-gpu-qTESLA main.cuh:61-70), implemented over the C ABI of qtesla_b200.h.
 *
 * A maintainer of the reference deletes the bodies of NTT.cu:1987-2443, includes this header
 * instead of the prototypes in main.cuh:61-70 and links libqtesla_b200.so; main.cu:203-226 then
 * compiles unchanged.  Same argument lists as the reference (X, Y, Z, tf0, ti0, nfg0, nig0, Ni are
 * accepted and ignored exactly as the reference ignores them, NTT.cu:2105-2127); same behaviour:
 * x and y are overwritten with ones (NTT.cu:2010, 2099, 2183, 2273, 2360), the product is left in z
 * (Z for the Stockham driver, NTT.cu:2078), time and "Multiplications per second" are printed
 * (NTT.cu:2083, 2167).  What the reference fixes with macros is set at run time:
 */
#ifndef QTESLA_B200_REFERENCE_API_H
#define QTESLA_B200_REFERENCE_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* replaces #define BATCH / NTTSIZE / P (main.cuh:7-21); defaults: QT_SET_III, batch 2, device 0 */
int qt_ref_configure(int param_set, uint64_t batch, int device);
/* 0 = reference behaviour (operands overwritten with ones, result dumped under DEBUG);
 * 1 = keep the caller's x and y, no dump */
void qt_ref_keep_operands(int keep);

void test_NTT_Stockham_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t* X, uint32_t* Y, uint32_t* Z,
                                uint32_t* tf0, uint32_t* ti0, uint32_t fg0, uint32_t ig0, uint32_t Ni);
void test_NTT_GS_CT_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t* X, uint32_t* Y, uint32_t* Z,
                             uint32_t* tf0, uint32_t* ti0, uint32_t nfg0, uint32_t nig0, uint32_t Ni);
void test_NTT_CT_CT_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t* X, uint32_t* Y, uint32_t* Z,
                             uint32_t* tf0, uint32_t* ti0, uint32_t nfg0, uint32_t nig0, uint32_t Ni);
void test_NTT_GS_GS_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t* X, uint32_t* Y, uint32_t* Z,
                             uint32_t* tf0, uint32_t* ti0, uint32_t nfg0, uint32_t nig0, uint32_t Ni);
void test_NTT_CT_GS_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t* X, uint32_t* Y, uint32_t* Z,
                             uint32_t* tf0, uint32_t* ti0, uint32_t nfg0, uint32_t nig0, uint32_t Ni);
/* GPU, batched counterpart of the reference's CPU-only single-polynomial driver (NTT.cu:1987-2005) */
void test_nussbaumer(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t* X, uint32_t* Y, uint32_t* Z);

#ifdef __cplusplus
}
#endif
#endif
