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

#include "qtesla_b200.h" /* QT_SET_*, error codes */
#ifdef __cplusplus
extern "C" {
#endif

/* ---- explicit handle: what the macros BATCH / NTTSIZE / P fix at compile time (main.cuh:7-21), plus the GPU ---- */
typedef struct qt_ref_session qt_ref_session;
#define QT_REF_STOCKHAM 0   /* test_NTT_Stockham_nega_gpu  NTT.cu:2008 */
#define QT_REF_GS_CT 1      /* test_NTT_GS_CT_nega_gpu     NTT.cu:2097 */
#define QT_REF_CT_CT 2      /* test_NTT_CT_CT_nega_gpu     NTT.cu:2181 */
#define QT_REF_GS_GS 3      /* test_NTT_GS_GS_nega_gpu     NTT.cu:2271 */
#define QT_REF_CT_GS 4      /* test_NTT_CT_GS_nega_gpu     NTT.cu:2358 */
#define QT_REF_NUSSBAUMER 5 /* test_nussbaumer             NTT.cu:1987 (batched, on the GPU) */
int qt_ref_open(int param_set, uint64_t batch, int device, qt_ref_session** out);
int qt_ref_close(qt_ref_session* s);
int qt_ref_session_keep_operands(qt_ref_session* s, int keep);
/* one driver run on a session: same prints, x = y = 1 fill (unless kept) and result placement as the reference's
 * function of that name; returns 0 or an error code instead of the reference's void */
int qt_ref_run(qt_ref_session* s, int driver, uint32_t* x, uint32_t* y, uint32_t* out);

/* ---- the reference's own signatures.  They cannot carry a handle (main.cuh:61-70 fixes the argument lists), so they
 *      share ONE default session (QT_SET_III, batch 2, device 0 — the reference's compile-time configuration) that
 *      qt_ref_configure replaces; calls are serialised by a mutex.  On an error they print the message and exit(1),
 *      the only channel a void driver has. ---- */
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
