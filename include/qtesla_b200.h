/*
 * qtesla_b200.h — C ABI of the B200-native batched qTESLA negacyclic polynomial-multiplication
 * engine (libqtesla_b200.so).  Plain pointers and sizes only; no C++/torch types.
 *
 * This is the drop-in boundary for the hot path of benlwk/ntt-gpu-qTESLA.  The reference has no
 * FFI: its boundary is the C++ functions declared in main.cuh:52-71 that main.cu:158-226 drives,
 * plus its stage kernels.  Each entry point below names the reference interface it replaces.
 *
 * Conventions (same as the reference unless stated):
 *   - coefficient arrays are batch-major uint32_t a[B*n]; coefficient i of polynomial b is
 *     a[b*n+i] (NTT.cu:975, 1159, 1576).  int32 and uint32 bit patterns coincide (values < 2^30).
 *   - inputs must be canonical, in [0,q); outputs are canonical, in [0,q).  (The reference's
 *     variants disagree with each other for inputs >= q, NTT.cu:446; that case is undefined here.)
 *   - "NTT domain" = bit-reversed order with psi merged: position i holds x(psi^(2*brv(i)+1)),
 *     bit-identical to the reference's Phi-scale + radix2NTTGS (NTT.cu:1866-1876).
 *   - every function returns 0 on success, a positive cudaError_t value for a CUDA failure, or a
 *     negative QT_ERR_* code.  (The reference returns void and checks nothing.)
 *   - d_* pointers are device pointers on the context's GPU; calls taking a context are
 *     asynchronous on the context's stream; a context is not thread-safe; one context per GPU.
 *   - There is no CPU fallback: without a usable CUDA device every call fails with an error.
 */
#ifndef QTESLA_B200_H
#define QTESLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* parameter sets; QT_SET_III is the reference's compile-time set (main.cuh:13-21) */
#define QT_SET_I 0     /* qTESLA-I     n=512  q=4205569   */
#define QT_SET_III 1   /* qTESLA-III   n=1024 q=8404993   */
#define QT_SET_P_I 2   /* qTESLA-p-I   n=1024 q=343576577 */
#define QT_SET_P_III 3 /* qTESLA-p-III n=2048 q=856145921 */
#define QT_NUM_SETS 4

#define QT_ERR_BAD_SET (-1)
#define QT_ERR_BAD_ARG (-2)
#define QT_ERR_NO_DEVICE (-3)
#define QT_ERR_UNSUPPORTED (-4)
#define QT_ERR_NOMEM (-5)

/* tables of constants.h:3-35 */
#define QT_TABLE_BITREV 0 /* bitrev_tbl / bitrev_tbl_gpu : brv_logn(i)      */
#define QT_TABLE_PHI 1    /* Phi / Phi_gpu               : psi^i            */
#define QT_TABLE_INVPHI 2 /* invPhi / invPhi_gpu         : n^-1 * psi^-i    */
#define QT_TABLE_TF0 3    /* tf0_gpu (tf0 of main.cu:119-124) : omega^i     */
#define QT_TABLE_TI0 4    /* ti0_gpu (ti0 of main.cu:126-129) : omega^-i    */

/* NTT-domain orderings */
#define QT_ORDER_BITREV 0  /* what GS-forward leaves (NTT.cu:2127-2136); the engine's native order */
#define QT_ORDER_NATURAL 1 /* what Stockham leaves (NTT.cu:2040-2049) */

/* Nussbaumer rings */
#define QT_RING_2P32M1 0 /* Z/(2^32-1): bit-exact with nussbaumer_fft (NTT.cu:167-277) */
#define QT_RING_MODQ 1   /* Z_q: equals the NTT product, canonical */
/* Z_q operands THROUGH the ring 2^32-1 (the sparse / small-operand path): canonical residues are
 * centred to (-q/2, q/2], multiplied in Z/(2^32-1) by the kernel of QT_RING_2P32M1 (ring macros NTT.cu:102-134), and the
 * signed lift of the ring value is reduced mod q.  EXACTNESS PRECONDITION: every integer coefficient of the negacyclic
 * product of the centred operands has magnitude < 2^31 (sufficient: max_k sum_i |x~_i| |y~_(k-i)| < 2^31 — qTESLA's
 * s*c and e*c with a small secret and a weight-h ternary challenge for every parameter set; uniform x * weight-h
 * ternary y while h*q/2 < 2^31, i.e. any h <= 255 for qTESLA-I/III, h <= 12 for p-I, h <= 5 for p-III).  Then the
 * result equals qt_polymul bit for bit.  The library cannot check the precondition; outside it (e.g. two uniform
 * operands) the output is a canonical residue but NOT the Z_q product. */
#define QT_RING_2P32M1_LIFT_Q 2

typedef struct qt_ctx qt_ctx;

/* run-time replacement of the macros of main.cuh:13-21 */
typedef struct {
    int set;
    uint32_t n, logn, q;
    uint32_t psi, psi_inv;     /* primitive 2n-th root (nfg0/nig0 role, main.cu:26) */
    uint32_t omega, omega_inv; /* fg0 / ig0, main.cu:26 */
    uint32_t n_inv;            /* Ni, main.cu:26 */
    uint32_t qinv_neg;         /* PARAM_QINV, main.cuh:15 */
    uint32_t barrett_mu48;     /* MIU, main.cuh:20 */
} qt_params;

/* ---- library / parameters (no GPU needed) ---- */
const char* qt_version(void);
const char* qt_error_string(int code);
int qt_get_params(int param_set, qt_params* out);
/* copies one of the constants.h tables (n words) to host memory */
int qt_get_table(int param_set, int which, uint32_t* out_host);
int qt_device_count(int* out);

/* ---- context: owns the device tables and a stream (replaces the per-call cudaMalloc/cudaFree
 *      and __constant__ tables of NTT.cu:2105-2118, constants.h) ---- */
int qt_create(int param_set, int device, qt_ctx** out);
int qt_destroy(qt_ctx* ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream */
int qt_set_stream(qt_ctx* ctx, void* cuda_stream);
int qt_synchronize(qt_ctx* ctx);
/* fused-kernel data path: 0 = automatic, 1 = direct coalesced loads, 2 = TMA bulk copies staged
 * through shared memory with an mbarrier (needs 16-byte aligned operands), 3 = n=2048 only: TMA-staged
 * with the polynomial processed as two 1024-point halves by one warp, 4 = n=2048 only: the same two halves by a
 * PAIR of warps side by side (what "automatic" picks for qTESLA-p-III), 5 = the 23-bit moduli only (qTESLA-I, -III): TMA-staged
 * with the butterflies' quotient estimate taken from the FP64 pipe instead of a mul.hi (bit-identical results; measured
 * slower than 2 on B200 and therefore never picked automatically — DESIGN.md 10); QT_ERR_UNSUPPORTED where a variant does
 * not exist for the context's parameter set */
int qt_set_fused_variant(qt_ctx* ctx, int variant);
/* row products of the Z_q Nussbaumer kernels: 0 = automatic, 1 = schoolbook (the structure of the reference's
 * `naive`, NTT.cu:147-165), 2 = recursive (the 2m length-r products are split once more, 32 = 4*8 / 64 = 8*8),
 * 3 = schoolbook on the FP64 pipe with exact double-precision accumulation (q < 2^25 only, else
 * QT_ERR_UNSUPPORTED).  Results are identical; the ring 2^32-1 always uses the reference's schoolbook order.
 * + 16 (flag): the whole-polynomial kernels instead of the block-pass warp kernel that serves n = 1024 / 2048 by
 * default (kept for A/B measurements; results are identical). */
int qt_set_nussbaumer_variant(qt_ctx* ctx, int variant);
/* Programmatic dependent launch of the TMA-staged kernels: the set-up of a launch (barriers, twiddle table) overlaps
 * the tail of the previous kernel of the stream; operands are touched only after that kernel has completed.
 * 0 = automatic: on when the ctx stream is a non-blocking stream (the own stream; a torch side stream), off on the
 * legacy default stream and other blocking streams, whose implicit ordering against the default stream the early
 * launch is not documented to keep; 1 = never; 2 = always. */
int qt_set_launch_overlap(qt_ctx* ctx, int mode);
int qt_device_malloc(qt_ctx* ctx, size_t bytes, void** out_dev);
int qt_device_free(qt_ctx* ctx, void* dev);
/* pinned host memory for the host-pointer entry points */
int qt_host_alloc(size_t bytes, void** out_host);
int qt_host_free(void* host);
int qt_memcpy_h2d(qt_ctx* ctx, void* dev, const void* host, size_t bytes); /* async on ctx stream */
int qt_memcpy_d2h(qt_ctx* ctx, void* host, const void* dev, size_t bytes); /* async on ctx stream */

/* ---- the hot path, device pointers ---- */
/* forward NTT in place, natural -> NTT domain.  Replaces 10 launches of GS_radix2NTT_gpu0/1/2
 * (NTT.cu:953-1031, driver 2127-2136) plus the Phi scale the GPU drivers forgot (NTT.cu:969-971);
 * result == CPU Phi-scale + radix2NTTGS (NTT.cu:1866-1876). */
int qt_ntt_forward(qt_ctx* ctx, uint32_t* d_a, size_t batch);
/* inverse NTT in place, NTT domain -> natural, includes n^-1 psi^-i.  Replaces radix2INTT_gpu0/1/2
 * (NTT.cu:1374-1433, driver 2151-2160); result == radix2INTT + invPhi scale (NTT.cu:1845-1849). */
int qt_ntt_inverse(qt_ctx* ctx, uint32_t* d_a, size_t batch);
/* The same two transforms with the NTT domain in NATURAL order (position k holds x(psi^(2k+1))) —
 * the ordering the reference's Stockham pipeline works in: forward == Phi scale + radix2NTTStock
 * (NTT.cu:1162-1191; GPU NTTStock_gpu0/1/2 1085-1153, driver 2040-2049), inverse == radix2INTTStock
 * + invPhi scale (NTT.cu:1339-1370; GPU INTTStock_gpu0/1/2 1268-1337, driver 2058-2067).  In place,
 * one launch, no scratch array; equal to qt_ntt_forward followed by qt_bitrev_copy (resp. preceded). */
int qt_ntt_forward_natural(qt_ctx* ctx, uint32_t* d_a, size_t batch);
int qt_ntt_inverse_natural(qt_ctx* ctx, uint32_t* d_a, size_t batch);
/* c = a*b mod q, element-wise.  Replaces pointwise_mult (NTT.cu:1155-1160). d_c may alias.  16-byte aligned
 * operands use 128-bit accesses; any 4-byte aligned pointer is accepted (word-wise kernel). */
int qt_pointwise(qt_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, uint32_t* d_c, size_t batch);
/* z = x*y mod (X^n+1, q), fused forward -> pointwise -> inverse in ONE launch; HBM is touched once
 * per operand.  Replaces the 31-34 launches of test_NTT_{Stockham,GS_CT,CT_CT,GS_GS,CT_GS}_nega_gpu
 * (NTT.cu:2008-2443, between the memcpys).  d_z may alias d_x or d_y. */
int qt_polymul(qt_ctx* ctx, const uint32_t* d_x, const uint32_t* d_y, uint32_t* d_z, size_t batch);
/* z[b] = a[b]*y[b] with NTT(a) supplied: d_a_hat is what qt_ntt_forward left (canonical, bit-reversed
 * order).  broadcast != 0: ONE polynomial a_hat (n words) multiplies every y[b] — qTESLA's own shape
 * (public a, many secrets/challenges; the caller of the reference's path pays the transform of a once
 * instead of per product).  broadcast == 0: one a_hat per product (batch*n words).  Operands must be
 * 16-byte aligned.  d_z may alias d_y. */
int qt_polymul_ntt(qt_ctx* ctx, const uint32_t* d_a_hat, int broadcast, const uint32_t* d_y, uint32_t* d_z,
                   size_t batch);
/* reorder between the two NTT-domain orderings; replaces bit_reverse_copy_tbl_gpu (NTT.cu:487-492).
 * d_out must not alias d_in. */
int qt_bitrev_copy(qt_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, size_t batch);
/* Nussbaumer negacyclic product (NTT-free).  ring = QT_RING_2P32M1 reproduces nussbaumer_fft
 * (NTT.cu:167-277, CPU-only and single-polynomial in the reference) bit for bit, batched;
 * ring = QT_RING_MODQ runs the same structure over Z_q and equals qt_polymul;
 * ring = QT_RING_2P32M1_LIFT_Q: see the definition above (exact only under its precondition). */
int qt_nussbaumer(qt_ctx* ctx, const uint32_t* d_x, const uint32_t* d_y, uint32_t* d_z, size_t batch,
                  int ring);
/* synthetic operands: a[i] = splitmix64(seed + first_index + i) % q, i in [0, count) */
int qt_fill_uniform(qt_ctx* ctx, uint32_t* d_a, size_t count, uint64_t seed, uint64_t first_index);

/* ---- harness-equivalent entry points, HOST pointers (what main.cu:203-210 calls) ----
 * x, y, z are caller-owned host arrays of batch*n words.  H2D of x and y, the fused kernel and D2H
 * of z are pipelined in chunks on the context's GPU; returns after z is complete.  Pinned arrays
 * (qt_host_alloc / cudaHostAlloc / cudaHostRegister) are transferred in place (71 GB/s over PCIe on
 * the B200 box); ordinary malloc'd arrays — what the reference's main.cu passes — go through the
 * library's own pinned staging buffers with several copy threads (49 GB/s; a plain cudaMemcpyAsync
 * of pageable memory gives 13 GB/s).  Each of x, y, z may be of either kind.  Same role as test_NTT_*_nega_gpu (main.cuh:66-70) without the
 * fixed x=y=1 fill, the timing prints and the per-call allocation. */
int qt_polymul_host(qt_ctx* ctx, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch);
/* In-process multi-GPU form: the batch is sharded contiguously over the first ngpus devices
 * (ngpus <= 0: all), GPU g owns polynomials [g*B/G, (g+1)*B/G); one context + one persistent host thread per
 * device, each bound to its GPU's NUMA node; no collective, no peer access.  A handle is not thread-safe; two
 * callers use two handles (the library keeps no process-global state). */
typedef struct qt_multi qt_multi;
int qt_multi_create(int param_set, int ngpus, qt_multi** out);
int qt_multi_destroy(qt_multi* m);
int qt_multi_gpus(qt_multi* m, int* out);
int qt_multi_polymul_host(qt_multi* m, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch);
/* one-shot convenience: qt_multi_create + qt_multi_polymul_host + qt_multi_destroy (pays the set-up every call) */
int qt_polymul_host_multi(int param_set, const uint32_t* x, const uint32_t* y, uint32_t* z,
                          size_t batch, int ngpus);
/* no-op, kept for ABI compatibility (earlier versions cached per-device contexts process-wide) */
int qt_shutdown(void);
int qt_nussbaumer_host(qt_ctx* ctx, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch,
                       int ring);

/* ---- CUDA graphs for launch-bound batches (no reference counterpart: the reference launches 31-34 kernels per
 *      product from the host, NTT.cu:2127-2160).  Between qt_graph_begin and qt_graph_end the device-pointer entry
 *      points of this context (qt_polymul, qt_polymul_ntt, qt_ntt_*, qt_pointwise, qt_bitrev_copy, qt_nussbaumer,
 *      qt_fill_uniform, qt_memcpy_*) are RECORDED on the context's stream instead of executed; qt_graph_launch replays
 *      the whole sequence with one host call.  The pointers and batch sizes are baked in.  The host-pointer forms
 *      (qt_polymul_host...) and qt_synchronize must not be called while recording. ---- */
typedef struct qt_graph qt_graph;
int qt_graph_begin(qt_ctx* ctx);
int qt_graph_end(qt_ctx* ctx, qt_graph** out);
int qt_graph_launch(qt_ctx* ctx, qt_graph* graph); /* asynchronous on the ctx stream */
int qt_graph_destroy(qt_graph* graph);
int qt_graph_kernel_count(qt_graph* graph, uint64_t* out); /* kernels one replay launches */

/* ---- host placement (no reference counterpart: the reference drives one GPU from pageable buffers,
 *      NTT.cu:2105-2124).  Where a GPU hangs in the machine, and binding the CALLING host thread to the CPUs
 *      of that GPU's NUMA node, so that pinned buffers the thread allocates and touches afterwards sit on the
 *      memory controller next to the GPU's PCIe root.  Call it before qt_host_alloc / before creating the
 *      worker threads that feed this GPU.  Unknown topology (sysfs says -1, e.g. inside a VM) is not an error:
 *      the thread is left alone and *cpus_out = 0. ---- */
int qt_device_pci_bus_id(int device, char* out, size_t len); /* "0000:1b:00.0", lower case; len >= 13 */
int qt_device_numa_node(int device, int* node_out);          /* -1 when the platform does not say */
int qt_bind_thread_to_device(int device, int* cpus_out);     /* *cpus_out = CPUs the thread is now bound to, 0 = unchanged */

/* ---- introspection for benchmarks ---- */
/* kernels launched by this context since creation (the bench's gpu_launches claim) */
int qt_launch_count(qt_ctx* ctx, uint64_t* out);
/* grid/block/shared-memory/occupancy of the fused kernel for this context */
int qt_kernel_info(qt_ctx* ctx, int* grid, int* block, int* smem_bytes, int* blocks_per_sm,
                   int* num_sms);

#ifdef __cplusplus
}
#endif
#endif /* QTESLA_B200_H */
