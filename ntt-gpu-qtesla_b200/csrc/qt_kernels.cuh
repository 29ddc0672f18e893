// qt_kernels.cuh — __global__ kernels built from the warp-tile engine (qt_tile.cuh).
//
// All kernels are persistent: grid = (#SMs x resident CTAs), every warp walks tiles
// tile = blockIdx.x*WARPS + warp, += gridDim.x*WARPS.  The per-lane twiddle tables are copied
// from global to shared memory once per CTA.
#pragma once
#include <cuda_runtime.h>

#include "qt_tile.cuh"

namespace qt {

constexpr int WARPS_PER_CTA = 8;   // direct-load kernels
// The warp index as a value the compiler KNOWS to be the same in all lanes (the shuffle of lane 0's copy): tile indices,
// buffer and barrier addresses and the bulk-copy descriptors derived from it live in uniform registers, the tile loop is
// uniform control flow, and ptxas stops duplicating the first forward pass around the barrier wait (k_polymul_tma<III>:
// 2 304 instead of 2 840 static instructions).  Measured per kernel (run r02F): fused n=512 512.1 vs 504.6, p-I 220.0 vs 214.6,
// n=2048 (pair) 92.4 vs 89.8, n=1024 245.8 vs 245.1 M polymul/s, single forward transform 641.8 vs 629.4 M polynomials/s,
// Nussbaumer ring 52.5 vs 49.7 — but the natural-order transforms LOSE 12 % (541.9 vs 614.5) and the split-tile cached product
// 7 % (109.7 vs 117.9), so those kernels (and the direct-load ones, which were not measured) keep the plain index.
#ifndef QT_UNIFORM_WARP
#define QT_UNIFORM_WARP 1
#endif
template <bool UNIFORM> __device__ __forceinline__ uint32_t warp_index() {
    return (UNIFORM && QT_UNIFORM_WARP) ? __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) : threadIdx.x >> 5;
}
// Geometry of the TMA-staged fused kernel, chosen by measurement on B200 (tools/ab.py):
// ONE 16-warp CTA per SM beat 2x8, 3x8, 2x12 and 5x4 warps (221 vs 208-215 M polymul/s at n=1024).
#ifndef QT_TMA_WARPS
#define QT_TMA_WARPS 16            // warps per CTA (32 coefficients per thread)
#endif
#ifndef QT_TMA_WARPS_E64
#define QT_TMA_WARPS_E64 12        // n=2048 (64 coefficients per thread): 16 KB of staging per warp
#endif
#ifndef QT_TMA_MINB
#define QT_TMA_MINB 1              // resident CTAs per SM the kernel is compiled for (register budget)
#endif

// launch geometry of the TMA-staged fused kernel per parameter set
// Programmatic dependent launch (QT_PDL): the TMA-staged kernels let the NEXT launch of the stream start its CTAs
// as soon as this grid's CTAs leave their SMs (pdl_launch_dependents at the top); everything a kernel does before
// pdl_wait() touches only shared memory and the constant twiddle table, so barrier set-up and the table copy
// overlap the tail of the previous launch.  pdl_wait() returns when the previous grid has completed and its
// writes are visible; it is a no-op for a launch without the attribute.  Measured (run r01s): n=1024 batch 65 536
// 240.8 -> 244.5 M polymul/s, batch 1 024 10.3 -> 7.5 us per launch.
#ifndef QT_PDL
#define QT_PDL 1
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if QT_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if QT_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Warps per CTA of the FUSED kernel, per parameter set (run r02t; only multiples of four — the four schedulers of an SM
// must get the same number of warps: 18 and 22 lose 8-10 %):
//   qTESLA-III : 16 / 20 / 24 warps -> 243.3 / 245.5 / 243.1 M polymul/s
//   qTESLA-p-I : 16 / 20 / 24 warps -> 209.3 / 211.0 / 214.0  (24 warps = 80 registers, 16 bytes of spill, still faster)
//   qTESLA-I   : 16 / 20 / 24 warps -> 505.5 / 505.7 / 493.5
#ifndef QT_FUSED_WARPS_III
#define QT_FUSED_WARPS_III 20
#endif
#ifndef QT_TMA_WARPS_N1024
#define QT_TMA_WARPS_N1024 20
#endif
#ifndef QT_FUSED_WARPS_P_I
#define QT_FUSED_WARPS_P_I 24
#endif
#ifndef QT_FUSED_WARPS_I
#define QT_FUSED_WARPS_I QT_TMA_WARPS
#endif
template <int SET> struct TmaCfg {
    // single transforms and the cached-transform product (run r02w, 16 / 20 / 24 warps): n=1024 forward 617 / 625 / 479 M
    // polynomials/s (p-I 585 / 602 / 501), cached product 359 / 360 / 356; n=512 forward 1206 / 1182 / 1175, cached 737 / 708 / 702
    static constexpr int WARPS = (Cfg<SET>::E == 64) ? QT_TMA_WARPS_E64 : (Cfg<SET>::N == 512 ? QT_TMA_WARPS : QT_TMA_WARPS_N1024);
    static constexpr int FUSED_WARPS = (Cfg<SET>::E == 64) ? QT_TMA_WARPS_E64
                                       : SET == SET_III ? QT_FUSED_WARPS_III : SET == SET_P_I ? QT_FUSED_WARPS_P_I : QT_FUSED_WARPS_I;
    static constexpr int MAX_WARPS = WARPS > FUSED_WARPS ? WARPS : FUSED_WARPS;
    static constexpr int MINB = (Cfg<SET>::E == 64) ? 1 : QT_TMA_MINB;  // 64 coefficients per thread need the registers
};

template <int SET> struct KernelShape {
    using T = Tile<SET>;
    static constexpr size_t TW_QUADS = T::TABLE_QUADS;  // the kernel's table block (qt_tile.cuh: lane_ptrs)
    static constexpr size_t TW_BYTES = TW_QUADS * sizeof(TwQuad);
    static constexpr size_t BUF_BYTES = (size_t)WARPS_PER_CTA * T::C::TILE_WORDS * sizeof(uint32_t);
    static constexpr size_t SMEM_DIRECT = TW_BYTES + BUF_BYTES;
};

__device__ __forceinline__ void copy_table_to_smem(TwQuad* dst, const TwQuad* __restrict__ src, size_t quads) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (size_t i = threadIdx.x; i < quads; i += blockDim.x) d[i] = __ldg(s + i);
}

// z = x*y mod (X^n+1, q): forward(x), forward(y), pointwise, inverse — one launch, HBM touched
// once per operand (12n bytes per product).  Direct-load variant: coalesced LDG/STG, both operands
// in registers.  Used when the operands are not 16-byte aligned (the TMA variant needs that).
template <int SET>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_polymul(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    __syncthreads();

    const uint32_t warp = warp_index<false>(), lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;

    for (size_t tile = (size_t)blockIdx.x * WARPS_PER_CTA + warp; tile < ntiles;
         tile += (size_t)gridDim.x * WARPS_PER_CTA) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        uint32_t vx[T::E], vy[T::E];
        T::load_rows(vx, x + base, lane, valid);
        T::load_rows(vy, y + base, lane, valid);
        T::fwd_rows(vx);
        T::sts_rows(vx, buf, lane);
        __syncwarp();
        T::lds_cols(vx, buf, lane);
        __syncwarp();
        T::fwd_rows(vy);
        T::sts_rows(vy, buf, lane);
        T::fwd_cols(vx, P.fwd);
        __syncwarp();
        T::lds_cols(vy, buf, lane);
        T::fwd_cols(vy, P.fwd);
        T::pointwise_mont(vy, vx);
        T::inv_cols(vy, P.inv);
        __syncwarp();
        T::sts_cols(vy, buf, lane);
        __syncwarp();
        T::lds_rows(vy, buf, lane);
        __syncwarp();
        T::template inv_rows<UNI_INV_FUSED>(vy, P);
        T::store_rows(vy, z + base, lane, valid);
    }
}

// ---- TMA-staged variant (the default) ------------------------------------------------------------
// Same arithmetic.  Each warp owns two shared-memory buffers A and B of one tile each and two
// mbarriers.  The operands of the warp's NEXT tile are fetched by 1-D bulk copies (cp.async.bulk =
// the TMA unit, SASS UBLKCP) that complete on the mbarriers while the current tile is computed, so
// no register is tied up by a load in flight and the HBM latency is off the critical path.  The
// buffers are recycled within a tile:
//     A: x staging -> transposition scratch of x -> stash of NTT(x)      -> (free) next x
//     B: y staging -> transposition scratch of y and of the inverse pass -> (free) next y
// so a warp needs 2 tiles of shared memory and ~E+temporaries registers, which is what lets three
// 8-warp CTAs (24 warps) share an SM.  The two forward transforms run through ONE copy of the code
// (a 2-trip loop), halving the instruction-cache footprint.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni QT_DONE;\n"
        "bra.uni QT_WAIT;\n"
        "QT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses of this thread are ordered before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A/B switch: write z with ONE bulk copy per tile from a third per-warp shared-memory buffer
// (cp.async.bulk.global.shared::cta) instead of 32 coalesced 32-bit global stores per thread.
#ifndef QT_TMA_STORE
#define QT_TMA_STORE 0
#endif
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int SET> struct StageShape {
    using T = Tile<SET>;
    static constexpr uint32_t PAD = (T::PPW == 2) ? 16 : 0;  // de-conflicts the rows of the two polynomials
    static constexpr uint32_t POLY_STRIDE = T::N + PAD;      // words
    static constexpr uint32_t WORDS = T::PPW * POLY_STRIDE;  // one buffer (>= TILE_WORDS)
    static constexpr uint32_t BUFS = 2 + QT_TMA_STORE;       // x staging, y staging (+ z staging for the bulk-store variant)
    static constexpr size_t SMEM = KernelShape<SET>::TW_BYTES + (size_t)TmaCfg<SET>::MAX_WARPS * BUFS * WORDS * sizeof(uint32_t) +
                                   (size_t)TmaCfg<SET>::MAX_WARPS * 2 * sizeof(uint64_t);
    static constexpr size_t SMEM_BCAST = SMEM + T::N * sizeof(uint32_t);  // + the broadcast a_hat of k_polymul_ntt
    static __device__ __forceinline__ uint32_t off(uint32_t lane, uint32_t r) {
        return (lane / T::LPP) * POLY_STRIDE + (lane % T::LPP) + T::LPP * r;
    }
};

template <int SET>
__global__ void __launch_bounds__(TmaCfg<SET>::FUSED_WARPS * 32, TmaCfg<SET>::MINB)
k_polymul_tma(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    using G = StageShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    const int NW = (int)(blockDim.x >> 5);  // <= TmaCfg<SET>::WARPS; small batches are launched with fewer warps (less shared memory: the next launch's CTA fits beside this one)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NW * G::BUFS * G::WORDS);

    const uint32_t warp = warp_index<true>(), lane = threadIdx.x & 31;
    uint32_t* A = s_stage + warp * G::BUFS * G::WORDS;
    uint32_t* B = A + G::WORDS;
    uint64_t* bar_a = s_bar + 2 * warp;
    uint64_t* bar_b = bar_a + 1;
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    const size_t stride = (size_t)gridDim.x * NW;
    size_t tile = (size_t)warp * gridDim.x + blockIdx.x;  // SM-interleaved: a partial last round spreads over all SMs

    auto issue = [&](const uint32_t* g, uint32_t* st, uint64_t* bar, size_t t) {  // one lane
        const size_t p0 = t * T::PPW;
        const uint32_t np = (uint32_t)((batch - p0 < T::PPW) ? batch - p0 : T::PPW);
        mbar_expect_tx(bar, np * T::N * (uint32_t)sizeof(uint32_t));
        if (G::PAD == 0) {
            bulk_g2s(st, g + p0 * T::N, np * T::N * (uint32_t)sizeof(uint32_t), bar);
        } else {
            for (uint32_t p = 0; p < np; p++)
                bulk_g2s(st + p * G::POLY_STRIDE, g + (p0 + p) * T::N, T::N * (uint32_t)sizeof(uint32_t), bar);
        }
    };

    pdl_launch_dependents();
    if (lane == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    pdl_wait();  // the operands may be the previous launch's output
    if (lane == 0 && tile < ntiles) {
        issue(x, A, bar_a, tile);
        issue(y, B, bar_b, tile);
    }
    __syncthreads();

    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    uint32_t phase = 0;
    for (; tile < ntiles; tile += stride, phase ^= 1) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        const bool more = tile + stride < ntiles;
        uint32_t v[T::E];
#pragma unroll 1
        for (int op = 0; op < 2; op++) {  // one copy of the forward-transform code for x and y
            uint32_t* st = op ? B : A;
            mbar_wait(op ? bar_b : bar_a, phase);
#pragma unroll
            for (uint32_t r = 0; r < T::E; r++) v[r] = st[G::off(lane, r)];
            __syncwarp();
            T::fwd_rows(v);
            T::sts_rows(v, st, lane);
            __syncwarp();
            T::lds_cols(v, st, lane);
            T::fwd_cols(v, P.fwd);
            if (op == 0) {
                __syncwarp();             // every lane has read its columns before A is overwritten
                T::sts_cols(v, A, lane);  // stash NTT(x); each lane reads back only what it wrote
            }
        }
        T::pointwise_mont_stash(v, A, lane);
        fence_proxy_async();
        __syncwarp();                     // A is free: fetch the next tile's x into it
        if (more && lane == 0) issue(x, A, bar_a, tile + stride);
        T::inv_cols(v, P.inv);
        T::sts_cols(v, B, lane);          // (all lanes passed the __syncwarp above after reading B)
        __syncwarp();
        T::lds_rows(v, B, lane);
        fence_proxy_async();
        __syncwarp();                     // B is free: fetch the next tile's y into it
        if (more && lane == 0) issue(y, B, bar_b, tile + stride);
        T::template inv_rows<UNI_INV_FUSED>(v, P);
#if QT_TMA_STORE
        {
            uint32_t* Cz = B + G::WORDS;
            if (lane == 0) bulk_wait_read();  // the previous tile's bulk store has read the buffer
            __syncwarp();
#pragma unroll
            for (uint32_t r = 0; r < T::E; r++) Cz[G::off(lane, r)] = v[r];
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const size_t p0 = tile * T::PPW;
                const uint32_t np = (uint32_t)((batch - p0 < T::PPW) ? batch - p0 : T::PPW);
                for (uint32_t p = 0; p < np; p++)
                    bulk_s2g(z + (p0 + p) * T::N, Cz + p * G::POLY_STRIDE, T::N * (uint32_t)sizeof(uint32_t));
            }
        }
#else
        T::store_rows(v, z + base, lane, valid);
#endif
    }
#if QT_TMA_STORE
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory must outlive the copies
#endif
}

// ---- FP64-quotient variant of the fused kernel (signed-lazy sets: qTESLA-I, qTESLA-III) ----------------------------
// Same staging, same buffers, same transposition pattern as k_polymul_tma; the butterflies take their quotient estimate
// from the FP64 pipe (Tile::ct_dq) instead of a mul.hi, which halves the multiply-pipe time of a product.  Values are
// register pairs whose high half is the zero high half of an FP64 result (qt_tile.cuh); the table block holds the
// twiddles as w in [0, q) (TwU2) and w / q (TwW2).
#ifndef QT_DQ_WARPS
#define QT_DQ_WARPS 16
#endif
template <int SET> struct DqShape {
    using T = Tile<SET>;
    using G = StageShape<SET>;
    static constexpr int WARPS = QT_DQ_WARPS;
    static constexpr size_t QUADS = KernelShape<SET>::TW_QUADS;
    static constexpr size_t TABLE_BYTES = QUADS * (sizeof(TwW2) + sizeof(TwU2));
    static constexpr size_t WARP_BYTES = 2 * G::WORDS * sizeof(uint32_t) + 2 * sizeof(uint64_t);
    static constexpr size_t SMEM = TABLE_BYTES + (size_t)WARPS * WARP_BYTES;
};

template <int SET>
__global__ void __launch_bounds__(DqShape<SET>::WARPS * 32, 1)
k_polymul_dq(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, const TwU2* __restrict__ g_laneU, const TwW2* __restrict__ g_laneW) {
    using T = Tile<SET>;
    using D = DqShape<SET>;
    using G = StageShape<SET>;
    using P64 = typename T::P64;
    static_assert(T::LAZY && QT_TMA_STORE == 0, "FP64-quotient kernel: signed-lazy sets");
    extern __shared__ uint4 smem_raw[];
    __shared__ double s_tiny;  // the denormal 2^-1074: the seeds of the register pairs are products with it
    TwW2* s_twW = reinterpret_cast<TwW2*>(smem_raw);
    TwU2* s_twU = reinterpret_cast<TwU2*>(s_twW + D::QUADS);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_twU + D::QUADS);
    const int NW = (int)(blockDim.x >> 5);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NW * 2 * G::WORDS);

    const uint32_t warp = warp_index<true>(), lane = threadIdx.x & 31;
    uint32_t* A = s_stage + warp * 2 * G::WORDS;
    uint32_t* B = A + G::WORDS;
    uint64_t* bar_a = s_bar + 2 * warp;
    uint64_t* bar_b = bar_a + 1;
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    const size_t stride = (size_t)gridDim.x * NW;
    size_t tile = (size_t)warp * gridDim.x + blockIdx.x;

    auto issue = [&](const uint32_t* g, uint32_t* st, uint64_t* bar, size_t t) {  // one lane
        const size_t p0 = t * T::PPW;
        const uint32_t np = (uint32_t)((batch - p0 < T::PPW) ? batch - p0 : T::PPW);
        mbar_expect_tx(bar, np * T::N * (uint32_t)sizeof(uint32_t));
        if (G::PAD == 0) {
            bulk_g2s(st, g + p0 * T::N, np * T::N * (uint32_t)sizeof(uint32_t), bar);
        } else {
            for (uint32_t p = 0; p < np; p++)
                bulk_g2s(st + p * G::POLY_STRIDE, g + (p0 + p) * T::N, T::N * (uint32_t)sizeof(uint32_t), bar);
        }
    };

    pdl_launch_dependents();
    if (lane == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (threadIdx.x == 0) s_tiny = __longlong_as_double(1ll);
    {
        const uint4* sw = reinterpret_cast<const uint4*>(g_laneW);
        uint4* dw = reinterpret_cast<uint4*>(s_twW);
        for (size_t i = threadIdx.x; i < D::QUADS; i += blockDim.x) dw[i] = __ldg(sw + i);
        const uint2* su = reinterpret_cast<const uint2*>(g_laneU);
        uint2* du = reinterpret_cast<uint2*>(s_twU);
        for (size_t i = threadIdx.x; i < D::QUADS; i += blockDim.x) du[i] = __ldg(su + i);
    }
    pdl_wait();
    if (lane == 0 && tile < ntiles) {
        issue(x, A, bar_a, tile);
        issue(y, B, bar_b, tile);
    }
    __syncthreads();

    const typename T::LanePtrsDQ P = T::lane_ptrs_dq(s_twU, s_twW, lane);
    // seed of register r: the (zero) high half of an FP64 product made here and now — a volatile operand keeps the
    // compiler from hoisting the sixteen products of a pass out of the loops and copying their results into place
#define QT_DQ_SEED(r) (P64{tiny * uni_W<SET, UNI_FWD>(r)})
    uint32_t phase = 0;
    for (; tile < ntiles; tile += stride, phase ^= 1) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        const bool more = tile + stride < ntiles;
        P64 v[T::E];
#pragma unroll 1
        for (int op = 0; op < 2; op++) {  // one copy of the forward-transform code for x and y
            uint32_t* st = op ? B : A;
            mbar_wait(op ? bar_b : bar_a, phase);
            {
                const double tiny = *(volatile double*)&s_tiny;
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) {
                    v[r] = (QT_DQ_SWAP || r >= T::E / 2) ? QT_DQ_SEED(r).with_lo(st[G::off(lane, r)]) : P64::make(st[G::off(lane, r)], 0u);
                }
            }
            __syncwarp();
            T::fwd_rows_dq(v);
            if (T::PPW == 2) {
                const typename T::RowBases RB = T::row_bases(lane);
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) st[T::rows_addr(RB, r)] = v[r].lo();
            } else {
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) st[T::swz(T::row_off(lane, r))] = v[r].lo();
            }
            __syncwarp();
            {
                const double tiny = *(volatile double*)&s_tiny;
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) {
                    v[r] = (QT_DQ_SWAP || T::dq_cols_first_y(r)) ? QT_DQ_SEED(r).with_lo(st[T::swz(T::E * lane + r)]) : P64::make(st[T::swz(T::E * lane + r)], 0u);
                }
            }
            T::fwd_cols_dq(v, P);
            if (op == 0) {
                __syncwarp();             // every lane has read its columns before A is overwritten
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) A[T::swz(T::E * lane + r)] = v[r].lo();  // stash NTT(x); each lane reads back only what it wrote
            }
        }
        T::pointwise_dq_stash(v, A, lane);
        fence_proxy_async();
        __syncwarp();                     // A is free: fetch the next tile's x into it
        if (more && lane == 0) issue(x, A, bar_a, tile + stride);
        T::inv_cols_dq(v);
#pragma unroll
        for (uint32_t r = 0; r < T::E; r++) B[T::swz(T::E * lane + r)] = v[r].lo();
        __syncwarp();
        {
            const double tiny = *(volatile double*)&s_tiny;
            uint32_t ld[T::E];
            if (T::PPW == 2) {
                const typename T::RowBases RB = T::row_bases(lane);
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) ld[r] = B[T::rows_addr(RB, r)];
            } else {
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) ld[r] = B[T::swz(T::row_off(lane, r))];
            }
#pragma unroll
            for (uint32_t r = 0; r < T::E; r++) v[r] = (QT_DQ_SWAP || T::dq_rows_first_y(r)) ? QT_DQ_SEED(r).with_lo(ld[r]) : P64::make(ld[r], 0u);
        }
        fence_proxy_async();
        __syncwarp();                     // B is free: fetch the next tile's y into it
        if (more && lane == 0) issue(y, B, bar_b, tile + stride);
        uint32_t out[T::E];
        T::inv_rows_dq(v, out, P);
        T::store_rows(out, z + base, lane, valid);
    }
#undef QT_DQ_SEED
}

// Fused product for n = 2048 (qTESLA-p-III) on the SPLIT tile: a polynomial is two 1024-point halves
// joined by one level, and every transform runs the 32-coefficients-per-thread half code twice (a rolled
// loop) instead of one 64-coefficients-per-thread pass.  Why: the 64-wide kernel is 7360 instructions
// (118 KB) of straight-line code, at the edge of the instruction cache — ncu shows 0.5 "no instruction"
// stall per issue against 0.02 for the n=1024 kernels (profiles/ncu_fused_p3_r01n.json, tools/icache.cu) —
// and needs 167 registers.  Same buffers as k_polymul_tma: A = x staging -> scratch / stash of the second
// half -> NTT(x) -> next x;  B = y staging -> scratch / stashes -> next y.  Each 1024-word half of a buffer
// is used with the half tile's own (swizzled) rows/cols patterns.
#ifndef QT_SPLIT_WARPS
#define QT_SPLIT_WARPS 12
#endif
struct SplitShape {
    using T = Tile<SET_P_III_H>;
    static constexpr int WARPS = QT_SPLIT_WARPS;
    static constexpr uint32_t HALF = T::N, WORDS = 2 * T::N;          // one buffer = one polynomial
    static constexpr size_t TW_BYTES = (size_t)T::TABLE_QUADS * sizeof(TwQuad);
    static constexpr size_t STAGE_BYTES = (size_t)WARPS * 2 * WORDS * sizeof(uint32_t);
    static constexpr size_t BAR_BYTES = (size_t)WARPS * 2 * sizeof(uint64_t);
    static constexpr size_t SMEM = TW_BYTES + STAGE_BYTES + BAR_BYTES;
    static constexpr size_t SMEM_BCAST = SMEM + WORDS * sizeof(uint32_t);  // + the broadcast a_hat (MODE 1)
};

// MODE 0: z = x*y (x in the first argument).  MODE 1 / 2: the first argument is NTT(a) as qt_ntt_forward
// leaves it (qt_polymul_ntt): 1 = one a_hat for the whole batch, kept in shared memory; 2 = one per product.
template <int MODE>
__global__ void __launch_bounds__(SplitShape::WARPS * 32, 1)
k_polymul_split(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = SplitShape::T;
    using G = SplitShape;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tw + T::TABLE_QUADS);
    const int NW = (int)(blockDim.x >> 5);  // <= SplitShape::WARPS
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NW * 2 * G::WORDS);
    const uint32_t warp = warp_index<false>(), lane = threadIdx.x & 31;
    uint32_t* A = s_stage + warp * 2 * G::WORDS;
    uint32_t* B = A + G::WORDS;
    uint64_t* bar_a = s_bar + 2 * warp;
    uint64_t* bar_b = bar_a + 1;
    const size_t stride = (size_t)gridDim.x * NW;
    size_t tile = (size_t)warp * gridDim.x + blockIdx.x;  // SM-interleaved: a partial last round spreads over all SMs  // one polynomial per tile
    auto issue = [&](const uint32_t* g, uint32_t* st, uint64_t* bar, size_t t) {  // one lane
        mbar_expect_tx(bar, G::WORDS * (uint32_t)sizeof(uint32_t));
        bulk_g2s(st, g + t * G::WORDS, G::WORDS * (uint32_t)sizeof(uint32_t), bar);
    };
    pdl_launch_dependents();
    if (lane == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    copy_table_to_smem(s_tw, g_lane, T::TABLE_QUADS);
    pdl_wait();  // the operands may be the previous launch's output
    if (lane == 0 && tile < batch) {
        if (MODE == 0) issue(x, A, bar_a, tile);
        issue(y, B, bar_b, tile);
    }
    // MODE 1: the one a_hat stays in shared memory (behind the barriers), each half in the half tile's
    // swizzled cols layout
    uint32_t* s_ahat = reinterpret_cast<uint32_t*>(s_bar + 2 * NW);
    if (MODE == 1) {
        for (uint32_t i = threadIdx.x; i < G::WORDS / 4; i += blockDim.x)
            *reinterpret_cast<uint4*>(s_ahat + (4 * i / G::HALF) * G::HALF + T::swz(4 * i % G::HALF)) =
                __ldg(reinterpret_cast<const uint4*>(x) + i);
    }
    __syncthreads();

    uint32_t phase = 0;
    for (uint32_t k = 0; tile < batch; tile += stride, phase ^= 1, k++) {
        const bool more = tile + stride < batch;
        // y staging.  MODE 0: always B (x's transform hides the copy).  MODE 1/2: no x, so the two buffers
        // alternate and the next y is requested a whole tile ahead.
        uint32_t* ybuf = (MODE == 0 || !(k & 1)) ? B : A;
        uint64_t* ybar = (MODE == 0 || !(k & 1)) ? bar_b : bar_a;
        const uint32_t yphase = (MODE == 0) ? phase : ((k >> 1) & 1);
        if (MODE != 0 && more && lane == 0)  // (the other buffer was released at the end of the previous tile)
            issue(y, (k & 1) ? B : A, (k & 1) ? bar_b : bar_a, tile + stride);
        const uint32_t* ah = x + tile * G::WORDS + T::E * lane;  // MODE 2: this lane's line of each half
        if (MODE == 2) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ah));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ah + G::HALF));
        }
        uint32_t v[T::E];
#pragma unroll 1
        for (int op = (MODE == 0 ? 0 : 1); op < 2; op++) {  // x, then y (which continues into the inverse)
            uint32_t* st = op ? ybuf : A;
            mbar_wait(op ? ybar : bar_a, op ? yphase : phase);
            {
                uint32_t hi[T::E];
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) {
                    v[r] = st[lane + 32 * r];
                    hi[r] = st[G::HALF + lane + 32 * r];
                }
                __syncwarp();  // the staging buffer becomes scratch
                T::split_fwd(v, hi);
                T::sts_rows(hi, st + G::HALF, lane);  // parked; read back by the same lane
            }
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                uint32_t* sh = st + h * G::HALF;
                if (h) T::lds_rows(v, sh, lane);
                T::fwd_rows(v, 32 * h);
                T::sts_rows(v, sh, lane);
                __syncwarp();
                T::lds_cols(v, sh, lane);
                T::fwd_cols(v, s_tw + h * T::TW_QUADS + lane);
                if (MODE == 0 && op == 0) {
                    __syncwarp();                // every lane has read its columns
                    T::sts_cols(v, sh, lane);    // NTT(x) half h; each lane reads back only what it wrote
                    continue;
                }
                if (MODE == 0) {
                    T::pointwise_mont_stash(v, A + h * G::HALF, lane);
                    if (h) {                     // A is free: fetch the next x
                        fence_proxy_async();
                        __syncwarp();
                        if (more && lane == 0) issue(x, A, bar_a, tile + stride);
                    }
                } else {
#pragma unroll
                    for (uint32_t c = 0; c < T::E / 4; c++) {
                        const uint4 u = (MODE == 1)
                            ? *reinterpret_cast<const uint4*>(s_ahat + h * G::HALF + T::swz(T::E * lane + 4 * c))
                            : __ldg(reinterpret_cast<const uint4*>(ah + h * G::HALF) + c);
                        v[4 * c] = T::pw_canonical(v[4 * c], u.x);
                        v[4 * c + 1] = T::pw_canonical(v[4 * c + 1], u.y);
                        v[4 * c + 2] = T::pw_canonical(v[4 * c + 2], u.z);
                        v[4 * c + 3] = T::pw_canonical(v[4 * c + 3], u.w);
                    }
                }
                T::inv_cols(v, s_tw + (1 - h) * T::TW_QUADS + (T::BLOCKS - 1 - lane));  // mirrored table of the OTHER half
                __syncwarp();
                T::sts_cols(v, sh, lane);
                __syncwarp();
                T::lds_rows(v, sh, lane);
                T::template inv_rows<UNI_INV_FUSED>(v, typename T::LanePtrs{nullptr, nullptr, nullptr}, 32 * h);
                if (h == 0) T::sts_rows(v, sh, lane);  // parked in this lane's own slots until half 1 is done
            }
        }
        // join the halves: last inverse level + scale, canonical
        uint32_t* zt = z + tile * G::WORDS;
#pragma unroll
        for (uint32_t r = 0; r < T::E; r++) {
            uint32_t a = ybuf[T::swz(T::row_off(lane, r))];
            T::template split_inv<UNI_INV_FUSED>(a, v[r]);
            zt[lane + 32 * r] = a;
            zt[G::HALF + lane + 32 * r] = v[r];
        }
        fence_proxy_async();
        __syncwarp();  // the y buffer is free
        if (MODE == 0 && more && lane == 0) issue(y, B, bar_b, tile + stride);
    }
}

// Fused product for n = 2048 with TWO WARPS per polynomial (run r02h).  Same split tile, same buffers per polynomial
// as k_polymul_split<0>, but after the level that joins the halves each warp of the pair transforms ONE half, so the
// two halves run side by side instead of one after the other: 16 resident warps per SM on the same 16 KiB of staging per
// polynomial (the one-warp form tops out at 12), and no second half waiting in shared memory.  The pair meets at a
// named barrier (bar.sync 1 + pair, 64) around the two places where data crosses the halves: the split level (each
// warp computes the butterflies of 16 of the 32 register rows and leaves both outputs in the buffer) and the join at
// the end (likewise).  One lane of the pair issues the bulk copies; both warps wait on the mbarriers.
#ifndef QT_PAIR_P
#define QT_PAIR_P 8  // pairs per CTA
#endif
struct PairShape {
    using T = Tile<SET_P_III_H>;
    static constexpr int PAIRS = QT_PAIR_P;
    static constexpr uint32_t HALF = T::N, WORDS = 2 * T::N;
    static constexpr size_t TW_BYTES = (size_t)T::TABLE_QUADS * sizeof(TwQuad);
    static constexpr size_t PAIR_BYTES = 2 * WORDS * sizeof(uint32_t) + 2 * sizeof(uint64_t);
    static constexpr size_t SMEM = TW_BYTES + (size_t)PAIRS * PAIR_BYTES;
};
__device__ __forceinline__ void pair_sync(uint32_t pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

__global__ void __launch_bounds__(PairShape::PAIRS * 64, 1)
k_polymul_pair(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = PairShape::T;
    using G = PairShape;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tw + T::TABLE_QUADS);
    const int NP = (int)(blockDim.x >> 6);  // pairs in this CTA (<= PairShape::PAIRS)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NP * 2 * G::WORDS);
    const uint32_t warp = warp_index<true>(), lane = threadIdx.x & 31, pair = warp >> 1, h = warp & 1;
    uint32_t* A = s_stage + pair * 2 * G::WORDS;
    uint32_t* B = A + G::WORDS;
    uint64_t* bar_a = s_bar + 2 * pair;
    uint64_t* bar_b = bar_a + 1;
    const size_t stride = (size_t)gridDim.x * NP;
    size_t tile = (size_t)pair * gridDim.x + blockIdx.x;  // SM-interleaved; one polynomial per tile
    const bool issuer = h == 0 && lane == 0;
    auto issue = [&](const uint32_t* g, uint32_t* st, uint64_t* bar, size_t t) {
        mbar_expect_tx(bar, G::WORDS * (uint32_t)sizeof(uint32_t));
        bulk_g2s(st, g + t * G::WORDS, G::WORDS * (uint32_t)sizeof(uint32_t), bar);
    };
    pdl_launch_dependents();
    if (issuer) {
        mbar_init(bar_a, 1);
        mbar_init(bar_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    copy_table_to_smem(s_tw, g_lane, T::TABLE_QUADS);
    pdl_wait();  // the operands may be the previous launch's output
    if (issuer && tile < batch) {
        issue(x, A, bar_a, tile);
        issue(y, B, bar_b, tile);
    }
    __syncthreads();

    constexpr uint32_t HR = T::E / 2;  // register rows of the split / join level this warp works on: [HR h, HR h + HR)
    uint32_t phase = 0;
    for (; tile < batch; tile += stride, phase ^= 1) {
        const bool more = tile + stride < batch;
        uint32_t v[T::E];
#pragma unroll 1
        for (int op = 0; op < 2; op++) {  // x, then y (which continues into the inverse)
            uint32_t* st = op ? B : A;
            mbar_wait(op ? bar_b : bar_a, phase);
            {
                uint32_t lo[HR], hi[HR];
#pragma unroll
                for (uint32_t k = 0; k < HR; k++) {
                    lo[k] = st[lane + 32 * (HR * h + k)];
                    hi[k] = st[G::HALF + lane + 32 * (HR * h + k)];
                }
                pair_sync(pair);  // every staging word has been read: the buffer becomes scratch
#pragma unroll
                for (uint32_t k = 0; k < HR; k++) {
                    T::ct(lo[k], hi[k], uni_tw<SET_P_III_H, UNI_FWD>(0));
                    const uint32_t o = T::swz(T::row_off(lane, HR * h + k));
                    st[o] = lo[k];
                    st[G::HALF + o] = hi[k];
                }
            }
            pair_sync(pair);
            uint32_t* sh = st + h * G::HALF;  // from here on each warp is alone with its half
            T::lds_rows(v, sh, lane);
            T::fwd_rows(v, 32 * h);
            T::sts_rows(v, sh, lane);
            __syncwarp();
            T::lds_cols(v, sh, lane);
            T::fwd_cols(v, s_tw + h * T::TW_QUADS + lane);
            if (op == 0) {
                __syncwarp();              // every lane has read its columns
                T::sts_cols(v, sh, lane);  // NTT(x) half h; each lane reads back only what it wrote
            }
        }
        T::pointwise_mont_stash(v, A + h * G::HALF, lane);
        fence_proxy_async();
        pair_sync(pair);  // both halves of NTT(x) are consumed: A is free, fetch the next x
        if (more && issuer) issue(x, A, bar_a, tile + stride);
        uint32_t* sh = B + h * G::HALF;
        T::inv_cols(v, s_tw + (1 - h) * T::TW_QUADS + (T::BLOCKS - 1 - lane));  // mirrored table of the OTHER half
        __syncwarp();
        T::sts_cols(v, sh, lane);
        __syncwarp();
        T::lds_rows(v, sh, lane);
        T::template inv_rows<UNI_INV_FUSED>(v, typename T::LanePtrs{nullptr, nullptr, nullptr}, 32 * h);
        T::sts_rows(v, sh, lane);  // this lane's own slots; the partner warp reads them after the barrier
        pair_sync(pair);
        // join the halves: last inverse level + scale, canonical; this warp's 16 register rows of BOTH halves
        uint32_t* zt = z + tile * G::WORDS;
#pragma unroll
        for (uint32_t k = 0; k < HR; k++) {
            const uint32_t r = HR * h + k, o = T::swz(T::row_off(lane, r));
            uint32_t a = B[o], b = B[G::HALF + o];
            T::template split_inv<UNI_INV_FUSED>(a, b);
            zt[lane + 32 * r] = a;
            zt[G::HALF + lane + 32 * r] = b;
        }
        fence_proxy_async();
        pair_sync(pair);  // B is free
        if (more && issuer) issue(y, B, bar_b, tile + stride);
    }
}

// Single transform for n = 2048 on the split tile (qt_ntt_forward / qt_ntt_inverse of qTESLA-p-III), in
// place, NTT domain in bit-reversed order.  Two one-polynomial buffers per warp alternate between staging
// of the next tile and scratch of the current one, as in k_ntt_tma.
template <bool INVERSE>
__global__ void __launch_bounds__(SplitShape::WARPS * 32, 1)
k_ntt_split(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = SplitShape::T;
    using G = SplitShape;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tw + T::TABLE_QUADS);
    const int NW = (int)(blockDim.x >> 5);  // <= SplitShape::WARPS
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NW * 2 * G::WORDS);
    const uint32_t warp = warp_index<false>(), lane = threadIdx.x & 31;
    uint32_t* buf0 = s_stage + warp * 2 * G::WORDS;
    uint64_t* bar0 = s_bar + 2 * warp;
    const size_t stride = (size_t)gridDim.x * NW;
    size_t tile = (size_t)warp * gridDim.x + blockIdx.x;
    auto issue = [&](uint32_t* st, uint64_t* bar, size_t t) {
        mbar_expect_tx(bar, G::WORDS * (uint32_t)sizeof(uint32_t));
        bulk_g2s(st, a + t * G::WORDS, G::WORDS * (uint32_t)sizeof(uint32_t), bar);
    };
    pdl_launch_dependents();
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    copy_table_to_smem(s_tw, g_lane, T::TABLE_QUADS);
    pdl_wait();
    if (lane == 0 && tile < batch) issue(buf0, bar0, tile);
    __syncthreads();
    for (uint32_t k = 0; tile < batch; tile += stride, k++) {
        uint32_t* st = buf0 + (k & 1) * G::WORDS;
        // the other buffer was released (fence + __syncwarp) at the end of the previous tile
        if (tile + stride < batch && lane == 0) issue(buf0 + ((k & 1) ^ 1) * G::WORDS, bar0 + ((k & 1) ^ 1), tile + stride);
        mbar_wait(bar0 + (k & 1), (k >> 1) & 1);
        uint32_t* gt = a + tile * G::WORDS;
        uint32_t v[T::E];
        if (!INVERSE) {
            {
                uint32_t hi[T::E];
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) {
                    v[r] = st[lane + 32 * r];
                    hi[r] = st[G::HALF + lane + 32 * r];
                }
                __syncwarp();
                T::split_fwd(v, hi);
                T::sts_rows(hi, st + G::HALF, lane);
            }
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                uint32_t* sh = st + h * G::HALF;
                if (h) T::lds_rows(v, sh, lane);
                T::fwd_rows(v, 32 * h);
                T::sts_rows(v, sh, lane);
                __syncwarp();
                T::lds_cols(v, sh, lane);
                T::fwd_cols(v, s_tw + h * T::TW_QUADS + lane);
                T::canon_fwd(v);
                T::sts_cols(v, sh, lane);  // re-layout only, so that the global store is coalesced
                __syncwarp();
                T::lds_rows(v, sh, lane);
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) gt[h * G::HALF + lane + 32 * r] = v[r];
            }
        } else {
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                uint32_t* sh = st + h * G::HALF;
#pragma unroll
                for (uint32_t r = 0; r < T::E; r++) v[r] = sh[lane + 32 * r];
                __syncwarp();              // the unswizzled staging words of this half are all read
                T::sts_rows(v, sh, lane);  // NTT-domain data is consumed in the cols layout
                __syncwarp();
                T::lds_cols(v, sh, lane);
                T::inv_cols(v, s_tw + (1 - h) * T::TW_QUADS + (T::BLOCKS - 1 - lane));
                T::sts_cols(v, sh, lane);
                __syncwarp();
                T::lds_rows(v, sh, lane);
                T::template inv_rows<UNI_INV_PLAIN>(v, typename T::LanePtrs{nullptr, nullptr, nullptr}, 32 * h);
                if (h == 0) T::sts_rows(v, sh, lane);  // parked in this lane's own slots
            }
#pragma unroll
            for (uint32_t r = 0; r < T::E; r++) {
                uint32_t lo = st[T::swz(T::row_off(lane, r))];
                T::template split_inv<UNI_INV_PLAIN>(lo, v[r]);
                gt[lane + 32 * r] = lo;
                gt[G::HALF + lane + 32 * r] = v[r];
            }
        }
        fence_proxy_async();
        __syncwarp();  // this buffer is free for the tile after next
    }
}

// z = a*y with NTT(a) supplied by the caller (qTESLA's own use: one public polynomial a, transformed
// once, multiplied by many secrets / sparse challenges).  Two transforms instead of three.
// a_hat is in the NTT domain exactly as qt_ntt_forward leaves it (canonical, bit-reversed order);
// BCAST: one a_hat for the whole batch, otherwise one per product.  y is staged by TMA like above.
template <int SET, bool BCAST>
__global__ void __launch_bounds__(TmaCfg<SET>::WARPS * 32, 1)
k_polymul_ntt(const uint32_t* __restrict__ a_hat, const uint32_t* y, uint32_t* z, size_t batch,
              const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    using G = StageShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    const int NW = (int)(blockDim.x >> 5);  // <= TmaCfg<SET>::WARPS; small batches are launched with fewer warps (less shared memory: the next launch's CTA fits beside this one)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NW * 2 * G::WORDS);
    const uint32_t warp = warp_index<true>(), lane = threadIdx.x & 31;
    // One y buffer per warp, refilled as soon as the inverse has left it (measured: requesting the next y a
    // whole tile ahead into a second buffer is not faster for these sets, and slower for n=512)
    uint32_t* B = s_stage + warp * 2 * G::WORDS;
    uint64_t* bar0 = s_bar + 2 * warp;
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    const size_t stride = (size_t)gridDim.x * NW;
    size_t tile = (size_t)warp * gridDim.x + blockIdx.x;  // SM-interleaved: a partial last round spreads over all SMs
    auto issue = [&](uint32_t* st, uint64_t* bar, size_t t) {
        const size_t p0 = t * T::PPW;
        const uint32_t np = (uint32_t)((batch - p0 < T::PPW) ? batch - p0 : T::PPW);
        mbar_expect_tx(bar, np * T::N * (uint32_t)sizeof(uint32_t));
        if (G::PAD == 0) {
            bulk_g2s(st, y + p0 * T::N, np * T::N * (uint32_t)sizeof(uint32_t), bar);
        } else {
            for (uint32_t p = 0; p < np; p++)
                bulk_g2s(st + p * G::POLY_STRIDE, y + (p0 + p) * T::N, T::N * (uint32_t)sizeof(uint32_t), bar);
        }
    };
    pdl_launch_dependents();
    if (lane == 0) {
        mbar_init(bar0, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    pdl_wait();  // y and a_hat may be the previous launch's output
    if (lane == 0 && tile < ntiles) issue(B, bar0, tile);
    // BCAST: the one a_hat lives in shared memory for the whole kernel (behind the barriers), stored with
    // the tile's swizzle so that the 128-bit reads below are conflict-free
    uint32_t* s_ahat = reinterpret_cast<uint32_t*>(s_bar + 2 * NW);
    if (BCAST) {
        for (uint32_t i = threadIdx.x; i < T::N / 4; i += blockDim.x)
            *reinterpret_cast<uint4*>(s_ahat + T::swz(4 * i)) = __ldg(reinterpret_cast<const uint4*>(a_hat) + i);
    }
    __syncthreads();
    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    for (uint32_t k = 0; tile < ntiles; tile += stride, k++) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        // this lane's E words of a_hat in the cols layout (clamped for the missing polynomial of a tail tile)
        const uint4* ah = reinterpret_cast<const uint4*>(a_hat + (valid ? base + (size_t)T::E * lane : 0));
        if (!BCAST) asm volatile("prefetch.global.L2 [%0];" ::"l"(ah));  // this lane's 128-byte line, needed after the forward transform
        uint32_t v[T::E];
        mbar_wait(bar0, k & 1);
#pragma unroll
        for (uint32_t r = 0; r < T::E; r++) v[r] = B[G::off(lane, r)];
        __syncwarp();
        T::fwd_rows(v);
        T::sts_rows(v, B, lane);
        __syncwarp();
        T::lds_cols(v, B, lane);
        T::fwd_cols(v, P.fwd);
#pragma unroll
        for (uint32_t c = 0; c < T::E / 4; c++) {
            const uint4 u = BCAST ? *reinterpret_cast<const uint4*>(s_ahat + T::swz(T::E * (lane % T::BLOCKS) + 4 * c))
                                  : __ldg(ah + c);
            const uint32_t b[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (uint32_t k2 = 0; k2 < 4; k2++)
                v[4 * c + k2] = T::pw_canonical(v[4 * c + k2], b[k2]);
        }
        T::inv_cols(v, P.inv);
        __syncwarp();
        T::sts_cols(v, B, lane);
        __syncwarp();
        T::lds_rows(v, B, lane);
        fence_proxy_async();
        __syncwarp();  // B is free: fetch the next tile's y into it
        if (tile + stride < ntiles && lane == 0) issue(B, bar0, tile + stride);
        T::template inv_rows<UNI_INV_FUSED>(v, P);
        T::store_rows(v, z + base, lane, valid);
    }
}

// forward NTT in place: natural -> NTT domain (bit-reversed, psi merged), canonical
template <int SET>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_ntt_forward(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    __syncthreads();
    const uint32_t warp = warp_index<false>(), lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    for (size_t tile = (size_t)blockIdx.x * WARPS_PER_CTA + warp; tile < ntiles;
         tile += (size_t)gridDim.x * WARPS_PER_CTA) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        uint32_t v[T::E];
        T::load_rows(v, a + base, lane, valid);
        T::fwd_rows(v);
        T::sts_rows(v, buf, lane);
        __syncwarp();
        T::lds_cols(v, buf, lane);
        T::fwd_cols(v, P.fwd);
        T::canon_fwd(v);
        __syncwarp();
        T::sts_cols(v, buf, lane);  // re-layout only, so that the global store is coalesced
        __syncwarp();
        T::lds_rows(v, buf, lane);
        __syncwarp();
        T::store_rows(v, a + base, lane, valid);
    }
}

// inverse NTT in place: NTT domain -> natural, n^-1 psi^-i included, canonical
template <int SET>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_ntt_inverse(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    __syncthreads();
    const uint32_t warp = warp_index<false>(), lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    for (size_t tile = (size_t)blockIdx.x * WARPS_PER_CTA + warp; tile < ntiles;
         tile += (size_t)gridDim.x * WARPS_PER_CTA) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        uint32_t v[T::E];
        T::load_rows(v, a + base, lane, valid);  // coalesced read, re-layout through smem
        T::sts_rows(v, buf, lane);
        __syncwarp();
        T::lds_cols(v, buf, lane);
        T::inv_cols(v, P.inv);
        __syncwarp();
        T::sts_cols(v, buf, lane);
        __syncwarp();
        T::lds_rows(v, buf, lane);
        __syncwarp();
        T::template inv_rows<UNI_INV_PLAIN>(v, P);
        T::store_rows(v, a + base, lane, valid);
    }
}

// Natural-order NTT domain (what the reference's Stockham pipeline produces, NTT.cu:2040-2049):
// forward = natural -> natural, inverse = natural -> natural.  Same arithmetic as above; the
// bit-reversal is folded into the global access of the cols layout (Tile::nat_off), which saves one
// shared-memory transposition per transform instead of costing a separate permutation pass.
template <int SET, bool INVERSE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_ntt_natural(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET, !INVERSE>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    __syncthreads();
    const uint32_t warp = warp_index<false>(), lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    for (size_t tile = (size_t)blockIdx.x * WARPS_PER_CTA + warp; tile < ntiles;
         tile += (size_t)gridDim.x * WARPS_PER_CTA) {
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        uint32_t v[T::E];
        if (!INVERSE) {
            T::load_rows(v, a + base, lane, valid);
            T::fwd_rows(v);
            T::sts_rows(v, buf, lane);
            __syncwarp();  // also orders every lane's loads before the in-place stores below
            T::lds_cols(v, buf, lane);
            T::fwd_cols(v, P.fwd);
            T::canon_fwd(v);
            T::store_cols_natural(v, a + base, lane, valid);
        } else {
            T::load_cols_natural(v, a + base, lane, valid);
            T::inv_cols(v, P.inv);
            T::sts_cols(v, buf, lane);
            __syncwarp();
            T::lds_rows(v, buf, lane);
            T::template inv_rows<UNI_INV_PLAIN>(v, P);
            T::store_rows(v, a + base, lane, valid);
        }
        __syncwarp();  // the buffer is rewritten by the next tile
    }
}

// TMA-staged single transform (forward or inverse), in place.  Two one-tile buffers per warp alternate
// between "staging of the next tile" (bulk copy in flight) and "current tile: staging, then transposition
// scratch".  Same arithmetic as k_ntt_forward / k_ntt_inverse, which remain the fallback for operands
// that are not 16-byte aligned.
template <int SET, bool INVERSE>
__global__ void __launch_bounds__(TmaCfg<SET>::WARPS * 32, 1)
k_ntt_tma(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    using G = StageShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tw + S::TW_QUADS);
    const int NW = (int)(blockDim.x >> 5);  // <= TmaCfg<SET>::WARPS; small batches are launched with fewer warps (less shared memory: the next launch's CTA fits beside this one)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + NW * 2 * G::WORDS);
    const uint32_t warp = warp_index<true>(), lane = threadIdx.x & 31;
    uint32_t* buf0 = s_stage + warp * 2 * G::WORDS;
    uint64_t* bar0 = s_bar + 2 * warp;
    const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
    const size_t stride = (size_t)gridDim.x * NW;
    size_t tile = (size_t)warp * gridDim.x + blockIdx.x;  // SM-interleaved: a partial last round spreads over all SMs
    auto issue = [&](uint32_t* st, uint64_t* bar, size_t t) {
        const size_t p0 = t * T::PPW;
        const uint32_t np = (uint32_t)((batch - p0 < T::PPW) ? batch - p0 : T::PPW);
        mbar_expect_tx(bar, np * T::N * (uint32_t)sizeof(uint32_t));
        if (G::PAD == 0) {
            bulk_g2s(st, a + p0 * T::N, np * T::N * (uint32_t)sizeof(uint32_t), bar);
        } else {
            for (uint32_t p = 0; p < np; p++)
                bulk_g2s(st + p * G::POLY_STRIDE, a + (p0 + p) * T::N, T::N * (uint32_t)sizeof(uint32_t), bar);
        }
    };
    pdl_launch_dependents();
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    copy_table_to_smem(s_tw, g_lane, S::TW_QUADS);
    pdl_wait();
    if (lane == 0 && tile < ntiles) issue(buf0, bar0, tile);
    __syncthreads();
    const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
    for (uint32_t k = 0; tile < ntiles; tile += stride, k++) {
        uint32_t* cur = buf0 + (k & 1) * G::WORDS;
        uint32_t* nxt = buf0 + ((k & 1) ^ 1) * G::WORDS;
        const size_t base = tile * T::C::TILE_WORDS;
        const bool valid = tile * T::PPW + lane / T::LPP < batch;
        // the other buffer was last read (as scratch) in the previous iteration, which ended with a __syncwarp
        if (tile + stride < ntiles && lane == 0) {
            fence_proxy_async();
            issue(nxt, bar0 + ((k & 1) ^ 1), tile + stride);
        }
        mbar_wait(bar0 + (k & 1), (k >> 1) & 1);
        uint32_t v[T::E];
#pragma unroll
        for (uint32_t r = 0; r < T::E; r++) v[r] = cur[G::off(lane, r)];
        __syncwarp();
        if (!INVERSE) {
            T::fwd_rows(v);
            T::sts_rows(v, cur, lane);
            __syncwarp();
            T::lds_cols(v, cur, lane);
            T::fwd_cols(v, P.fwd);
            T::canon_fwd(v);
        } else {
            T::sts_rows(v, cur, lane);  // re-layout: NTT-domain data is consumed in the cols layout
            __syncwarp();
            T::lds_cols(v, cur, lane);
            T::inv_cols(v, P.inv);
        }
        __syncwarp();
        T::sts_cols(v, cur, lane);
        __syncwarp();
        T::lds_rows(v, cur, lane);
        fence_proxy_async();
        __syncwarp();
        if (INVERSE) T::template inv_rows<UNI_INV_PLAIN>(v, P);
        T::store_rows(v, a + base, lane, valid);
    }
}

// c = a*b mod q element-wise (HBM-bound: 12 bytes per coefficient)
template <int SET>
__global__ void __launch_bounds__(256)
k_pointwise(const uint32_t* a, const uint32_t* b, uint32_t* c, size_t words) {
    using T = Tile<SET>;
    const TwPair r2{T::C::R_MODQ, (uint32_t)(((uint64_t)T::C::R_MODQ << 32) / T::Q)};
    const size_t quads = words / 4;  // n is a multiple of 4
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint4 ua = reinterpret_cast<const uint4*>(a)[i];
        const uint4 ub = reinterpret_cast<const uint4*>(b)[i];
        uint4 uc;
        // mont gives a*b*R^-1 in [0,2q); multiplying by R (Shoup) restores a*b
        uc.x = T::csub(T::mul_shoup(T::mul_mont(ua.x, ub.x), r2), T::Q);
        uc.y = T::csub(T::mul_shoup(T::mul_mont(ua.y, ub.y), r2), T::Q);
        uc.z = T::csub(T::mul_shoup(T::mul_mont(ua.z, ub.z), r2), T::Q);
        uc.w = T::csub(T::mul_shoup(T::mul_mont(ua.w, ub.w), r2), T::Q);
        reinterpret_cast<uint4*>(c)[i] = uc;
    }
}

// same for operands that are only 4-byte aligned (a view into a larger array): one word per thread, still coalesced
template <int SET>
__global__ void __launch_bounds__(256)
k_pointwise_scalar(const uint32_t* a, const uint32_t* b, uint32_t* c, size_t words) {
    using T = Tile<SET>;
    const TwPair r2{T::C::R_MODQ, (uint32_t)(((uint64_t)T::C::R_MODQ << 32) / T::Q)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x)
        c[i] = T::csub(T::mul_shoup(T::mul_mont(a[i], b[i]), r2), T::Q);
}

// out[b*n + j] = in[b*n + brv(j)]   (bit_reverse_copy_tbl_gpu, NTT.cu:487-492)
// One warp per polynomial.  Lane L reads the coefficients L + 32 r (one 128-byte line per warp
// instruction); their bit-reversed positions brv(L + 32 r) = R*brv5(L) + brv_logR(r), R = n/32, form ONE
// contiguous run of R words, so the permutation inside the run is a compile-time register renaming.  The
// runs are turned back into coalesced rows through a padded shared-memory tile (row stride R+1), so both
// the global read and the global write are full lines — no scattered gather (the reference's
// table-driven gather uses 4 bytes of every 32-byte sector it touches).
template <int SET> struct BitrevShape {
    static constexpr uint32_t R = Cfg<SET>::N / 32, LOGR = Cfg<SET>::LOGN - 5, WARPS = 8;
    static constexpr uint32_t WARP_WORDS = 32 * (R + 1);
    static constexpr size_t SMEM = (size_t)WARPS * WARP_WORDS * sizeof(uint32_t);
};
template <int SET>
__global__ void __launch_bounds__(256)
k_bitrev_copy(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t batch) {
    using C = Cfg<SET>;
    using S = BitrevShape<SET>;
    constexpr uint32_t R = S::R, LOGR = S::LOGR;
    extern __shared__ uint4 smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = warp_index<false>();
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw) + warp * S::WARP_WORDS;
    const uint32_t rl = __brev(lane) >> 27;  // brv5(lane): the output run this lane's inputs belong to
    const size_t warps = (size_t)gridDim.x * S::WARPS;
    for (size_t p = (size_t)blockIdx.x * S::WARPS + warp; p < batch; p += warps) {
        const uint32_t* src = in + p * C::N + lane;
        uint32_t v[R];
#pragma unroll
        for (uint32_t r = 0; r < R; r++) v[r] = src[32 * r];
#pragma unroll
        for (uint32_t k = 0; k < R; k++) buf[rl * (R + 1) + k] = v[c_bitrev(k, LOGR)];  // out[rl*R + k]
        __syncwarp();
        uint32_t* dst = out + p * C::N + lane;
#pragma unroll
        for (uint32_t r = 0; r < R; r++) {
            const uint32_t o = lane + 32 * r;                                            // coalesced output index
            dst[32 * r] = buf[(o / R) * (R + 1) + (o % R)];
        }
        __syncwarp();
    }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// synthetic operands: a[i] = splitmix64(seed + first + i) % q
__global__ void __launch_bounds__(256)
k_fill_uniform(uint32_t* a, size_t count, uint64_t seed, uint64_t first, uint32_t q) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (size_t)gridDim.x * blockDim.x)
        a[i] = (uint32_t)(splitmix64(seed + first + i) % q);
}

}  // namespace qt
