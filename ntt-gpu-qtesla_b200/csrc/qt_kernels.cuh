// qt_kernels.cuh — __global__ kernels built from the warp-tile engine (qt_tile.cuh).
//
// All kernels are persistent: grid = (#SMs x resident CTAs), every warp walks tiles
// tile = blockIdx.x*WARPS + warp, += gridDim.x*WARPS.  The per-lane twiddle tables are copied
// from global to shared memory once per CTA.
#pragma once
#include <cuda_runtime.h>

#include "qt_tile.cuh"

namespace qt {

constexpr int WARPS_PER_CTA = 8;

template <int SET> struct KernelShape {
    using T = Tile<SET>;
    static constexpr size_t TW_QUADS = (size_t)T::SLOT_PAIRS * T::BLOCKS;  // per direction
    static constexpr size_t TW_BYTES = TW_QUADS * sizeof(TwQuad);
    static constexpr size_t BUF_BYTES = (size_t)WARPS_PER_CTA * T::C::TILE_WORDS * sizeof(uint32_t);
    static constexpr size_t SMEM_FUSED = 2 * TW_BYTES + BUF_BYTES;
    static constexpr size_t SMEM_ONE = TW_BYTES + BUF_BYTES;
};

__device__ __forceinline__ void copy_table_to_smem(TwQuad* dst, const TwQuad* __restrict__ src, size_t quads) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (size_t i = threadIdx.x; i < quads; i += blockDim.x) d[i] = __ldg(s + i);
}

// z = x*y mod (X^n+1, q): forward(x), forward(y), pointwise, inverse — one launch, HBM touched
// once per operand (12n bytes per product).
template <int SET>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_polymul(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch,
          const TwQuad* __restrict__ g_lane_fwd, const TwQuad* __restrict__ g_lane_inv) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_fwd = reinterpret_cast<TwQuad*>(smem_raw);
    TwQuad* s_inv = s_fwd + S::TW_QUADS;
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_inv + S::TW_QUADS);
    copy_table_to_smem(s_fwd, g_lane_fwd, S::TW_QUADS);
    copy_table_to_smem(s_inv, g_lane_inv, S::TW_QUADS);
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const TwQuad* tw_f = s_fwd + (lane % T::BLOCKS);
    const TwQuad* tw_i = s_inv + (lane % T::BLOCKS);
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
        T::fwd_cols(vx, tw_f);
        __syncwarp();
        T::lds_cols(vy, buf, lane);
        T::fwd_cols(vy, tw_f);
        T::pointwise_mont(vy, vx);
        T::inv_cols(vy, tw_i);
        __syncwarp();
        T::sts_cols(vy, buf, lane);
        __syncwarp();
        T::lds_rows(vy, buf, lane);
        __syncwarp();
        T::template inv_rows<UNI_INV_FUSED>(vy);
        T::store_rows(vy, z + base, lane, valid);
    }
}

// forward NTT in place: natural -> NTT domain (bit-reversed, psi merged), canonical
template <int SET>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_ntt_forward(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane_fwd) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_fwd = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_fwd + S::TW_QUADS);
    copy_table_to_smem(s_fwd, g_lane_fwd, S::TW_QUADS);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const TwQuad* tw_f = s_fwd + (lane % T::BLOCKS);
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
        T::fwd_cols(v, tw_f);
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
k_ntt_inverse(uint32_t* a, size_t batch, const TwQuad* __restrict__ g_lane_inv) {
    using T = Tile<SET>;
    using S = KernelShape<SET>;
    extern __shared__ uint4 smem_raw[];
    TwQuad* s_inv = reinterpret_cast<TwQuad*>(smem_raw);
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(s_inv + S::TW_QUADS);
    copy_table_to_smem(s_inv, g_lane_inv, S::TW_QUADS);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* buf = s_buf + warp * T::C::TILE_WORDS;
    const TwQuad* tw_i = s_inv + (lane % T::BLOCKS);
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
        T::inv_cols(v, tw_i);
        __syncwarp();
        T::sts_cols(v, buf, lane);
        __syncwarp();
        T::lds_rows(v, buf, lane);
        __syncwarp();
        T::template inv_rows<UNI_INV_PLAIN>(v);
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

// out[b*n + j] = in[b*n + brv(j)]   (bit_reverse_copy_tbl_gpu, NTT.cu:487-492)
template <int SET>
__global__ void __launch_bounds__(256)
k_bitrev_copy(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t words) {
    using C = Cfg<SET>;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(i & (C::N - 1));
        const uint32_t rj = __brev(j) >> (32 - C::LOGN);
        out[i] = in[i - j + rj];
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
