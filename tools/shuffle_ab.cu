// shuffle_ab.cu — A/B of the two ways to run the five lane-crossing butterfly levels of a 1024-point transform held
// by ONE WARP (32 registers per lane, register r of lane j = coefficient j + 32 r):
//   A  "transpose": 32 STS (swizzled) + 8 LDS.128 turn the lane-crossing levels into thread-local ones, then 80
//      butterflies per thread (what qt_tile.cuh does);
//   B  "shuffle":   every level exchanges each register with the partner lane (shfl.bfly), the lane holding x forms
//      x + w y, its partner x - w y — both must multiply (SIMT), so the level costs one Shoup product per ELEMENT
//      instead of one per butterfly (the north-star's "last log2(32) stages by warp shuffle").
// Same signed-lazy butterfly arithmetic in both (3 multiply-pipe instructions + adds), 16 warps per SM, no global
// traffic inside the timed loop.  Prints clocks per warp for the five levels and the ratio.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/shuffle_ab tools/shuffle_ab.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

constexpr uint32_t Q = 8404993u;
struct alignas(16) U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ void ct(uint32_t& x, uint32_t& y, uint32_t w, uint32_t ws) {
    const uint32_t hi = (uint32_t)__mulhi((int)y, (int)ws), u = y * w + x, xn = u - hi * Q;
    y = x + x - xn;
    x = xn;
}
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 5) & 7u) << 2); }

template <int MODE> __global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, const uint32_t* tw, int iters) {
    extern __shared__ uint4 smem_raw[];
    uint32_t (*buf)[1024] = reinterpret_cast<uint32_t (*)[1024]>(smem_raw);
    uint32_t* stw = reinterpret_cast<uint32_t*>(smem_raw) + 16 * 1024;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) stw[threadIdx.x] = tw[threadIdx.x];
    __syncthreads();
    uint32_t v[32];
#pragma unroll
    for (int r = 0; r < 32; r++) v[r] = (lane * 977u + r * 131u + warp) % Q;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (uint32_t r = 0; r < 32; r++) buf[warp][swz(lane + 32 * r)] = v[r];
            __syncwarp();
#pragma unroll
            for (uint32_t c = 0; c < 8; c++) {
                const U4 u = *reinterpret_cast<const U4*>(&buf[warp][swz(32 * lane + 4 * c)]);
                v[4 * c] = u.x; v[4 * c + 1] = u.y; v[4 * c + 2] = u.z; v[4 * c + 3] = u.w;
            }
            __syncwarp();
#pragma unroll
            for (uint32_t k = 0; k < 5; k++) {
                const uint32_t half = 16u >> k;
#pragma unroll
                for (uint32_t i = 0; i < 16; i++) {
                    const uint32_t g = i / half, j = i % half;
                    ct(v[2 * g * half + j], v[2 * g * half + j + half], stw[2 * (g + k) & 63], stw[(2 * (g + k) + 1) & 63]);
                }
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 5; k++) {
                const uint32_t d = 16u >> k;
                const bool hi = (lane & d) != 0;
                const uint32_t w = stw[(2 * k) & 63], ws = stw[(2 * k + 1) & 63];
#pragma unroll
                for (uint32_t r = 0; r < 32; r++) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], d);
                    uint32_t x = hi ? o : v[r], y = hi ? v[r] : o;  // the pair as (x, y); both lanes multiply
                    ct(x, y, w, ws);
                    v[r] = hi ? y : x;
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < 32; r++) s += v[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> double run(int sms, uint32_t* out, long long* cyc, const uint32_t* tw) {
    const int iters = 2000;
    const size_t smem = (16 * 1024 + 64) * sizeof(uint32_t);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<sms, 512, smem>>>(out, cyc, tw, 10);
    cudaDeviceSynchronize();
    k<MODE><<<sms, 512, smem>>>(out, cyc, tw, iters);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    return (double)mx / iters / 16.0;  // SM clocks per warp-transform-part (16 warps share the SM)
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    uint32_t *out, *tw;
    long long* cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 512 * 4);
    cudaMalloc(&cyc, (size_t)p.multiProcessorCount * 8);
    cudaMalloc(&tw, 64 * 4);
    std::vector<uint32_t> h(64);
    for (int i = 0; i < 64; i++) h[i] = (uint32_t)(i * 2654435761u) % Q;
    cudaMemcpy(tw, h.data(), 64 * 4, cudaMemcpyHostToDevice);
    const double a = run<0>(p.multiProcessorCount, out, cyc, tw), b = run<1>(p.multiProcessorCount, out, cyc, tw);
    printf("{\"device\": \"%s\", \"what\": \"five lane-crossing levels of a 1024-point transform per warp, 16 warps per SM\", "
           "\"transpose_then_local_clk_per_warp\": %.1f, \"shuffle_levels_clk_per_warp\": %.1f, \"shuffle_over_transpose\": %.2f}\n",
           p.name, a, b, b / a);
    return 0;
}
