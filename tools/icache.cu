// icache.cu — does straight-line code size limit the butterfly rate?  REPEAT unrolled Harvey levels over
// 32 registers inside a rolled loop; the same work per thread for every REPEAT, only the code footprint
// changes.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/icache tools/icache.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ uint2 c_tw[4096];
constexpr uint32_t Q = 856145921u;

__device__ __forceinline__ void bf(uint32_t& x, uint32_t& y, uint2 w) {
    uint32_t xr = min(x, x - 2 * Q);
    uint32_t hi = __umulhi(y, w.y);
    uint32_t t = y * w.x - hi * Q;
    x = xr + t;
    y = xr - t + 2 * Q;
}

template <int REPEAT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(uint32_t* out, int total_levels, long long* cyc) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = threadIdx.x * 33u + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < total_levels / REPEAT; it++) {
#pragma unroll
        for (int l = 0; l < REPEAT; l++) {
            const int half = 1 << (l % 5);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int g = i / half, j = i % half;
                bf(v[2 * g * half + j], v[2 * g * half + j + half], c_tw[(l * 16 + g) & 4095]);
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int REPEAT, int WARPS> void run(uint32_t* out, long long* cyc, int sms) {
    const int total = 3840;  // levels per thread, divisible by every REPEAT below
    k<REPEAT, WARPS><<<sms, WARPS * 32>>>(out, total, cyc);
    cudaDeviceSynchronize();
    k<REPEAT, WARPS><<<sms, WARPS * 32>>>(out, total, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k<REPEAT, WARPS>);
    // per SMSP: WARPS/4 warps, each total*16 butterflies
    double clk = (double)c / ((double)total * 16 * WARPS / 4);
    printf("  {\"unrolled_levels\": %d, \"warps_per_sm\": %d, \"approx_code_kb\": %.0f, \"regs\": %d, \"clk_per_warp_butterfly_per_smsp\": %.3f},\n",
           REPEAT, WARPS, (REPEAT * 108 + 100) * 16 / 1024.0, a.numRegs, clk);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint2* h = new uint2[4096];
    for (int i = 0; i < 4096; i++) { uint32_t w = (i * 2654435761u) % Q; h[i] = make_uint2(w, (uint32_t)(((uint64_t)w << 32) / Q)); }
    cudaMemcpyToSymbol(c_tw, h, sizeof(uint2) * 4096);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&cyc, 8);
    printf("[\n");
    run<4, 12>(out, cyc, sms); run<8, 12>(out, cyc, sms); run<16, 12>(out, cyc, sms); run<24, 12>(out, cyc, sms);
    run<32, 12>(out, cyc, sms); run<48, 12>(out, cyc, sms); run<64, 12>(out, cyc, sms); run<96, 12>(out, cyc, sms);
    run<4, 16>(out, cyc, sms); run<16, 16>(out, cyc, sms); run<32, 16>(out, cyc, sms); run<48, 16>(out, cyc, sms); run<64, 16>(out, cyc, sms);
    printf("  {}\n]\n");
    return 0;
}
