// synth.cu — how fast does the compiler-generated butterfly code of qt_tile.cuh run when nothing else
// is in the way?  Register-only loops over Tile<SET_III>::fwd_rows / inv_rows (uniform twiddles) and
// fwd_cols / inv_cols (twiddles from shared memory), 16 warps per SM like the fused kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/synth tools/synth.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../ntt-gpu-qtesla_b200/csrc/qt_tile.cuh"
namespace qt { TwPair h_uni[NUM_TILE_SETS][UNI_KINDS][UNI_MAX]; }
using namespace qt;
using T = Tile<SET_III>;

// issue-slot competition probe: the forward transform plus EXTRA independent ALU instructions per butterfly
template <int EXTRA> __device__ __forceinline__ void fwd_rows_padded(uint32_t (&v)[T::E], uint32_t (&d)[8]) {
#pragma unroll
    for (uint32_t l = 0; l < T::LB1; l++) {
        const uint32_t half = T::E >> (l + 1);
#pragma unroll
        for (uint32_t i = 0; i < T::E / 2; i++) {
            const uint32_t g = i / half, j = i % half;
            T::ct(v[2 * g * half + j], v[2 * g * half + j + half], uni_tw<SET_III, UNI_FWD>((1u << l) + g));
#pragma unroll
            for (int e = 0; e < EXTRA; e++)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(d[(i + e) & 7]) : "r"(d[(i + e + 3) & 7]));
        }
    }
}

// probe: signed Cooley-Tukey butterfly with the add folded into the multiply-add:
//   hi = mulhi.s32(y, ws); u = y*w + x; x' = u - hi*q; y' = 2x - x'      (3 multiply-pipe + 1 ALU instruction)
template <int EXTRA> __device__ __forceinline__ void fwd_rows_signed(uint32_t (&vu)[T::E], uint32_t (&d)[8]) {
    int (&v)[T::E] = reinterpret_cast<int (&)[T::E]>(vu);
    constexpr int Q = (int)T::Q;
#pragma unroll
    for (uint32_t l = 0; l < T::LB1; l++) {
        const uint32_t half = T::E >> (l + 1);
#pragma unroll
        for (uint32_t i = 0; i < T::E / 2; i++) {
            const uint32_t g = i / half, j = i % half;
            const TwPair t = uni_tw<SET_III, UNI_FWD>((1u << l) + g);
            int& x = v[2 * g * half + j];
            int& y = v[2 * g * half + j + half];
            const int hi = __mulhi(y, (int)t.ws);
            const int u = y * (int)t.w + x;
            const int xn = u - hi * Q;
            y = x + x - xn;
            x = xn;
#pragma unroll
            for (int e = 0; e < EXTRA; e++)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(d[(i + e) & 7]) : "r"(d[(i + e + 3) & 7]));
        }
    }
}

template <int MODE> __global__ void __launch_bounds__(768, 1) k(uint32_t* out, const TwQuad* g_tw, int iters, long long* cyc) {
    extern __shared__ uint4 sm[];
    TwQuad* s_tw = reinterpret_cast<TwQuad*>(sm);
    for (int i = threadIdx.x; i < (int)T::TABLE_QUADS; i += blockDim.x) s_tw[i] = g_tw[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t v[T::E];
    for (uint32_t r = 0; r < T::E; r++) v[r] = threadIdx.x * 33 + r;
    uint32_t dmy[8];
    for (int r = 0; r < 8; r++) dmy[r] = threadIdx.x + r;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) T::fwd_rows(v);
        if (MODE == 20) fwd_rows_signed<0>(v, dmy);
        if (MODE == 21) fwd_rows_signed<1>(v, dmy);
        if (MODE == 22) fwd_rows_signed<2>(v, dmy);
        if (MODE == 10) fwd_rows_padded<1>(v, dmy);
        if (MODE == 11) fwd_rows_padded<2>(v, dmy);
        if (MODE == 12) fwd_rows_padded<4>(v, dmy);
        const typename T::LanePtrs P = T::lane_ptrs(s_tw, lane);
        if (MODE == 1) T::fwd_cols(v, P.fwd);
        if (MODE == 2) { T::fwd_rows(v); T::fwd_cols(v, P.fwd); }
        if (MODE == 3) { T::inv_cols(v, P.inv); T::inv_rows<UNI_INV_FUSED>(v, P); }
        if (MODE == 4) { T::fwd_rows(v); T::fwd_cols(v, P.fwd); T::inv_cols(v, P.inv); T::inv_rows<UNI_INV_FUSED>(v, P); }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (uint32_t r = 0; r < T::E; r++) s += v[r];
    for (int r = 0; r < 8; r++) s += dmy[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static int g_threads = 512;
template <int MODE> void run(const char* name, double bf_per_iter, int sms, uint32_t* out, const TwQuad* tw, long long* cyc) {
    const int iters = 400, smem = T::TABLE_QUADS * sizeof(TwQuad);
    k<MODE><<<sms, g_threads, smem>>>(out, tw, 10, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<sms, g_threads, smem>>>(out, tw, iters, cyc);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
    // per SMSP: 4 warps, each does bf_per_iter warp-butterflies per iteration
    const double clk_per_warp_bf = (double)mx / (iters * bf_per_iter * (g_threads / 128.0));
    printf("  \"%s\": {\"clk_per_warp_butterfly_per_smsp\": %.3f, \"fraction_of_8clk_model\": %.3f},\n", name, clk_per_warp_bf, 8.0 / clk_per_warp_bf);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    HostTables tab; build_tables(SET_III, &tab);
    cudaMemcpyToSymbol(c_uni, tab.uni, sizeof(tab.uni), (size_t)SET_III * sizeof(tab.uni));
    TwQuad* tw; cudaMalloc(&tw, tab.block[1].size() * sizeof(TwQuad));
    cudaMemcpy(tw, tab.block[1].data(), tab.block[1].size() * sizeof(TwQuad), cudaMemcpyHostToDevice);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 768 * 4); cudaMalloc(&cyc, p.multiProcessorCount * 8);
    printf("{\n");
    for (int th : {128, 256, 384, 768}) {  // how many warps in butterfly code does the pipe need?
        g_threads = th;
        char nm[64];
        snprintf(nm, sizeof nm, "forward transform (160), %d warps per SM", th / 32);
        run<2>(nm, 160, p.multiProcessorCount, out, tw, cyc);
    }
    g_threads = 512;
    run<20>("signed fwd_rows, add folded into IMAD (3 FMA + 1 ALU)", 80, p.multiProcessorCount, out, tw, cyc);
    run<21>("signed fwd_rows + 1 extra ALU instr per butterfly", 80, p.multiProcessorCount, out, tw, cyc);
    run<22>("signed fwd_rows + 2 extra ALU instr per butterfly", 80, p.multiProcessorCount, out, tw, cyc);
    run<10>("fwd_rows + 1 extra ALU instr per butterfly (3 non-FMA per butterfly)", 80, p.multiProcessorCount, out, tw, cyc);
    run<11>("fwd_rows + 2 extra ALU instr per butterfly (4 non-FMA per butterfly)", 80, p.multiProcessorCount, out, tw, cyc);
    run<12>("fwd_rows + 4 extra ALU instr per butterfly (6 non-FMA per butterfly)", 80, p.multiProcessorCount, out, tw, cyc);
    run<0>("fwd_rows (80 butterflies, uniform twiddles)", 80, p.multiProcessorCount, out, tw, cyc);
    run<1>("fwd_cols (80 butterflies, smem twiddles)", 80, p.multiProcessorCount, out, tw, cyc);
    run<2>("forward transform (160)", 160, p.multiProcessorCount, out, tw, cyc);
    run<3>("inverse transform (160 + 16 + folds)", 160, p.multiProcessorCount, out, tw, cyc);
    run<4>("forward + inverse (320)", 320, p.multiProcessorCount, out, tw, cyc);
    printf("  \"note\": \"16 warps per SM, registers only\"\n}\n");
    return 0;
}
