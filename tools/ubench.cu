// ubench.cu — integer-pipe micro-benchmark for the roofline denominator (SURVEY.md 8d asks for the
// IMAD rate to be confirmed on the box).  Measures sustained warp-instruction throughput of the
// integer multiply forms a Shoup/Montgomery butterfly can be built from, and of the ALU ops beside
// them, as ops/clk/SM (clock64) and ops/s (CUDA events).  Prints one JSON object.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int CHAINS = 8, UNROLL = 16;

template <int OP> __device__ __forceinline__ void step(uint32_t (&r)[CHAINS], uint64_t (&w)[CHAINS], uint32_t b, uint32_t c) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
        if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
        if (OP == 2) {  // multiplicand depends on the accumulator so ptxas cannot hoist the product
            uint32_t lo = (uint32_t)w[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(b));
        }
        if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
        if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c));
        if (OP == 5) asm volatile("min.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
        if (OP == 6) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
        if (OP == 7) {  // Shoup butterfly: 1 mul.hi + 2 mad.lo + 2 add (the fused kernel's inner op)
            uint32_t hi, t;
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(r[i]), "r"(c));
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(r[i]), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(hi), "r"(0u - 8404993u));
            uint32_t x = (uint32_t)w[i];
            asm volatile("add.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(x), "r"(t));
            asm volatile("sub.u32 %0, %1, %2;" : "=r"(x) : "r"(x), "r"(t));
            w[i] = x;
        }
        if (OP == 8) {  // 1 mad.lo + 1 add interleaved: can both pipes issue every cycle?
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
            uint32_t x = (uint32_t)w[i];
            asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
            w[i] = x;
        }
        if (OP == 9) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(r[i]));
        if (OP == 10) {  // fp64 fma on the accumulator pair
            double d = __longlong_as_double((long long)w[i]);
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5));
            w[i] = (uint64_t)__double_as_longlong(d);
        }
        if (OP == 11) {  // fp32 fma
            float f = __uint_as_float(r[i]);
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
            r[i] = __float_as_uint(f);
        }
        if (OP == 12) {  // fp64 fma + integer mad.lo interleaved: do the two pipes overlap?
            double d = __longlong_as_double((long long)w[i]);
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5));
            w[i] = (uint64_t)__double_as_longlong(d);
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
        }
        if (OP == 14 || OP == 15 || OP == 16) {  // fp64 Shoup-style butterfly: 5 DP ops, exact for |y|<2^30
            double yv = __longlong_as_double((long long)w[i]);   // operand y (also plays x)
            const double W = 2083362.0, WQ = 2083362.0 / 8404993.0, Qd = 8404993.0;
            const double MAGIC = 6755399441055744.0, QM = 8404993.0 * 6755399441055744.0;
            double hm, u, t, xp, yp;
            asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(hm) : "d"(yv), "d"(WQ), "d"(MAGIC));
            asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(u) : "d"(-Qd), "d"(hm), "d"(QM));
            asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(t) : "d"(yv), "d"(W), "d"(u));
            asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(xp) : "d"(yv), "d"(t));
            asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(yp) : "d"(xp), "d"(t));
            w[i] = (uint64_t)__double_as_longlong(yp);
        }
        if (OP == 15 || OP == 16) {  // + integer Shoup butterflies in the same thread (other polynomial)
            for (int rep = 0; rep < (OP == 16 ? 2 : 1); rep++) {
                uint32_t hi, t;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(r[i]), "r"(c));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(r[i]), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(hi), "r"(0u - 8404993u));
                uint32_t x = r[i] ^ 0x5555u;
                asm volatile("add.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(x), "r"(t));
                asm volatile("sub.u32 %0, %1, %2;" : "=r"(x) : "r"(r[i]), "r"(t));
                r[i] ^= x;
            }
        }
        if (OP == 13) {  // mul.hi + fp32 fma interleaved: does FFMA find room beside IMAD.HI?
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            float f = __uint_as_float((uint32_t)w[i]);
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
            w[i] = __float_as_uint(f);
        }
    }
}

template <int OP> __global__ void __launch_bounds__(256) k(uint32_t* out, long long* cyc, uint32_t b, uint32_t c, int iters) {
    uint32_t r[CHAINS];
    uint64_t w[CHAINS];
    for (int i = 0; i < CHAINS; i++) { r[i] = threadIdx.x * 7 + i + b; w[i] = r[i] + c; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) step<OP>(r, w, b, c);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < CHAINS; i++) s += r[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int instr_per_step, int sms, uint32_t* out, long long* cyc, bool last) {
    const int ctas_per_sm = 4, block = 256, iters = 2000;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, block>>>(out, cyc, 12345u, 67891u, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<grid, block>>>(out, cyc, 12345u, 67891u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const double thread_instr = (double)iters * UNROLL * CHAINS * instr_per_step;
    const double total = thread_instr * grid * block;
    const double per_clk_sm = thread_instr * ctas_per_sm * block / (double)mx;  // lane-ops / clk / SM
    printf("  \"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"tera_lane_ops_per_s\": %.3f, \"ms\": %.3f, \"sm_mhz_effective\": %.0f}%s\n",
           name, per_clk_sm, total / (ms * 1e-3) / 1e12, ms, mx / (ms * 1e-3) / 1e6, last ? "" : ",");
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 256 * sizeof(uint32_t));
    cudaMalloc(&cyc, (size_t)p.multiProcessorCount * 4 * sizeof(long long));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("mad_lo_u32", 1, p.multiProcessorCount, out, cyc, false);
    run<1>("mul_hi_u32", 1, p.multiProcessorCount, out, cyc, false);
    run<6>("mad_hi_u32", 1, p.multiProcessorCount, out, cyc, false);
    run<2>("mad_wide_u32", 1, p.multiProcessorCount, out, cyc, false);
    run<3>("add_u32", 1, p.multiProcessorCount, out, cyc, false);
    run<4>("lop3", 1, p.multiProcessorCount, out, cyc, false);
    run<5>("min_u32", 1, p.multiProcessorCount, out, cyc, false);
    run<9>("shfl_bfly", 1, p.multiProcessorCount, out, cyc, false);
    run<10>("fma_f64", 1, p.multiProcessorCount, out, cyc, false);
    run<11>("fma_f32", 1, p.multiProcessorCount, out, cyc, false);
    run<12>("fma_f64_plus_mad_lo (2 instr)", 2, p.multiProcessorCount, out, cyc, false);
    run<13>("mul_hi_plus_fma_f32 (2 instr)", 2, p.multiProcessorCount, out, cyc, false);
    run<14>("fp64_butterfly (5 instr = 1 butterfly)", 5, p.multiProcessorCount, out, cyc, false);
    run<15>("fp64_butterfly + int_butterfly (10 instr = 2 butterflies)", 10, p.multiProcessorCount, out, cyc, false);
    run<16>("fp64_butterfly + 2 int_butterflies (15 instr = 3 butterflies)", 15, p.multiProcessorCount, out, cyc, false);
    run<8>("mad_lo_plus_add (2 instr)", 2, p.multiProcessorCount, out, cyc, false);
    run<7>("shoup_butterfly (5 instr, 3 mul)", 5, p.multiProcessorCount, out, cyc, true);
    printf("}\n");
    return 0;
}
