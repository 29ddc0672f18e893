// dq_ubench.cu — micro-benchmark behind the "FP64 quotient" butterfly.
//
// A Shoup butterfly spends 8 of its multiply-pipe clocks per warp on three integer multiplies, 4 of them on the
// mul.hi that estimates the quotient floor(y w / q).  B200 has a full-rate FP64 pipe (64 lanes/clk/SM) that runs
// beside the integer multiply-add (profiles/ubench_r01s.json: fma_f64_plus_mad_lo 123.6 lanes/clk/SM), so the
// quotient can come from ONE DFMA instead:
//     t    = fma(D(y), w/q * 2^1000, 1.5 * 2^-22)        D(y) = the double whose BIT PATTERN is {lo = y, hi = 0},
//     qest = low word of t's bit pattern                   i.e. the denormal y * 2^-1074 — no int->double conversion
//     x'   = y*w + x - qest*q,  y' = 2x - x'               (two mad.lo + one add, all mod 2^32)
// For 0 <= y < 2^32 and |w/q| <= 1/2 the sum is exact before the single rounding, which lands on an integer multiple
// of 2^-74 = one unit of the low mantissa word, so qest = rint(y w / q) (+-1 from the rounding of w/q) and
// |y w - qest q| <= q/2 + eps.
// This tool measures (a) whether DFMA with a denormal operand runs at full rate, (b) the butterfly's throughput
// against the integer Shoup butterfly in the same 32-values-per-thread network the fused kernel uses, and
// (c) exactness on random operands.   Prints one JSON object.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dq_ubench tools/dq_ubench.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr uint32_t Q = 8404993u;
constexpr int NTW = 32;
__constant__ double c_W[NTW];      // (w centred) / q * 2^1000
__constant__ uint32_t c_w[NTW];    // w centred, two's complement
__constant__ uint32_t c_ws[NTW];   // signed Shoup companion floor(w 2^32 / q)

__device__ __forceinline__ void ct_dq(uint64_t& X, uint64_t& Y, double Ww, uint32_t w) {
    const double M = 3.5762786865234375e-07;  // 1.5 * 2^-22
    const uint32_t x = (uint32_t)X, y = (uint32_t)Y;
    const double t = fma(__longlong_as_double((long long)Y), Ww, M);
    const uint32_t qe = (uint32_t)__double2loint(t);
    const uint32_t u = y * w + x;
    const uint32_t xn = u - qe * Q;
    X = xn;
    Y = x + x - xn;
}
__device__ __forceinline__ void ct_shoup(uint32_t& x, uint32_t& y, uint32_t w, uint32_t ws) {
    const uint32_t hi = (uint32_t)__mulhi((int)y, (int)ws);
    const uint32_t u = y * w + x;
    const uint32_t xn = u - hi * Q;
    y = x + x - xn;
    x = xn;
}

// 32 values per thread, 5 levels x 16 butterflies, uniform twiddles: the rows pass of the fused kernel
template <int KIND> __global__ void __launch_bounds__(512) k_net(const uint32_t* in, uint32_t* out, long long* cyc, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (KIND == 0) {
        uint64_t v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = in[(tid + 977 * i) & 0xFFFF];
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int l = 0; l < 5; l++) {
                const int half = 16 >> l;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int g = i / half, j = i % half;
                    ct_dq(v[2 * g * half + j], v[2 * g * half + j + half], c_W[(1 << l) + g], c_w[(1 << l) + g]);
                }
            }
        }
        t1 = clock64();
#pragma unroll
        for (int i = 0; i < 32; i++) acc += (uint32_t)v[i];
    } else {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = in[(tid + 977 * i) & 0xFFFF];
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int l = 0; l < 5; l++) {
                const int half = 16 >> l;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int g = i / half, j = i % half;
                    ct_shoup(v[2 * g * half + j], v[2 * g * half + j + half], c_w[(1 << l) + g], c_ws[(1 << l) + g]);
                }
            }
        }
        t1 = clock64();
#pragma unroll
        for (int i = 0; i < 32; i++) acc += v[i];
    }
    out[tid] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// DFMA chains: operand denormal ({lo, 0}) or normal ({lo, 0x43300000})
template <int DEN> __global__ void __launch_bounds__(256) k_dfma(uint32_t* out, long long* cyc, int iters, double W) {
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 7 + i + 12345u;
    const double M = DEN ? 3.5762786865234375e-07 : 6755399441055744.0;
    const uint32_t hi = DEN ? 0u : 0x43300000u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double d = fma(__hiloint2double((int)hi, (int)r[i]), W, M);
                r[i] = (uint32_t)__double2loint(d);
            }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// exactness: one butterfly per (x, y, w) triple; flags: bit0 residue mismatch, bit1 |remainder| > q/2 + 1
__global__ void k_check(const uint32_t* xs, const uint32_t* ys, const int32_t* ws, unsigned long long* bad, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t w = ws[i];
    const double Ww = ldexp((double)w / (double)Q, 1000);
    uint64_t X = xs[i], Y = ys[i];
    const uint32_t x = xs[i], y = ys[i];
    ct_dq(X, Y, Ww, (uint32_t)w);
    const long long r = (long long)(int32_t)((uint32_t)X - x);          // y w - qest q
    const long long want = ((long long)y * w) % (long long)Q;             // same residue class
    long long d = (r - want) % (long long)Q;
    if (d != 0) atomicAdd(&bad[0], 1ull);
    if (r > (long long)Q / 2 + 1 || r < -(long long)Q / 2 - 1) atomicAdd(&bad[1], 1ull);
    if ((uint32_t)Y != x + x - (uint32_t)X) atomicAdd(&bad[2], 1ull);
}

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int sms = p.multiProcessorCount;
    uint64_t seed = 42;
    double hW[NTW]; uint32_t hw[NTW], hws[NTW];
    for (int i = 0; i < NTW; i++) {
        int64_t w = (int64_t)(sm64(seed) % Q); if (w > Q / 2) w -= Q;
        hw[i] = (uint32_t)(int32_t)w; hW[i] = ldexp((double)w / (double)Q, 1000);
        hws[i] = (uint32_t)(int32_t)floor(ldexp((double)w, 32) / (double)Q);
    }
    cudaMemcpyToSymbol(c_W, hW, sizeof hW); cudaMemcpyToSymbol(c_w, hw, sizeof hw); cudaMemcpyToSymbol(c_ws, hws, sizeof hws);
    uint32_t *in, *out; long long* cyc;
    std::vector<uint32_t> hin(65536);
    for (auto& v : hin) v = 256u * Q + (uint32_t)(sm64(seed) % Q);
    cudaMalloc(&in, 65536 * 4); cudaMalloc(&out, (size_t)sms * 4 * 512 * 4); cudaMalloc(&cyc, (size_t)sms * 4 * 8);
    cudaMemcpy(in, hin.data(), 65536 * 4, cudaMemcpyHostToDevice);
    printf("{\n  \"device\": \"%s\", \"sms\": %d,\n", p.name, sms);
    auto report = [&](const char* name, double per_thread, int block, int grid, float ms) {
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("  \"%s\": {\"per_clk_per_sm\": %.2f, \"clk_per_warp_op_per_smsp\": %.3f, \"ms\": %.3f},\n", name,
               per_thread * block * (grid / sms) / (double)mx, (double)mx / (per_thread * (block / 32) * (grid / sms) / 4.0), ms);
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int warps : {8, 12, 16}) {
        const int iters = 400, block = warps * 32, grid = sms;
        char name[96];
        k_net<0><<<grid, block>>>(in, out, cyc, 4); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_net<0><<<grid, block>>>(in, out, cyc, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        snprintf(name, sizeof name, "butterfly_fp64_quotient_%dwarps (butterflies)", warps); report(name, 80.0 * iters, block, grid, ms);
        k_net<1><<<grid, block>>>(in, out, cyc, 4); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_net<1><<<grid, block>>>(in, out, cyc, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        snprintf(name, sizeof name, "butterfly_shoup_%dwarps (butterflies)", warps); report(name, 80.0 * iters, block, grid, ms);
    }
    {
        const int iters = 2000, block = 256, grid = sms * 4;
        k_dfma<1><<<grid, block>>>(out, cyc, 4, hW[0]); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<1><<<grid, block>>>(out, cyc, iters, hW[0]); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        report("dfma_denormal_operand", 128.0 * iters, block, grid, ms);
        k_dfma<0><<<grid, block>>>(out, cyc, 4, 0.37); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<0><<<grid, block>>>(out, cyc, iters, 0.37); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        report("dfma_normal_operand", 128.0 * iters, block, grid, ms);
    }
    {   // exactness on 2^24 random triples: y over the whole 32-bit range, x in the offset window, w centred
        const int n = 1 << 24;
        std::vector<uint32_t> hx(n), hy(n); std::vector<int32_t> hwv(n);
        for (int i = 0; i < n; i++) {
            hx[i] = 240u * Q + (uint32_t)(sm64(seed) % (32ull * Q));
            hy[i] = (i & 1) ? (uint32_t)sm64(seed) : 240u * Q + (uint32_t)(sm64(seed) % (32ull * Q));
            int64_t w = (int64_t)(sm64(seed) % Q); if (w > Q / 2) w -= Q;
            if ((i & 1023) == 0) w = (i & 1024) ? (int64_t)(Q / 2) : -(int64_t)(Q / 2);
            hwv[i] = (int32_t)w;
        }
        uint32_t *dx, *dy; int32_t* dw; unsigned long long* bad;
        cudaMalloc(&dx, n * 4ull); cudaMalloc(&dy, n * 4ull); cudaMalloc(&dw, n * 4ull); cudaMalloc(&bad, 24); cudaMemset(bad, 0, 24);
        cudaMemcpy(dx, hx.data(), n * 4ull, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy.data(), n * 4ull, cudaMemcpyHostToDevice);
        cudaMemcpy(dw, hwv.data(), n * 4ull, cudaMemcpyHostToDevice);
        k_check<<<n / 256, 256>>>(dx, dy, dw, bad, n);
        unsigned long long hb[3]; cudaMemcpy(hb, bad, 24, cudaMemcpyDeviceToHost);
        printf("  \"exactness\": {\"triples\": %d, \"residue_mismatches\": %llu, \"remainder_out_of_range\": %llu, \"difference_mismatches\": %llu},\n", n, hb[0], hb[1], hb[2]);
    }
    printf("  \"error\": \"%s\"\n}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
