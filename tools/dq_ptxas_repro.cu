// dq_ptxas_repro.cu — level-by-level GPU-vs-host comparison of the inverse cols pass of the FP64-quotient kernel (qTESLA-I).
// Written to find why k_polymul_dq was bit-exact on the host emulator and for qTESLA-III but wrong for qTESLA-I on the GPU
// (run r02B): with the quotient written as a plain product  t = D(y) * W  ptxas 12.9 emitted ONE  DMUL W, {-q, 4}  per level
// and  x' = y * lo(that) + u  for every butterfly (cuobjdump -sass), which is not what the PTX says; with
// fma.rn.f64 t, D(y), W, 0  (qt_tile.cuh: dq_quot) the SASS is one DFMA per butterfly and GPU == host.  Build with
// -DQT_DQ_PLAIN_MUL to see the failing form.
//   nvcc -std=c++17 -O3 -diag-suppress=128 -gencode arch=compute_100a,code=sm_100a -o tools/dq_ptxas_repro tools/dq_ptxas_repro.cu
#include <cstdio>
#include "../ntt-gpu-qtesla_b200/csrc/qt_tile.cuh"
namespace qt { TwPair h_uni[NUM_TILE_SETS][UNI_KINDS][UNI_MAX]; double h_uniW[NUM_SETS][UNI_KINDS][UNI_MAX]; uint32_t h_uniU[NUM_SETS][UNI_KINDS][UNI_MAX]; }
using namespace qt;
template <int SET, int LV> QT_HD void levels(typename Tile<SET>::P64 (&v)[32]) {
    using T = Tile<SET>;
#pragma unroll
    for (uint32_t s_ = 0; s_ < LV; s_++) {
        const uint32_t l = 1u << s_;
#pragma unroll
        for (uint32_t i = 0; i < 16; i++) {
            const uint32_t u = i / l, j = i % l;
            auto& x = v[2 * l * u + j];
            auto& y = v[2 * l * u + j + l];
            if (j == 0) {
                const uint32_t a = x.lo(), b = y.lo();
                x = x.with_lo(a + b - T::DQ_OFF);
                y = y.with_lo(a - b + T::DQ_OFF);
            } else {
                T::ct_dq(x, y, uni_U<SET, UNI_INV_PLAIN>(l + j), uni_W<SET, UNI_INV_PLAIN>(l + j));
            }
        }
    }
}
template <int SET, int LV> __global__ void k(uint32_t c, double tiny, uint32_t* out) {
    using T = Tile<SET>;
    typename T::P64 v[32];
#pragma unroll
    for (uint32_t r = 0; r < 32; r++) v[r] = typename T::P64{tiny * uni_W<SET, UNI_FWD>(r)}.with_lo(c + T::DQ_OFF + (threadIdx.x == 77 ? r : 0));
    levels<SET, LV>(v);
#pragma unroll
    for (uint32_t r = 0; r < 32; r++) out[threadIdx.x * 32 + r] = v[r].lo() - T::DQ_OFF;
}
template <int SET, int LV> void run() {
    using T = Tile<SET>;
    uint32_t* d; cudaMalloc(&d, 32 * 32 * 4);
    const uint64_t one = 1; double tiny; memcpy(&tiny, &one, 8);
    k<SET, LV><<<1, 32>>>(1052805u, tiny, d);
    uint32_t h[32 * 32]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    typename T::P64 v[32];
    for (uint32_t r = 0; r < 32; r++) v[r] = typename T::P64{tiny * h_uniW[SET][UNI_FWD][r]}.with_lo(1052805u + T::DQ_OFF);
    levels<SET, LV>(v);
    printf("SET %d levels %d\n  gpu :", SET, LV);
    for (int r = 0; r < 16; r++) printf(" %d", (int)h[r]);
    printf("\n  host:");
    for (int r = 0; r < 16; r++) printf(" %d", (int)(v[r].lo() - T::DQ_OFF));
    printf("\n");
}
int main() {
    HostTables Tb; build_tables(SET_I, &Tb);
    cudaMemcpyToSymbol(c_uniW, Tb.uniW, sizeof(Tb.uniW), 0); cudaMemcpyToSymbol(c_uniU, Tb.uniU, sizeof(Tb.uniU), 0);
    memcpy(h_uniW[0], Tb.uniW, sizeof(Tb.uniW)); memcpy(h_uniU[0], Tb.uniU, sizeof(Tb.uniU));
    run<SET_I, 1>(); run<SET_I, 2>(); run<SET_I, 3>(); run<SET_I, 4>();
    return 0;
}
