// pipe_probe.cu — which stream structure gets a chunked H2D(x,y) -> kernel -> D2H(z) pipeline closest to the bare-copy
// ceiling?  No arithmetic: the "kernel" is a device-to-device copy of the chunk (x -> z buffer), so only the copy
// scheduling is measured.  512 MiB in, 256 MiB out, pinned host memory (one e2e step of the bench).
//   A  per-slot streams: each slot's stream does H2D x, H2D y, kernel, D2H (what qt_polymul_host does), S slots
//   B  three role streams (H2D / compute / D2H) chained by events, S device slots
//   C  bare: all H2D on one stream, all D2H on another, no dependencies (the ceiling)
//   nvcc -O3 -std=c++17 -o tools/pipe_probe tools/pipe_probe.cu
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
static const size_t MB = 1u << 20, TOTAL = 256 * MB;  // per operand
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void k_touch(const uint4* x, const uint4* y, uint4* z, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 a = x[i], b = y[i];
        z[i] = make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w);
    }
}

int main() {
    char *hx, *hy, *hz;
    CK(cudaHostAlloc(&hx, TOTAL, 0)); CK(cudaHostAlloc(&hy, TOTAL, 0)); CK(cudaHostAlloc(&hz, TOTAL, 0));
    const int MAXS = 6;
    for (size_t chunk : {4 * MB, 8 * MB, 16 * MB, 32 * MB}) {
        for (int S : {3, 4, 6}) {
            std::vector<char*> dx(S), dy(S);
            std::vector<cudaStream_t> st(S);
            std::vector<cudaEvent_t> ev_in(S), ev_k(S), ev_out(S);
            for (int i = 0; i < S; i++) {
                CK(cudaMalloc(&dx[i], chunk)); CK(cudaMalloc(&dy[i], chunk));
                CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
            }
            cudaStream_t s_in, s_k, s_out;
            CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
            const size_t nch = TOTAL / chunk;
            auto runA = [&]() {
                for (size_t c = 0; c < nch; c++) {
                    const int i = (int)(c % S);
                    CK(cudaMemcpyAsync(dx[i], hx + c * chunk, chunk, cudaMemcpyHostToDevice, st[i]));
                    CK(cudaMemcpyAsync(dy[i], hy + c * chunk, chunk, cudaMemcpyHostToDevice, st[i]));
                    k_touch<<<296, 256, 0, st[i]>>>((uint4*)dx[i], (uint4*)dy[i], (uint4*)dx[i], chunk / 16);
                    CK(cudaMemcpyAsync(hz + c * chunk, dx[i], chunk, cudaMemcpyDeviceToHost, st[i]));
                }
                for (int i = 0; i < S; i++) CK(cudaStreamSynchronize(st[i]));
            };
            auto runB = [&]() {
                for (size_t c = 0; c < nch; c++) {
                    const int i = (int)(c % S);
                    if (c >= (size_t)S) CK(cudaStreamWaitEvent(s_in, ev_out[i], 0));  // slot free again
                    CK(cudaMemcpyAsync(dx[i], hx + c * chunk, chunk, cudaMemcpyHostToDevice, s_in));
                    CK(cudaMemcpyAsync(dy[i], hy + c * chunk, chunk, cudaMemcpyHostToDevice, s_in));
                    CK(cudaEventRecord(ev_in[i], s_in));
                    CK(cudaStreamWaitEvent(s_k, ev_in[i], 0));
                    k_touch<<<296, 256, 0, s_k>>>((uint4*)dx[i], (uint4*)dy[i], (uint4*)dx[i], chunk / 16);
                    CK(cudaEventRecord(ev_k[i], s_k));
                    CK(cudaStreamWaitEvent(s_out, ev_k[i], 0));
                    CK(cudaMemcpyAsync(hz + c * chunk, dx[i], chunk, cudaMemcpyDeviceToHost, s_out));
                    CK(cudaEventRecord(ev_out[i], s_out));
                }
                CK(cudaStreamSynchronize(s_out)); CK(cudaStreamSynchronize(s_in)); CK(cudaStreamSynchronize(s_k));
            };
            auto timeit = [&](auto&& f) { f(); const double t0 = now(); for (int r = 0; r < 5; r++) f(); return (now() - t0) / 5 * 1e3; };
            const double a = timeit(runA), b = timeit(runB);
            printf("{\"chunk_MiB\": %zu, \"slots\": %d, \"A_per_slot_streams_ms\": %.3f, \"B_role_streams_ms\": %.3f}\n", chunk / MB, S, a, b);
            fflush(stdout);
            for (int i = 0; i < S; i++) { cudaFree(dx[i]); cudaFree(dy[i]); cudaStreamDestroy(st[i]); cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_k[i]); cudaEventDestroy(ev_out[i]); }
            cudaStreamDestroy(s_in); cudaStreamDestroy(s_k); cudaStreamDestroy(s_out);
        }
    }
    {   // C: the ceiling
        char *dX, *dZ;
        CK(cudaMalloc(&dX, 2 * TOTAL)); CK(cudaMalloc(&dZ, TOTAL));
        cudaStream_t s1, s2;
        CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        auto runC = [&]() {
            CK(cudaMemcpyAsync(dX, hx, TOTAL, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(dX + TOTAL, hy, TOTAL, cudaMemcpyHostToDevice, s1));
            CK(cudaMemcpyAsync(hz, dZ, TOTAL, cudaMemcpyDeviceToHost, s2));
            CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
        };
        runC();
        const double t0 = now();
        for (int r = 0; r < 5; r++) runC();
        printf("{\"bare_copies_ms\": %.3f}\n", (now() - t0) / 5 * 1e3);
    }
    (void)MAXS;
    return 0;
}
