// qt_capi.cu — the extern "C" boundary (include/qtesla_b200.h) and the host-side runtime:
// contexts (device tables + stream), launch geometry, chunked H2D/compute/D2H pipeline for the
// harness-equivalent host-pointer entry points, and contiguous batch sharding over GPUs.
//
// There is NO CPU fallback here: every compute entry point launches a CUDA kernel or returns an
// error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/qtesla_b200.h"
#include "qt_kernels.cuh"
#include "qt_nussbaumer.cuh"

namespace qt {
TwPair h_uni[NUM_TILE_SETS][UNI_KINDS][UNI_MAX];
double h_uniW[NUM_SETS][UNI_KINDS][UNI_MAX];
uint32_t h_uniU[NUM_SETS][UNI_KINDS][UNI_MAX];
}

using namespace qt;

#ifndef QT_AUTO_PREFERS_TMA
#define QT_AUTO_PREFERS_TMA 1
#endif
#ifndef QT_AUTO_PREFERS_DQ
#define QT_AUTO_PREFERS_DQ 0  // signed-lazy sets: the FP64-quotient fused kernel (k_polymul_dq) rather than k_polymul_tma
#endif
#ifndef QT_AUTO_PREFERS_PAIR
#define QT_AUTO_PREFERS_PAIR 1  // n = 2048: two warps per polynomial (k_polymul_pair) rather than one (k_polymul_split)
#endif

#define QT_CUDA(call)                                \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

struct qt_ctx {
    int set = -1;
    int device = -1;
    RtParams p{};
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    TwQuad* d_tab[2] = {nullptr, nullptr};  // kernel table blocks: [0] plain output scale, [1] fused (x 2^32)
    int num_sms = 0;
    int grid_fused = 0, grid_fwd = 0, grid_inv = 0, grid_nuss = 0, grid_tma = 0;
    int occ_fused = 0, occ_tma = 0, tma_warps = 0;
    TwQuad* d_tab_split = nullptr;          // n=2048 only: tables of the split tile (k_polymul_split)
    TwW2* d_tabW = nullptr;                 // signed-lazy sets: the twiddles of d_tab[1] as w / q and as w in [0, q) (k_polymul_dq)
    TwU2* d_tabU = nullptr;
    bool dq_ok = false;
    bool split_ok = false;
    bool pair_ok = false;                   // n=2048 only: two warps per polynomial (k_polymul_pair)
    int variant = 0;  // 0 auto, 1 direct loads, 2 TMA-staged, 3 split tile (n=2048)
    int nuss_variant = 0;  // 0 auto, 1 schoolbook row products, 2 recursive row products (Z_q)
    int overlap = 0;       // programmatic dependent launch: 0 auto (non-blocking streams only), 1 never, 2 always
    cudaStream_t overlap_probe = nullptr;  // stream the cached answer below belongs to
    bool overlap_ok = false, overlap_probed = false;
    size_t smem_fused = 0, smem_one = 0, smem_tma = 0;
    std::atomic<uint64_t> launches{0};
    bool capturing = false;             // between qt_graph_begin and qt_graph_end
    uint64_t capture_launches0 = 0;
    // host pipeline (qt_polymul_host): lazily created
    static constexpr int PIPE = 8;  // maximum number of pipeline slots
    int pipe_slots = 4;             // slots in use
    bool pipe_ramp = true;          // small chunks at both ends of a large batch (QT_PIPE_RAMP=0 switches it off)
    cudaStream_t pipe_stream[PIPE] = {};
    uint32_t* pipe_buf[PIPE] = {};  // x | y per slot, z overwrites x
    // role streams of the pinned pipeline (H2D / kernels / D2H) and the per-slot events that chain them
    cudaStream_t role_stream[3] = {};
    cudaEvent_t ev_in[PIPE] = {}, ev_k[PIPE] = {}, ev_out[PIPE] = {};
    bool pipe_roles = true;         // QT_PIPE_ROLES=0: one stream per slot (the round-1 structure)
    size_t pipe_polys = 0;
    bool pipe_ready = false;
    // staged pipeline for PAGEABLE host buffers (what a malloc-ing caller such as the reference's main.cu
    // passes): worker threads copy chunks through pinned staging buffers, lazily created
    static constexpr int STAGE = 16;  // maximum number of workers
    int stage_workers = 0;           // 0 = not decided yet
    int stage_share = 1;             // contexts sharing the host cores (qt_polymul_host_multi)
    cudaStream_t stage_stream[STAGE] = {};
    uint32_t* stage_dev[STAGE] = {};   // x | y per worker (z overwrites x)
    uint32_t* stage_host[STAGE] = {};  // pinned mirror of the same
    size_t stage_polys = 0;
    bool stage_ready = false;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// every entry point that touches the device: switch to the context's GPU or fail (never run on whatever device
// happens to be current, with table pointers that belong to another one)
#define QT_GUARD(dev)            \
    DeviceGuard g__(dev);        \
    if (!g__.ok) return QT_ERR_NO_DEVICE

std::mutex g_uni_mutex;
bool g_uni_uploaded[64][NUM_TILE_SETS];  // per device

template <int SET> int setup_set(qt_ctx* c, const HostTables& T) {
    using S = KernelShape<SET>;
    c->smem_fused = S::SMEM_DIRECT;
    c->smem_one = S::SMEM_DIRECT;
    QT_CUDA(cudaFuncSetAttribute(k_polymul<SET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_DIRECT));
    QT_CUDA(cudaFuncSetAttribute(k_ntt_forward<SET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_DIRECT));
    QT_CUDA(cudaFuncSetAttribute(k_ntt_inverse<SET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_DIRECT));
    QT_CUDA(cudaFuncSetAttribute(k_ntt_natural<SET, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_DIRECT));
    QT_CUDA(cudaFuncSetAttribute(k_ntt_natural<SET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_DIRECT));
    int occ = 0;
    QT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_polymul<SET>, WARPS_PER_CTA * 32, S::SMEM_DIRECT));
    if (occ < 1) return QT_ERR_UNSUPPORTED;
    c->occ_fused = occ;
    c->grid_fused = occ * c->num_sms;
    c->smem_tma = StageShape<SET>::SMEM;
    c->tma_warps = TmaCfg<SET>::FUSED_WARPS;
    c->occ_tma = 0;
    if (cudaFuncSetAttribute(k_polymul_tma<SET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StageShape<SET>::SMEM) == cudaSuccess) {
        int o2 = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_polymul_tma<SET>, TmaCfg<SET>::FUSED_WARPS * 32, StageShape<SET>::SMEM) == cudaSuccess)
            c->occ_tma = o2;
    } else {
        (void)cudaGetLastError();
    }
    c->grid_tma = c->occ_tma * c->num_sms;
    if constexpr (Cfg<SET>::LAZY) {
        int o3 = 0;
        if (c->d_tabW && c->d_tabU && cudaFuncSetAttribute(k_polymul_dq<SET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DqShape<SET>::SMEM) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_polymul_dq<SET>, DqShape<SET>::WARPS * 32, DqShape<SET>::SMEM) == cudaSuccess)
            c->dq_ok = o3 > 0;
        else
            (void)cudaGetLastError();
    }
    (void)cudaFuncSetAttribute(k_ntt_tma<SET, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StageShape<SET>::SMEM);
    (void)cudaFuncSetAttribute(k_ntt_tma<SET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StageShape<SET>::SMEM);
    (void)cudaFuncSetAttribute(k_bitrev_copy<SET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BitrevShape<SET>::SMEM);
    (void)cudaFuncSetAttribute(k_polymul_ntt<SET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StageShape<SET>::SMEM_BCAST);
    (void)cudaFuncSetAttribute(k_polymul_ntt<SET, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StageShape<SET>::SMEM);
    (void)cudaGetLastError();
    QT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ntt_forward<SET>, WARPS_PER_CTA * 32, S::SMEM_DIRECT));
    c->grid_fwd = std::max(1, occ) * c->num_sms;
    QT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ntt_inverse<SET>, WARPS_PER_CTA * 32, S::SMEM_DIRECT));
    c->grid_inv = std::max(1, occ) * c->num_sms;
    int rc = nuss_setup<SET>(c->num_sms, &c->grid_nuss);
    if (rc) return rc;
    (void)T;
    return 0;
}

int upload_tables(qt_ctx* c) {
    HostTables T;
    build_tables(c->set, &T);
    for (int k = 0; k < 2; k++) {
        const size_t bytes = T.block[k].size() * sizeof(TwQuad);
        QT_CUDA(cudaMalloc(&c->d_tab[k], bytes));
        QT_CUDA(cudaMemcpy(c->d_tab[k], T.block[k].data(), bytes, cudaMemcpyHostToDevice));
    }
    if (!T.blockW[1].empty()) {
        const size_t bytes = T.blockW[1].size() * sizeof(TwW2);
        QT_CUDA(cudaMalloc(&c->d_tabW, bytes));
        QT_CUDA(cudaMemcpy(c->d_tabW, T.blockW[1].data(), bytes, cudaMemcpyHostToDevice));
        const size_t bytes_u = T.blockU[1].size() * sizeof(TwU2);
        QT_CUDA(cudaMalloc(&c->d_tabU, bytes_u));
        QT_CUDA(cudaMemcpy(c->d_tabU, T.blockU[1].data(), bytes_u, cudaMemcpyHostToDevice));
    }
    {
        std::lock_guard<std::mutex> lk(g_uni_mutex);
        memcpy(h_uni[c->set], T.uni, sizeof(T.uni));
        memcpy(h_uniW[c->set], T.uniW, sizeof(T.uniW));
        memcpy(h_uniU[c->set], T.uniU, sizeof(T.uniU));
        if (c->device >= 64 || !g_uni_uploaded[c->device][c->set]) {
            QT_CUDA(cudaMemcpyToSymbol(c_uni, T.uni, sizeof(T.uni), (size_t)c->set * sizeof(T.uni)));
            QT_CUDA(cudaMemcpyToSymbol(c_uniW, T.uniW, sizeof(T.uniW), (size_t)c->set * sizeof(T.uniW)));
            QT_CUDA(cudaMemcpyToSymbol(c_uniU, T.uniU, sizeof(T.uniU), (size_t)c->set * sizeof(T.uniU)));
            if (c->device < 64) g_uni_uploaded[c->device][c->set] = true;
        }
    }
    if (c->set == SET_P_III) {  // the split tile of the fused n=2048 kernel
        HostTables H;
        build_tables_split(&H);
        const size_t bytes = H.block[1].size() * sizeof(TwQuad);
        QT_CUDA(cudaMalloc(&c->d_tab_split, bytes));
        QT_CUDA(cudaMemcpy(c->d_tab_split, H.block[1].data(), bytes, cudaMemcpyHostToDevice));
        {
            std::lock_guard<std::mutex> lk(g_uni_mutex);
            memcpy(h_uni[SET_P_III_H], H.uni, sizeof(H.uni));
            if (c->device >= 64 || !g_uni_uploaded[c->device][SET_P_III_H]) {
                QT_CUDA(cudaMemcpyToSymbol(c_uni, H.uni, sizeof(H.uni), (size_t)SET_P_III_H * sizeof(H.uni)));
                if (c->device < 64) g_uni_uploaded[c->device][SET_P_III_H] = true;
            }
        }
        int occ = 0;
        if (cudaFuncSetAttribute(k_polymul_split<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SplitShape::SMEM) == cudaSuccess &&
            cudaFuncSetAttribute(k_polymul_split<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SplitShape::SMEM_BCAST) == cudaSuccess &&
            cudaFuncSetAttribute(k_polymul_split<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SplitShape::SMEM) == cudaSuccess &&
            cudaFuncSetAttribute(k_ntt_split<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SplitShape::SMEM) == cudaSuccess &&
            cudaFuncSetAttribute(k_ntt_split<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SplitShape::SMEM) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_polymul_split<0>, SplitShape::WARPS * 32, SplitShape::SMEM) == cudaSuccess)
            c->split_ok = occ > 0;
        else
            (void)cudaGetLastError();
        if (c->split_ok && cudaFuncSetAttribute(k_polymul_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairShape::SMEM) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_polymul_pair, PairShape::PAIRS * 64, PairShape::SMEM) == cudaSuccess)
            c->pair_ok = occ > 0;
        else
            (void)cudaGetLastError();
    }
    switch (c->set) {
    case SET_I: return setup_set<SET_I>(c, T);
    case SET_III: return setup_set<SET_III>(c, T);
    case SET_P_I: return setup_set<SET_P_I>(c, T);
    default: return setup_set<SET_P_III>(c, T);
    }
}

// (TMA kernels: grid = min(#SMs, tiles) — with the SM-interleaved tile order a small batch spreads over all
//  SMs, a few warps each, instead of filling 16 warps of a few SMs: B=1024 at n=1024 takes 7 warps on 148 SMs.)
inline int grid_for(int max_grid, size_t tiles) {
    size_t ctas = (tiles + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)max_grid, ctas));
}

// Launch of a kernel that contains pdl_wait() (qt_kernels.cuh) with programmatic stream serialization: its
// prologue may overlap the tail of the previous kernel of the stream.  Only such kernels may be launched this way.
// "Automatic" enables it on non-blocking streams only: a blocking stream is implicitly ordered against the legacy
// default stream, an ordering the early launch is not documented to keep.
// `known`: OVL_YES for the streams the library creates itself (pipeline / staging streams are always non-blocking; their
// launches come from worker threads, which must not touch the cache below), OVL_PROBE for the ctx stream, which
// only the API caller's thread launches on.
enum { OVL_PROBE = -1, OVL_YES = 1 };
static bool overlap_allowed(qt_ctx* c, cudaStream_t s, int known) {
    if (!QT_PDL || c->overlap == 1) return false;
    if (c->overlap == 2 || known == OVL_YES) return true;
    if (!c->overlap_probed || c->overlap_probe != s) {
        unsigned flags = 0;
        c->overlap_ok = s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread &&
                        cudaStreamGetFlags(s, &flags) == cudaSuccess && (flags & cudaStreamNonBlocking) != 0;
        (void)cudaGetLastError();
        c->overlap_probe = s;
        c->overlap_probed = true;
    }
    return c->overlap_ok;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(qt_ctx* c, int known, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = overlap_allowed(c, s, known) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Launch geometry of the TMA-staged kernels for `tiles` warp tiles: grid = min(#SMs, tiles) CTAs of just enough warps
// (<= max_warps).  Small batches therefore take less shared memory per CTA (table block + 2 buffers + 2 barriers per
// warp), and the CTA of the NEXT launch of the stream fits on the SM beside a running one: with programmatic dependent
// launch its prologue (barrier set-up, twiddle-table copy) then overlaps the predecessor's arithmetic.
struct StageGeom { int grid, warps; size_t smem; };
static StageGeom stage_geom(const qt_ctx* c, size_t tiles, int max_warps, size_t table_bytes, size_t warp_bytes, size_t extra = 0) {
    StageGeom g;
    g.grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->num_sms, tiles));
    g.warps = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_warps, (tiles + g.grid - 1) / g.grid));
    g.smem = table_bytes + (size_t)g.warps * warp_bytes + extra;
    return g;
}
template <int SET> static StageGeom stage_geom_set(const qt_ctx* c, size_t tiles, size_t extra = 0, int max_warps = TmaCfg<SET>::WARPS) {
    using G = StageShape<SET>;
    return stage_geom(c, tiles, max_warps, KernelShape<SET>::TW_BYTES, G::BUFS * G::WORDS * sizeof(uint32_t) + 2 * sizeof(uint64_t), extra);
}
static StageGeom stage_geom_split(const qt_ctx* c, size_t tiles, size_t extra = 0) {
    return stage_geom(c, tiles, SplitShape::WARPS, SplitShape::TW_BYTES, 2 * SplitShape::WORDS * sizeof(uint32_t) + 2 * sizeof(uint64_t), extra);
}

template <int SET> int launch_polymul(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B,
                                      cudaStream_t s, int known = OVL_PROBE) {
    const size_t tiles = (B + Cfg<SET>::PPW - 1) / Cfg<SET>::PPW;
    const bool aligned = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;  // bulk copies need 16-byte alignment
    const bool tma = c->occ_tma > 0 && aligned && (c->variant == 2 || (c->variant == 0 && QT_AUTO_PREFERS_TMA));
    if (SET == SET_P_III && c->pair_ok && aligned && (c->variant == 4 || (c->variant == 0 && QT_AUTO_PREFERS_PAIR))) {
        // two warps per polynomial; grid = min(#SMs, B) CTAs of just enough pairs
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->num_sms, B));
        const int pairs = (int)std::max<size_t>(1, std::min<size_t>((size_t)PairShape::PAIRS, (B + grid - 1) / grid));
        cudaError_t e = launch_pdl(c, known, k_polymul_pair, grid, pairs * 64,
                                   PairShape::TW_BYTES + (size_t)pairs * PairShape::PAIR_BYTES, s, x, y, z, B, c->d_tab_split);
        if (e != cudaSuccess) return (int)e;
    } else if (SET == SET_P_III && c->split_ok && aligned && (c->variant == 3 || c->variant == 0)) {
        const StageGeom g = stage_geom_split(c, B);
        cudaError_t e = launch_pdl(c, known, k_polymul_split<0>, g.grid, g.warps * 32, g.smem, s, x, y, z, B, c->d_tab_split);
        if (e != cudaSuccess) return (int)e;
    } else if (Cfg<SET>::LAZY && c->dq_ok && aligned && (c->variant == 5 || (c->variant == 0 && QT_AUTO_PREFERS_DQ))) {
        if constexpr (Cfg<SET>::LAZY) {
            const StageGeom g = stage_geom(c, tiles, DqShape<SET>::WARPS, DqShape<SET>::TABLE_BYTES, DqShape<SET>::WARP_BYTES);
            cudaError_t e = launch_pdl(c, known, k_polymul_dq<SET>, g.grid, g.warps * 32, g.smem, s, x, y, z, B, c->d_tabU, c->d_tabW);
            if (e != cudaSuccess) return (int)e;
        }
    } else if (tma) {
        const StageGeom g = stage_geom_set<SET>(c, tiles, 0, TmaCfg<SET>::FUSED_WARPS);
        cudaError_t e = launch_pdl(c, known, k_polymul_tma<SET>, g.grid, g.warps * 32, g.smem, s, x, y, z, B, c->d_tab[1]);
        if (e != cudaSuccess) return (int)e;
    }
    else
        k_polymul<SET><<<grid_for(c->grid_fused, tiles), WARPS_PER_CTA * 32, c->smem_fused, s>>>(
            x, y, z, B, c->d_tab[1]);
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET> int launch_polymul_ntt(qt_ctx* c, const uint32_t* ahat, bool bcast, const uint32_t* y, uint32_t* z, size_t B) {
    const size_t tiles = (B + Cfg<SET>::PPW - 1) / Cfg<SET>::PPW;
    if (c->occ_tma < 1) return QT_ERR_UNSUPPORTED;
    if ((((uintptr_t)ahat | (uintptr_t)y) & 15) != 0) return QT_ERR_BAD_ARG;  // 128-bit / bulk-copy alignment
    if (SET == SET_P_III && c->split_ok && c->variant != 2) {
        const StageGeom g = stage_geom_split(c, B, bcast ? SplitShape::WORDS * sizeof(uint32_t) : 0);
        const cudaError_t e = bcast ? launch_pdl(c, OVL_PROBE, k_polymul_split<1>, g.grid, g.warps * 32, g.smem, c->stream, ahat, y, z, B, c->d_tab_split)
                                    : launch_pdl(c, OVL_PROBE, k_polymul_split<2>, g.grid, g.warps * 32, g.smem, c->stream, ahat, y, z, B, c->d_tab_split);
        if (e != cudaSuccess) return (int)e;
        c->launches++;
        return (int)cudaGetLastError();
    }
    const StageGeom g = stage_geom_set<SET>(c, tiles, bcast ? Cfg<SET>::N * sizeof(uint32_t) : 0);
    const cudaError_t e = bcast ? launch_pdl(c, OVL_PROBE, k_polymul_ntt<SET, true>, g.grid, g.warps * 32, g.smem, c->stream, ahat, y, z, B, c->d_tab[1])
                                : launch_pdl(c, OVL_PROBE, k_polymul_ntt<SET, false>, g.grid, g.warps * 32, g.smem, c->stream, ahat, y, z, B, c->d_tab[1]);
    if (e != cudaSuccess) return (int)e;
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET, bool INV> int launch_ntt_tma(qt_ctx* c, uint32_t* a, size_t B) {
    const size_t tiles = (B + Cfg<SET>::PPW - 1) / Cfg<SET>::PPW;
    const StageGeom g = stage_geom_set<SET>(c, tiles);
    const cudaError_t e = launch_pdl(c, OVL_PROBE, k_ntt_tma<SET, INV>, g.grid, g.warps * 32, g.smem, c->stream, a, B, c->d_tab[0]);
    if (e != cudaSuccess) return (int)e;
    c->launches++;
    return (int)cudaGetLastError();
}
template <bool INV> int launch_ntt_split(qt_ctx* c, uint32_t* a, size_t B) {
    const StageGeom g = stage_geom_split(c, B);
    const cudaError_t e = launch_pdl(c, OVL_PROBE, k_ntt_split<INV>, g.grid, g.warps * 32, g.smem, c->stream, a, B, c->d_tab_split);
    if (e != cudaSuccess) return (int)e;
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET> int launch_forward(qt_ctx* c, uint32_t* a, size_t B) {
    if (SET == SET_P_III && c->split_ok && c->variant != 1 && c->variant != 2 && ((uintptr_t)a & 15) == 0)
        return launch_ntt_split<false>(c, a, B);
    if (c->occ_tma > 0 && c->variant != 1 && ((uintptr_t)a & 15) == 0) return launch_ntt_tma<SET, false>(c, a, B);
    const size_t tiles = (B + Cfg<SET>::PPW - 1) / Cfg<SET>::PPW;
    k_ntt_forward<SET><<<grid_for(c->grid_fwd, tiles), WARPS_PER_CTA * 32, c->smem_one, c->stream>>>(a, B, c->d_tab[0]);
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET> int launch_inverse(qt_ctx* c, uint32_t* a, size_t B) {
    if (SET == SET_P_III && c->split_ok && c->variant != 1 && c->variant != 2 && ((uintptr_t)a & 15) == 0)
        return launch_ntt_split<true>(c, a, B);
    if (c->occ_tma > 0 && c->variant != 1 && ((uintptr_t)a & 15) == 0) return launch_ntt_tma<SET, true>(c, a, B);
    const size_t tiles = (B + Cfg<SET>::PPW - 1) / Cfg<SET>::PPW;
    k_ntt_inverse<SET><<<grid_for(c->grid_inv, tiles), WARPS_PER_CTA * 32, c->smem_one, c->stream>>>(a, B, c->d_tab[0]);
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET> int launch_natural(qt_ctx* c, uint32_t* a, size_t B, bool inverse) {
    const size_t tiles = (B + Cfg<SET>::PPW - 1) / Cfg<SET>::PPW;
    if (inverse)
        k_ntt_natural<SET, true><<<grid_for(c->grid_inv, tiles), WARPS_PER_CTA * 32, c->smem_one, c->stream>>>(a, B, c->d_tab[0]);
    else
        k_ntt_natural<SET, false><<<grid_for(c->grid_fwd, tiles), WARPS_PER_CTA * 32, c->smem_one, c->stream>>>(a, B, c->d_tab[0]);
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET> int launch_pointwise(qt_ctx* c, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t B) {
    const size_t words = B * Cfg<SET>::N;
    if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)o) & 15) != 0) {  // 128-bit accesses need 16-byte alignment
        const int g1 = (int)std::min<size_t>((size_t)c->num_sms * 8, (words + 255) / 256);
        k_pointwise_scalar<SET><<<std::max(1, g1), 256, 0, c->stream>>>(a, b, o, words);
        c->launches++;
        return (int)cudaGetLastError();
    }
    const int grid = (int)std::min<size_t>((size_t)c->num_sms * 8, (words / 4 + 255) / 256);
    k_pointwise<SET><<<std::max(1, grid), 256, 0, c->stream>>>(a, b, o, words);
    c->launches++;
    return (int)cudaGetLastError();
}
template <int SET> int launch_bitrev(qt_ctx* c, const uint32_t* in, uint32_t* out, size_t B) {
    using S = BitrevShape<SET>;
    const int grid = (int)std::min<size_t>((size_t)c->num_sms * 6, (B + S::WARPS - 1) / S::WARPS);
    k_bitrev_copy<SET><<<std::max(1, grid), S::WARPS * 32, S::SMEM, c->stream>>>(in, out, B);
    c->launches++;
    return (int)cudaGetLastError();
}

#define QT_DISPATCH(c, fn, ...)                                  \
    ((c)->set == SET_I       ? fn<SET_I>(__VA_ARGS__)            \
     : (c)->set == SET_III   ? fn<SET_III>(__VA_ARGS__)          \
     : (c)->set == SET_P_I   ? fn<SET_P_I>(__VA_ARGS__)          \
                             : fn<SET_P_III>(__VA_ARGS__))

void release_pipe(qt_ctx* c) {
    for (int i = 0; i < qt_ctx::PIPE; i++) {
        if (c->pipe_buf[i]) cudaFree(c->pipe_buf[i]);
        if (c->pipe_stream[i]) cudaStreamDestroy(c->pipe_stream[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
        c->pipe_buf[i] = nullptr;
        c->pipe_stream[i] = nullptr;
        c->ev_in[i] = c->ev_k[i] = c->ev_out[i] = nullptr;
    }
    for (int i = 0; i < 3; i++) {
        if (c->role_stream[i]) cudaStreamDestroy(c->role_stream[i]);
        c->role_stream[i] = nullptr;
    }
    c->pipe_ready = false;
}

void release_stage(qt_ctx* c) {
    for (int i = 0; i < qt_ctx::STAGE; i++) {
        if (c->stage_dev[i]) cudaFree(c->stage_dev[i]);
        if (c->stage_host[i]) cudaFreeHost(c->stage_host[i]);
        if (c->stage_stream[i]) cudaStreamDestroy(c->stage_stream[i]);
        c->stage_dev[i] = nullptr;
        c->stage_host[i] = nullptr;
        c->stage_stream[i] = nullptr;
    }
    c->stage_ready = false;
}

// decides the worker count (once) and creates the staging resources of workers [0, want) that do not exist yet
int ensure_stage(qt_ctx* c, size_t want) {
    if (!c->stage_ready) {
        // 8 MiB per operand per chunk; three workers per four host cores, at most STAGE.  Measured on the
        // 16-core GPU box (768 MiB per batch): 2 / 4 / 8 / 12 / 16 workers -> 38.8 / 25.4 / 17.5 / 16.3 / 17.4 ms
        // (driver-staged cudaMemcpyAsync: 61 ms; pinned buffers: 11.3 ms)
        c->stage_polys = std::max<size_t>(1, (size_t)(2u << 20) / c->p.n);
        int w = (int)std::thread::hardware_concurrency() * 3 / (4 * std::max(1, c->stage_share));
        if (const char* e = getenv("QT_STAGE_THREADS")) w = atoi(e);  // tuning aid
        c->stage_workers = std::min((int)qt_ctx::STAGE, std::max(1, w));
        c->stage_ready = true;
    }
    const size_t bytes = 2 * c->stage_polys * c->p.n * sizeof(uint32_t);
    for (size_t i = 0; i < std::min<size_t>(want, (size_t)c->stage_workers); i++) {
        if (c->stage_host[i]) continue;
        cudaError_t e = cudaStreamCreateWithFlags(&c->stage_stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&c->stage_dev[i], bytes);
        if (e == cudaSuccess) e = cudaMallocHost(&c->stage_host[i], bytes);
        if (e != cudaSuccess) {
            release_stage(c);
            return (int)e;
        }
    }
    return 0;
}

// true when cudaMemcpyAsync from/to p would be staged by the driver (ordinary malloc/new memory)
bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

int ensure_pipe(qt_ctx* c) {
    if (c->pipe_ready) return 0;
    // chunk: 4096 polynomials of n=1024 (16 MiB per operand) — large enough for PCIe efficiency,
    // small enough that three slots overlap H2D, compute and D2H
    c->pipe_polys = (size_t)(4u << 20) / c->p.n;
    if (const char* e = getenv("QT_PIPE_CHUNK_WORDS")) {  // tuning aid: coefficients per operand per chunk
        const size_t w = strtoull(e, nullptr, 10);
        if (w >= c->p.n) c->pipe_polys = w / c->p.n;
    }
    if (const char* e = getenv("QT_PIPE_RAMP")) c->pipe_ramp = atoi(e) != 0;  // tuning aid
    if (const char* e = getenv("QT_PIPE_SLOTS")) {  // tuning aid
        const int k = atoi(e);
        if (k >= 1 && k <= qt_ctx::PIPE) c->pipe_slots = k;
    }
    if (const char* e = getenv("QT_PIPE_ROLES")) c->pipe_roles = atoi(e) != 0;  // tuning aid
    for (int i = 0; i < c->pipe_slots; i++) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->pipe_stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&c->pipe_buf[i], 2 * c->pipe_polys * c->p.n * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {  // leave no half-built pipeline behind
            release_pipe(c);
            return (int)e;
        }
    }
    for (int i = 0; i < 3; i++) {
        const cudaError_t e = cudaStreamCreateWithFlags(&c->role_stream[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            release_pipe(c);
            return (int)e;
        }
    }
    c->pipe_ready = true;
    return 0;
}

}  // namespace

extern "C" {

const char* qt_version(void) { return "qtesla_b200 0.1 (sm_100a)"; }

const char* qt_error_string(int code) {
    if (code == 0) return "ok";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
    case QT_ERR_BAD_SET: return "unknown parameter set";
    case QT_ERR_BAD_ARG: return "bad argument";
    case QT_ERR_NO_DEVICE: return "no usable CUDA device";
    case QT_ERR_UNSUPPORTED: return "unsupported on this device/parameter set";
    case QT_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
    }
}

int qt_get_params(int set, qt_params* out) {
    RtParams p;
    if (!out) return QT_ERR_BAD_ARG;
    if (!rt_params(set, &p)) return QT_ERR_BAD_SET;
    out->set = set; out->n = p.n; out->logn = p.logn; out->q = p.q;
    out->psi = p.psi; out->psi_inv = p.psi_inv; out->omega = p.omega; out->omega_inv = p.omega_inv;
    out->n_inv = p.n_inv; out->qinv_neg = p.qinv_neg;
    out->barrett_mu48 = (uint32_t)((1ull << 48) / p.q);
    return 0;
}

int qt_get_table(int set, int which, uint32_t* out) {
    RtParams p;
    if (!out) return QT_ERR_BAD_ARG;
    if (!rt_params(set, &p)) return QT_ERR_BAD_SET;
    HostTables T;
    build_tables(set, &T);
    const std::vector<uint32_t>* src = which == QT_TABLE_BITREV ? &T.bitrev : which == QT_TABLE_PHI ? &T.Phi
                                     : which == QT_TABLE_INVPHI ? &T.invPhi : which == QT_TABLE_TF0 ? &T.tf0
                                     : which == QT_TABLE_TI0 ? &T.ti0 : nullptr;
    if (!src) return QT_ERR_BAD_ARG;
    memcpy(out, src->data(), p.n * sizeof(uint32_t));
    return 0;
}

int qt_device_count(int* out) {
    if (!out) return QT_ERR_BAD_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return (int)e; }
    *out = n;
    return 0;
}

int qt_create(int set, int device, qt_ctx** out) {
    RtParams p;
    if (!out) return QT_ERR_BAD_ARG;
    *out = nullptr;
    if (!rt_params(set, &p)) return QT_ERR_BAD_SET;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return QT_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return QT_ERR_BAD_ARG;
    DeviceGuard g(device);
    if (!g.ok) return QT_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    QT_CUDA(cudaGetDeviceProperties(&prop, device));
    // the library carries ONE cubin, for sm_100a: it loads on compute capability 10.0 and nowhere else
    // (sm_103 / sm_12x parts would fail later with "no kernel image")
    if (prop.major != 10 || prop.minor != 0) return QT_ERR_UNSUPPORTED;
    qt_ctx* c = new (std::nothrow) qt_ctx();
    if (!c) return QT_ERR_NOMEM;
    c->set = set; c->device = device; c->p = p; c->num_sms = prop.multiProcessorCount;
    int rc = (int)cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (!rc) { c->stream = c->own_stream; rc = upload_tables(c); }
    if (rc) { qt_destroy(c); return rc; }
    *out = c;
    return 0;
}

int qt_destroy(qt_ctx* c) {
    if (!c) return 0;
    QT_GUARD(c->device);
    release_pipe(c);
    release_stage(c);
    for (int k = 0; k < 2; k++)
        if (c->d_tab[k]) cudaFree(c->d_tab[k]);
    if (c->d_tab_split) cudaFree(c->d_tab_split);
    if (c->d_tabW) cudaFree(c->d_tabW);
    if (c->d_tabU) cudaFree(c->d_tabU);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

int qt_set_stream(qt_ctx* c, void* s) {
    if (!c) return QT_ERR_BAD_ARG;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

int qt_set_fused_variant(qt_ctx* c, int variant) {
    if (!c || variant < 0 || variant > 5) return QT_ERR_BAD_ARG;
    if (variant == 5 && !c->dq_ok) return QT_ERR_UNSUPPORTED;
    if (variant == 2 && c->occ_tma < 1) return QT_ERR_UNSUPPORTED;
    if (variant == 3 && !c->split_ok) return QT_ERR_UNSUPPORTED;
    if (variant == 4 && !c->pair_ok) return QT_ERR_UNSUPPORTED;
    c->variant = variant;
    return 0;
}

int qt_set_launch_overlap(qt_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return QT_ERR_BAD_ARG;
    c->overlap = mode;
    return 0;
}

int qt_set_nussbaumer_variant(qt_ctx* c, int variant) {
    if (!c || (variant & ~NUSS_WHOLE) < NUSS_AUTO || (variant & ~NUSS_WHOLE) > NUSS_FP64) return QT_ERR_BAD_ARG;
    if ((variant & ~NUSS_WHOLE) == NUSS_FP64 && !QT_DISPATCH(c, nuss_has_f64)) return QT_ERR_UNSUPPORTED;
    c->nuss_variant = variant;
    return 0;
}

int qt_synchronize(qt_ctx* c) {
    if (!c) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    QT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int qt_device_malloc(qt_ctx* c, size_t bytes, void** out) {
    if (!c || !out) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    QT_CUDA(cudaMalloc(out, bytes));
    return 0;
}
int qt_device_free(qt_ctx* c, void* d) {
    if (!c) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    QT_CUDA(cudaFree(d));
    return 0;
}
int qt_host_alloc(size_t bytes, void** out) {
    if (!out) return QT_ERR_BAD_ARG;
    QT_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return 0;
}
int qt_host_free(void* h) {
    QT_CUDA(cudaFreeHost(h));
    return 0;
}
int qt_memcpy_h2d(qt_ctx* c, void* d, const void* h, size_t bytes) {
    if (!c) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    QT_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}
int qt_memcpy_d2h(qt_ctx* c, void* h, const void* d, size_t bytes) {
    if (!c) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    QT_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

int qt_ntt_forward(qt_ctx* c, uint32_t* a, size_t B) {
    if (!c || (!a && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_forward, c, a, B);
}
int qt_ntt_inverse(qt_ctx* c, uint32_t* a, size_t B) {
    if (!c || (!a && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_inverse, c, a, B);
}
int qt_ntt_forward_natural(qt_ctx* c, uint32_t* a, size_t B) {
    if (!c || (!a && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_natural, c, a, B, false);
}
int qt_ntt_inverse_natural(qt_ctx* c, uint32_t* a, size_t B) {
    if (!c || (!a && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_natural, c, a, B, true);
}
int qt_pointwise(qt_ctx* c, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t B) {
    if (!c || ((!a || !b || !o) && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_pointwise, c, a, b, o, B);
}
int qt_polymul(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    if (!c || ((!x || !y || !z) && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_polymul, c, x, y, z, B, c->stream);
}
int qt_polymul_ntt(qt_ctx* c, const uint32_t* ahat, int broadcast, const uint32_t* y, uint32_t* z, size_t B) {
    if (!c || ((!ahat || !y || !z) && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_polymul_ntt, c, ahat, broadcast != 0, y, z, B);
}
int qt_bitrev_copy(qt_ctx* c, const uint32_t* in, uint32_t* out, size_t B) {
    if (!c || ((!in || !out) && B) || (in == out && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    return QT_DISPATCH(c, launch_bitrev, c, in, out, B);
}
int qt_nussbaumer(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int ring) {
    if (!c || ((!x || !y || !z) && B) || (ring != QT_RING_2P32M1 && ring != QT_RING_MODQ && ring != QT_RING_2P32M1_LIFT_Q)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    int rc = QT_DISPATCH(c, nuss_launch, c->grid_nuss, x, y, z, B, ring, c->nuss_variant, c->stream);
    if (rc == 0) c->launches++;
    return rc;
}
int qt_fill_uniform(qt_ctx* c, uint32_t* a, size_t count, uint64_t seed, uint64_t first) {
    if (!c || (!a && count)) return QT_ERR_BAD_ARG;
    if (!count) return 0;
    QT_GUARD(c->device);
    const int grid = (int)std::min<size_t>((size_t)c->num_sms * 8, (count + 255) / 256);
    k_fill_uniform<<<std::max(1, grid), 256, 0, c->stream>>>(a, count, seed, first, c->p.q);
    c->launches++;
    return (int)cudaGetLastError();
}

// Pageable operands: cudaMemcpyAsync would fall back to the driver's own single staging path
// (measured 13 GB/s, 61 ms for the 768 MiB of one n=1024 batch of 65 536).  Instead a few worker threads
// each take chunks end to end — copy x, y into their pinned buffer, H2D, kernel, D2H, copy z out — so the
// host-side copies of one worker overlap the transfers and kernels of the others.  Operands that ARE
// pinned are transferred in place.
static int staged_pipeline(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int nuss_ring,
                           bool px, bool py, bool pz) {
    int rc0 = ensure_stage(c, 0);  // decides the worker count
    if (rc0) return rc0;
    // chunk: at most the staging buffer; smaller for small batches so that every worker gets ~2 chunks
    // (overlap), but not below 1 MiB per operand (a worker thread costs ~50 us to start)
    const size_t n = c->p.n, cap = c->stage_polys;
    const size_t chunk = std::min(cap, std::max<size_t>(std::max<size_t>(1, (256u << 10) / n),
                                                         (B + 2 * c->stage_workers - 1) / (2 * c->stage_workers)));
    const size_t nchunks = (B + chunk - 1) / chunk;
    const int workers = (int)std::min<size_t>((size_t)c->stage_workers, nchunks);
    rc0 = ensure_stage(c, (size_t)workers);
    if (rc0) return rc0;
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    auto work = [&](int w) {
        if (cudaSetDevice(c->device) != cudaSuccess) { err = (int)cudaGetLastError(); return; }
        cudaStream_t s = c->stage_stream[w];
        uint32_t* dx = c->stage_dev[w];
        uint32_t* dy = dx + cap * n;
        uint32_t* hx = c->stage_host[w];
        uint32_t* hy = hx + cap * n;
        for (size_t i = next++; i < nchunks && !err; i = next++) {
            const size_t off = i * chunk * n, cnt = std::min(chunk, B - i * chunk), bytes = cnt * n * sizeof(uint32_t);
            int rc = 0;
            if (px) memcpy(hx, x + off, bytes);
            rc = (int)cudaMemcpyAsync(dx, px ? hx : x + off, bytes, cudaMemcpyHostToDevice, s);
            if (py) memcpy(hy, y + off, bytes);
            if (!rc) rc = (int)cudaMemcpyAsync(dy, py ? hy : y + off, bytes, cudaMemcpyHostToDevice, s);
            if (!rc) {
                if (nuss_ring < 0) rc = QT_DISPATCH(c, launch_polymul, c, dx, dy, dx, cnt, s, OVL_YES);
                else { rc = QT_DISPATCH(c, nuss_launch, c->grid_nuss, dx, dy, dx, cnt, nuss_ring, c->nuss_variant, s); if (!rc) c->launches++; }
            }
            if (!rc) rc = (int)cudaMemcpyAsync(pz ? hx : z + off, dx, bytes, cudaMemcpyDeviceToHost, s);
            const cudaError_t e = cudaStreamSynchronize(s);  // always: nothing of this chunk may stay in flight
            if (!rc && e != cudaSuccess) rc = (int)e;
            if (!rc && pz) memcpy(z + off, hx, bytes);
            if (rc) err = rc;
        }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < workers; w++) th.emplace_back(work, w);
    work(0);
    for (auto& t : th) t.join();
    return err;
}

// Host-pointer product: chunks of pipe_polys polynomials rotate through PIPE device slots; each
// slot's stream does H2D(x,y) -> kernel -> D2H(z), so copies of neighbouring chunks overlap with
// compute and with each other (PCIe is full duplex).
static int host_pipeline(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int nuss_ring) {
    if (!c || ((!x || !y || !z) && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    QT_GUARD(c->device);
    const bool px = is_pageable(x), py = is_pageable(y), pz = is_pageable(z);
    if (px || py || pz) return staged_pipeline(c, x, y, z, B, nuss_ring, px, py, pz);
    int rc = ensure_pipe(c);
    if (rc) return rc;
    const size_t n = c->p.n, cap = c->pipe_polys;
    // smaller chunks for small batches (about two per slot) so that H2D, kernel and D2H still overlap
    const size_t chunk = std::min(cap, std::max<size_t>(std::max<size_t>(1, (128u << 10) / n),
                                                         (B + 2 * c->pipe_slots - 1) / (2 * c->pipe_slots)));
    // Large batches: the pipeline is full only between the first kernel and the last copy back, so the first chunks
    // are small (the first kernel starts after 1/8 of a chunk has arrived instead of a whole one) and so are the
    // last (the final device-to-host copy is short): cap/8, cap/8, cap/4, cap/2, cap ... cap, cap/2, cap/4, cap/8, cap/8.
    // Measured on the 768 MiB step of the bench: 11.27 -> see DESIGN.md (bare concurrent copies: 10.3 ms).
    const bool ramp = c->pipe_ramp && chunk == cap && cap >= 8 && B >= 6 * cap;
    size_t done = 0, chunk_no = 0;
    int slot = 0;
    const bool roles = c->pipe_roles;
    while (done < B && !rc) {
        size_t cnt = std::min(chunk, B - done);
        if (ramp) {
            const size_t left = B - done;
            const size_t front = done < cap / 4 ? cap / 8 : done < cap / 2 ? cap / 4 : done < cap ? cap / 2 : cap;
            const size_t back = left <= cap / 4 ? cap / 8 : left <= cap / 2 ? cap / 4 : left <= cap ? cap / 2 : cap;
            cnt = std::min(std::min(front, back), left);
        }
        const size_t bytes = cnt * n * sizeof(uint32_t);
        uint32_t* dx = c->pipe_buf[slot];
        uint32_t* dy = dx + cap * n;
        // One stream per ROLE (all host-to-device copies in order on the first, the kernels on the second, the copies
        // back on the third), chained per slot by events: the copy engines then see two plain queues.  With one stream
        // per slot a host-to-device copy that waits for its slot sits in front of copies that could go
        // (tools/pipe_probe.cu: 11.6 -> 11.0 ms for the 768 MiB step without any arithmetic; bare copies 10.25 ms).
        cudaStream_t s_in = roles ? c->role_stream[0] : c->pipe_stream[slot];
        cudaStream_t s_k = roles ? c->role_stream[1] : s_in, s_out = roles ? c->role_stream[2] : s_in;
        if (roles && chunk_no >= (size_t)c->pipe_slots) rc = (int)cudaStreamWaitEvent(s_in, c->ev_out[slot], 0);  // slot free again
        if (!rc) rc = (int)cudaMemcpyAsync(dx, x + done * n, bytes, cudaMemcpyHostToDevice, s_in);
        if (!rc) rc = (int)cudaMemcpyAsync(dy, y + done * n, bytes, cudaMemcpyHostToDevice, s_in);
        if (!rc && roles) {
            rc = (int)cudaEventRecord(c->ev_in[slot], s_in);
            if (!rc) rc = (int)cudaStreamWaitEvent(s_k, c->ev_in[slot], 0);
        }
        if (!rc) {
            if (nuss_ring < 0) rc = QT_DISPATCH(c, launch_polymul, c, dx, dy, dx, cnt, s_k, OVL_YES);
            else { rc = QT_DISPATCH(c, nuss_launch, c->grid_nuss, dx, dy, dx, cnt, nuss_ring, c->nuss_variant, s_k); if (!rc) c->launches++; }
        }
        if (!rc && roles) {
            rc = (int)cudaEventRecord(c->ev_k[slot], s_k);
            if (!rc) rc = (int)cudaStreamWaitEvent(s_out, c->ev_k[slot], 0);
        }
        if (!rc) rc = (int)cudaMemcpyAsync(z + done * n, dx, bytes, cudaMemcpyDeviceToHost, s_out);
        if (!rc && roles) rc = (int)cudaEventRecord(c->ev_out[slot], s_out);
        done += cnt;
        chunk_no++;
        slot = (slot + 1) % c->pipe_slots;
    }
    // always drain: the caller's buffers must not be touched after we return, error or not
    for (int i = 0; i < c->pipe_slots; i++) {
        const cudaError_t e = cudaStreamSynchronize(c->pipe_stream[i]);
        if (!rc && e != cudaSuccess) rc = (int)e;
    }
    for (int i = 0; i < 3; i++) {
        const cudaError_t e = cudaStreamSynchronize(c->role_stream[i]);
        if (!rc && e != cudaSuccess) rc = (int)e;
    }
    return rc;
}

int qt_polymul_host(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    return host_pipeline(c, x, y, z, B, -1);
}
int qt_nussbaumer_host(qt_ctx* c, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int ring) {
    if (ring != QT_RING_2P32M1 && ring != QT_RING_MODQ && ring != QT_RING_2P32M1_LIFT_Q) return QT_ERR_BAD_ARG;
    return host_pipeline(c, x, y, z, B, ring);
}

// ---- in-process multi-GPU: contiguous slices [g*B/G, (g+1)*B/G), no collective -------------------
// A qt_multi owns one context and one persistent host thread per device.  Each thread binds itself next to its GPU
// (qt_bind_thread_to_device) BEFORE it creates the context, so the pinned staging buffers of the pageable path are
// allocated and first touched on that NUMA node, then sleeps until a job arrives.  No process-global state: two
// callers use two handles.  Like a context, one handle is not thread-safe.
struct qt_multi {
    int set = -1, ngpus = 0;
    RtParams p{};
    std::vector<qt_ctx*> ctx;
    std::vector<std::thread> th;
    std::vector<int> rc;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    uint64_t job_id = 0;
    int pending = 0;
    bool quit = false;
    const uint32_t *x = nullptr, *y = nullptr;
    uint32_t* z = nullptr;
    size_t B = 0;
};

static void multi_worker(qt_multi* mh, int g) {
    int cpus = 0;
    (void)qt_bind_thread_to_device(g, &cpus);
    int rc = qt_create(mh->set, g, &mh->ctx[g]);
    if (!rc) mh->ctx[g]->stage_share = mh->ngpus;  // the devices share the host cores of the staged (pageable) path
    uint64_t seen = 0;
    {
        std::lock_guard<std::mutex> lk(mh->m);
        mh->rc[g] = rc;
        if (--mh->pending == 0) mh->cv_done.notify_all();
    }
    for (;;) {
        std::unique_lock<std::mutex> lk(mh->m);
        mh->cv_job.wait(lk, [&] { return mh->quit || mh->job_id != seen; });
        if (mh->quit) break;
        seen = mh->job_id;
        const size_t lo = mh->B * (size_t)g / mh->ngpus, hi = mh->B * (size_t)(g + 1) / mh->ngpus, n = mh->p.n;
        const uint32_t *x = mh->x, *y = mh->y;
        uint32_t* z = mh->z;
        lk.unlock();
        int r = mh->ctx[g] ? 0 : QT_ERR_NO_DEVICE;
        if (!r && hi > lo) r = qt_polymul_host(mh->ctx[g], x + lo * n, y + lo * n, z + lo * n, hi - lo);
        lk.lock();
        mh->rc[g] = r;
        if (--mh->pending == 0) mh->cv_done.notify_all();
    }
    if (mh->ctx[g]) { qt_destroy(mh->ctx[g]); mh->ctx[g] = nullptr; }
}

int qt_multi_create(int set, int ngpus, qt_multi** out) {
    RtParams p;
    if (!out) return QT_ERR_BAD_ARG;
    *out = nullptr;
    if (!rt_params(set, &p)) return QT_ERR_BAD_SET;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return QT_ERR_NO_DEVICE;
    if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
    qt_multi* mh = new (std::nothrow) qt_multi();
    if (!mh) return QT_ERR_NOMEM;
    mh->set = set; mh->ngpus = ngpus; mh->p = p;
    mh->ctx.assign(ngpus, nullptr);
    mh->rc.assign(ngpus, 0);
    mh->pending = ngpus;
    for (int g = 0; g < ngpus; g++) mh->th.emplace_back(multi_worker, mh, g);
    {
        std::unique_lock<std::mutex> lk(mh->m);
        mh->cv_done.wait(lk, [&] { return mh->pending == 0; });
    }
    for (int g = 0; g < ngpus; g++)
        if (mh->rc[g]) {
            const int rc = mh->rc[g];
            qt_multi_destroy(mh);
            return rc;
        }
    *out = mh;
    return 0;
}

int qt_multi_destroy(qt_multi* mh) {
    if (!mh) return 0;
    {
        std::lock_guard<std::mutex> lk(mh->m);
        mh->quit = true;
    }
    mh->cv_job.notify_all();
    for (auto& t : mh->th) t.join();
    delete mh;
    return 0;
}

int qt_multi_gpus(qt_multi* mh, int* out) {
    if (!mh || !out) return QT_ERR_BAD_ARG;
    *out = mh->ngpus;
    return 0;
}

int qt_multi_polymul_host(qt_multi* mh, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B) {
    if (!mh || ((!x || !y || !z) && B)) return QT_ERR_BAD_ARG;
    if (!B) return 0;
    std::unique_lock<std::mutex> lk(mh->m);
    mh->x = x; mh->y = y; mh->z = z; mh->B = B;
    mh->pending = mh->ngpus;
    mh->job_id++;
    mh->cv_job.notify_all();
    mh->cv_done.wait(lk, [&] { return mh->pending == 0; });
    for (int rc : mh->rc)
        if (rc) return rc;
    return 0;
}

// one-shot form: a handle for the duration of the call (contexts, streams and staging buffers are set up and torn
// down every time — callers that multiply more than once keep a qt_multi)
int qt_polymul_host_multi(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t B, int ngpus) {
    RtParams p;
    if (!rt_params(set, &p)) return QT_ERR_BAD_SET;
    if ((!x || !y || !z) && B) return QT_ERR_BAD_ARG;
    qt_multi* mh = nullptr;
    int rc = qt_multi_create(set, ngpus, &mh);
    if (rc) return rc;
    rc = qt_multi_polymul_host(mh, x, y, z, B);
    qt_multi_destroy(mh);
    return rc;
}

int qt_shutdown(void) { return 0; }  // kept for ABI compatibility: the library holds no process-global contexts any more

// ---- CUDA graphs for launch-bound batches -------------------------------------------------------------------------
// A caller with a fixed sequence of small launches (B <= 4096: the kernels take a few microseconds, the host-side
// launch path costs as much) records the sequence once and replays it: between qt_graph_begin and qt_graph_end every
// device-pointer entry point of this context is CAPTURED on the ctx stream instead of executed.  Programmatic dependent
// launches are kept: inside the graph they become programmatic edges, so the prologue of node k+1 still overlaps
// the tail of node k.
struct qt_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int device = -1;
    uint64_t kernels = 0;
};

int qt_graph_begin(qt_ctx* c) {
    if (!c || c->capturing) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    (void)overlap_allowed(c, c->stream, OVL_PROBE);  // stream-flag query now, not inside the capture
    QT_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    c->capture_launches0 = c->launches;
    return 0;
}

int qt_graph_end(qt_ctx* c, qt_graph** out) {
    if (!c || !out || !c->capturing) return QT_ERR_BAD_ARG;
    *out = nullptr;
    QT_GUARD(c->device);
    c->capturing = false;
    const uint64_t recorded = c->launches - c->capture_launches0;
    c->launches = c->capture_launches0;  // recorded, not run
    cudaGraph_t g = nullptr;
    QT_CUDA(cudaStreamEndCapture(c->stream, &g));
    qt_graph* h = new (std::nothrow) qt_graph();
    if (!h) { cudaGraphDestroy(g); return QT_ERR_NOMEM; }
    h->graph = g; h->device = c->device; h->kernels = recorded;
    const cudaError_t e = cudaGraphInstantiate(&h->exec, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); delete h; return (int)e; }
    *out = h;
    return 0;
}

int qt_graph_launch(qt_ctx* c, qt_graph* g) {
    if (!c || !g || !g->exec || c->capturing || g->device != c->device) return QT_ERR_BAD_ARG;
    QT_GUARD(c->device);
    QT_CUDA(cudaGraphLaunch(g->exec, c->stream));
    c->launches += g->kernels;
    return 0;
}

int qt_graph_destroy(qt_graph* g) {
    if (!g) return 0;
    DeviceGuard dg(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return 0;
}

int qt_graph_kernel_count(qt_graph* g, uint64_t* out) {
    if (!g || !out) return QT_ERR_BAD_ARG;
    *out = g->kernels;
    return 0;
}

int qt_launch_count(qt_ctx* c, uint64_t* out) {
    if (!c || !out) return QT_ERR_BAD_ARG;
    *out = c->launches;
    return 0;
}

int qt_kernel_info(qt_ctx* c, int* grid, int* block, int* smem, int* per_sm, int* sms) {
    if (!c) return QT_ERR_BAD_ARG;
    if (c->pair_ok && (c->variant == 4 || (c->variant == 0 && QT_AUTO_PREFERS_PAIR))) {
        if (grid) *grid = c->num_sms;
        if (block) *block = PairShape::PAIRS * 64;
        if (smem) *smem = (int)PairShape::SMEM;
        if (per_sm) *per_sm = 1;
        if (sms) *sms = c->num_sms;
        return 0;
    }
    if (c->split_ok && (c->variant == 3 || c->variant == 0)) {
        if (grid) *grid = c->num_sms;
        if (block) *block = SplitShape::WARPS * 32;
        if (smem) *smem = (int)SplitShape::SMEM;
        if (per_sm) *per_sm = 1;
        if (sms) *sms = c->num_sms;
        return 0;
    }
    const bool tma = c->occ_tma > 0 && (c->variant == 2 || (c->variant == 0 && QT_AUTO_PREFERS_TMA));
    if (grid) *grid = tma ? c->grid_tma : c->grid_fused;
    if (block) *block = (tma ? c->tma_warps : WARPS_PER_CTA) * 32;
    if (smem) *smem = (int)(tma ? c->smem_tma : c->smem_fused);
    if (per_sm) *per_sm = tma ? c->occ_tma : c->occ_fused;
    if (sms) *sms = c->num_sms;
    return 0;
}

}  // extern "C"
