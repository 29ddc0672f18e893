// qt_emu.cpp — TEST INFRASTRUCTURE: lane-by-lane CPU execution of the warp-tile engine.
//
// Compiles ntt-gpu-qtesla_b200/csrc/qt_tile.cuh for the host and runs the same phase sequence as
// the CUDA kernels of qt_kernels.cuh, with the 32 lanes of a warp executed one after another between
// the kernels' __syncwarp points.  This checks the index arithmetic, the lazy-reduction bounds
// (true 32-bit wrap-around) and the twiddle tables without a GPU.  It is never loaded by the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../ntt-gpu-qtesla_b200/csrc/qt_tile.cuh"
#include "../../ntt-gpu-qtesla_b200/csrc/qt_nussbaumer.cuh"

namespace qt {
TwPair h_uni[NUM_TILE_SETS][UNI_KINDS][UNI_MAX];
double h_uniW[NUM_SETS][UNI_KINDS][UNI_MAX];
uint32_t h_uniU[NUM_SETS][UNI_KINDS][UNI_MAX];
}
using namespace qt;

namespace {

template <int SET> struct Emu {
    using T = Tile<SET>;
    static constexpr uint32_t E = T::E;
    HostTables tab;
    Emu() {
        build_tables(SET, &tab);
        memcpy(h_uni[SET], tab.uni, sizeof(tab.uni));
        if (SET < NUM_SETS) {
            memcpy(h_uniW[SET < NUM_SETS ? SET : 0], tab.uniW, sizeof(tab.uniW));
            memcpy(h_uniU[SET < NUM_SETS ? SET : 0], tab.uniU, sizeof(tab.uniU));
        }
    }
    typename T::LanePtrs ptrs(uint32_t lane, int kind) const { return T::lane_ptrs(tab.block[kind].data(), lane); }

    // k_polymul_dq (FP64-quotient butterflies, signed-lazy sets): same phases; the denormal arithmetic runs on the host FPU.
    // `stats` (optional): [0] largest |true value| / q seen at the end of a pass (offset form removed); 1e9 if the high
    // half of a register pair was ever non-zero
    void polymul_dq(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, double* stats) {
        if constexpr (T::LAZY) {
            using P64 = typename T::P64;
            alignas(16) static uint32_t bufA[64 * 32], bufB[64 * 32];
            const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
            std::vector<P64> V(32 * E);
            auto v = [&](uint32_t lane) -> P64(&)[E] { return *reinterpret_cast<P64(*)[E]>(&V[lane * E]); };
            double worst = 0;
            const uint64_t one = 1;
            double tiny;
            memcpy(&tiny, &one, sizeof tiny);  // the denormal 2^-1074
            auto seed = [&](uint32_t r) { return P64{tiny * h_uniW[SET < NUM_SETS ? SET : 0][UNI_FWD][r]}; };
            auto track = [&](uint32_t lane) {
                for (uint32_t r = 0; r < E; r++) {
                    if (V[lane * E + r].hi() != 0) worst = 1e9;  // the high half must stay zero
                    const double t = std::fabs((double)((int64_t)V[lane * E + r].lo() - (int64_t)T::DQ_OFF)) / (double)T::Q;
                    worst = t > worst ? t : worst;
                }
            };
            for (size_t tile = 0; tile < ntiles; tile++) {
                const size_t base = tile * T::C::TILE_WORDS;
                auto valid = [&](uint32_t lane) { return tile * T::PPW + lane / T::LPP < batch; };
                for (int op = 0; op < 2; op++) {
                    uint32_t* st = op ? bufB : bufA;
                    const uint32_t* g = (op ? y : x) + base;
                    for (uint32_t l = 0; l < 32; l++) {
                        for (uint32_t r = 0; r < E; r++) v(l)[r] = seed(r).with_lo(valid(l) ? g[T::row_off(l, r)] : 0u);
                        T::fwd_rows_dq(v(l));
                        track(l);
                    }
                    for (uint32_t l = 0; l < 32; l++)
                        for (uint32_t r = 0; r < E; r++) st[T::swz(T::row_off(l, r))] = v(l)[r].lo();
                    for (uint32_t l = 0; l < 32; l++)
                        for (uint32_t r = 0; r < E; r++) v(l)[r] = seed(r).with_lo(st[T::swz(E * l + r)]);
                    for (uint32_t l = 0; l < 32; l++) {
                        T::fwd_cols_dq(v(l), T::lane_ptrs_dq(tab.blockU[1].data(), tab.blockW[1].data(), l));
                        track(l);
                        if (op == 0)
                            for (uint32_t r = 0; r < E; r++) bufA[T::swz(E * l + r)] = v(l)[r].lo();
                    }
                }
                for (uint32_t l = 0; l < 32; l++) {
                    T::pointwise_dq_stash(v(l), bufA, l);
                    T::inv_cols_dq(v(l));
                    track(l);
                }
                for (uint32_t l = 0; l < 32; l++)
                    for (uint32_t r = 0; r < E; r++) bufB[T::swz(E * l + r)] = v(l)[r].lo();
                for (uint32_t l = 0; l < 32; l++)
                    for (uint32_t r = 0; r < E; r++) v(l)[r] = seed(r).with_lo(bufB[T::swz(T::row_off(l, r))]);
                for (uint32_t l = 0; l < 32; l++) {
                    uint32_t out[E];
                    T::inv_rows_dq(v(l), out, T::lane_ptrs_dq(tab.blockU[1].data(), tab.blockW[1].data(), l));
                    T::store_rows(out, z + base, l, valid(l));
                }
            }
            if (stats) stats[0] = worst;
        }
    }

    void polymul(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
        alignas(16) static uint32_t buf[64 * 32];
        const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
        std::vector<uint32_t> VX(32 * E), VY(32 * E);
        auto vx = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&VX[lane * E]); };
        auto vy = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&VY[lane * E]); };
        for (size_t tile = 0; tile < ntiles; tile++) {
            const size_t base = tile * T::C::TILE_WORDS;
            auto valid = [&](uint32_t lane) { return tile * T::PPW + lane / T::LPP < batch; };
            for (uint32_t l = 0; l < 32; l++) {
                T::load_rows(vx(l), x + base, l, valid(l));
                T::load_rows(vy(l), y + base, l, valid(l));
                T::fwd_rows(vx(l));
                T::sts_rows(vx(l), buf, l);
            }
            for (uint32_t l = 0; l < 32; l++) T::lds_cols(vx(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) {
                T::fwd_rows(vy(l));
                T::sts_rows(vy(l), buf, l);
                T::fwd_cols(vx(l), ptrs(l, 1).fwd);
            }
            for (uint32_t l = 0; l < 32; l++) {
                T::lds_cols(vy(l), buf, l);
                T::fwd_cols(vy(l), ptrs(l, 1).fwd);
                T::pointwise_mont(vy(l), vx(l));
                T::inv_cols(vy(l), ptrs(l, 1).inv);
            }
            for (uint32_t l = 0; l < 32; l++) T::sts_cols(vy(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) T::lds_rows(vy(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) {
                T::template inv_rows<UNI_INV_FUSED>(vy(l), ptrs(l, 1));
                T::store_rows(vy(l), z + base, l, valid(l));
            }
        }
    }

    void forward(uint32_t* a, size_t batch) {
        alignas(16) static uint32_t buf[64 * 32];
        const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
        std::vector<uint32_t> V(32 * E);
        auto v = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&V[lane * E]); };
        for (size_t tile = 0; tile < ntiles; tile++) {
            const size_t base = tile * T::C::TILE_WORDS;
            auto valid = [&](uint32_t lane) { return tile * T::PPW + lane / T::LPP < batch; };
            for (uint32_t l = 0; l < 32; l++) {
                T::load_rows(v(l), a + base, l, valid(l));
                T::fwd_rows(v(l));
                T::sts_rows(v(l), buf, l);
            }
            for (uint32_t l = 0; l < 32; l++) {
                T::lds_cols(v(l), buf, l);
                T::fwd_cols(v(l), ptrs(l, 0).fwd);
                T::canon_fwd(v(l));
            }
            for (uint32_t l = 0; l < 32; l++) T::sts_cols(v(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) T::lds_rows(v(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) T::store_rows(v(l), a + base, l, valid(l));
        }
    }

    void inverse(uint32_t* a, size_t batch) {
        alignas(16) static uint32_t buf[64 * 32];
        const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
        std::vector<uint32_t> V(32 * E);
        auto v = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&V[lane * E]); };
        for (size_t tile = 0; tile < ntiles; tile++) {
            const size_t base = tile * T::C::TILE_WORDS;
            auto valid = [&](uint32_t lane) { return tile * T::PPW + lane / T::LPP < batch; };
            for (uint32_t l = 0; l < 32; l++) {
                T::load_rows(v(l), a + base, l, valid(l));
                T::sts_rows(v(l), buf, l);
            }
            for (uint32_t l = 0; l < 32; l++) {
                T::lds_cols(v(l), buf, l);
                T::inv_cols(v(l), ptrs(l, 0).inv);
            }
            for (uint32_t l = 0; l < 32; l++) T::sts_cols(v(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) T::lds_rows(v(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) {
                T::template inv_rows<UNI_INV_PLAIN>(v(l), ptrs(l, 0));
                T::store_rows(v(l), a + base, l, valid(l));
            }
        }
    }

    // natural-order NTT domain (k_ntt_natural): the permutation lives in the global access pattern
    void forward_natural(uint32_t* a, size_t batch) {
        alignas(16) static uint32_t buf[64 * 32];
        const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
        std::vector<uint32_t> V(32 * E);
        auto v = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&V[lane * E]); };
        for (size_t tile = 0; tile < ntiles; tile++) {
            const size_t base = tile * T::C::TILE_WORDS;
            auto valid = [&](uint32_t lane) { return tile * T::PPW + lane / T::LPP < batch; };
            for (uint32_t l = 0; l < 32; l++) {
                T::load_rows(v(l), a + base, l, valid(l));
                T::fwd_rows(v(l));
                T::sts_rows(v(l), buf, l);
            }
            for (uint32_t l = 0; l < 32; l++) {
                T::lds_cols(v(l), buf, l);
                T::fwd_cols(v(l), ptrs(l, 0).fwd);
                T::canon_fwd(v(l));
                T::store_cols_natural(v(l), a + base, l, valid(l));
            }
        }
    }
    void inverse_natural(uint32_t* a, size_t batch) {
        alignas(16) static uint32_t buf[64 * 32];
        const size_t ntiles = (batch + T::PPW - 1) / T::PPW;
        std::vector<uint32_t> V(32 * E);
        auto v = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&V[lane * E]); };
        for (size_t tile = 0; tile < ntiles; tile++) {
            const size_t base = tile * T::C::TILE_WORDS;
            auto valid = [&](uint32_t lane) { return tile * T::PPW + lane / T::LPP < batch; };
            for (uint32_t l = 0; l < 32; l++) {
                T::load_cols_natural(v(l), a + base, l, valid(l));
                T::inv_cols(v(l), ptrs(l, 0).inv);
                T::sts_cols(v(l), buf, l);
            }
            for (uint32_t l = 0; l < 32; l++) T::lds_rows(v(l), buf, l);
            for (uint32_t l = 0; l < 32; l++) {
                T::template inv_rows<UNI_INV_PLAIN>(v(l), ptrs(l, 0));
                T::store_rows(v(l), a + base, l, valid(l));
            }
        }
    }

    // worst bank-conflict degree of the four shared-memory access patterns
    // (32-bit: 32 lanes per wavefront; 128-bit: 8 lanes per wavefront, 4 banks each)
    int bank_conflicts() {
        int worst = 1;
        for (uint32_t r = 0; r < E; r++) {
            int cnt[32] = {0};
            for (uint32_t l = 0; l < 32; l++) cnt[T::swz(T::row_off(l, r)) % 32]++;
            for (int b = 0; b < 32; b++) worst = cnt[b] > worst ? cnt[b] : worst;
        }
        for (uint32_t c = 0; c < E / 4; c++)
            for (uint32_t q8 = 0; q8 < 4; q8++) {
                int cnt[32] = {0};
                for (uint32_t l = q8 * 8; l < q8 * 8 + 8; l++) {
                    const uint32_t p = T::swz(E * l + 4 * c);
                    if (p % 4) return -1;  // 128-bit access must stay aligned
                    for (int k = 0; k < 4; k++) cnt[(p + k) % 32]++;
                }
                for (int b = 0; b < 32; b++) worst = cnt[b] > worst ? cnt[b] : worst;
            }
        // the base + offset form of the rows addresses (two polynomials per warp) must equal the swizzle
        if (T::PPW == 2)
            for (uint32_t l = 0; l < 32; l++) {
                const typename T::RowBases B = T::row_bases(l);
                for (uint32_t r = 0; r < E; r++)
                    if (T::rows_addr(B, r) != T::swz(T::row_off(l, r))) return -3;
            }
        // the swizzle must be a permutation of the tile
        std::vector<uint8_t> seen(T::C::TILE_WORDS, 0);
        for (uint32_t o = 0; o < T::C::TILE_WORDS; o++) {
            const uint32_t p = T::swz(o);
            if (p >= T::C::TILE_WORDS || seen[p]) return -2;
            seen[p] = 1;
        }
        return worst;
    }
};

// k_polymul_split lane by lane: n=2048 as two 1024-point halves on the SET_P_III_H tile
struct EmuSplit {
    using T = Tile<SET_P_III_H>;
    static constexpr uint32_t E = T::E, HALF = T::N;
    HostTables tab;
    EmuSplit() {
        build_tables_split(&tab);
        memcpy(h_uni[SET_P_III_H], tab.uni, sizeof(tab.uni));
    }
    void polymul(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
        alignas(16) static uint32_t A[2 * 1024], B[2 * 1024];
        const TwQuad* s_tw = tab.block[1].data();
        std::vector<uint32_t> V(32 * E), H(32 * E);
        auto v = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&V[lane * E]); };
        auto hi = [&](uint32_t lane) -> uint32_t(&)[E] { return *reinterpret_cast<uint32_t(*)[E]>(&H[lane * E]); };
        for (size_t tile = 0; tile < batch; tile++) {
            for (int op = 0; op < 2; op++) {
                uint32_t* st = op ? B : A;
                memcpy(st, (op ? y : x) + tile * 2 * HALF, 2 * HALF * sizeof(uint32_t));  // the bulk copy
                for (uint32_t l = 0; l < 32; l++)
                    for (uint32_t r = 0; r < E; r++) { v(l)[r] = st[l + 32 * r]; hi(l)[r] = st[HALF + l + 32 * r]; }
                for (uint32_t l = 0; l < 32; l++) {
                    T::split_fwd(v(l), hi(l));
                    T::sts_rows(hi(l), st + HALF, l);
                }
                for (uint32_t h = 0; h < 2; h++) {
                    uint32_t* sh = st + h * HALF;
                    for (uint32_t l = 0; l < 32; l++) {
                        if (h) T::lds_rows(v(l), sh, l);
                        T::fwd_rows(v(l), 32 * h);
                        T::sts_rows(v(l), sh, l);
                    }
                    for (uint32_t l = 0; l < 32; l++) {
                        T::lds_cols(v(l), sh, l);
                        T::fwd_cols(v(l), s_tw + h * T::TW_QUADS + l);
                    }
                    if (op == 0) {
                        for (uint32_t l = 0; l < 32; l++) T::sts_cols(v(l), sh, l);
                        continue;
                    }
                    for (uint32_t l = 0; l < 32; l++) {
                        T::pointwise_mont_stash(v(l), A + h * HALF, l);
                        T::inv_cols(v(l), s_tw + (1 - h) * T::TW_QUADS + (T::BLOCKS - 1 - l));
                    }
                    for (uint32_t l = 0; l < 32; l++) T::sts_cols(v(l), sh, l);
                    for (uint32_t l = 0; l < 32; l++) {
                        T::lds_rows(v(l), sh, l);
                        T::template inv_rows<UNI_INV_FUSED>(v(l), typename T::LanePtrs{nullptr, nullptr, nullptr}, 32 * h);
                        if (h == 0) T::sts_rows(v(l), sh, l);
                    }
                }
            }
            uint32_t* zt = z + tile * 2 * HALF;
            for (uint32_t l = 0; l < 32; l++)
                for (uint32_t r = 0; r < E; r++) {
                    uint32_t a = B[T::swz(T::row_off(l, r))];
                    T::template split_inv<UNI_INV_FUSED>(a, v(l)[r]);
                    zt[l + 32 * r] = a;
                    zt[HALF + l + 32 * r] = v(l)[r];
                }
        }
    }
};

template <int SET> Emu<SET>& emu() {
    static Emu<SET> e;
    return e;
}

}  // namespace

// Nussbaumer: same phase sequence as k_nussbaumer; warps run read-all-lanes then write-all-lanes,
// threads of a phase run one after another between the kernel's __syncthreads points.
template <int SET, int RING, bool REC = false> int emu_nussbaumer(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    using K = NussCfg<SET>;
    using NU = Nuss<SET, RING>;
    std::vector<uint32_t> smem(K::P * K::POLY_WORDS);
    constexpr uint32_t WARPS = K::THREADS / 32;
    const size_t ngroups = (batch + K::P - 1) / K::P;
    for (size_t grp = 0; grp < ngroups; grp++) {
        const size_t p0 = grp * K::P;
        const uint32_t np = (uint32_t)((batch - p0 < K::P) ? batch - p0 : K::P);
        const uint32_t per = K::THREADS / K::P;
        for (uint32_t tid = 0; tid < K::THREADS; tid++) {
            const uint32_t p = tid / per;
            if (p < np)
                NU::load(tid % per, per, x + (p0 + p) * K::N, y + (p0 + p) * K::N, smem.data() + p * K::POLY_WORDS,
                         smem.data() + p * K::POLY_WORDS + K::X_WORDS);
        }
        for (int j = (int)K::LOGM - 1; j >= 0; j--)
            for (uint32_t warp = 0; warp < WARPS; warp++)
                for (uint32_t w = warp; w < K::P * 2 * K::M; w += WARPS) {
                    const uint32_t p = w / (2 * K::M), op = (w / K::M) & 1, bf = w % K::M;
                    if (p >= np) continue;
                    uint32_t* v = smem.data() + p * K::POLY_WORDS + (op ? K::X_WORDS : 0);
                    const uint32_t stride = op ? K::YS : K::XS;
                    typename NU::Regs rg[32];
                    for (uint32_t lane = 0; lane < 32; lane++) NU::fwd_read(lane, (uint32_t)j, bf, v, stride, rg[lane]);
                    for (uint32_t lane = 0; lane < 32; lane++) NU::fwd_write(lane, (uint32_t)j, bf, v, stride, rg[lane]);
                }
        for (uint32_t tid = 0; tid < K::THREADS; tid++) {
            const uint32_t p = tid / K::ROWS, row = tid % K::ROWS;
            if (p < np) {
                uint32_t* xr = smem.data() + p * K::POLY_WORDS + row * K::XS;
                uint32_t* yr = smem.data() + p * K::POLY_WORDS + K::X_WORDS + row * K::YS;
                if (REC) NU::product_recursive(xr, yr);
                else NU::product(xr, yr);
            }
        }
        for (uint32_t j = 0; j <= K::LOGM; j++)
            for (uint32_t warp = 0; warp < WARPS; warp++)
                for (uint32_t w = warp; w < K::P * K::M; w += WARPS) {
                    const uint32_t p = w / K::M, bf = w % K::M;
                    if (p >= np) continue;
                    uint32_t* zr = smem.data() + p * K::POLY_WORDS;
                    typename NU::Regs rg[32];
                    for (uint32_t lane = 0; lane < 32; lane++) NU::inv_read(lane, j, bf, zr, rg[lane]);
                    for (uint32_t lane = 0; lane < 32; lane++) NU::inv_write(lane, j, bf, zr, rg[lane]);
                }
        for (uint32_t tid = 0; tid < K::THREADS; tid++) {
            const uint32_t p = tid / per;
            if (p < np) NU::store(tid % per, per, smem.data() + p * K::POLY_WORDS, z + (p0 + p) * K::N);
        }
    }
    return 0;
}

// The signed-lazy flavour of NussInner as the warp-resident kernel calls it (length 32, q < 2^25): x, y are
// two's-complement words with |x| <= bx, |y| <= by; returns z * 1 (fix = the plain-product constant) in
// [-q/2, 3q/2).  rows products at once.
template <int SET> int emu_inner_lazy(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t rows) {
    using IN = NussInner<SET, 32, true>;
    const TwPair fix = tw_signed_c(IN::fix_value(1u), Cfg<SET>::Q);
    for (size_t r = 0; r < rows; r++) {
        uint32_t xa[32], ya[32], za[32];
        for (int j = 0; j < 32; j++) { xa[j] = x[32 * r + j]; ya[j] = y[32 * r + j]; }
        IN::product(xa, ya, za, fix);
        for (int j = 0; j < 32; j++) z[32 * r + j] = za[j];
    }
    return 0;
}

// NussRowF64 (row products on the FP64 pipe) on operands in [-q/2, 3q/2)
template <int SET> int emu_row_f64(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t rows) {
    for (size_t r = 0; r < rows; r++) {
        uint32_t xa[32], ya[32];
        for (int j = 0; j < 32; j++) { xa[j] = x[32 * r + j]; ya[j] = y[32 * r + j]; }
        NussRowF64<SET>::product(xa, ya);  // z over x
        for (int j = 0; j < 32; j++) z[32 * r + j] = xa[j];
    }
    return 0;
}

#define EMU_DISPATCH(set, call)                  \
    switch (set) {                               \
    case SET_I: emu<SET_I>().call; break;        \
    case SET_III: emu<SET_III>().call; break;    \
    case SET_P_I: emu<SET_P_I>().call; break;    \
    case SET_P_III: emu<SET_P_III>().call; break; \
    default: return -1;                          \
    }

extern "C" {
// static range plan of the Harvey path (qt_tile.cuh: HarveyPlan) for qTESLA-p-I: out = {ok, corrections in one forward
// transform, in the pointwise product of two forward outputs, in the inverse transform, largest forward-output bound,
// bound carried across the forward transposition, bound carried across the inverse transposition, 2^32 / q}
int qtemu_harvey_plan(uint32_t* out) {
    using T = qt::Tile<qt::SET_P_I>;
    static_assert(T::HLAZY, "qTESLA-p-I runs the range plan");
    constexpr auto p = T::HPlan::make();
    uint32_t f = 0, w = 0, v = 0, m = 0;
    for (uint32_t l = 0; l < T::LB1; l++) for (uint32_t i = 0; i < T::E / 2; i++) f += p.fr[l][i] != 0;
    for (uint32_t k = 0; k < T::LB2; k++) for (uint32_t i = 0; i < T::E / 2; i++) for (int c = 0; c < 3; c++) f += p.fc[k][i][c] != 0;
    for (uint32_t r = 0; r < T::E; r++) {
        for (int c = 0; c < 3; c++) w += (p.pwa[r][c] != 0) + (p.pwb[r][c] != 0);
        m = p.fout[r] > m ? p.fout[r] : m;
    }
    for (uint32_t k = 0; k < T::LB2; k++) for (uint32_t i = 0; i < T::E / 2; i++) v += (p.ic[k][i][0] != 0) + (p.ic[k][i][1] != 0);
    for (uint32_t k = 0; k < T::LB1; k++) for (uint32_t i = 0; i < T::E / 2; i++) v += (p.ir[k][i][0] != 0) + (p.ir[k][i][1] != 0);
    out[0] = p.ok; out[1] = f; out[2] = w; out[3] = v; out[4] = m; out[5] = p.rows_out; out[6] = p.icols_out; out[7] = T::QCAP;
    return 0;
}

int qtemu_polymul(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    EMU_DISPATCH(set, polymul(x, y, z, batch));
    return 0;
}
// FP64-quotient fused kernel (fused variant 5); stats[0] = largest |value| / q met on the way
int qtemu_polymul_dq(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, double* stats) {
    if (set != SET_I && set != SET_III) return -4;
    EMU_DISPATCH(set, polymul_dq(x, y, z, batch, stats));
    return 0;
}
// one FP64-quotient product y*w - rint(y w / q) q for every (y, w) pair, w in [0, q): out = the signed remainder
int qtemu_dq_remainder(int set, const uint32_t* y, const uint32_t* w, int32_t* out, size_t count) {
    if (set != SET_I && set != SET_III) return -4;
    for (size_t i = 0; i < count; i++) {
        const uint32_t q = set == SET_I ? Cfg<SET_I>::Q : Cfg<SET_III>::Q;
        const double W = dq_companion(w[i], q);
        const auto yy = Tile<SET_III>::P64::make(y[i], 0u);
        const auto t = Tile<SET_III>::dq_quot(yy, W);
        if (t.hi() != 0) return -5;  // the quotient must come out as {integer, 0}
        out[i] = (int32_t)(y[i] * w[i] - t.lo() * q);
    }
    return 0;
}
int qtemu_polymul_split(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    static EmuSplit e;
    e.polymul(x, y, z, batch);
    return 0;
}
int qtemu_forward(int set, uint32_t* a, size_t batch) {
    EMU_DISPATCH(set, forward(a, batch));
    return 0;
}
int qtemu_inverse(int set, uint32_t* a, size_t batch) {
    EMU_DISPATCH(set, inverse(a, batch));
    return 0;
}
int qtemu_forward_natural(int set, uint32_t* a, size_t batch) {
    EMU_DISPATCH(set, forward_natural(a, batch));
    return 0;
}
int qtemu_inverse_natural(int set, uint32_t* a, size_t batch) {
    EMU_DISPATCH(set, inverse_natural(a, batch));
    return 0;
}
int qtemu_nussbaumer(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, int ring) {
    switch (set * 2 + ring) {
    case 0: return emu_nussbaumer<SET_I, 0>(x, y, z, batch);
    case 1: return emu_nussbaumer<SET_I, 1>(x, y, z, batch);
    case 2: return emu_nussbaumer<SET_III, 0>(x, y, z, batch);
    case 3: return emu_nussbaumer<SET_III, 1>(x, y, z, batch);
    case 4: return emu_nussbaumer<SET_P_I, 0>(x, y, z, batch);
    case 5: return emu_nussbaumer<SET_P_I, 1>(x, y, z, batch);
    case 6: return emu_nussbaumer<SET_P_III, 0>(x, y, z, batch);
    case 7: return emu_nussbaumer<SET_P_III, 1>(x, y, z, batch);
    default: return -1;
    }
}
// recursive row products (k_nussbaumer<SET, 1, true>): same phases, NussInner instead of the schoolbook
int qtemu_nussbaumer_recursive(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    switch (set) {
    case SET_I: return emu_nussbaumer<SET_I, 1, true>(x, y, z, batch);
    case SET_III: return emu_nussbaumer<SET_III, 1, true>(x, y, z, batch);
    case SET_P_I: return emu_nussbaumer<SET_P_I, 1, true>(x, y, z, batch);
    case SET_P_III: return emu_nussbaumer<SET_P_III, 1, true>(x, y, z, batch);
    default: return -1;
    }
}
int qtemu_inner_lazy(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t rows) {
    switch (set) {
    case SET_I: return emu_inner_lazy<SET_I>(x, y, z, rows);
    case SET_III: return emu_inner_lazy<SET_III>(x, y, z, rows);
    default: return -1;
    }
}
int qtemu_row_f64(int set, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t rows) {
    switch (set) {
    case SET_I: return emu_row_f64<SET_I>(x, y, z, rows);
    case SET_III: return emu_row_f64<SET_III>(x, y, z, rows);
    default: return -1;
    }
}
int qtemu_bank_conflicts(int set) {
    int r = 0;
    switch (set) {
    case SET_I: r = emu<SET_I>().bank_conflicts(); break;
    case SET_III: r = emu<SET_III>().bank_conflicts(); break;
    case SET_P_I: r = emu<SET_P_I>().bank_conflicts(); break;
    case SET_P_III: r = emu<SET_P_III>().bank_conflicts(); break;
    default: return -9;
    }
    return r;
}
}
