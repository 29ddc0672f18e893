// qt_tile.cuh — the warp-tile negacyclic NTT engine (device code, also compilable for the host).
//
// Replaces, for the hot path, the reference's one-launch-per-level stage kernels
// (GS_radix2NTT_gpu0/1/2 NTT.cu:953-1031, radix2NTT_gpu0/1 1436-1470, NTTStock_gpu* 1085-1153,
// radix2INTT_gpu0/1/2 1374-1433, GS_radix2INTT_gpu* 1224-1240/1033-1056, pointwise_mult 1155-1160,
// bit_reverse_copy_tbl_*_gpu 487-509) with one fused pass in which a polynomial never leaves the SM.
//
// Design:
//  * One WARP owns a tile of E*32 consecutive words = PPW whole polynomials (n=512: 2, n=1024: 1,
//    n=2048: 1 with E=64).  No __syncthreads in the steady state, only __syncwarp.
//  * Merged-psi Cooley-Tukey forward / Gentleman-Sande inverse: zeta[k] = psi^brv(k).  The forward
//    output at position i is x(psi^(2 brv(i)+1)) — bit-for-bit what the reference's
//    Phi-scale + radix2NTTGS (NTT.cu:1866-1876) leaves — with no psi pass and no bit-reverse pass.
//  * Two register-resident passes per transform, joined by one shared-memory transposition:
//      "rows" layout : register r of lane (p,j) holds coefficient j + LPP*r of polynomial p
//                      -> the log2(E) levels with the largest strides are thread-local and their
//                         twiddles are identical for all lanes (constant-bank operands);
//      "cols" layout : register r of lane L holds tile word E*L + r (E consecutive coefficients)
//                      -> the remaining levels are thread-local; twiddles are per lane (shared memory).
//  * Modular arithmetic: Shoup multiplication by precomputed constants (1 mul.hi + 2 mul.lo),
//    result in [0,2q).  For q < 2^25 ("LAZY") butterflies carry no correction at all: the forward
//    transform grows values by 2q per level (< 21q), the inverse doubles per level with one
//    mid-transform Barrett fold of the few registers that could overflow.  For the 29/30-bit moduli
//    Harvey's [0,4q) butterflies are used.  The single final store is canonical in [0,q).
//
// Everything here is __host__ __device__ so that tests/emu can execute the identical index
// arithmetic and 32-bit wrap-around behaviour lane by lane on a CPU (test infrastructure only).
#pragma once
#include <cstddef>
#include <cstdint>

#include "qt_params.h"
#include "qt_tables.h"

#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#define QT_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define QT_HD inline
#define QT_CONSTEXPR_HD constexpr
#endif

namespace qt {

#if defined(__CUDACC__)
// Uniform twiddles of the rows pass.  Indexed only with compile-time constants after unrolling,
// so they are consumed as constant-bank operands of IMAD (no load instruction).
__constant__ TwPair c_uni[NUM_SETS][UNI_KINDS][UNI_MAX];
#endif
#if !defined(__CUDA_ARCH__)
extern TwPair h_uni[NUM_SETS][UNI_KINDS][UNI_MAX];  // host mirror (table upload / emulation)
#endif

template <int SET, int KIND> QT_HD TwPair uni_tw(int k) {
#if defined(__CUDA_ARCH__)
    return c_uni[SET][KIND][k];
#else
    return h_uni[SET][KIND][k];
#endif
}

QT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
QT_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

struct alignas(16) U4 { uint32_t x, y, z, w; };  // 128-bit shared-memory access unit

template <int SET> struct Tile {
    using C = Cfg<SET>;
    static constexpr uint32_t N = C::N, Q = C::Q, LOGN = C::LOGN, E = C::E, LOGE = C::LOGE;
    static constexpr uint32_t PPW = C::PPW, LPP = C::LPP, LB1 = C::LB1, LB2 = C::LB2;
    static constexpr uint32_t BLOCKS = N / E, SLOT_PAIRS = C::SLOT_PAIRS;
    static constexpr uint32_t G0 = E >> LB2;  // groups per thread at the first cols level
    static constexpr bool LAZY = C::LAZY;
    static constexpr uint32_t TWO_Q = 2 * Q;

    // ---- lazy-reduction budget (all bounds in units of q) -------------------------------------
    static constexpr uint32_t QCAP = (uint32_t)(0xFFFFFFFFull / Q);  // values < QCAP*q fit 32 bits
    static constexpr uint32_t FWD_BOUND = LAZY ? (2 * LOGN + 1) : 4;  // forward output < FWD_BOUND*q
    // largest power of two MID with MID * 2^LB1 <= QCAP: bound allowed when the rows inverse starts
    static QT_CONSTEXPR_HD uint32_t mid_bound() {
        uint32_t m = 1;
        while ((uint64_t)(2 * m) << LB1 <= QCAP) m *= 2;
        return m;
    }
    static constexpr uint32_t MID = LAZY ? mid_bound() : 2;
    static_assert(!LAZY || (uint64_t)FWD_BOUND * FWD_BOUND * Q < (1ull << 32),
                  "pointwise Montgomery needs a*b < q*2^32");
    static_assert(!LAZY || ((uint64_t)(2u << LB2)) <= QCAP, "cols inverse must fit 32 bits");
    static_assert(!LAZY || MID >= 2, "no room for the lazy inverse");
    // bound of register r after the cols inverse pass (inputs < 2q, a-path doubles, b-path -> 2q)
    static QT_CONSTEXPR_HD uint32_t bound_after_cols_inverse(uint32_t r) {
        uint32_t low = r & ((1u << LB2) - 1);
        if (low == 0) return 2u << LB2;
        uint32_t h = 0;
        while ((low >> (h + 1)) != 0) h++;
        return 2u << (LB2 - 1 - h);
    }

    // ---- modular arithmetic ---------------------------------------------------------------------
    // y*w mod q for ANY 32-bit y, result in [0,2q)   (Shoup / Harvey)
    static QT_HD uint32_t mul_shoup(uint32_t y, TwPair t) { return y * t.w - mulhi32(y, t.ws) * Q; }
    // a*b*2^-32 mod q, result in [0,2q); requires a*b < q*2^32
    static QT_HD uint32_t mul_mont(uint32_t a, uint32_t b) {
        uint64_t t = (uint64_t)a * b;
        uint32_t m = (uint32_t)t * C::QINV_NEG;
        return (uint32_t)((t + (uint64_t)m * Q) >> 32);
    }
    static QT_HD uint32_t fold2q(uint32_t a) { return a - mulhi32(a, C::MU32) * Q; }  // any a -> [0,2q)
    static QT_HD uint32_t csub(uint32_t a, uint32_t m) { return umin32(a, a - m); }    // [0,2m) -> [0,m)

    // forward (Cooley-Tukey) butterfly: (x, y) -> (x + w y, x - w y)
    static QT_HD void ct(uint32_t& x, uint32_t& y, TwPair t) {
        uint32_t wy = mul_shoup(y, t);
        uint32_t xx = LAZY ? x : csub(x, TWO_Q);
        x = xx + wy;
        y = xx - wy + TWO_Q;
    }
    // inverse (Gentleman-Sande) butterfly: (a, b) -> (a + b, (a - b) w); inputs < bound*q
    static QT_HD void gs(uint32_t& a, uint32_t& b, TwPair t, uint32_t bound) {
        uint32_t s = a + b;
        uint32_t d = a - b + bound * Q;
        a = LAZY ? s : csub(s, TWO_Q);
        b = mul_shoup(d, t);
    }

    // ---- transforms on one lane's registers ---------------------------------------------------
    // rows layout, forward levels 0..LB1-1 (register distance E/2 .. 1)
    static QT_HD void fwd_rows(uint32_t (&v)[E]) {
#pragma unroll
        for (uint32_t l = 0; l < LB1; l++) {
            const uint32_t half = E >> (l + 1);
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {  // flat butterfly index: constant trip count
                const uint32_t g = i / half, j = i % half;
                ct(v[2 * g * half + j], v[2 * g * half + j + half], uni_tw<SET, UNI_FWD>((1u << l) + g));
            }
        }
    }

    static QT_HD TwPair lane_slot(const TwQuad* tw, uint32_t slot) {
        const TwQuad qd = tw[(size_t)(slot >> 1) * BLOCKS];
        return (slot & 1) ? TwPair{qd.w1, qd.ws1} : TwPair{qd.w0, qd.ws0};
    }

    // cols layout, forward levels LB1..LOGN-1 (register distance N>>(l+1)); tw = &lane_fwd[block]
    static QT_HD void fwd_cols(uint32_t (&v)[E], const TwQuad* tw) {
#pragma unroll
        for (uint32_t k = 0; k < LB2; k++) {
            const uint32_t half = (E >> 1) >> (k + LB1 + LOGE - LOGN);  // N >> (l+1), l = LB1 + k
            const uint32_t G = E / (2 * half);
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t g = i / half, j = i % half;
                ct(v[2 * g * half + j], v[2 * g * half + j + half], lane_slot(tw, G - G0 + g));
            }
        }
    }

    // cols layout, inverse levels LOGN-1..LB1; inputs < 2q.  Ends with the mid-transform fold.
    // No inverse table: zeta[k]^-1 = -zeta[k'] with k' the mirror of k inside its level
    // (psi^-e = -psi^(n-e)), so (a-b)*zeta^-1 = (b-a)*zeta[k'].  The mirror of lane-block j, group g
    // is lane-block BLOCKS-1-j, group G-1-g: `tw_mirror` = &lane_fwd[BLOCKS-1-block].
    static QT_HD void inv_cols(uint32_t (&v)[E], const TwQuad* tw_mirror) {
#pragma unroll
        for (uint32_t k = 0; k < LB2; k++) {
            const uint32_t half = 1u << k;
            const uint32_t G = E / (2 * half);
            const uint32_t bound = LAZY ? (2u << k) : 2u;
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t g = i / half, j = i % half;
                uint32_t& a = v[2 * g * half + j];
                uint32_t& b = v[2 * g * half + j + half];
                const TwPair t = lane_slot(tw_mirror, G - G0 + (G - 1 - g));
                const uint32_t s_ = a + b, d = b - a + bound * Q;
                a = LAZY ? s_ : csub(s_, TWO_Q);
                b = mul_shoup(d, t);
            }
        }
        if (LAZY) {
#pragma unroll
            for (uint32_t r = 0; r < E; r++)
                if (bound_after_cols_inverse(r) > MID) v[r] = fold2q(v[r]);
        }
    }

    // rows layout, inverse levels LB1-1..0, output scale K folded into the last level, canonical
    template <int KIND> static QT_HD void inv_rows(uint32_t (&v)[E]) {
#pragma unroll
        for (uint32_t k = 0; k < LB1; k++) {
            const uint32_t l = LB1 - 1 - k;
            const uint32_t half = 1u << k;
            const uint32_t bound = LAZY ? (MID << k) : 2u;
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t g = i / half, j = i % half;
                uint32_t& a = v[2 * g * half + j];
                uint32_t& b = v[2 * g * half + j + half];
                const TwPair t = uni_tw<SET, KIND>((1u << l) + g);
                if (l != 0) {
                    gs(a, b, t, bound);
                } else {  // last level: both outputs are multiplied (K resp. K*zeta^-1)
                    const uint32_t s = a + b, d = a - b + bound * Q;
                    a = csub(mul_shoup(s, uni_tw<SET, KIND>(0)), Q);
                    b = csub(mul_shoup(d, t), Q);
                }
            }
        }
    }

    // NTT-domain product of two forward outputs (each < FWD_BOUND*q): a*b*2^-32 mod q in [0,2q)
    static QT_HD void pointwise_mont(uint32_t (&a)[E], const uint32_t (&b)[E]) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++) {
            if (LAZY) a[r] = mul_mont(a[r], b[r]);
            else a[r] = mul_mont(csub(a[r], TWO_Q), csub(b[r], TWO_Q));
        }
    }

    // same, the second operand read back from a stash written with sts_cols (keeps it out of registers)
    static QT_HD void pointwise_mont_stash(uint32_t (&a)[E], const uint32_t* stash, uint32_t lane) {
#pragma unroll
        for (uint32_t c = 0; c < E / 4; c++) {
            const U4 u = *reinterpret_cast<const U4*>(stash + swz(E * lane + 4 * c));
            const uint32_t b[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                if (LAZY) a[4 * c + k] = mul_mont(a[4 * c + k], b[k]);
                else a[4 * c + k] = mul_mont(csub(a[4 * c + k], TWO_Q), csub(b[k], TWO_Q));
            }
        }
    }

    // forward output -> canonical [0,q) (only the unfused forward entry point needs it)
    static QT_HD void canon_fwd(uint32_t (&v)[E]) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++)
            v[r] = LAZY ? csub(fold2q(v[r]), Q) : csub(csub(v[r], TWO_Q), Q);
    }

    // ---- data movement ----------------------------------------------------------------------------
    // tile word of register r in the rows layout
    static QT_HD uint32_t row_off(uint32_t lane, uint32_t r) {
        return (lane / LPP) * N + (lane % LPP) + LPP * r;
    }
    // bank swizzle of the transposition buffer: keeps 4-word groups intact, conflict-free for both
    // the 32-bit rows pattern and the 128-bit cols pattern
    static QT_HD uint32_t swz(uint32_t off) {
        uint32_t s = off ^ (((off >> LOGE) & 7u) << 2);
        if (PPW == 2) s ^= ((off >> 9) & 1u) << 4;
        return s;
    }
    static QT_HD void load_rows(uint32_t (&v)[E], const uint32_t* g_tile, uint32_t lane, bool valid) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++) v[r] = valid ? g_tile[row_off(lane, r)] : 0u;
    }
    static QT_HD void store_rows(const uint32_t (&v)[E], uint32_t* g_tile, uint32_t lane, bool valid) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++)
            if (valid) g_tile[row_off(lane, r)] = v[r];
    }
    static QT_HD void sts_rows(const uint32_t (&v)[E], uint32_t* buf, uint32_t lane) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++) buf[swz(row_off(lane, r))] = v[r];
    }
    static QT_HD void lds_rows(uint32_t (&v)[E], const uint32_t* buf, uint32_t lane) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++) v[r] = buf[swz(row_off(lane, r))];
    }
    static QT_HD void sts_cols(const uint32_t (&v)[E], uint32_t* buf, uint32_t lane) {
#pragma unroll
        for (uint32_t c = 0; c < E / 4; c++) {
            U4 u{v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]};
            *reinterpret_cast<U4*>(buf + swz(E * lane + 4 * c)) = u;
        }
    }
    static QT_HD void lds_cols(uint32_t (&v)[E], const uint32_t* buf, uint32_t lane) {
#pragma unroll
        for (uint32_t c = 0; c < E / 4; c++) {
            const U4 u = *reinterpret_cast<const U4*>(buf + swz(E * lane + 4 * c));
            v[4 * c] = u.x; v[4 * c + 1] = u.y; v[4 * c + 2] = u.z; v[4 * c + 3] = u.w;
        }
    }
};

}  // namespace qt
