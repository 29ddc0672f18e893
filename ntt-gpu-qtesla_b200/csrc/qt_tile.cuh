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
//  * Forward: merged-psi Cooley-Tukey, zeta[k] = psi^brv(k).  The output at position i is
//    x(psi^(2 brv(i)+1)) — bit-for-bit what the reference's Phi-scale + radix2NTTGS (NTT.cu:1866-1876)
//    leaves — with no psi pass and no bit-reverse pass.  Inverse: cyclic decimation-in-time + invPhi
//    scale (small moduli) or merged Gentleman-Sande (29/30-bit moduli); both consume that order.
//  * Two register-resident passes per transform, joined by one shared-memory transposition:
//      "rows" layout : register r of lane (p,j) holds coefficient j + LPP*r of polynomial p
//                      -> the log2(E) levels with the largest strides are thread-local and their
//                         twiddles are identical for all lanes (constant-bank operands);
//      "cols" layout : register r of lane L holds tile word E*L + r (E consecutive coefficients)
//                      -> the remaining levels are thread-local; twiddles are per lane (shared memory).
//  * Modular arithmetic: Shoup multiplication by precomputed constants (1 mul.hi + 2 mul.lo).
//    q < 2^25 ("LAZY"): signed residues, no correction anywhere; the Cooley-Tukey butterfly is
//    3 multiply-pipe instructions + 1 add (the sum rides on the multiply-add), and the inverse is a
//    cyclic decimation-in-time transform (also Cooley-Tukey butterflies, 31 of 160 per thread free of
//    multiplications) followed by the reference's own invPhi scale.  29/30-bit moduli: Harvey's
//    [0,4q) butterflies, merged Gentleman-Sande inverse.  The single final store is canonical.
//
// Everything here is __host__ __device__ so that tests/emu can execute the identical index
// arithmetic and 32-bit wrap-around behaviour lane by lane on a CPU (test infrastructure only).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <utility>

#include "qt_params.h"
#include "qt_tables.h"

#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#else
#define QT_HD inline
#endif

namespace qt {

#if defined(__CUDACC__)
// Uniform twiddles of the rows pass.  Indexed only with compile-time constants after unrolling,
// so they are consumed as constant-bank operands of IMAD (no load instruction).
__constant__ TwPair c_uni[NUM_TILE_SETS][UNI_KINDS][UNI_MAX];
// uniform twiddles of the FP64-quotient kernels (LAZY sets; qt_tables.h): w in [0, q) and W = w / q
__constant__ double c_uniW[NUM_SETS][UNI_KINDS][UNI_MAX];
__constant__ uint32_t c_uniU[NUM_SETS][UNI_KINDS][UNI_MAX];
#endif
#if !defined(__CUDA_ARCH__)
extern TwPair h_uni[NUM_TILE_SETS][UNI_KINDS][UNI_MAX];  // host mirror (table upload / emulation)
extern double h_uniW[NUM_SETS][UNI_KINDS][UNI_MAX];
extern uint32_t h_uniU[NUM_SETS][UNI_KINDS][UNI_MAX];
#endif

template <int SET, int KIND> QT_HD TwPair uni_tw(int k) {
#if defined(__CUDA_ARCH__)
    return c_uni[SET][KIND][k];
#else
    return h_uni[SET][KIND][k];
#endif
}

template <int SET, int KIND> QT_HD double uni_W(int k) {
#if defined(__CUDA_ARCH__)
    return c_uniW[SET][KIND][k];
#else
    return h_uniW[SET][KIND][k];
#endif
}

template <int SET, int KIND> QT_HD uint32_t uni_U(int k) {
#if defined(__CUDA_ARCH__)
    return c_uniU[SET][KIND][k];
#else
    return h_uniU[SET][KIND][k];
#endif
}

QT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
QT_HD int32_t mulhi32s(uint32_t a, uint32_t b) {  // signed high word of two two's-complement patterns
#if defined(__CUDA_ARCH__)
    return __mulhi((int)a, (int)b);
#else
    return (int32_t)(((int64_t)(int32_t)a * (int64_t)(int32_t)b) >> 32);
#endif
}
QT_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

struct alignas(16) U4 { uint32_t x, y, z, w; };  // 128-bit shared-memory access unit

// ---- static range plan of the Harvey path when the modulus leaves head-room --------------------------------------
// Harvey's butterfly keeps everything in [0, 4q) by one conditional subtraction per butterfly (2 of its 4 ALU
// instructions).  For qTESLA-p-I 2^32 / q = 12.5, so a value may grow to 12 q before it has to be touched, and WHICH
// registers grow is a function of the register index only: the x input of a Cooley-Tukey butterfly gains 2q per
// level, the y input is reset to [0, 2q) by its Shoup multiplication; in the Gentleman-Sande inverse the sum output
// adds the two bounds and the product output is reset.  The plan below runs that bookkeeping at COMPILE TIME over the
// unrolled butterfly network of one thread and records where a conditional subtraction is really needed (and by how
// many q): 276 instead of 576 per product and thread for n = 1024.  Bounds are in units of q ("b" means value < b q);
// at the shared-memory transpositions the worst register's bound is taken for all (a thread's registers there come
// from different registers of 32 other threads).
template <uint32_t E, uint32_t LB1, uint32_t LB2, uint32_t CAP> struct HarveyPlan {
    struct P {
        uint8_t fr[LB1][E / 2];     // forward, rows levels: csub modulus (x q) applied to x before butterfly i, 0 = none
        uint8_t fc[LB2][E / 2][3];  // forward, cols levels: a chain of up to three (the last level reduces x to [0, 2q), see make())
        uint8_t fout[E];            // bounds of the forward output
        uint8_t pwa[E][3], pwb[E][3];  // pointwise product of two forward outputs: csub chains of either operand
        uint8_t cf[E][4];           // forward output -> canonical: csub chain down to 1
        uint8_t ic[LB2][E / 2][2];  // inverse, cols levels: csub modulus of a / of b before the butterfly
        uint8_t ick[LB2][E / 2];    // ... bound of a when the difference b - a + K q is formed
        uint8_t ir[LB1][E / 2][2];  // inverse, rows levels
        uint8_t irk[LB1][E / 2];    // ... bound of b (difference a - b + K q)
        uint8_t rows_out, icols_out, ok;
    };
    static QT_CHD uint32_t halve(uint32_t b) { return (b + 1) / 2; }  // csub(v, m q) with m = ceil(b / 2): [0, b q) -> [0, m q)
    static QT_CHD P make() {
        P p{};
        p.ok = 1;
        uint32_t b[E] = {};
        for (uint32_t r = 0; r < E; r++) b[r] = 1;  // canonical input
        for (uint32_t l = 0; l < LB1; l++) {
            const uint32_t half = E >> (l + 1);
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t x = 2 * (i / half) * half + i % half, y = x + half;
                if (b[x] + 2 > CAP) { p.fr[l][i] = (uint8_t)halve(b[x]); b[x] = halve(b[x]); }
                if (b[x] + 2 > CAP) p.ok = 0;
                b[x] += 2; b[y] = b[x];
            }
        }
        uint32_t u = 0;
        for (uint32_t r = 0; r < E; r++) u = b[r] > u ? b[r] : u;
        p.rows_out = (uint8_t)u;
        for (uint32_t r = 0; r < E; r++) b[r] = u;
        for (uint32_t k = 0; k < LB2; k++) {
            const uint32_t half = (E >> 1) >> k;
            // the LAST level brings x down to [0, 2q) first: both outputs are then below 4q and the pointwise product
            // needs one correction per pair instead of four (64 + 32 instead of 32 + 128 corrections per thread)
            const uint32_t lim = (k + 1 == LB2) ? 2 : CAP - 2;
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t x = 2 * (i / half) * half + i % half, y = x + half;
                uint32_t n = 0;
                while (b[x] > lim && n < 3) { p.fc[k][i][n++] = (uint8_t)halve(b[x]); b[x] = halve(b[x]); }
                if (b[x] > lim) p.ok = 0;
                b[x] += 2; b[y] = b[x];
            }
        }
        for (uint32_t r = 0; r < E; r++) {
            p.fout[r] = (uint8_t)b[r];
            uint32_t ba = b[r], bb = b[r], na = 0, nb = 0;  // a b < q 2^32  <=  ba bb <= CAP
            while (ba * bb > CAP) {
                if (ba >= bb) { ba = halve(ba); if (na < 3) p.pwa[r][na] = (uint8_t)ba; na++; }
                else { bb = halve(bb); if (nb < 3) p.pwb[r][nb] = (uint8_t)bb; nb++; }
            }
            if (na > 3 || nb > 3) p.ok = 0;
            uint32_t c = b[r], nc = 0;
            while (c > 1) { c = halve(c); if (nc < 4) p.cf[r][nc] = (uint8_t)c; nc++; }
            if (nc > 4) p.ok = 0;
        }
        for (uint32_t r = 0; r < E; r++) b[r] = 2;  // Montgomery product / canonical NTT-domain input
        for (uint32_t k = 0; k < LB2; k++) {
            const uint32_t half = 1u << k;
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t a = 2 * (i / half) * half + i % half, c = a + half;
                if (b[a] + b[c] > CAP) {
                    if (b[a] >= b[c]) { p.ic[k][i][0] = (uint8_t)halve(b[a]); b[a] = halve(b[a]); }
                    else { p.ic[k][i][1] = (uint8_t)halve(b[c]); b[c] = halve(b[c]); }
                }
                if (b[a] + b[c] > CAP) {  // second step on the other / the still larger operand
                    if (b[a] >= b[c] && !p.ic[k][i][0]) { p.ic[k][i][0] = (uint8_t)halve(b[a]); b[a] = halve(b[a]); }
                    else if (!p.ic[k][i][1]) { p.ic[k][i][1] = (uint8_t)halve(b[c]); b[c] = halve(b[c]); }
                }
                if (b[a] + b[c] > CAP) p.ok = 0;
                p.ick[k][i] = (uint8_t)b[a];
                b[a] += b[c]; b[c] = 2;
            }
        }
        u = 0;
        for (uint32_t r = 0; r < E; r++) u = b[r] > u ? b[r] : u;
        p.icols_out = (uint8_t)u;
        for (uint32_t r = 0; r < E; r++) b[r] = u;
        for (uint32_t k = 0; k < LB1; k++) {
            const uint32_t half = 1u << k;
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t a = 2 * (i / half) * half + i % half, c = a + half;
                if (b[a] + b[c] > CAP) {
                    if (b[a] >= b[c]) { p.ir[k][i][0] = (uint8_t)halve(b[a]); b[a] = halve(b[a]); }
                    else { p.ir[k][i][1] = (uint8_t)halve(b[c]); b[c] = halve(b[c]); }
                }
                if (b[a] + b[c] > CAP) {
                    if (b[a] >= b[c] && !p.ir[k][i][0]) { p.ir[k][i][0] = (uint8_t)halve(b[a]); b[a] = halve(b[a]); }
                    else if (!p.ir[k][i][1]) { p.ir[k][i][1] = (uint8_t)halve(b[c]); b[c] = halve(b[c]); }
                }
                if (b[a] + b[c] > CAP) p.ok = 0;
                p.irk[k][i] = (uint8_t)b[c];
                b[a] += b[c]; b[c] = 2;  // (the last level multiplies both outputs: any bound is fine for a Shoup product)
            }
        }
        return p;
    }
};


// SHIFT_OK = false switches the shift-add form of hi*q off (see ct(); it costs the natural-order inverse
// kernel 12 % while every other kernel gains or is unaffected).
template <int SET, bool SHIFT_OK = true> struct Tile {
    using C = Cfg<SET>;
    static constexpr uint32_t N = C::N, Q = C::Q, LOGN = C::LOGN, E = C::E, LOGE = C::LOGE;
    static constexpr uint32_t PPW = C::PPW, LPP = C::LPP, LB1 = C::LB1, LB2 = C::LB2;
    static constexpr uint32_t BLOCKS = N / E, SLOT_PAIRS = C::SLOT_PAIRS;
    static constexpr uint32_t G0 = E >> LB2;  // groups per thread at the first cols level
    static constexpr bool LAZY = C::LAZY;
    static constexpr uint32_t TWO_Q = 2 * Q;

    // ---- shared-memory table block of one kernel ---------------------------------------------------
    // [forward per-lane twiddles][LAZY only: inverse per-lane twiddles][LAZY only: output scale]
    static constexpr uint32_t TW_QUADS = SLOT_PAIRS * BLOCKS;
    static constexpr uint32_t INV_SLOTS = E - 1;                      // rows levels of the DIT inverse: 1+2+..+E/2
    static constexpr uint32_t INV_QUADS = LAZY ? ((INV_SLOTS + 1) / 2) * LPP : 0;
    static constexpr uint32_t SCALE_QUADS = LAZY ? (E / 2) * LPP : 0;
    static constexpr uint32_t TABLE_QUADS = C::HALVES * TW_QUADS + INV_QUADS + SCALE_QUADS;
    struct LanePtrs { const TwQuad *fwd, *inv, *scale; };
    static QT_HD LanePtrs lane_ptrs(const TwQuad* tab, uint32_t lane) {
        if (LAZY) return LanePtrs{tab + lane % BLOCKS, tab + TW_QUADS + lane % LPP, tab + TW_QUADS + INV_QUADS + lane % LPP};
        return LanePtrs{tab + lane % BLOCKS, tab + (BLOCKS - 1 - lane % BLOCKS), nullptr};  // Harvey sets: mirrored table
    }

    // ---- range bookkeeping ------------------------------------------------------------------------------
    static constexpr uint32_t QCAP = (uint32_t)(0xFFFFFFFFull / Q);  // values < QCAP*q fit 32 bits
    // LAZY ("signed lazy", q < 2^25): values are two's-complement residues.  A signed Shoup product lies in
    // [-q/2, 3q/2), so a Cooley-Tukey level moves |v| by at most 1.5 q and NO correction is ever needed:
    //   forward : |v| < (1 + 1.5 log2 n) q <= 17.5 q
    //   inverse : decimation-in-time; its first LB2 levels contain the multiplication-free butterflies
    //             (twiddle 1), which may double: |v| < 2^LB2 q, then + 1.5 q per remaining level
    static_assert(!LAZY || (uint64_t)(2 + 3 * LOGN) * Q < (1ull << 32), "forward range");           // 2*(1+1.5 logn) q < 2^32
    static_assert(!LAZY || ((uint64_t)(4u << LB2) + 3 * LB1) * Q < (1ull << 32), "inverse range");  // (2*2^LB2 + 1.5 LB1) q < 2^31

    // ---- modular arithmetic ---------------------------------------------------------------------
    // y*w mod q for ANY 32-bit y, result in [0,2q)   (Shoup / Harvey), w with floor(w*2^32/q)
    static QT_HD uint32_t mul_shoup(uint32_t y, TwPair t) { return y * t.w - mulhi32(y, t.ws) * Q; }
    // a*b*2^-32 mod q, result in [0,2q); requires a*b < q*2^32
    static QT_HD uint32_t mul_mont(uint32_t a, uint32_t b) {
        uint64_t t = (uint64_t)a * b;
        uint32_t m = (uint32_t)t * C::QINV_NEG;
        return (uint32_t)((t + (uint64_t)m * Q) >> 32);
    }
    static QT_HD uint32_t fold2q(uint32_t a) { return a - mulhi32(a, C::MU32) * Q; }  // any a -> [0,2q)
    static QT_HD uint32_t csub(uint32_t a, uint32_t m) { return umin32(a, a - m); }    // [0,2m) -> [0,m)

    // signed variants: operands are two's-complement bit patterns; t = (w centred in (-q/2, q/2],
    // floor(w*2^32/q) as a signed word).  Result in [-q/2, 3q/2) for any |y| < 2^31.
    static QT_HD uint32_t smul_shoup(uint32_t y, TwPair t) { return y * t.w - (uint32_t)mulhi32s(y, t.ws) * Q; }
    // signed Montgomery: a*b*2^-32 mod q, |result| < q/2 + |a*b|/2^32 + 1
    static QT_HD uint32_t smul_mont(uint32_t a, uint32_t b) {
        const int64_t t = (int64_t)(int32_t)a * (int32_t)b;
        const uint32_t m = (uint32_t)t * (0u - C::QINV_NEG);  // lo(t) * q^-1: t - m*q has a zero low word
        return (uint32_t)(t >> 32) - (uint32_t)mulhi32s(m, Q);
    }
    // [-q/2, 3q/2) -> [0, q)
    static QT_HD uint32_t scanon(uint32_t r) { return csub(umin32(r, r + Q), Q); }

    // forward (Cooley-Tukey) butterfly: (x, y) -> (x + w y, x - w y)
    // q = 2^23 + 2^14 + 1 (qTESLA-III): hi*q as two shift-adds (LEA).  Moves one multiply-add of a butterfly
    // from the multiply pipe (the binding unit) to the ALU at the price of two more issue slots, so it pays
    // only for a fraction of the butterflies: every QT_SHIFT_MOD-th one (0 = never).  Measured at n=1024,
    // batch 65 536: never 238.9, every 2nd < 238, 3rd 238.9, 4th 240.9, 5th 240.2, 8th 240.5 M polymul/s.
    // (Inline PTX: written in C the compiler folds the shifts back into one IMAD.)
    // Re-measured on the final kernel (uniform warp index, run r02I/r02J): never 244.5, every 2nd 240.3, 3rd 248.2, 4th 246.4,
    // 5th 248.0, 6th 247.6 M polymul/s; the single transforms prefer 4 or 5 (3rd: -3 %), the cached product 5 -> 5.
#ifndef QT_SHIFT_MOD
#define QT_SHIFT_MOD 5
#endif
    static constexpr bool SHIFT_Q = SHIFT_OK && (Q == (1u << 23) + (1u << 14) + 1u) && (QT_SHIFT_MOD != 0);
    static QT_HD void ct(uint32_t& x, uint32_t& y, TwPair t, uint32_t idx = 1) {
        if (LAZY) {
            // 3 multiply-pipe instructions + ONE add: the sum rides on the multiply-add's addend,
            // the difference is 2x - x'
            const uint32_t hi = (uint32_t)mulhi32s(y, t.ws);
            const uint32_t u = y * t.w + x;
            uint32_t xn;
            if (SHIFT_Q && idx % (QT_SHIFT_MOD ? QT_SHIFT_MOD : 1) == 0) {
#if defined(__CUDA_ARCH__)
                uint32_t m;
                asm("{\n\t.reg .u32 a, b;\n\tshl.b32 a, %1, 14;\n\tadd.u32 a, a, %1;\n\tshl.b32 b, %1, 23;\n\tadd.u32 %0, a, b;\n\t}" : "=r"(m) : "r"(hi));
#else
                const uint32_t m = ((hi << 14) + hi) + (hi << 23);
#endif
                xn = u - m;
            } else {
                xn = u - hi * Q;
            }
            y = x + x - xn;
            x = xn;
        } else {
#ifndef QT_HARVEY_FUSED_ADD
#define QT_HARVEY_FUSED_ADD 1  // run r02q: n=2048 two-warp kernel 87.8 -> 90.6, one-warp split kernel 84.4 -> 87.9 M polymul/s
#endif
            const uint32_t xx = csub(x, TWO_Q);
            if (QT_HARVEY_FUSED_ADD) {
                // the sum rides on the multiply-add (as in the lazy path), the difference is 2 xx - x' + 2q: the separate
                // addition of x' — which ptxas likes to issue on the MULTIPLY pipe as IMAD.IADD — disappears
                const uint32_t hi = mulhi32(y, t.ws);
                const uint32_t u = y * t.w + xx;
                const uint32_t xn = u - hi * Q;
                y = xx + xx - xn + TWO_Q;
                x = xn;
            } else {
                const uint32_t wy = mul_shoup(y, t);
                x = xx + wy;
                y = xx - wy + TWO_Q;
            }
        }
    }
    // inverse (Gentleman-Sande) butterfly of the Harvey sets: (a, b) -> (a + b, (b - a) w'), in/out < 2q
    static QT_HD void gs(uint32_t& a, uint32_t& b, TwPair t_mirror) {
        const uint32_t s_ = a + b, d = b - a + TWO_Q;
        a = csub(s_, TWO_Q);
        b = mul_shoup(d, t_mirror);
    }

    // ---- Harvey path with a static range plan (HarveyPlan above; qTESLA-p-I) ---------------------------------
    static constexpr bool HLAZY = !LAZY && !C::SPLIT && QCAP >= 8;
    static_assert(!HLAZY || (LB1 + LOGE == LOGN && PPW == 1), "range plan: written for one polynomial per warp");
    using HPlan = HarveyPlan<E, LB1, LB2, QCAP>;
    static_assert(!HLAZY || HPlan::make().ok == 1, "range plan: a value would leave 32 bits");
    template <uint32_t M_> static QT_HD uint32_t csub_q(uint32_t a) {  // [0, 2 M q) -> [0, M q); M = 0: nothing
        if constexpr (M_ == 0) return a;
        else return umin32(a, a - M_ * Q);
    }
    template <uint32_t L, uint32_t I> static QT_HD void h_fr(uint32_t (&v)[E], uint32_t ub) {
        constexpr uint32_t half = E >> (L + 1), g = I / half, X = 2 * g * half + I % half, Y = X + half;
        constexpr uint32_t m = HPlan::make().fr[L][I];
        const uint32_t wy = mul_shoup(v[Y], uni_tw<SET, UNI_FWD>(ub + (1u << L) + g)), xx = csub_q<m>(v[X]);
        v[X] = xx + wy;
        v[Y] = xx - wy + TWO_Q;
    }
    template <uint32_t K, uint32_t I> static QT_HD void h_fc(uint32_t (&v)[E], const TwQuad* tw) {
        constexpr uint32_t half = (E >> 1) >> K, G = E / (2 * half), g = I / half, X = 2 * g * half + I % half, Y = X + half;
        constexpr auto pl = HPlan::make();
        const uint32_t wy = mul_shoup(v[Y], lane_slot(tw, G - G0 + g, BLOCKS));
        const uint32_t xx = csub_q<pl.fc[K][I][2]>(csub_q<pl.fc[K][I][1]>(csub_q<pl.fc[K][I][0]>(v[X])));
        v[X] = xx + wy;
        v[Y] = xx - wy + TWO_Q;
    }
    template <uint32_t K, uint32_t I> static QT_HD void h_ic(uint32_t (&v)[E], const TwQuad* tw) {
        constexpr uint32_t half = 1u << K, G = E / (2 * half), g = I / half, A = 2 * g * half + I % half, B = A + half;
        constexpr auto pl = HPlan::make();
        constexpr uint32_t ma = pl.ic[K][I][0], mb = pl.ic[K][I][1], kq = pl.ick[K][I];
        const uint32_t a = csub_q<ma>(v[A]), b = csub_q<mb>(v[B]);
        v[A] = a + b;
        v[B] = mul_shoup(b - a + kq * Q, lane_slot(tw, G - G0 + (G - 1 - g), BLOCKS));
    }
    template <int KIND, uint32_t K, uint32_t I> static QT_HD void h_ir(uint32_t (&v)[E], uint32_t ub) {
        constexpr uint32_t l = LB1 - 1 - K, half = 1u << K, g = I / half, A = 2 * g * half + I % half, B = A + half;
        constexpr auto pl = HPlan::make();
        constexpr uint32_t ma = pl.ir[K][I][0], mb = pl.ir[K][I][1], kq = pl.irk[K][I];
        const uint32_t a = csub_q<ma>(v[A]), b = csub_q<mb>(v[B]);
        const TwPair t = uni_tw<SET, KIND>(ub + (1u << l) + g);
        const uint32_t s_ = a + b, d = a - b + kq * Q;
        if constexpr (l != 0) {
            v[A] = s_;
            v[B] = mul_shoup(d, t);
        } else {  // last level: both outputs are multiplied (K resp. K*zeta^-1), canonical
            v[A] = csub(mul_shoup(s_, uni_tw<SET, KIND>(0)), Q);
            v[B] = csub(mul_shoup(d, t), Q);
        }
    }
    template <uint32_t R_> static QT_HD uint32_t h_pw(uint32_t a, uint32_t b) {  // two forward outputs
        constexpr auto pl = HPlan::make();
        a = csub_q<pl.pwa[R_][2]>(csub_q<pl.pwa[R_][1]>(csub_q<pl.pwa[R_][0]>(a)));
        b = csub_q<pl.pwb[R_][2]>(csub_q<pl.pwb[R_][1]>(csub_q<pl.pwb[R_][0]>(b)));
        return mul_mont(a, b);
    }
    template <uint32_t R_> static QT_HD uint32_t h_canon(uint32_t a) {
        constexpr auto pl = HPlan::make();
        return csub_q<pl.cf[R_][3]>(csub_q<pl.cf[R_][2]>(csub_q<pl.cf[R_][1]>(csub_q<pl.cf[R_][0]>(a))));
    }
    // (compile-time loops: the plan entries must be constant expressions, which loop variables are not)
    template <uint32_t L, uint32_t... Is> static QT_HD void h_fr_level(uint32_t (&v)[E], uint32_t ub, std::integer_sequence<uint32_t, Is...>) { (h_fr<L, Is>(v, ub), ...); }
    template <uint32_t... Ls> static QT_HD void h_fr_all(uint32_t (&v)[E], uint32_t ub, std::integer_sequence<uint32_t, Ls...>) { (h_fr_level<Ls>(v, ub, std::make_integer_sequence<uint32_t, E / 2>{}), ...); }
    template <uint32_t K, uint32_t... Is> static QT_HD void h_fc_level(uint32_t (&v)[E], const TwQuad* tw, std::integer_sequence<uint32_t, Is...>) { (h_fc<K, Is>(v, tw), ...); }
    template <uint32_t... Ks> static QT_HD void h_fc_all(uint32_t (&v)[E], const TwQuad* tw, std::integer_sequence<uint32_t, Ks...>) { (h_fc_level<Ks>(v, tw, std::make_integer_sequence<uint32_t, E / 2>{}), ...); }
    template <uint32_t K, uint32_t... Is> static QT_HD void h_ic_level(uint32_t (&v)[E], const TwQuad* tw, std::integer_sequence<uint32_t, Is...>) { (h_ic<K, Is>(v, tw), ...); }
    template <uint32_t... Ks> static QT_HD void h_ic_all(uint32_t (&v)[E], const TwQuad* tw, std::integer_sequence<uint32_t, Ks...>) { (h_ic_level<Ks>(v, tw, std::make_integer_sequence<uint32_t, E / 2>{}), ...); }
    template <int KIND, uint32_t K, uint32_t... Is> static QT_HD void h_ir_level(uint32_t (&v)[E], uint32_t ub, std::integer_sequence<uint32_t, Is...>) { (h_ir<KIND, K, Is>(v, ub), ...); }
    template <int KIND, uint32_t... Ks> static QT_HD void h_ir_all(uint32_t (&v)[E], uint32_t ub, std::integer_sequence<uint32_t, Ks...>) { (h_ir_level<KIND, Ks>(v, ub, std::make_integer_sequence<uint32_t, E / 2>{}), ...); }
    template <uint32_t... Rs> static QT_HD void h_pw_all(uint32_t (&a)[E], const uint32_t (&b)[E], std::integer_sequence<uint32_t, Rs...>) { ((a[Rs] = h_pw<Rs>(a[Rs], b[Rs])), ...); }
    template <uint32_t... Rs> static QT_HD void h_canon_all(uint32_t (&v)[E], std::integer_sequence<uint32_t, Rs...>) { ((v[Rs] = h_canon<Rs>(v[Rs])), ...); }
    template <uint32_t... Cs> static QT_HD void h_pw_stash(uint32_t (&a)[E], const uint32_t* stash, uint32_t lane, std::integer_sequence<uint32_t, Cs...>) {
        (([&] {
             const U4 u = *reinterpret_cast<const U4*>(stash + swz(E * lane + 4 * Cs));
             a[4 * Cs] = h_pw<4 * Cs>(a[4 * Cs], u.x);
             a[4 * Cs + 1] = h_pw<4 * Cs + 1>(a[4 * Cs + 1], u.y);
             a[4 * Cs + 2] = h_pw<4 * Cs + 2>(a[4 * Cs + 2], u.z);
             a[4 * Cs + 3] = h_pw<4 * Cs + 3>(a[4 * Cs + 3], u.w);
         }()),
         ...);
    }

    // ---- FP64-quotient ("DQ") arithmetic for the signed-lazy sets ------------------------------------------------
    // The Shoup butterfly above keeps the multiply pipe busy for 8 clocks per warp, 4 of them for the mul.hi that
    // estimates the quotient floor(y w / q).  B200's FP64 pipe runs beside the integer multiply-add at the same rate
    // (64 lanes/clk/SM each, profiles/ubench_r01s.json), so here the quotient comes from ONE FP64 multiplication:
    //     t = D(y) * (w / q),      D(y) = the double whose BIT PATTERN is {lo = y, hi = 0}: the DENORMAL y 2^-1074.
    // The product of a denormal and a number below 1 is a denormal again, i.e. it is rounded to a multiple of 2^-1074:
    // t's bit pattern is {rint(y w / q), 0} — an integer quotient without any conversion instruction, for every
    // UNSIGNED 32-bit y and twiddle w in [0, q) — and   y w - rint(y w / q) q   lies in [-q/2 - 1, q/2 + 1] (the
    // rounding of w/q to a double moves the product by < 2^-21).  Butterfly = DMUL + 2 mad.lo + 1 add: 4 clocks of the
    // multiply pipe and 2 of the FP64 pipe per warp (tools/dq_ubench.cu: exactness and rates).
    //  * unsigned y: every value v travels in OFFSET FORM v + DQ_OFF, DQ_OFF a multiple of q around 2^30: the same
    //    residue, never negative.  A butterfly keeps the form by itself (x' = y w + x - qe q inherits x's offset,
    //    y' = 2x - x' too); only additions of two values and the very first level have to mind it.
    //  * the FP64 unit reads a register PAIR.  A value is a P64 {lo, hi} whose hi is the zero high half of SOME earlier
    //    FP64 result: x' = mad(qe, -q, u) overwrites qe where the DMUL left it, so {x', hi(t)} is a pair with a zero high
    //    half by construction, and y' overwrites y in y's own pair.  Only the 16 y operands of a pass's first level
    //    need a seed (dq_seed).  Were the high halves written as constants, ptxas would re-create them with a MOV per
    //    butterfly; were they loaded, it copies them around.
    // Ranges (true values): forward |v| < q + LOGN (q/2 + 1); inverse < 2^LB2 q + LB1 (q/2 + 1); both < 40 q.
    static constexpr uint32_t DQ_OFF = 128u * Q;
    static_assert(!LAZY || ((128ull + 40) * Q < (1ull << 32) && (1u << LB2) + (LB1 + 1) / 2 + 1 < 40 && 1 + (LOGN + 1) / 2 < 40), "DQ offset form");
    // A value: the 64-bit register pair as ONE object (a double, so that the FP64 unit reads it where it lies); lo() is a
    // sub-register read, with_lo() the same pair with the low half replaced.
    struct P64 {
        double d;
        QT_HD uint32_t lo() const {
#if defined(__CUDA_ARCH__)
            return (uint32_t)__double2loint(d);
#else
            uint64_t b; memcpy(&b, &d, sizeof b); return (uint32_t)b;
#endif
        }
        QT_HD uint32_t hi() const {
#if defined(__CUDA_ARCH__)
            return (uint32_t)__double2hiint(d);
#else
            uint64_t b; memcpy(&b, &d, sizeof b); return (uint32_t)(b >> 32);
#endif
        }
        QT_HD P64 with_lo(uint32_t v) const { return make(v, hi()); }
        static QT_HD P64 make(uint32_t lo_, uint32_t hi_) {
#if defined(__CUDA_ARCH__)
            return P64{__hiloint2double((int)hi_, (int)lo_)};
#else
            const uint64_t b = ((uint64_t)hi_ << 32) | lo_;
            P64 r; memcpy(&r.d, &b, sizeof b); return r;
#endif
        }
    };
    // {rint(y W), 0} for W = w / q in [0, 1)
    static QT_HD P64 dq_quot(P64 y, double W) {
#if defined(__CUDA_ARCH__) && !defined(QT_DQ_PLAIN_MUL)
        // fma with a +0 addend, NOT a plain product: given mul.f64 followed by mad.lo(lo(t), -q, u), ptxas 12.9 emitted ONE
        // DMUL of W with a constant pair {-q, 4} per level and x' = y * lo(that) + u for every butterfly — it moved the integer
        // multiplication through the FP64 product as if D(y) W were linear in y.  Wrong results for qTESLA-I (run r02B,
        // tools/dq_ptxas_repro.cu: GPU vs host level by level); the fma form is compiled as written (DFMA Rt, Ry, W, RZ).
        double t;
        asm("fma.rn.f64 %0, %1, %2, 0d0000000000000000;" : "=d"(t) : "d"(y.d), "d"(W));
        return P64{t};
#else
        return P64{y.d * W};
#endif
    }
    // (x, y) -> (x + w y, x - w y); x in offset form or not (the outputs inherit it), y any unsigned representative in a pair
    static QT_HD void ct_dq(P64& X, P64& Y, uint32_t w, double W) {
        const P64 t = dq_quot(Y, W);
        const uint32_t x = X.lo(), y = Y.lo();
        const uint32_t u = y * w + x;
        const uint32_t xn = u - t.lo() * Q;
#ifndef QT_DQ_SWAP
#define QT_DQ_SWAP 0
#endif
        if (QT_DQ_SWAP) Y = X.with_lo(x + x - xn);  // y' over x, in x's pair (every register of a pass then needs a seed)
        else Y = Y.with_lo(x + x - xn);             // y' over y, in y's pair
        X = t.with_lo(xn);                          // x' over the quotient, in the pair the FP64 unit wrote
    }
    struct TwDQ { uint32_t w; double W; };
    // per-lane tables of the DQ kernel: w in [0, q) and W = w / q, two slots per access
    static QT_HD TwDQ lane_slot_dq(const TwU2* tw, const TwW2* twW, uint32_t slot, uint32_t stride) {
        const TwU2 qd = tw[(size_t)(slot >> 1) * stride];
        const TwW2 wd = twW[(size_t)(slot >> 1) * stride];
        return (slot & 1) ? TwDQ{qd.w1, wd.W1} : TwDQ{qd.w0, wd.W0};
    }
    struct LanePtrsDQ { const TwU2 *fwd, *inv, *scale; const TwW2 *fwdW, *invW, *scaleW; };
    static QT_HD LanePtrsDQ lane_ptrs_dq(const TwU2* tab, const TwW2* tabW, uint32_t lane) {
        return LanePtrsDQ{tab + lane % BLOCKS, tab + TW_QUADS + lane % LPP, tab + TW_QUADS + INV_QUADS + lane % LPP,
                          tabW + lane % BLOCKS, tabW + TW_QUADS + lane % LPP, tabW + TW_QUADS + INV_QUADS + lane % LPP};
    }
    // forward, rows layout; input: canonical coefficients in lo (NOT in offset form); hi of v[E/2 ..] seeded by the caller
    static QT_HD void fwd_rows_dq(P64 (&v)[E]) {
#pragma unroll
        for (uint32_t r = 0; r < E / 2; r++) v[r] = v[r].with_lo(v[r].lo() + DQ_OFF);  // the x inputs of level 0 carry the offset in
#pragma unroll
        for (uint32_t l = 0; l < LB1; l++) {
            const uint32_t half = E >> (l + 1);
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t g = i / half, j = i % half;
                ct_dq(v[2 * g * half + j], v[2 * g * half + j + half], uni_U<SET, UNI_FWD>((1u << l) + g), uni_W<SET, UNI_FWD>((1u << l) + g));
            }
        }
    }
    // forward, cols layout; input in offset form; hi of the first level's y operands seeded by the caller
    static QT_HD void fwd_cols_dq(P64 (&v)[E], const LanePtrsDQ& p) {
#pragma unroll
        for (uint32_t k = 0; k < LB2; k++) {
            const uint32_t half = (E >> 1) >> (k + LB1 + LOGE - LOGN);
            const uint32_t G = E / (2 * half);
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t g = i / half, j = i % half;
                const TwDQ t = lane_slot_dq(p.fwd, p.fwdW, G - G0 + g, BLOCKS);
                ct_dq(v[2 * g * half + j], v[2 * g * half + j + half], t.w, t.W);
            }
        }
    }
    // y operands of the first level of fwd_cols_dq / of inv_rows_dq: the registers the caller has to seed
    static QT_CHD bool dq_cols_first_y(uint32_t r) { return ((r / ((E >> 1) >> (LB1 + LOGE - LOGN))) & 1u) != 0; }
    static QT_CHD bool dq_rows_first_y(uint32_t r) { return (r & 1u) != 0; }
    // NTT-domain product with the stashed first operand (both in offset form): a b 2^-32, offset form
    static QT_HD void pointwise_dq_stash(P64 (&a)[E], const uint32_t* stash, uint32_t lane) {
#pragma unroll
        for (uint32_t c = 0; c < E / 4; c++) {
            const U4 u = *reinterpret_cast<const U4*>(stash + swz(E * lane + 4 * c));
            a[4 * c] = a[4 * c].with_lo(smul_mont(a[4 * c].lo() - DQ_OFF, u.x - DQ_OFF) + DQ_OFF);
            a[4 * c + 1] = a[4 * c + 1].with_lo(smul_mont(a[4 * c + 1].lo() - DQ_OFF, u.y - DQ_OFF) + DQ_OFF);
            a[4 * c + 2] = a[4 * c + 2].with_lo(smul_mont(a[4 * c + 2].lo() - DQ_OFF, u.z - DQ_OFF) + DQ_OFF);
            a[4 * c + 3] = a[4 * c + 3].with_lo(smul_mont(a[4 * c + 3].lo() - DQ_OFF, u.w - DQ_OFF) + DQ_OFF);
        }
    }
    // inverse, cols layout: cyclic decimation-in-time (see inv_cols); the multiplication-free butterflies re-centre the
    // offset.  Every register is still in the pair the forward cols pass left it in.
    static QT_HD void inv_cols_dq(P64 (&v)[E]) {
#pragma unroll
        for (uint32_t s_ = 0; s_ < LB2; s_++) {
            const uint32_t l = 1u << s_;
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t u = i / l, j = i % l;
                P64& x = v[2 * l * u + j];
                P64& y = v[2 * l * u + j + l];
                if (j == 0) {
                    const uint32_t a = x.lo(), b = y.lo();
                    x = x.with_lo(a + b - DQ_OFF);
                    y = y.with_lo(a - b + DQ_OFF);
                } else {
                    ct_dq(x, y, uni_U<SET, UNI_INV_PLAIN>(l + j), uni_W<SET, UNI_INV_PLAIN>(l + j));
                }
            }
        }
    }
    // inverse, rows layout, and the output scale; out canonical in [0, q); hi of the odd registers seeded by the caller
    static QT_HD void inv_rows_dq(P64 (&v)[E], uint32_t (&out)[E], const LanePtrsDQ& p) {
#pragma unroll
        for (uint32_t k = 0; k < LB1; k++) {
            const uint32_t G = 1u << k;
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t u = i / G, g = i % G;
                const TwDQ t = lane_slot_dq(p.inv, p.invW, G - 1 + g, LPP);
                ct_dq(v[2 * G * u + g], v[2 * G * u + g + G], t.w, t.W);
            }
        }
#pragma unroll
        for (uint32_t r = 0; r < E; r++) {
            const TwDQ t = lane_slot_dq(p.scale, p.scaleW, r, LPP);
            const uint32_t m = v[r].lo() * t.w - dq_quot(v[r], t.W).lo() * Q;  // [-q/2 - 1, q/2 + 1]
            out[r] = umin32(m, m + Q);
        }
    }

    // ---- transforms on one lane's registers ---------------------------------------------------
    // rows layout, forward levels 0..LB1-1 (register distance E/2 .. 1)
    // ub: offset into the uniform tables (split tiles: 32 * half, a run-time value; otherwise 0 and
    // every index is a compile-time constant)
    static QT_HD void fwd_rows(uint32_t (&v)[E], uint32_t ub = 0) {
        if constexpr (HLAZY) {
            h_fr_all(v, ub, std::make_integer_sequence<uint32_t, LB1>{});
            return;
        }
#pragma unroll
        for (uint32_t l = 0; l < LB1; l++) {
            const uint32_t half = E >> (l + 1);
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {  // flat butterfly index: constant trip count
                const uint32_t g = i / half, j = i % half;
                ct(v[2 * g * half + j], v[2 * g * half + j + half], uni_tw<SET, UNI_FWD>(ub + (1u << l) + g), i);
            }
        }
    }

    static QT_HD TwPair lane_slot(const TwQuad* tw, uint32_t slot, uint32_t stride) {
        const TwQuad qd = tw[(size_t)(slot >> 1) * stride];
        return (slot & 1) ? TwPair{qd.w1, qd.ws1} : TwPair{qd.w0, qd.ws0};
    }

    // cols layout, forward levels LB1..LOGN-1 (register distance N>>(l+1)); tw = lane_ptrs().fwd
    static QT_HD void fwd_cols(uint32_t (&v)[E], const TwQuad* tw) {
        if constexpr (HLAZY) {
            h_fc_all(v, tw, std::make_integer_sequence<uint32_t, LB2>{});
            return;
        }
#pragma unroll
        for (uint32_t k = 0; k < LB2; k++) {
            const uint32_t half = (E >> 1) >> (k + LB1 + LOGE - LOGN);  // N >> (l+1), l = LB1 + k
            const uint32_t G = E / (2 * half);
#pragma unroll
            for (uint32_t i = 0; i < E / 2; i++) {
                const uint32_t g = i / half, j = i % half;
                ct(v[2 * g * half + j], v[2 * g * half + j + half], lane_slot(tw, G - G0 + g, BLOCKS), i);
            }
        }
    }

    // Inverse, cols layout (the first LB2 levels), input in NTT-domain (bit-reversed) order.
    //  LAZY  : cyclic decimation-in-time with omega^-1 (the structure of the reference's radix2INTT,
    //          NTT.cu:1473-1494); the twiddle of level s depends only on the position inside the 2^(s+1)
    //          group, so it is the same for every lane (constant bank), and position 0 needs no multiply.
    //          `tw` unused.  Input |v| < q.
    //  Harvey: merged Gentleman-Sande; tw = lane_ptrs().inv (the forward table mirrored: zeta[k]^-1 =
    //          -zeta[k'], k' the mirror of k in its level).  Input < 2q.
    static QT_HD void inv_cols(uint32_t (&v)[E], const TwQuad* tw) {
        if constexpr (HLAZY) {
            h_ic_all(v, tw, std::make_integer_sequence<uint32_t, LB2>{});
            return;
        }
        if (LAZY) {
#pragma unroll
            for (uint32_t s_ = 0; s_ < LB2; s_++) {
                const uint32_t l = 1u << s_;
#pragma unroll
                for (uint32_t i = 0; i < E / 2; i++) {
                    const uint32_t u = i / l, j = i % l;
                    uint32_t& x = v[2 * l * u + j];
                    uint32_t& y = v[2 * l * u + j + l];
                    if (j == 0) {
                        const uint32_t a = x, b = y;
                        x = a + b;
                        y = a - b;
                    } else {
                        ct(x, y, uni_tw<SET, UNI_INV_PLAIN>(l + j), i);
                    }
                }
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < LB2; k++) {
                const uint32_t half = 1u << k;
                const uint32_t G = E / (2 * half);
#pragma unroll
                for (uint32_t i = 0; i < E / 2; i++) {
                    const uint32_t g = i / half, j = i % half;
                    gs(v[2 * g * half + j], v[2 * g * half + j + half], lane_slot(tw, G - G0 + (G - 1 - g), BLOCKS));
                }
            }
        }
    }

    // Inverse, rows layout (the last LB1 levels) and the output scale; result canonical in [0,q).
    //  LAZY  : DIT levels with per-lane twiddles (p.inv), then every coefficient i is multiplied by
    //          n^-1 psi^-i — the reference's invPhi table (NTT.cu:1846-1849) — from p.scale; the FUSED
    //          scale table also carries the 2^32 of the pointwise Montgomery product.
    //  Harvey: merged Gentleman-Sande with uniform twiddles, scale folded into the last level.
    template <int KIND> static QT_HD void inv_rows(uint32_t (&v)[E], const LanePtrs& p, uint32_t ub = 0) {
        if constexpr (HLAZY) {
            h_ir_all<KIND>(v, ub, std::make_integer_sequence<uint32_t, LB1>{});
            return;
        }
        if (LAZY) {
#pragma unroll
            for (uint32_t k = 0; k < LB1; k++) {
                const uint32_t G = 1u << k;
#pragma unroll
                for (uint32_t i = 0; i < E / 2; i++) {
                    const uint32_t u = i / G, g = i % G;
                    ct(v[2 * G * u + g], v[2 * G * u + g + G], lane_slot(p.inv, G - 1 + g, LPP), i);
                }
            }
#pragma unroll
            for (uint32_t r = 0; r < E; r++) v[r] = scanon(smul_shoup(v[r], lane_slot(p.scale, r, LPP)));
        } else {
#pragma unroll
            for (uint32_t k = 0; k < LB1; k++) {
                const uint32_t l = LB1 - 1 - k;
                const uint32_t half = 1u << k;
#pragma unroll
                for (uint32_t i = 0; i < E / 2; i++) {
                    const uint32_t g = i / half, j = i % half;
                    uint32_t& a = v[2 * g * half + j];
                    uint32_t& b = v[2 * g * half + j + half];
                    const TwPair t = uni_tw<SET, KIND>(ub + (1u << l) + g);
                    if (l != 0 || C::SPLIT) {  // (split tiles: the last level is split_inv's)
                        const uint32_t s_ = a + b, d = a - b + TWO_Q;
                        a = csub(s_, TWO_Q);
                        b = mul_shoup(d, t);
                    } else {  // last level: both outputs are multiplied (K resp. K*zeta^-1)
                        const uint32_t s_ = a + b, d = a - b + TWO_Q;
                        a = csub(mul_shoup(s_, uni_tw<SET, KIND>(0)), Q);
                        b = csub(mul_shoup(d, t), Q);
                    }
                }
            }
        }
    }

    // ---- split tiles (C::SPLIT): the level that joins the two halves --------------------------------
    // forward: (lo[i], hi[i]) -> (lo + zeta1 hi, lo - zeta1 hi), the first Cooley-Tukey level of the
    // full-size transform; afterwards lo / hi are the inputs of the two independent half transforms
    static QT_HD void split_fwd(uint32_t (&lo)[E], uint32_t (&hi)[E]) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++) ct(lo[r], hi[r], uni_tw<SET, UNI_FWD>(0));
    }
    // inverse: the last Gentleman-Sande level with the output scale K folded in; inputs < 2q, outputs canonical
    template <int KIND> static QT_HD void split_inv(uint32_t& a, uint32_t& b) {
        const uint32_t s_ = a + b, d = a - b + TWO_Q;
        a = csub(mul_shoup(s_, uni_tw<SET, KIND>(0)), Q);
        b = csub(mul_shoup(d, uni_tw<SET, KIND>(32)), Q);
    }

    // NTT-domain product of two forward outputs: a*b*2^-32 mod q (LAZY: signed, |.| < q; Harvey: [0,2q))
    static QT_HD uint32_t pw(uint32_t a, uint32_t b) {
        return LAZY ? smul_mont(a, b) : mul_mont(csub(a, TWO_Q), csub(b, TWO_Q));
    }
    static QT_HD void pointwise_mont(uint32_t (&a)[E], const uint32_t (&b)[E]) {
        if constexpr (HLAZY) {
            h_pw_all(a, b, std::make_integer_sequence<uint32_t, E>{});
            return;
        }
#pragma unroll
        for (uint32_t r = 0; r < E; r++) a[r] = pw(a[r], b[r]);
    }
    // same, the second operand read back from a stash written with sts_cols (keeps it out of registers)
    static QT_HD void pointwise_mont_stash(uint32_t (&a)[E], const uint32_t* stash, uint32_t lane) {
        if constexpr (HLAZY) {
            h_pw_stash(a, stash, lane, std::make_integer_sequence<uint32_t, E / 4>{});
            return;
        }
#pragma unroll
        for (uint32_t c = 0; c < E / 4; c++) {
            const U4 u = *reinterpret_cast<const U4*>(stash + swz(E * lane + 4 * c));
            a[4 * c] = pw(a[4 * c], u.x);
            a[4 * c + 1] = pw(a[4 * c + 1], u.y);
            a[4 * c + 2] = pw(a[4 * c + 2], u.z);
            a[4 * c + 3] = pw(a[4 * c + 3], u.w);
        }
    }
    // forward output times a canonical NTT-domain operand (qt_polymul_ntt)
    static QT_HD uint32_t pw_canonical(uint32_t a, uint32_t b_canonical) {
        if constexpr (HLAZY) return mul_mont(a, b_canonical);  // a < QCAP q, b < q: a b < q 2^32 as it is
        return LAZY ? smul_mont(a, b_canonical) : mul_mont(csub(a, TWO_Q), b_canonical);
    }

    // forward output -> canonical [0,q) (only the unfused forward entry points need it).
    // LAZY: |v| < (1 + 1.5 log2 n) q and q = 2^QS + d with a small d, so k = round(v / 2^QS) is within the
    // stated bound of round(v / q): r = v - k q satisfies |r| <= 2^(QS-1) + |k| d < q and one conditional add
    // finishes — 1 multiply-pipe instruction per coefficient instead of the 3 of a Shoup reduction.
    static constexpr uint32_t QS = C::QBITS - 1;                                      // floor(log2 q)
    static constexpr uint64_t KMAX = ((2 + 3 * (uint64_t)LOGN) * Q >> (QS + 1)) + 2;  // |k| bound
    static_assert(!LAZY || ((1ull << (QS - 1)) + KMAX * (Q - (1u << QS)) < Q), "canon_fwd: shift-based quotient");
    static QT_HD void canon_fwd(uint32_t (&v)[E]) {
        if constexpr (HLAZY) {
            h_canon_all(v, std::make_integer_sequence<uint32_t, E>{});
            return;
        }
#pragma unroll
        for (uint32_t r = 0; r < E; r++) {
            if (LAZY) {
                const int32_t k = ((int32_t)v[r] + (int32_t)(1u << (QS - 1))) >> QS;
                const uint32_t t = v[r] - (uint32_t)k * Q;
                v[r] = umin32(t, t + Q);
            } else {
                v[r] = csub(csub(v[r], TWO_Q), Q);
            }
        }
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
    // Natural-order view of the NTT domain.  A thread of the cols layout holds the bit-reversed-order
    // words E*l+r (l = lane within its polynomial); their natural-order positions are
    // brv(E*l+r) = brv_LOGE(r)*LPP + brv_log2(LPP)(l): for a fixed register the lanes of one polynomial
    // cover one contiguous LPP-word run, so the accesses stay coalesced without a transposition.
    static QT_HD uint32_t nat_off(uint32_t lane, uint32_t r) {
        const uint32_t l = lane % LPP;
        uint32_t bl = 0;
#pragma unroll
        for (uint32_t b = 0; b < C::LOGN - LOGE; b++) bl |= ((l >> b) & 1u) << (C::LOGN - LOGE - 1 - b);
        return (lane / LPP) * N + c_bitrev(r, LOGE) * LPP + bl;
    }
    static QT_HD void load_cols_natural(uint32_t (&v)[E], const uint32_t* g_tile, uint32_t lane, bool valid) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++) v[r] = valid ? g_tile[nat_off(lane, r)] : 0u;
    }
    static QT_HD void store_cols_natural(const uint32_t (&v)[E], uint32_t* g_tile, uint32_t lane, bool valid) {
#pragma unroll
        for (uint32_t r = 0; r < E; r++)
            if (valid) g_tile[nat_off(lane, r)] = v[r];
    }
    // swz(row_off(lane, r)) as (one of 8 per-lane bases) + (compile-time offset), for the two-polynomials-per-
    // warp tile.  With one polynomial per warp the compiler finds this form by itself (the XORed bits come from
    // the lane term only); here bit 4 mixes r with the run-time polynomial index p = lane / LPP:
    //   (p N + l + 16 r) ^ (c << 2) ^ (p << 4),  c = (r >> 1) & 7 = clo + 4 chi,  b = (r & 1) ^ chi
    //   = [p N + (l ^ (clo << 2)) + (b ? -16 p : +16 p)] + 16 b + 16 (r & ~1)
    // which saves ~2 address instructions on each of the 96 row accesses of a product.
    struct RowBases { uint32_t b[4][2]; };
    static QT_HD RowBases row_bases(uint32_t lane) {
        static_assert(PPW != 2 || (N == 512 && LPP == 16 && LOGE == 5), "row_bases is written for n = 512");
        RowBases B;
        const uint32_t p = lane / LPP, l = lane % LPP;
#pragma unroll
        for (uint32_t clo = 0; clo < 4; clo++) {
            const uint32_t L = p * N + (l ^ (clo << 2));
            B.b[clo][0] = L + 16 * p;
            B.b[clo][1] = L - 16 * p;
        }
        return B;
    }
    static QT_HD uint32_t rows_addr(const RowBases& B, uint32_t r) {
        const uint32_t c = (r >> 1) & 7u, bit = (r & 1u) ^ (c >> 2);
        return B.b[c & 3u][bit] + 16 * bit + 16 * (r & ~1u);
    }
    static QT_HD void sts_rows(const uint32_t (&v)[E], uint32_t* buf, uint32_t lane) {
        if (PPW == 2) {
            const RowBases B = row_bases(lane);
#pragma unroll
            for (uint32_t r = 0; r < E; r++) buf[rows_addr(B, r)] = v[r];
        } else {
#pragma unroll
            for (uint32_t r = 0; r < E; r++) buf[swz(row_off(lane, r))] = v[r];
        }
    }
    static QT_HD void lds_rows(uint32_t (&v)[E], const uint32_t* buf, uint32_t lane) {
        if (PPW == 2) {
            const RowBases B = row_bases(lane);
#pragma unroll
            for (uint32_t r = 0; r < E; r++) v[r] = buf[rows_addr(B, r)];
        } else {
#pragma unroll
            for (uint32_t r = 0; r < E; r++) v[r] = buf[swz(row_off(lane, r))];
        }
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
