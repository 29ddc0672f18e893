// qt_params.h — compile-time description of the four qTESLA parameter sets.
//
// The reference fixes ONE set with macros (main.cuh:13-21: P, PARAM_QINV, NTTSIZE, MIU);
// here every set is a constexpr record so that each kernel instantiation sees q, n and the
// reduction constants as immediates.  Derived values are computed with constexpr functions
// and static_assert'ed against the reference's literals where the reference has them.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define QT_CHD __host__ __device__ constexpr
#else
#define QT_CHD constexpr
#endif

namespace qt {

enum : int { SET_I = 0, SET_III = 1, SET_P_I = 2, SET_P_III = 3, NUM_SETS = 4 };
// Internal tile configuration (not a parameter set of the API): qTESLA-p-III seen as TWO 1024-point
// sub-transforms joined by one split level (X^2048+1 = (X^1024 - zeta)(X^1024 + zeta)).  The fused
// n=2048 kernel runs the 32-coefficients-per-thread tile code twice per transform instead of a
// 64-coefficients-per-thread tile (qt_kernels.cuh: k_polymul_split).
enum : int { SET_P_III_H = 4, NUM_TILE_SETS = 5 };

QT_CHD uint32_t c_mulmod(uint32_t a, uint32_t b, uint32_t q) {
    return (uint32_t)((uint64_t)a * b % q);
}
QT_CHD uint32_t c_powmod(uint32_t b, uint64_t e, uint32_t q) {
    uint64_t r = 1, x = b % q;
    while (e) {
        if (e & 1) r = r * x % q;
        x = x * x % q;
        e >>= 1;
    }
    return (uint32_t)r;
}
QT_CHD uint32_t c_log2(uint32_t n) {
    uint32_t l = 0;
    while ((1u << l) < n) l++;
    return l;
}
QT_CHD uint32_t c_neg_qinv32(uint32_t q) {  // -q^-1 mod 2^32 (PARAM_QINV, main.cuh:15)
    uint32_t inv = q;
    for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
    return 0u - inv;
}
QT_CHD uint32_t c_find_psi(uint32_t n, uint32_t q) {  // smallest g with g^((q-1)/2n) of order 2n
    for (uint32_t g = 2;; g++) {
        uint32_t psi = c_powmod(g, (q - 1) / (2 * n), q);
        if (c_powmod(psi, n, q) == q - 1) return psi;
    }
}
QT_CHD uint32_t c_bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

template <int SET> struct Params;

template <> struct Params<SET_I> {
    static constexpr uint32_t N = 512, Q = 4205569u;
    static constexpr uint32_t PSI = c_find_psi(N, Q);
};
template <> struct Params<SET_III> {
    static constexpr uint32_t N = 1024, Q = 8404993u;  // main.cuh:14,16
    static constexpr uint32_t PSI = 2083362u;          // Phi[1], constants.h:11; main.cu:26
};
template <> struct Params<SET_P_I> {
    static constexpr uint32_t N = 1024, Q = 343576577u;
    static constexpr uint32_t PSI = c_find_psi(N, Q);
};
template <> struct Params<SET_P_III> {
    static constexpr uint32_t N = 2048, Q = 856145921u;
    static constexpr uint32_t PSI = c_find_psi(N, Q);
};

template <> struct Params<SET_P_III_H> {
    static constexpr uint32_t N = 1024, Q = Params<SET_P_III>::Q;                  // one half of n = 2048
    static constexpr uint32_t PSI = c_mulmod(Params<SET_P_III>::PSI, Params<SET_P_III>::PSI, Q);
};

// Everything the kernels need, derived from (N, Q, PSI).
template <int SET> struct Cfg {
    using P = Params<SET>;
    static constexpr uint32_t N = P::N, Q = P::Q, PSI = P::PSI;
    static constexpr uint32_t LOGN = c_log2(N);
    static constexpr uint32_t PSI_INV = c_powmod(PSI, Q - 2, Q);
    static constexpr uint32_t N_INV = c_powmod(N, Q - 2, Q);
    static constexpr uint32_t QINV_NEG = c_neg_qinv32(Q);
    static constexpr uint32_t R_MODQ = (uint32_t)((1ull << 32) % Q);  // Montgomery radix mod q
    static constexpr uint32_t MU32 = (uint32_t)((1ull << 32) / Q);    // floor(2^32/q)
    static constexpr uint32_t QBITS = c_log2(Q);                      // ceil(log2 q)
    // Moduli below 2^25 leave >= 7 spare bits in a 32-bit word: butterflies run without any
    // per-level correction ("lazy").  The 29/30-bit moduli use Harvey's [0,4q) butterflies.
    static constexpr bool LAZY = QBITS <= 24;
    // SPLIT: this tile is one half of a polynomial twice its size; the level joining the halves
    // (and the output scale) is applied by the caller (Tile::split_fwd / split_inv)
    static constexpr bool SPLIT = (SET == SET_P_III_H);
    static constexpr uint32_t HALVES = SPLIT ? 2 : 1;

    // warp tile: one warp owns E*32 consecutive words = PPW whole polynomials
    static constexpr uint32_t E = (N == 2048) ? 64 : 32;  // coefficients per thread
    static constexpr uint32_t LOGE = c_log2(E);
    static constexpr uint32_t PPW = 32 * E / N;           // polynomials per warp (2 for n=512)
    static constexpr uint32_t LPP = 32 / PPW;             // lanes per polynomial
    static constexpr uint32_t LB1 = LOGE;                 // levels done in the strided layout
    static constexpr uint32_t LB2 = LOGN - LB1;           // levels done in the contiguous layout
    static constexpr uint32_t TILE_WORDS = 32 * E;
    // per-lane twiddles of the contiguous-layout pass: sum over its levels of E/(n>>l)
    static constexpr uint32_t SLOTS = E - (E >> LB2);
    static constexpr uint32_t SLOT_PAIRS = (SLOTS + 1) / 2;
    static constexpr uint32_t UNI = 1u << LB1;            // uniform twiddles: indices 1..UNI-1

    static_assert(PPW * N == TILE_WORDS, "tile must hold whole polynomials");
    static_assert(c_powmod(PSI, N, Q) == Q - 1, "psi must be a primitive 2n-th root");
    static_assert((uint64_t)Q * 5 < (1ull << 32), "Harvey butterflies need 4q < 2^32");
};

static_assert(Cfg<SET_III>::QINV_NEG == 4034936831u, "PARAM_QINV, main.cuh:15");
static_assert(Cfg<SET_III>::N_INV == 8396785u, "Ni, main.cu:26");
static_assert(c_mulmod(Cfg<SET_III>::PSI, Cfg<SET_III>::PSI, Cfg<SET_III>::Q) == 2893u, "fg0, main.cu:26");
static_assert(Cfg<SET_III>::PSI_INV == 5907167u, "psi^-1, main.cu:26 comment");
static_assert((uint32_t)((1ull << 48) / Cfg<SET_III>::Q) == 33489019u, "MIU, main.cuh:20");
static_assert(Cfg<SET_I>::PSI == 3353664u && Cfg<SET_P_I>::PSI == 249751876u &&
              Cfg<SET_P_III>::PSI == 89095543u, "derived psi (SURVEY.md 8c)");

// run-time view of the same data (C ABI qt_get_params, host table generation)
struct RtParams {
    int set;
    uint32_t n, logn, q, psi, psi_inv, omega, omega_inv, n_inv, qinv_neg, r_modq, mu32;
    uint32_t E, ppw, lb1, lb2, slots, slot_pairs, lazy;
};

template <int SET> constexpr RtParams make_rt() {
    using C = Cfg<SET>;
    return RtParams{SET, C::N, C::LOGN, C::Q, C::PSI, C::PSI_INV, c_mulmod(C::PSI, C::PSI, C::Q),
                    c_mulmod(C::PSI_INV, C::PSI_INV, C::Q), C::N_INV, C::QINV_NEG, C::R_MODQ, C::MU32,
                    C::E, C::PPW, C::LB1, C::LB2, C::SLOTS, C::SLOT_PAIRS, C::LAZY ? 1u : 0u};
}

inline bool rt_params(int set, RtParams* out) {
    switch (set) {
    case SET_I: *out = make_rt<SET_I>(); return true;
    case SET_III: *out = make_rt<SET_III>(); return true;
    case SET_P_I: *out = make_rt<SET_P_I>(); return true;
    case SET_P_III: *out = make_rt<SET_P_III>(); return true;
    default: return false;
    }
}

}  // namespace qt
