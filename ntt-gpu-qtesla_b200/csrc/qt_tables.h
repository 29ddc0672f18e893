// qt_tables.h — host-side table generation.
//
// (1) The five drop-in tables of the reference's constants.h:3-35 (bitrev_tbl, Phi, invPhi,
//     tf0_gpu, ti0_gpu), regenerated bit-identically from (n, q, psi) for every set.
// (2) The engine's own twiddle tables: merged-psi zetas in bit-reversed order,
//     zeta[k] = psi^brv_logn(k), each with its Shoup companion floor(w*2^32/q), arranged for
//     the two passes of the warp-tile kernel (qt_tile.cuh).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "qt_params.h"

namespace qt {

struct alignas(8) TwPair { uint32_t w, ws; };           // twiddle and floor(w * 2^32 / q)
struct alignas(16) TwQuad { uint32_t w0, ws0, w1, ws1; };  // two consecutive slots (one 128-bit load)

// Twiddles of the FP64-quotient ("DQ") butterflies (qt_tile.cuh): w as a representative in [0, q) and W = w / q as a
// double — D(y) * W, with D(y) the DENORMAL double whose bit pattern is {lo = y, hi = 0}, has the bit pattern
// {rint(y w / q), 0}.  Two consecutive slots per access, like TwQuad.
struct alignas(16) TwW2 { double W0, W1; };
struct alignas(8) TwU2 { uint32_t w0, w1; };
inline uint32_t dq_unsigned(uint32_t w_centred, uint32_t q) { return (int32_t)w_centred < 0 ? w_centred + q : w_centred; }
inline double dq_companion(uint32_t w_unsigned, uint32_t q) { return (double)w_unsigned / (double)q; }

enum : int { UNI_FWD = 0, UNI_INV_PLAIN = 1, UNI_INV_FUSED = 2, UNI_KINDS = 3 };
constexpr int UNI_MAX = 64;  // 2^LB1 for the largest tile (n=2048: 6 strided levels)

struct HostTables {
    RtParams p;
    // reference drop-in tables, n words each
    std::vector<uint32_t> bitrev, Phi, invPhi, tf0, ti0;
    // uniform (lane-independent) twiddles, index k in [1, 2^LB1).
    //   UNI_FWD            : merged zetas of the forward rows pass, zeta[k] = psi^brv(k)
    //   UNI_INV_* (Harvey) : inverse zetas of the rows pass; index 0 = output scale K, index 1 pre-multiplied by K
    //   UNI_INV_* (LAZY)   : cyclic DIT twiddles of the cols pass: index l + j = omega^-(j*n/(2l)), l = 2^s
    TwPair uni[UNI_KINDS][UNI_MAX];
    // kernel table blocks (one per output-scale kind): [fwd per-lane][LAZY: inverse per-lane][LAZY: scale]
    std::vector<TwQuad> block[2];  // 0: plain scale n^-1 psi^-i (unfused inverse), 1: fused scale (x 2^32)
    uint32_t fwd_quads, inv_quads, scale_quads;
    // LAZY sets: the same twiddles for the FP64-quotient kernels (same indexing as uni[][] / block[][])
    double uniW[UNI_KINDS][UNI_MAX];
    uint32_t uniU[UNI_KINDS][UNI_MAX];
    std::vector<TwW2> blockW[2];
    std::vector<TwU2> blockU[2];
};

// LAZY sets store twiddles for the SIGNED Shoup product: w centred in (-q/2, q/2] as a two's-complement word,
// companion floor(w * 2^32 / q) as a signed word.  Harvey sets: w in [0,q), companion floor(w*2^32/q).
inline uint32_t shoup(uint32_t w, uint32_t q) { return (uint32_t)(((uint64_t)w << 32) / q); }
inline TwPair tw_unsigned(uint32_t w, uint32_t q) { return TwPair{w, shoup(w, q)}; }
inline TwPair tw_signed(uint32_t w, uint32_t q) {
    const int64_t wc = (w > q / 2) ? (int64_t)w - (int64_t)q : (int64_t)w;
    const int64_t num = wc * (int64_t)(1ll << 32);
    int64_t fl = num / (int64_t)q;
    if (num % (int64_t)q != 0 && num < 0) fl -= 1;  // floor division
    return TwPair{(uint32_t)(int32_t)wc, (uint32_t)(int32_t)fl};
}

// the same as a compile-time constant usable in device code
QT_CHD TwPair tw_signed_c(uint32_t w, uint32_t q) {
    const int64_t wc = (w > q / 2) ? (int64_t)w - (int64_t)q : (int64_t)w;
    const int64_t num = wc * (int64_t)(1ll << 32);
    int64_t fl = num / (int64_t)q;
    if (num % (int64_t)q != 0 && num < 0) fl -= 1;
    return TwPair{(uint32_t)(int32_t)wc, (uint32_t)(int32_t)fl};
}

QT_CHD TwPair tw_unsigned_c(uint32_t w, uint32_t q) { return TwPair{w, (uint32_t)(((uint64_t)w << 32) / q)}; }

inline void put_slot(std::vector<TwQuad>& v, size_t base, uint32_t slot, uint32_t stride, uint32_t lane, TwPair t) {
    TwQuad& qd = v[base + (size_t)(slot / 2) * stride + lane];
    if (slot & 1) { qd.w1 = t.w; qd.ws1 = t.ws; }
    else          { qd.w0 = t.w; qd.ws0 = t.ws; }
}

inline void build_tables(int set, HostTables* T) {
    RtParams p;
    rt_params(set, &p);
    T->p = p;
    const uint32_t n = p.n, q = p.q;
    const bool lazy = p.lazy != 0;
    auto mk = [&](uint32_t w) { return lazy ? tw_signed(w, q) : tw_unsigned(w, q); };
    T->bitrev.resize(n); T->Phi.resize(n); T->invPhi.resize(n); T->tf0.resize(n); T->ti0.resize(n);
    std::vector<uint32_t> psi_pow(2 * n);  // psi^e, e in [0,2n)
    psi_pow[0] = 1;
    for (uint32_t e = 1; e < 2 * n; e++) psi_pow[e] = c_mulmod(psi_pow[e - 1], p.psi, q);
    for (uint32_t i = 0; i < n; i++) {
        T->bitrev[i] = c_bitrev(i, p.logn);
        T->Phi[i] = psi_pow[i];                                              // psi^i
        T->invPhi[i] = c_mulmod(p.n_inv, psi_pow[(2 * n - i) % (2 * n)], q);  // n^-1 psi^-i
        T->tf0[i] = psi_pow[2 * i];                                          // omega^i
        T->ti0[i] = psi_pow[(2 * n - 2 * i) % (2 * n)];                      // omega^-i
    }
    // merged zetas: zeta[k] = psi^brv(k); inverse zeta = psi^(2n - brv(k))
    auto zf = [&](uint32_t k) { return psi_pow[c_bitrev(k, p.logn)]; };
    auto zi = [&](uint32_t k) { return psi_pow[(2 * n - c_bitrev(k, p.logn)) % (2 * n)]; };
    const uint32_t uni = 1u << p.lb1;
    for (int kind = 0; kind < UNI_KINDS; kind++)
        for (int k = 0; k < UNI_MAX; k++) T->uni[kind][k] = TwPair{0, 0};
    const uint32_t K_plain = p.n_inv;
    const uint32_t K_fused = c_mulmod(p.n_inv, p.r_modq, q);  // absorbs the pointwise Montgomery 2^-32
    for (uint32_t k = 1; k < uni; k++) T->uni[UNI_FWD][k] = mk(zf(k));
    if (!lazy) {
        for (uint32_t k = 1; k < uni; k++) {
            uint32_t ip = zi(k), ifu = zi(k);
            if (k == 1) { ip = c_mulmod(ip, K_plain, q); ifu = c_mulmod(ifu, K_fused, q); }
            T->uni[UNI_INV_PLAIN][k] = tw_unsigned(ip, q);
            T->uni[UNI_INV_FUSED][k] = tw_unsigned(ifu, q);
        }
        T->uni[UNI_INV_PLAIN][0] = tw_unsigned(K_plain, q);
        T->uni[UNI_INV_FUSED][0] = tw_unsigned(K_fused, q);
    } else {
        // cyclic DIT (radix2INTT, NTT.cu:1478-1493): level s, half size l = 2^s, position j < l: ti0[j * n/(2l)]
        for (uint32_t s = 0; s < p.lb2; s++) {
            const uint32_t l = 1u << s;
            for (uint32_t j = 0; j < l; j++)
                T->uni[UNI_INV_PLAIN][l + j] = T->uni[UNI_INV_FUSED][l + j] = tw_signed(T->ti0[j * (n / (2 * l))], q);
        }
    }
    // ---- kernel table blocks -----------------------------------------------------------------------
    const uint32_t blocks = n / p.E;                 // lane blocks of the forward cols pass
    const uint32_t lpp = 32 / p.ppw;                 // lanes per polynomial (rows layout)
    T->fwd_quads = p.slot_pairs * blocks;
    T->inv_quads = lazy ? ((p.E - 1 + 1) / 2) * lpp : 0;
    T->scale_quads = lazy ? (p.E / 2) * lpp : 0;
    for (int kind = 0; kind < 2; kind++) {
        std::vector<TwQuad>& B = T->block[kind];
        B.assign((size_t)T->fwd_quads + T->inv_quads + T->scale_quads, TwQuad{0, 0, 0, 0});
        // forward per-lane table: block jb (first element E*jb) at level l uses zeta indices
        // 2^l + jb*G_l + g, G_l = E >> (logn - l); slots enumerate (l, g), l ascending
        for (uint32_t jb = 0; jb < blocks; jb++) {
            uint32_t slot = 0;
            for (uint32_t l = p.lb1; l < p.logn; l++) {
                const uint32_t G = p.E >> (p.logn - l);
                for (uint32_t g = 0; g < G; g++, slot++) put_slot(B, 0, slot, blocks, jb, mk(zf((1u << l) + jb * G + g)));
            }
        }
        if (!lazy) continue;
        // inverse rows pass (DIT levels s = lb2 .. logn-1): lane j holds positions j + lpp*r; level with
        // half size l = lpp*G pairs registers (r, r+G), twiddle ti0[(j + lpp*g) * n/(2l)], slot G-1+g
        for (uint32_t j = 0; j < lpp; j++)
            for (uint32_t k = 0; k < p.lb1; k++) {
                const uint32_t G = 1u << k, l = lpp * G;
                for (uint32_t g = 0; g < G; g++)
                    put_slot(B, T->fwd_quads, G - 1 + g, lpp, j, tw_signed(T->ti0[((j + lpp * g) * (n / (2 * l))) % n], q));
            }
        // output scale of coefficient i = j + lpp*r: n^-1 psi^-i (= invPhi[i]), fused kind times 2^32
        for (uint32_t j = 0; j < lpp; j++)
            for (uint32_t r = 0; r < p.E; r++) {
                uint32_t sc = T->invPhi[j + lpp * r];
                if (kind == 1) sc = c_mulmod(sc, p.r_modq, q);
                put_slot(B, (size_t)T->fwd_quads + T->inv_quads, r, lpp, j, tw_signed(sc, q));
            }
    }
    for (int kind = 0; kind < UNI_KINDS; kind++)
        for (int k = 0; k < UNI_MAX; k++) {
            T->uniU[kind][k] = lazy ? dq_unsigned(T->uni[kind][k].w, q) : 0u;
            T->uniW[kind][k] = lazy ? dq_companion(T->uniU[kind][k], q) : 0.0;
        }
    for (int kind = 0; kind < 2; kind++) {
        T->blockW[kind].clear();
        T->blockU[kind].clear();
        if (!lazy) continue;
        for (const TwQuad& qd : T->block[kind]) {
            const TwU2 u{dq_unsigned(qd.w0, q), dq_unsigned(qd.w1, q)};
            T->blockU[kind].push_back(u);
            T->blockW[kind].push_back(TwW2{dq_companion(u.w0, q), dq_companion(u.w1, q)});
        }
    }
}

// Tables of the split tile (SET_P_III_H): the n=2048 zeta table cut into the two 1024-point halves.
// Half h at sub-level l' (group g') uses the full-size zeta index 2^(l'+1) + h*2^l' + g'.
//   uni[UNI_FWD][32h + k]   k in [1,32): forward zetas of the rows pass of half h;   [0] = zeta[1] (split level)
//   uni[UNI_INV_*][32h + k] k in [1,32): inverse zetas of the rows pass of half h;   [0] = K, [32] = K * zeta[1]^-1
//   block[kind] = [per-lane forward table of half 0][of half 1]; the inverse cols pass of half h reads
//   the table of half 1-h mirrored (zeta[k]^-1 = -zeta[mirror(k)], and the mirror lives in the other half).
inline void build_tables_split(HostTables* T) {
    const RtParams p = make_rt<SET_P_III_H>();
    const RtParams f = make_rt<SET_P_III>();
    T->p = p;
    const uint32_t q = f.q, nf = f.n;
    std::vector<uint32_t> psi_pow(2 * nf);
    psi_pow[0] = 1;
    for (uint32_t e = 1; e < 2 * nf; e++) psi_pow[e] = c_mulmod(psi_pow[e - 1], f.psi, q);
    auto zf = [&](uint32_t k) { return psi_pow[c_bitrev(k, f.logn)]; };
    auto zi = [&](uint32_t k) { return psi_pow[(2 * nf - c_bitrev(k, f.logn)) % (2 * nf)]; };
    auto full_index = [&](uint32_t h, uint32_t ksub) {  // ksub = 2^l' + g'
        const uint32_t lp = c_log2(ksub + 1) - 1;     // floor(log2 ksub)
        return (2u << lp) + h * (1u << lp) + (ksub - (1u << lp));
    };
    for (int kind = 0; kind < UNI_KINDS; kind++)
        for (int k = 0; k < UNI_MAX; k++) T->uni[kind][k] = TwPair{0, 0};
    const uint32_t K_plain = f.n_inv, K_fused = c_mulmod(f.n_inv, f.r_modq, q);
    for (uint32_t h = 0; h < 2; h++)
        for (uint32_t k = 1; k < 32; k++) {
            T->uni[UNI_FWD][32 * h + k] = tw_unsigned(zf(full_index(h, k)), q);
            T->uni[UNI_INV_PLAIN][32 * h + k] = T->uni[UNI_INV_FUSED][32 * h + k] = tw_unsigned(zi(full_index(h, k)), q);
        }
    T->uni[UNI_FWD][0] = tw_unsigned(zf(1), q);
    T->uni[UNI_INV_PLAIN][0] = tw_unsigned(K_plain, q);
    T->uni[UNI_INV_FUSED][0] = tw_unsigned(K_fused, q);
    T->uni[UNI_INV_PLAIN][32] = tw_unsigned(c_mulmod(zi(1), K_plain, q), q);
    T->uni[UNI_INV_FUSED][32] = tw_unsigned(c_mulmod(zi(1), K_fused, q), q);
    const uint32_t blocks = p.n / p.E;
    T->fwd_quads = 2 * p.slot_pairs * blocks;
    T->inv_quads = T->scale_quads = 0;
    for (int kind = 0; kind < 2; kind++) {
        std::vector<TwQuad>& B = T->block[kind];
        B.assign(T->fwd_quads, TwQuad{0, 0, 0, 0});
        for (uint32_t h = 0; h < 2; h++)
            for (uint32_t jb = 0; jb < blocks; jb++) {
                uint32_t slot = 0;
                for (uint32_t l = p.lb1; l < p.logn; l++) {
                    const uint32_t G = p.E >> (p.logn - l);
                    for (uint32_t g = 0; g < G; g++, slot++)
                        put_slot(B, (size_t)h * p.slot_pairs * blocks, slot, blocks, jb,
                                 tw_unsigned(zf(full_index(h, (1u << l) + jb * G + g)), q));
                }
            }
    }
}

}  // namespace qt
