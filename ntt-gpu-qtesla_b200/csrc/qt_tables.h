// qt_tables.h — host-side table generation.
//
// (1) The five drop-in tables of the reference's constants.h:3-35 (bitrev_tbl, Phi, invPhi,
//     tf0_gpu, ti0_gpu), regenerated bit-identically from (n, q, psi) for every set.
// (2) The engine's own twiddle tables: merged-psi zetas in bit-reversed order,
//     zeta[k] = psi^brv_logn(k), each with its Shoup companion floor(w*2^32/q), arranged for
//     the two passes of the warp-tile kernel (qt_tile.cuh).
#pragma once
#include <cstdint>
#include <vector>

#include "qt_params.h"

namespace qt {

struct alignas(8) TwPair { uint32_t w, ws; };           // twiddle and floor(w * 2^32 / q)
struct alignas(16) TwQuad { uint32_t w0, ws0, w1, ws1; };  // two consecutive slots (one 128-bit load)

enum : int { UNI_FWD = 0, UNI_INV_PLAIN = 1, UNI_INV_FUSED = 2, UNI_KINDS = 3 };
constexpr int UNI_MAX = 64;  // 2^LB1 for the largest tile (n=2048: 6 strided levels)

struct HostTables {
    RtParams p;
    // reference drop-in tables, n words each
    std::vector<uint32_t> bitrev, Phi, invPhi, tf0, ti0;
    // uniform (lane-independent) twiddles of the strided pass: index k in [1, 2^LB1);
    // for the inverse kinds index 0 holds the output scale K and index 1 is pre-multiplied by K
    TwPair uni[UNI_KINDS][UNI_MAX];
    // per-lane twiddles of the contiguous pass: [pair][block], block = n/E lanes
    std::vector<TwQuad> lane_fwd;  // the inverse pass reads the same table mirrored (qt_tile.cuh inv_cols)
    uint32_t lane_blocks;  // n / E
};

inline uint32_t shoup(uint32_t w, uint32_t q) { return (uint32_t)(((uint64_t)w << 32) / q); }

inline void build_tables(int set, HostTables* T) {
    RtParams p;
    rt_params(set, &p);
    T->p = p;
    const uint32_t n = p.n, q = p.q;
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
    const uint32_t K_fused = c_mulmod(p.n_inv, p.r_modq, q);  // absorbs the pointwise Montgomery R^-1
    for (uint32_t k = 1; k < uni; k++) {
        uint32_t f = zf(k), ip = zi(k), ifu = zi(k);
        if (k == 1) { ip = c_mulmod(ip, K_plain, q); ifu = c_mulmod(ifu, K_fused, q); }
        T->uni[UNI_FWD][k] = TwPair{f, shoup(f, q)};
        T->uni[UNI_INV_PLAIN][k] = TwPair{ip, shoup(ip, q)};
        T->uni[UNI_INV_FUSED][k] = TwPair{ifu, shoup(ifu, q)};
    }
    T->uni[UNI_INV_PLAIN][0] = TwPair{K_plain, shoup(K_plain, q)};
    T->uni[UNI_INV_FUSED][0] = TwPair{K_fused, shoup(K_fused, q)};
    // per-lane tables.  Thread-block j' (first element E*j') at level l uses zeta indices
    // 2^l + j'*G_l + g, g in [0,G_l), G_l = E >> (logn - l); slots enumerate (l, g), l ascending.
    const uint32_t blocks = n / p.E;
    T->lane_blocks = blocks;
    T->lane_fwd.assign((size_t)p.slot_pairs * blocks, TwQuad{0, 0, 0, 0});
    for (uint32_t jb = 0; jb < blocks; jb++) {
        uint32_t slot = 0;
        for (uint32_t l = p.lb1; l < p.logn; l++) {
            const uint32_t G = p.E >> (p.logn - l);
            for (uint32_t g = 0; g < G; g++, slot++) {
                const uint32_t k = (1u << l) + jb * G + g;
                const uint32_t f = zf(k);
                TwQuad& qf = T->lane_fwd[(size_t)(slot / 2) * blocks + jb];
                if (slot & 1) { qf.w1 = f; qf.ws1 = shoup(f, q); }
                else          { qf.w0 = f; qf.ws0 = shoup(f, q); }
            }
        }
    }
}

}  // namespace qt
