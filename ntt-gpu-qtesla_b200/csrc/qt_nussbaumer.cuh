// qt_nussbaumer.cuh — batched Nussbaumer negacyclic convolution (the NTT-free alternative).
//
// The reference has this only as a single-polynomial CPU function (nussbaumer_fft, NTT.cu:167-277,
// ring macros NTT.cu:102-134); there is no reference GPU kernel.  Structure (n = m*r):
//   rows X_i[j] = x[m*j+i], rows m..2m-1 are copies         (NTT.cu:185-193)
//   log2(m) forward stages of rotate-and-add row butterflies  (NTT.cu:195-235)   — no multiplies
//   2m negacyclic schoolbook products of length r             (NTT.cu:237-239, naive 147-165)
//   log2(m)+1 inverse stages with halving                     (NTT.cu:241-269)
//   recombination with u^m = w                                (NTT.cu:271-276)
// Two rings:
//   RING 0: Z/(2^32-1) with the reference's end-around-carry macros, every operation in the
//           reference's order, so even the representation of zero (0 vs 0xFFFFFFFF) is bit-exact;
//   RING 1: Z_q — same index maps, arithmetic mod q, canonical output; equals the NTT product.
//
// Mapping: a CTA of 128 threads works on P = 128/(2m) polynomials at a time, all rows in shared
// memory (row stride r+1 words: conflict-free both along a row and down a column).
//   * stage phases: one WARP owns a whole row butterfly (lane = coefficient), reads then writes;
//   * product phase: one THREAD owns a whole row: x, y in registers, r*r multiply-accumulates with
//     no communication at all (IMAD.WIDE back to back);
//   * __syncthreads only between phases.
// The phase functions are __host__ __device__ so tests/emu can run them thread by thread.
//
// Kernels: k_nussbaumer (the mapping above; n = 2048) and k_nussbaumer_warp (one warp per polynomial, rows in
// registers, rotations as warp shuffles; n = 512, 1024).  Row products of the Z_q kernels come in three
// bit-identical forms (qt_set_nussbaumer_variant): schoolbook on the integer pipe, recursive (NussInner: the
// row product split once more) and schoolbook on the FP64 pipe with exact double-precision accumulation
// (NussRowF64, q < 2^25).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <type_traits>

#include "qt_tile.cuh"

namespace qt {

template <int SET> struct NussCfg {
    using C = Cfg<SET>;
    static constexpr uint32_t N = C::N, Q = C::Q;
    static constexpr uint32_t M = (N == 512) ? 16 : 32;
    static constexpr uint32_t R = N / M;           // 32, 32, 64
    static constexpr uint32_t LOGM = c_log2(M);
    static constexpr uint32_t ROWS = 2 * M;
    static constexpr uint32_t THREADS = 128;
    static constexpr uint32_t P = THREADS / ROWS;  // polynomials per CTA pass: 4, 2, 2
    static constexpr uint32_t XS = R + 1;          // X row stride (words)
    static constexpr uint32_t YS = (R == 64) ? 2 * R + 1 : R + 1;  // Y rows are doubled in place for r=64
    static constexpr uint32_t X_WORDS = ROWS * XS, Y_WORDS = ROWS * YS;
    static constexpr uint32_t POLY_WORDS = X_WORDS + Y_WORDS;
    static constexpr size_t SMEM_BYTES = (size_t)P * POLY_WORDS * sizeof(uint32_t);
    static constexpr uint32_t ROT_UNIT = R / M;    // w^(r/m) is the 2m-th root of unity
};

// ---- ring arithmetic --------------------------------------------------------------------------
template <int SET, int RING> struct NussOps;

template <int SET> struct NussOps<SET, 0> {  // Z/(2^32-1), NTT.cu:102-134
    // End-around-carry arithmetic.  On the device each operation is the two-instruction carry chain it
    // is (IADD3 with carry-out + IADD3.X); the C++ forms below are what the host emulator executes and
    // what the compiler otherwise turns into compare/select sequences twice as long.
    static QT_HD uint32_t add(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        uint32_t r;
        asm("add.cc.u32 %0, %1, %2;\n\taddc.u32 %0, %0, 0;" : "=r"(r) : "r"(a), "r"(b));
        return r;
#else
        uint32_t t = a + b;
        return t + (uint32_t)(t < a);
#endif
    }
    static QT_HD uint32_t sub(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        uint32_t r;
        asm("sub.cc.u32 %0, %1, %2;\n\tsubc.u32 %0, %0, 0;" : "=r"(r) : "r"(a), "r"(b));
        return r;
#else
        return (a - b) - (uint32_t)(b > a);
#endif
    }
    static QT_HD uint32_t norm(uint32_t a) {  // 0xFFFFFFFF -> 0 (the carry of a + 1)
#if defined(__CUDA_ARCH__)
        uint32_t r;
        asm("add.cc.u32 %0, %1, 1;\n\taddc.u32 %0, %1, 0;" : "=&r"(r) : "r"(a));
        return r;
#else
        return a + (uint32_t)(a == 0xFFFFFFFFu);
#endif
    }
    static QT_HD uint32_t neg(uint32_t a) { return norm(0xFFFFFFFFu - a); }
    // halving adds the (odd) modulus to an odd value first: (a + 2^32 - 1) / 2 = (a >> 1) + 2^31, i.e. a
    // rotation to the right by one bit (2^32 = 1 in this ring)
    static QT_HD uint32_t half(uint32_t a) {
        a = norm(a);
#if defined(__CUDA_ARCH__)
        return __funnelshift_r(a, a, 1);
#else
        return (uint32_t)(((uint64_t)a + (uint64_t)(uint32_t)(0u - (a & 1u))) >> 1);
#endif
    }
    static QT_HD uint32_t fold(uint64_t t) { return add((uint32_t)t, (uint32_t)(t >> 32)); }
    // ---- QT_RING_2P32M1_LIFT_Q: Z_q operands through the ring 2^32-1 -------------------------
    // A canonical residue v in [0, q) enters as its CENTRED integer representative v~ in (-q/2, q/2] (negative
    // values as 2^32-1 + v~); the ring product is the INTEGER negacyclic product as long as every integer
    // coefficient satisfies |c| < 2^31 (sum over i of |x~_i| |y~_(k-i)| < 2^31 is sufficient: e.g. qTESLA's s*c, e*c with a
    // small secret and a weight-h ternary challenge, or uniform * ternary for the 23-bit moduli); the signed lift of
    // the ring value reduced mod q is then the Z_q product.  Outside the precondition the result is NOT the Z_q product.
    static constexpr uint32_t Q = Cfg<SET>::Q;
    static QT_HD uint32_t lift_in(uint32_t v) { return v > Q / 2 ? v + (0xFFFFFFFFu - Q) : v; }
    static QT_HD uint32_t lift_out(uint32_t r) {
        r = norm(r);
        const bool negative = r >= 0x80000000u;
        const uint32_t mag = negative ? 0xFFFFFFFFu - r : r;                       // |signed lift| < 2^31
        uint32_t m = mag - mulhi32(mag, Cfg<SET>::MU32) * Q;                       // [0, 2q)
        m = umin32(m, m - Q);                                                      // [0, q)
        return (negative && m != 0) ? Q - m : m;
    }
};

template <int SET> struct NussOps<SET, 1> {  // Z_q, operands canonical
    static constexpr uint32_t Q = Cfg<SET>::Q;
    static QT_HD uint32_t csub(uint32_t a) { return umin32(a, a - Q); }
    static QT_HD uint32_t add(uint32_t a, uint32_t b) { return csub(a + b); }
    static QT_HD uint32_t sub(uint32_t a, uint32_t b) { return csub(a - b + Q); }
    static QT_HD uint32_t neg(uint32_t a) { return csub(Q - a); }
    static QT_HD uint32_t half(uint32_t a) { return (a + (Q & (0u - (a & 1u)))) >> 1; }
};

// Z_q for q < 2^25 inside the warp-resident kernel: two's-complement residues and NO reduction in the
// stage phases.  Forward: |v| <= 2^LOGM q.  Products accumulate in a signed 64-bit register
// (32 terms of < (2^LOGM q)^2 each); one signed Montgomery reduction and one signed Shoup multiplication by
// 2^32 * 2^-(LOGM+1) bring every Z row back to [-q/2, 3q/2) and pay for the halvings of ALL inverse stages
// in advance, so the inverse is additions only: |v| < 1.5 * 2^(LOGM+1) q, recombination doubles once more.
template <int SET> struct NussOpsLazy {
    static constexpr uint32_t Q = Cfg<SET>::Q;
    static QT_HD uint32_t add(uint32_t a, uint32_t b) { return a + b; }
    static QT_HD uint32_t sub(uint32_t a, uint32_t b) { return a - b; }
    static QT_HD uint32_t neg(uint32_t a) { return 0u - a; }
    static QT_HD uint32_t half(uint32_t a) { return a; }  // deferred: folded into the post-product constant
};

// ---- recursive row products -----------------------------------------------------------------------
// "Nussbaumer recursive": the 2m products of length r are themselves negacyclic products mod (w^r + 1), so
// the same split applies again instead of the schoolbook `naive` (NTT.cu:147-165; the reference does not
// recurse).  r = MI*RI with MI | RI (32 = 4*8, 64 = 8*8): 2*MI schoolbook products of length 8 replace one of
// length r — 512 instead of 1024 multiplications (r = 32), 1024 instead of 4096 (r = 64) — at the price of
// log2(MI) forward and log2(MI)+1 inverse stages of additions.  Everything is thread-private: after
// unrolling every index is a compile-time constant, so rows are registers, a rotation by w^sr is a renaming
// and its sign turns an add into a subtract.  Z_q only (both arithmetic flavours); the ring 2^32-1 keeps the
// reference's operation order and with it the reference's representation of zero.
//   LAZYQ (q < 2^25): two's-complement residues, no reduction in the stages.  |x| <= BX, |y| <= BY on entry
//     (static_assert'ed by the caller through fits()); a product accumulates in ONE signed 64-bit register,
//     is Montgomery-reduced and multiplied (signed Shoup) by `fix` = 2^32 * 2^-(LOGMI+1) * (whatever the caller
//     folds in), which pays for the halvings of the inner inverse in advance; the inner inverse and the
//     recombination are additions (|v| < 24 q), one signed Shoup multiplication by 1 brings the output back
//     to [-q/2, 3q/2).
//   canonical (29/30-bit q): operands and result in [0, q); compare-subtract additions, accumulators of
//     TERMS products, Montgomery reduction, the halvings and 2^32 in one final Shoup multiplication.
template <int SET, uint32_t LEN, bool LAZYQ> struct NussInner {
    using T = Tile<SET>;
    static constexpr uint32_t Q = Cfg<SET>::Q;
    static constexpr uint32_t MI = (LEN >= 64) ? 8 : 4, RI = LEN / MI, LOGMI = c_log2(MI), ROWS = 2 * MI;
    static constexpr uint32_t UNIT = RI / MI;  // w^(RI/MI) is the 2*MI-th root of unity
    static_assert(MI * RI == LEN && RI % MI == 0 && RI == 8, "inner split: MI | RI");
    static constexpr uint32_t TERMS = LAZYQ ? RI : (T::QCAP >= 8 ? 8 : 4);  // products per 64-bit accumulator
    static constexpr uint32_t INV_HALVES = c_powmod((Q + 1) / 2, LOGMI + 1, Q);  // 2^-(LOGMI+1)
    // lazy ranges: inner forward values in an int32, RI products in an int64
    static QT_CHD bool fits(uint64_t bx, uint64_t by) {
        return by * MI < (1ull << 31) && bx * MI < (1ull << 31) && (bx * MI) * (by * MI) < ((1ull << 63) / RI);
    }
    // the constant a caller passes as `fix` when it wants the plain product (canonical) / the product times
    // `extra` (LAZYQ, extra = whatever else it defers)
    static QT_CHD uint32_t fix_value(uint32_t extra) { return c_mulmod(c_mulmod(T::C::R_MODQ, INV_HALVES, Q), extra, Q); }

    static QT_HD uint32_t add(uint32_t a, uint32_t b) { return LAZYQ ? a + b : T::csub(a + b, Q); }
    static QT_HD uint32_t sub(uint32_t a, uint32_t b) { return LAZYQ ? a - b : T::csub(a - b + Q, Q); }
    static QT_HD uint32_t rot(uint32_t i, uint32_t j) { return (c_bitrev(i, LOGMI - j) << j) * UNIT; }

    // rows V_i[a] = v[MI*a + i], rows MI..2MI-1 copies, then log2(MI) rotate-and-add stages
    static QT_HD void forward(const uint32_t (&v)[LEN], uint32_t (&V)[ROWS][RI]) {
#pragma unroll
        for (uint32_t i = 0; i < MI; i++)
#pragma unroll
            for (uint32_t a = 0; a < RI; a++) V[i][a] = V[i + MI][a] = v[MI * a + i];
#pragma unroll
        for (int j = (int)LOGMI - 1; j >= 0; j--) {
#pragma unroll
            for (uint32_t bf = 0; bf < MI; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t I = (i << (j + 1)) + t, L = I + (1u << j), sr = rot(i, (uint32_t)j);
                uint32_t lo[RI], hi[RI];
#pragma unroll
                for (uint32_t a = 0; a < RI; a++) {
                    const uint32_t src = V[L][(a - sr) & (RI - 1)], vi = V[I][a];
                    const bool wrap = a < sr;  // T[a] = -src on wrap-around
                    lo[a] = wrap ? sub(vi, src) : add(vi, src);
                    hi[a] = wrap ? add(vi, src) : sub(vi, src);
                }
#pragma unroll
                for (uint32_t a = 0; a < RI; a++) { V[I][a] = lo[a]; V[L][a] = hi[a]; }
            }
        }
    }

    // one length-RI schoolbook product, output coefficient o
    static QT_HD uint32_t dot(const uint32_t (&xr)[RI], const uint32_t (&yr)[RI], const uint32_t (&nyr)[RI], uint32_t o,
                              TwPair fix) {
        if (LAZYQ) {
            int64_t acc = 0;
#pragma unroll
            for (uint32_t j = 0; j < RI; j++) {
                // wrapped terms enter negated (nyr = -yr)
                acc += (int64_t)(int32_t)xr[j] * (int64_t)(int32_t)((j <= o) ? yr[(o - j) & (RI - 1)] : nyr[(o - j) & (RI - 1)]);
            }
            const uint32_t m = (uint32_t)acc * (0u - T::C::QINV_NEG);               // lo(acc) * q^-1
            const uint32_t red = (uint32_t)(acc >> 32) - (uint32_t)mulhi32s(m, Q);  // acc * 2^-32, |red| < 2^31
            return T::smul_shoup(red, fix);                                         // [-q/2, 3q/2)
        }
        uint32_t r = 0;
#pragma unroll
        for (uint32_t g = 0; g < RI / TERMS; g++) {
            uint64_t acc = 0;
#pragma unroll
            for (uint32_t jj = 0; jj < TERMS; jj++) {
                const uint32_t j = g * TERMS + jj;
                acc += (uint64_t)xr[j] * ((j <= o) ? yr[(o - j) & (RI - 1)] : nyr[(o - j) & (RI - 1)]);
            }
            const uint32_t m = (uint32_t)acc * T::C::QINV_NEG;                // Montgomery: acc * 2^-32
            const uint32_t red = (uint32_t)((acc + (uint64_t)m * Q) >> 32);  // [0, 2q)
            r = (g == 0) ? red : r + red;
        }
        return (RI / TERMS == 1) ? T::csub(r, Q) : T::csub(T::fold2q(r), Q);
    }

    // z = x (*) y * fix * 2^-32 * 2^(LOGMI+1) mod (w^LEN + 1, q); LAZYQ: fix is a signed Shoup pair and z lies in
    // [-q/2, 3q/2); canonical: fix is an unsigned Shoup pair and z in [0, q).
    static QT_HD void product(const uint32_t (&x)[LEN], const uint32_t (&y)[LEN], uint32_t (&z)[LEN], TwPair fix) {
        uint32_t X[ROWS][RI], Y[ROWS][RI];
        forward(x, X);
        forward(y, Y);
#pragma unroll
        for (uint32_t k = 0; k < ROWS; k++) {
            uint32_t ny[RI], out[RI];
#pragma unroll
            for (uint32_t a = 0; a < RI; a++) ny[a] = LAZYQ ? 0u - Y[k][a] : Q - Y[k][a];
#pragma unroll
            for (uint32_t o = 0; o < RI; o++) out[o] = dot(X[k], Y[k], ny, o, fix);
#pragma unroll
            for (uint32_t o = 0; o < RI; o++) X[k][o] = out[o];  // Z_k over X_k
        }
        // inverse stages, halvings deferred (NTT.cu:241-269 applied to the inner split)
#pragma unroll
        for (uint32_t j = 0; j <= LOGMI; j++) {
#pragma unroll
            for (uint32_t bf = 0; bf < MI; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t A = (i << (j + 1)) + t, B = A + (1u << j), sr = (j == LOGMI) ? 0u : rot(i, j);
                uint32_t d[RI];
#pragma unroll
                for (uint32_t a = 0; a < RI; a++) {
                    const uint32_t za = X[A][a], zb = X[B][a];
                    X[A][a] = add(za, zb);
                    // Z_B = (Z_A - Z_B) * w^-sr: position a - sr takes the difference at a, negated on wrap-around
                    d[a] = (a >= sr) ? sub(za, zb) : sub(zb, za);
                }
#pragma unroll
                for (uint32_t a = 0; a < RI; a++) X[B][(a - sr) & (RI - 1)] = d[a];
            }
        }
        // recombination with u^MI = w (NTT.cu:271-276)
        const TwPair one{1u, T::C::MU32};
#pragma unroll
        for (uint32_t i = 0; i < MI; i++)
#pragma unroll
            for (uint32_t a = 0; a < RI; a++) {
                const uint32_t v = (a == 0) ? sub(X[i][0], X[MI + i][RI - 1]) : add(X[i][a], X[MI + i][a - 1]);
                z[MI * a + i] = LAZYQ ? T::smul_shoup(v, one) : T::csub(T::mul_shoup(v, fix), Q);
            }
    }
};

// ---- row products on the FP64 pipe ------------------------------------------------------------------
// B200 issues DFMA at 64 lanes/clk/SM: twice the rate of the 32x32->64 integer multiply-add (IMAD.WIDE, 32
// lanes/clk/SM, profiles/ubench_*.json) and on a pipe of its own, so the row products stop competing with the
// stage arithmetic for the integer multiply pipe.  Exactness: both operands are first brought to [-q/2, 3q/2)
// (one signed Shoup multiplication each, which also carries the deferred halvings); a product is then below
// 2.25 q^2 and ANY partial sum of the 32 terms of an output below 72 q^2 < 2^53 — every DFMA is exact, in any
// order.  The sum is reduced with the rounded quotient t = rint(acc / q) (magic-number rounding) and
// r = acc - t q, again exact, |r| <= q/2 + 1.  q < 2^25 only (the signed-lazy sets).
template <int SET> struct NussRowF64 {
    using T = Tile<SET>;
    static constexpr uint32_t Q = Cfg<SET>::Q;
    static constexpr bool OK = T::LAZY && 72.0 * (double)Q * (double)Q < 9007199254740992.0;
    // z = x (*) y mod (w^32 + 1, q) for two's-complement x, y in [-q/2, 3q/2); z in [-q/2 - 1, q/2 + 1].
    // xr, yr are the rows where they lie (shared memory); z overwrites x.  Outer-product order — step j multiplies
    // x_j into all accumulators, so consecutive DFMAs are independent — in two halves of 16 outputs: y (64
    // registers) + 16 accumulators (32) stay live instead of y + 32 accumulators, x_j is loaded and converted
    // when its step comes (twice in total).  The __syncwarp between the halves keeps the compiler from merging
    // them again.
    // VEC: the rows start on 16-byte boundaries (device, row stride 36): 128-bit loads and stores
    template <bool VEC = false> static QT_HD void product(uint32_t* xr, const uint32_t* yr) {
        double yd[32];
        uint32_t xw[32];
        if (VEC) {
#pragma unroll
            for (uint32_t c = 0; c < 8; c++) {
                const U4 u = reinterpret_cast<const U4*>(yr)[c];
                yd[4 * c] = (double)(int32_t)u.x; yd[4 * c + 1] = (double)(int32_t)u.y;
                yd[4 * c + 2] = (double)(int32_t)u.z; yd[4 * c + 3] = (double)(int32_t)u.w;
            }
        } else {
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) yd[j] = (double)(int32_t)yr[j];
        }
        const double INVQ = 1.0 / (double)Q, QD = (double)Q, MAGIC = 6755399441055744.0;  // 1.5 * 2^52
        uint32_t z0[16];
#pragma unroll
        for (uint32_t h = 0; h < 2; h++) {
            double acc[16];
#pragma unroll
            for (uint32_t kk = 0; kk < 16; kk++) acc[kk] = 0.0;
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) {
                if (VEC && j % 4 == 0) {
                    const U4 u = reinterpret_cast<const U4*>(xr)[j / 4];
                    xw[j] = u.x; xw[j + 1] = u.y; xw[j + 2] = u.z; xw[j + 3] = u.w;
                }
                const double xd = (double)(int32_t)(VEC ? xw[j] : xr[j]);
#pragma unroll
                for (uint32_t kk = 0; kk < 16; kk++) {  // wrapped terms enter negated
                    const uint32_t k = 16 * h + kk;
                    acc[kk] = (j <= k) ? fma(xd, yd[(k - j) & 31], acc[kk]) : fma(-xd, yd[(32 + k - j) & 31], acc[kk]);
                }
            }
#pragma unroll
            for (uint32_t kk = 0; kk < 16; kk++) {
                const double t = fma(acc[kk], INVQ, MAGIC) - MAGIC;  // rint(acc / q)
#if defined(__CUDA_ARCH__) && defined(QT_NUSS_F64_MAGIC_F2I)
                // A/B variant (off; run r02D: n=1024 97.3 vs 96.6, n=512 229.1 vs 232.2 M polymul/s — no gain): |r| <= q/2 + 1, so its
                // integer value is the low word of r + 1.5 * 2^52 — one more FP64 add instead of a conversion instruction
                const uint32_t r = (uint32_t)__double2loint(fma(-t, QD, acc[kk]) + MAGIC);
#else
                const uint32_t r = (uint32_t)(int32_t)fma(-t, QD, acc[kk]);
#endif
                if (h == 0) z0[kk] = r;
                else if (VEC) xw[kk] = r;
                else xr[16 + kk] = r;
            }
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
        }
        if (VEC) {
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                reinterpret_cast<U4*>(xr)[c] = U4{z0[4 * c], z0[4 * c + 1], z0[4 * c + 2], z0[4 * c + 3]};
                reinterpret_cast<U4*>(xr)[4 + c] = U4{xw[4 * c], xw[4 * c + 1], xw[4 * c + 2], xw[4 * c + 3]};
            }
        } else {
#pragma unroll
            for (uint32_t kk = 0; kk < 16; kk++) xr[kk] = z0[kk];
        }
    }
};

template <int SET, int RING> struct Nuss {
    using K = NussCfg<SET>;
    using O = NussOps<SET, RING>;
    static constexpr uint32_t M = K::M, R = K::R, LOGM = K::LOGM, ROWS = K::ROWS, Q = K::Q;
    static constexpr uint32_t EPL = R / 32;  // coefficients per lane in the stage phases (1 or 2)

    static QT_HD uint32_t brev(uint32_t x, uint32_t bits) {
#if defined(__CUDA_ARCH__)
        return bits ? __brev(x) >> (32u - bits) : 0u;  // one BREV + shift instead of a run-time loop per lane
#else
        uint32_t r = 0;
        for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
        return r;
#endif
    }
    // rotation exponent of stage j, group i  (sr of NTT.cu:200-203)
    static QT_HD uint32_t rot(uint32_t i, uint32_t j) { return (brev(i, LOGM - j) << j) * K::ROT_UNIT; }

    // phase 0: global -> rows (both copies).  tid strides over the n coefficients of polynomial p.
    static QT_HD void load(uint32_t tid, uint32_t nthreads, const uint32_t* gx, const uint32_t* gy,
                           uint32_t* sx, uint32_t* sy, bool lift = false) {
        for (uint32_t g = tid; g < K::N; g += nthreads) {
            const uint32_t i = g % M, j = g / M;
            uint32_t vx = gx[g], vy = gy[g];
            if (RING == 0 && lift) { vx = NussOps<SET, 0>::lift_in(vx); vy = NussOps<SET, 0>::lift_in(vy); }
            sx[i * K::XS + j] = vx; sx[(i + M) * K::XS + j] = vx;
            sy[i * K::YS + j] = vy; sy[(i + M) * K::YS + j] = vy;
        }
    }

    // forward stage j, one row butterfly bf (0..m-1) of one operand, executed by one warp:
    // read part (all lanes), then write part.
    struct Regs { uint32_t t[2], vi[2]; };
    static QT_HD void fwd_read(uint32_t lane, uint32_t j, uint32_t bf, const uint32_t* v, uint32_t stride, Regs& rg) {
        const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
        const uint32_t I = (i << (j + 1)) + t, L = I + (1u << j), sr = rot(i, j);
#pragma unroll
        for (uint32_t e = 0; e < EPL; e++) {
            const uint32_t a = lane + 32 * e;
            const uint32_t src = v[L * stride + ((a - sr) & (R - 1))];
            rg.t[e] = (a >= sr) ? src : O::neg(src);  // X_l rotated by w^sr (NTT.cu:210-215)
            rg.vi[e] = v[I * stride + a];
        }
    }
    static QT_HD void fwd_write(uint32_t lane, uint32_t j, uint32_t bf, uint32_t* v, uint32_t stride, const Regs& rg) {
        const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
        const uint32_t I = (i << (j + 1)) + t, L = I + (1u << j);
#pragma unroll
        for (uint32_t e = 0; e < EPL; e++) {
            const uint32_t a = lane + 32 * e;
            v[L * stride + a] = O::sub(rg.vi[e], rg.t[e]);  // NTT.cu:217-220
            v[I * stride + a] = O::add(rg.vi[e], rg.t[e]);
        }
    }

    // inverse stage j (0..logm), row butterfly bf, on Z (stored in the X rows)
    static QT_HD void inv_read(uint32_t lane, uint32_t j, uint32_t bf, const uint32_t* z, Regs& rg) {
        const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
        const uint32_t A = (i << (j + 1)) + t, B = A + (1u << j);
#pragma unroll
        for (uint32_t e = 0; e < EPL; e++) {
            const uint32_t a = lane + 32 * e;
            const uint32_t za = z[A * K::XS + a], zb = z[B * K::XS + a];
            rg.t[e] = O::half(O::sub(za, zb));   // NTT.cu:254-259
            rg.vi[e] = O::half(O::add(za, zb));
        }
    }
    static QT_HD void inv_write(uint32_t lane, uint32_t j, uint32_t bf, uint32_t* z, const Regs& rg) {
        const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
        const uint32_t A = (i << (j + 1)) + t, B = A + (1u << j), sr = (j == LOGM) ? 0u : rot(i, j);
#pragma unroll
        for (uint32_t e = 0; e < EPL; e++) {
            const uint32_t a = lane + 32 * e;
            z[A * K::XS + a] = rg.vi[e];
            // Z_B[a'] = T[a'+sr] (a' < r-sr) ; = -T[a'-(r-sr)] otherwise   (NTT.cu:262-267)
            z[B * K::XS + ((a - sr) & (R - 1))] = (a >= sr) ? rg.t[e] : O::neg(rg.t[e]);
        }
    }

    // product phase: thread owns row `row` of one polynomial; x, y rows -> z row (written over x)
    static QT_HD void product(uint32_t* xr /* row of X, becomes Z */, uint32_t* yr /* row of Y */) {
        if (R == 32) {
            uint32_t x[32], y[32];
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) { x[j] = xr[j]; y[j] = yr[j]; }
            if (RING == 0) {
                // naive, NTT.cu:147-165: chain A over j<=k, chain B over j>k, z = A - B
#pragma unroll
                for (uint32_t k = 0; k < 32; k++) {
                    uint32_t A = NussOps<SET, 0>::fold((uint64_t)x[0] * y[k]), B = 0;
#pragma unroll
                    for (uint32_t j = 1; j < 32; j++) {
                        if (j <= k) A = NussOps<SET, 0>::fold((uint64_t)x[j] * y[(k - j) & 31] + A);
                        else B = NussOps<SET, 0>::fold((uint64_t)x[j] * y[(32 + k - j) & 31] + B);
                    }
                    xr[k] = NussOps<SET, 0>::sub(A, B);
                }
            } else {
                uint32_t ny[32];
#pragma unroll
                for (uint32_t j = 0; j < 32; j++) ny[j] = Q - y[j];  // wrapped terms enter negated
#pragma unroll
                for (uint32_t k = 0; k < 32; k++) {
                    uint64_t acc0 = 0, acc1 = 0;  // 16 terms each: 16*q^2 < 2^64 even for the 30-bit q
#pragma unroll
                    for (uint32_t j = 0; j < 32; j++) {
                        const uint32_t yy = (j <= k) ? y[(k - j) & 31] : ny[(32 + k - j) & 31];
                        if (j < 16) acc0 += (uint64_t)x[j] * yy;
                        else acc1 += (uint64_t)x[j] * yy;
                    }
                    xr[k] = (uint32_t)((acc0 % Q + acc1 % Q) % Q);
                }
            }
        } else if (RING == 0) {  // R == 64, ring 2^32-1: x in registers, y indexed in shared memory
            uint32_t x[64];
#pragma unroll
            for (uint32_t j = 0; j < 64; j++) x[j] = xr[j];
            for (uint32_t k = 0; k < 64; k++) {  // naive (NTT.cu:147-165) with n = 64: chains A (j<=k), B (j>k) in j order
                uint32_t A = NussOps<SET, 0>::fold((uint64_t)x[0] * yr[k]), B = 0;
#pragma unroll
                for (uint32_t j = 1; j < 64; j++) {
                    const bool in_a = j <= k;
                    const uint32_t t = NussOps<SET, 0>::fold((uint64_t)x[j] * yr[(k - j) & 63u] + (in_a ? A : B));
                    A = in_a ? t : A;
                    B = in_a ? B : t;
                }
                xr[k] = NussOps<SET, 0>::sub(A, B);
            }
        } else {  // R == 64, Z_q: y row doubled in place [q - y | y], x in registers
            uint32_t x[64];
#pragma unroll
            for (uint32_t j = 0; j < 64; j++) x[j] = xr[j];
            for (uint32_t j = 0; j < 64; j++) { const uint32_t v = yr[j]; yr[64 + j] = v; yr[j] = Q - v; }
            for (uint32_t k = 0; k < 64; k++) {
                uint64_t acc[4] = {0, 0, 0, 0};
#pragma unroll
                for (uint32_t j = 0; j < 64; j++) acc[j >> 4] += (uint64_t)x[j] * yr[64 + k - j];
                xr[k] = (uint32_t)((acc[0] % Q + acc[1] % Q + acc[2] % Q + acc[3] % Q) % Q);
            }
        }
    }

    // product phase, recursive form (Z_q): the row product is split once more (NussInner)
    static QT_HD void product_recursive(uint32_t* xr, const uint32_t* yr) {
        using IN = NussInner<SET, R, false>;
        uint32_t x[R], y[R], zz[R];
#pragma unroll
        for (uint32_t j = 0; j < R; j++) { x[j] = xr[j]; y[j] = yr[j]; }
        IN::product(x, y, zz, tw_unsigned_c(IN::fix_value(1u), Q));
#pragma unroll
        for (uint32_t j = 0; j < R; j++) xr[j] = zz[j];
    }

    // final phase: recombination and coalesced store (NTT.cu:271-276)
    static QT_HD void store(uint32_t tid, uint32_t nthreads, const uint32_t* z, uint32_t* gz, bool lift = false) {
        for (uint32_t g = tid; g < K::N; g += nthreads) {
            const uint32_t i = g % M, j = g / M;
            uint32_t v;
            if (j == 0) v = O::sub(z[i * K::XS], z[(M + i) * K::XS + R - 1]);
            else v = O::add(z[i * K::XS + j], z[(M + i) * K::XS + j - 1]);
            if (RING == 0 && lift) v = NussOps<SET, 0>::lift_out(v);
            gz[g] = v;
        }
    }
};

#if defined(__CUDACC__)

// LIFT (ring 2^32-1 only): QT_RING_2P32M1_LIFT_Q — centred Z_q operands in, signed lift reduced mod q out
template <int SET, int RING, bool REC = false, bool LIFT = false>
__global__ void __launch_bounds__(NussCfg<SET>::THREADS)
k_nussbaumer(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    static_assert(!REC || RING == 1, "recursive row products exist for Z_q only");
    static_assert(!LIFT || RING == 0, "the lift belongs to the ring 2^32-1");
    using K = NussCfg<SET>;
    using NU = Nuss<SET, RING>;
    extern __shared__ uint4 nuss_smem_raw[];
    uint32_t* smem = reinterpret_cast<uint32_t*>(nuss_smem_raw);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t WARPS = K::THREADS / 32;
    const size_t ngroups = (batch + K::P - 1) / K::P;
    for (size_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const size_t p0 = grp * K::P;
        const uint32_t np = (uint32_t)((batch - p0 < K::P) ? batch - p0 : K::P);
        // load: the threads of the CTA are split evenly over the polynomials of the group
        {
            const uint32_t per = K::THREADS / K::P, p = tid / per;
            if (p < np)
                NU::load(tid % per, per, x + (p0 + p) * K::N, y + (p0 + p) * K::N, smem + p * K::POLY_WORDS,
                         smem + p * K::POLY_WORDS + K::X_WORDS, LIFT);
        }
        __syncthreads();
        // forward stages: P polys x 2 operands x m row butterflies per stage, one warp each
        for (int j = (int)K::LOGM - 1; j >= 0; j--) {
            for (uint32_t w = warp; w < K::P * 2 * K::M; w += WARPS) {
                const uint32_t p = w / (2 * K::M), op = (w / K::M) & 1, bf = w % K::M;
                if (p >= np) continue;
                uint32_t* v = smem + p * K::POLY_WORDS + (op ? K::X_WORDS : 0);
                const uint32_t stride = op ? K::YS : K::XS;
                typename NU::Regs rg;
                NU::fwd_read(lane, (uint32_t)j, bf, v, stride, rg);
                __syncwarp();
                NU::fwd_write(lane, (uint32_t)j, bf, v, stride, rg);
            }
            __syncthreads();
        }
        // products: one thread per row
        {
            const uint32_t p = tid / K::ROWS, row = tid % K::ROWS;
            if (p < np) {
                uint32_t* xr = smem + p * K::POLY_WORDS + row * K::XS;
                uint32_t* yr = smem + p * K::POLY_WORDS + K::X_WORDS + row * K::YS;
                if constexpr (REC) NU::product_recursive(xr, yr);
                else NU::product(xr, yr);
            }
        }
        __syncthreads();
        // inverse stages on Z (in the X rows)
        for (uint32_t j = 0; j <= K::LOGM; j++) {
            for (uint32_t w = warp; w < K::P * K::M; w += WARPS) {
                const uint32_t p = w / K::M, bf = w % K::M;
                if (p >= np) continue;
                uint32_t* zr = smem + p * K::POLY_WORDS;
                typename NU::Regs rg;
                NU::inv_read(lane, j, bf, zr, rg);
                __syncwarp();
                NU::inv_write(lane, j, bf, zr, rg);
            }
            __syncthreads();
        }
        {
            const uint32_t per = K::THREADS / K::P, p = tid / per;
            if (p < np) NU::store(tid % per, per, smem + p * K::POLY_WORDS, z + (p0 + p) * K::N, LIFT);
        }
        __syncthreads();
    }
}

// ---- warp-resident variant for the 32-column splits (n = 512, 1024) ------------------------------
// One WARP per polynomial.  Stage phases: lane = coefficient index a (0..31), registers = rows, so
// every row butterfly is register-to-register, the negacyclic rotation w^sr is ONE warp shuffle with a
// compile-time lane offset and the sign a lane predicate — no shared memory, no barrier.  Product
// phase: the rows are transposed through shared memory (stride 33, conflict-free both ways) so that
// lane = row; each lane multiplies its row(s) privately: 32x32 multiply-accumulates from registers.
// Z_q mode accumulates 64-bit products lazily and Montgomery-reduces once per accumulator; the common
// factor 2^-32 is removed by one Shoup multiplication per output coefficient at the very end.
// MODE: row products — 0 schoolbook on the integer pipe (the reference's structure), 1 recursive (NussInner),
// 2 schoolbook on the FP64 pipe (NussRowF64)
template <int SET, int RING, int MODE = 0> struct NussWarp {
    static constexpr bool REC = MODE == 1, F64 = MODE == 2;
    using K = NussCfg<SET>;
    using T = Tile<SET>;
    static_assert(MODE == 0 || RING == 1, "recursive / FP64 row products exist for Z_q only");
    static constexpr bool LAZYQ = (RING == 1) && T::LAZY;  // signed-lazy Z_q (see NussOpsLazy)
    static_assert(!F64 || (LAZYQ && NussRowF64<SET>::OK), "FP64 row products need q < 2^25");
    using O = typename std::conditional<LAZYQ, NussOpsLazy<SET>, NussOps<SET, RING>>::type;
    static constexpr uint32_t M = K::M, LOGM = K::LOGM, ROWS = K::ROWS, Q = K::Q;
    // ranges of the lazy variant: 32 products of two forward outputs in an int64; 3 * 2^(LOGM+1) q in an int32
    static_assert(!LAZYQ || (uint64_t)Q * Q < (1ull << (63 - 5 - 2 * LOGM)), "lazy product accumulator");
    static_assert(!LAZYQ || ((uint64_t)(3u << (LOGM + 1)) * Q < (1ull << 31)), "lazy inverse range");
#ifndef QT_NUSS_SIGN_MAD
#define QT_NUSS_SIGN_MAD 0  // measured (run r02w): 80.9 vs 87.5 M polymul/s at n=1024, 210.8 vs 219.4 at n=512 — rejected, kept for A/B
#endif
    static constexpr bool SIGN_MAD = QT_NUSS_SIGN_MAD && F64;            // signed rotations as multiply-adds (forward() below)
    // shift-based reduction before the FP64 row products (k_nussbaumer_warp): (2^LOGM + 2) (q - 2^QS) below q/2, and the inverse
    // stages (additions only, 2^(LOGM+2) times a row-product output of at most q/2 + 1) inside an int32
    static constexpr bool F64_SHIFT_OK = !F64 || ((((uint64_t)1 << LOGM) + 2) * (Q - (1u << T::QS)) < Q / 2 &&
                                                  ((uint64_t)4 << LOGM) * (Q / 2 + 2) < (1ull << 31));
#ifndef QT_NUSS_F64_RS
#define QT_NUSS_F64_RS 36
#endif
    // row stride in shared memory: 33 words = conflict-free along a row and down a column; the FP64 rows use 36 — rows start on
    // 16-byte boundaries, so a lane reads and writes its row with 128-bit accesses (a quarter of the shared-memory instructions),
    // which are conflict-free too (lane l covers banks 4l .. 4l+3 mod 32), as is the column access of the stage phases
    // (run r02D: n=1024 96.4 vs 92.9 M polymul/s; n=512, one row per lane: 230.5 vs 232.2 — the 64-row sets only)
#ifndef QT_NUSS_RS_ALL
#define QT_NUSS_RS_ALL 1  // every 64-row Z_q warp kernel (schoolbook 77.2 vs 75.3, recursive 79.2 vs 78.1, p-I recursive 57.3 vs 56.6 M
#endif                    // polymul/s), not only the FP64 rows; the ring 2^32-1 kernels lose (47.3 vs 50.0: spills at 168 registers)
    static constexpr uint32_t RS = ((F64 || (QT_NUSS_RS_ALL && RING == 1)) && ROWS == 64) ? QT_NUSS_F64_RS : 33;
    static constexpr uint32_t WARP_WORDS = 2 * ROWS * RS;                // X rows then Y rows
#ifndef QT_NUSS_WARPS
#define QT_NUSS_WARPS 12
#endif
    // 12 x 16.9 KiB of rows, <= 170 registers: one CTA per SM.  The recursive row products want more registers
    // (two 8x8 row blocks live at once): 8 warps with up to 255 registers measured faster for the 64-row sets
    // (qTESLA-III 78.5 vs 74.6 M polymul/s, p-I 57.7 vs 53.1; qTESLA-I, 32 rows: 182.8 vs 189.7).
#ifndef QT_NUSS_F64_WARPS
#define QT_NUSS_F64_WARPS 12  // 158 / 168 registers, no spills (8 warps: 78.6 vs 87.0 M polymul/s at n=1024)
#endif
    static constexpr uint32_t WARPS = F64 ? QT_NUSS_F64_WARPS : (REC && ROWS == 64) ? 8 : QT_NUSS_WARPS;
    static constexpr size_t SMEM_BYTES = (size_t)WARPS * WARP_WORDS * sizeof(uint32_t);
    static constexpr uint32_t RPL = ROWS / 32;                           // rows per lane in the product phase
    // terms a 64-bit accumulator may take before a Montgomery reduction: sum < q * 2^32
    static constexpr uint32_t TERMS = (T::QCAP >= 32) ? 32 : (T::QCAP >= 8 ? 8 : 4);

    static __device__ __forceinline__ uint32_t rot(uint32_t i, uint32_t j) {
        return (c_bitrev(i, LOGM - j) << j) * K::ROT_UNIT;
    }

    static __device__ __forceinline__ void forward(uint32_t (&v)[ROWS], uint32_t lane) {
#pragma unroll
        for (int j = (int)LOGM - 1; j >= 0; j--) {
#pragma unroll
            for (uint32_t bf = 0; bf < M; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t I = (i << (j + 1)) + t, L = I + (1u << j), sr = rot(i, (uint32_t)j);
                uint32_t tv = v[L];
                const uint32_t vi = v[I];
                if (SIGN_MAD && sr != 0) {
                    // A/B variant (off): the sign of the rotated row as a +-1 multiplier, a row butterfly = SHFL + 2 multiply-adds
                    // instead of SHFL + negate + select + add + subtract.  Fewer instructions, but slower (see QT_NUSS_SIGN_MAD).
                    const uint32_t src = __shfl_sync(0xffffffffu, tv, (lane - sr) & 31u);
                    const uint32_t sg = (lane >= sr) ? 1u : 0xFFFFFFFFu;
                    v[I] = src * sg + vi;
                    v[L] = src * (0u - sg) + vi;
                    continue;
                }
                if (sr != 0) {  // X_l * w^sr: coefficient a comes from a - sr, negated on wrap-around
                    const uint32_t src = __shfl_sync(0xffffffffu, tv, (lane - sr) & 31u);
                    tv = (lane >= sr) ? src : O::neg(src);
                }
                v[L] = O::sub(vi, tv);
                v[I] = O::add(vi, tv);
            }
        }
    }

    // both operands through the stages TOGETHER: one copy of the lane predicates and shuffle indices, twice the independent work
    // between a shuffle and its use (A/B: QT_NUSS_F64_PAIRED)
    static __device__ __forceinline__ void forward2(uint32_t (&v)[ROWS], uint32_t (&u)[ROWS], uint32_t lane) {
#pragma unroll
        for (int j = (int)LOGM - 1; j >= 0; j--) {
#pragma unroll
            for (uint32_t bf = 0; bf < M; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t I = (i << (j + 1)) + t, L = I + (1u << j), sr = rot(i, (uint32_t)j);
                uint32_t tv = v[L], tu = u[L];
                const uint32_t vi = v[I], ui = u[I];
                if (sr != 0) {
                    const uint32_t sv = __shfl_sync(0xffffffffu, tv, (lane - sr) & 31u), su = __shfl_sync(0xffffffffu, tu, (lane - sr) & 31u);
                    const bool keep = lane >= sr;
                    tv = keep ? sv : O::neg(sv);
                    tu = keep ? su : O::neg(su);
                }
                v[L] = O::sub(vi, tv); v[I] = O::add(vi, tv);
                u[L] = O::sub(ui, tu); u[I] = O::add(ui, tu);
            }
        }
    }

    static __device__ __forceinline__ void inverse(uint32_t (&z)[ROWS], uint32_t lane) {
#pragma unroll
        for (uint32_t j = 0; j <= LOGM; j++) {
#pragma unroll
            for (uint32_t bf = 0; bf < M; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t A = (i << (j + 1)) + t, B = A + (1u << j);
                const uint32_t sr = (j == LOGM) ? 0u : rot(i, j);
                const uint32_t za = z[A], zb = z[B];
                uint32_t tv = O::half(O::sub(za, zb));
                z[A] = O::half(O::add(za, zb));
                if (sr != 0) {  // Z_B = T * w^-sr: coefficient a takes T[a + sr], negated on wrap-around
                    const uint32_t src = __shfl_sync(0xffffffffu, tv, (lane + sr) & 31u);
                    if (SIGN_MAD) tv = src * ((lane < 32u - sr) ? 1u : 0xFFFFFFFFu);  // (lazy: half() is the identity)
                    else tv = (lane < 32u - sr) ? src : O::neg(src);
                }
                z[B] = tv;
            }
        }
    }

    // recursive row products (REC): ranges of the signed-lazy flavour.  The forward stages leave |v| <= 2^LOGM * |input|;
    // where two uncentred operands would overflow the 64-bit accumulator of the inner products, x is centred to
    // (-q/2, q/2] when it is loaded (CENTRE_X; two instructions per coefficient).
    using Inner = NussInner<SET, 32, LAZYQ>;
    static constexpr bool CENTRE_X = REC && LAZYQ && !Inner::fits((uint64_t)Q << LOGM, (uint64_t)Q << LOGM);
    static_assert(!(REC && LAZYQ) || Inner::fits((uint64_t)(CENTRE_X ? Q / 2 + 1 : Q) << LOGM, (uint64_t)Q << LOGM),
                  "inner product accumulator");

    // one row: z = x (*) y negacyclic, length 32; x, y, z are shared-memory rows (z overwrites x)
    // VEC: the rows start on 16-byte boundaries (this kernel's own row stride; the block-pass kernel has its own and passes false)
    template <bool VEC = (RS % 4 == 0)> static __device__ __forceinline__ void product_row(uint32_t* xr, const uint32_t* yr) {
        if constexpr (F64) NussRowF64<SET>::template product<VEC>(xr, yr);
        else product_row_int<VEC>(xr, yr);
    }
    template <bool VEC> static __device__ __forceinline__ void product_row_int(uint32_t* xr, const uint32_t* yr) {
        uint32_t x[32], y[32];
        if (VEC) {
#pragma unroll
            for (uint32_t c = 0; c < 8; c++) {
                const U4 u = reinterpret_cast<const U4*>(xr)[c], w = reinterpret_cast<const U4*>(yr)[c];
                x[4 * c] = u.x; x[4 * c + 1] = u.y; x[4 * c + 2] = u.z; x[4 * c + 3] = u.w;
                y[4 * c] = w.x; y[4 * c + 1] = w.y; y[4 * c + 2] = w.z; y[4 * c + 3] = w.w;
            }
        } else {
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) { x[j] = xr[j]; y[j] = yr[j]; }
        }
        uint32_t o4[4];  // outputs leave four at a time
#define QT_NUSS_PUT(k, val)                                                                             \
    do {                                                                                                \
        if (VEC) {                                                                                      \
            o4[(k) & 3] = (val);                                                                        \
            if (((k) & 3) == 3) reinterpret_cast<U4*>(xr)[(k) >> 2] = U4{o4[0], o4[1], o4[2], o4[3]};   \
        } else {                                                                                        \
            xr[k] = (val);                                                                              \
        }                                                                                               \
    } while (0)
        if (REC) {
            // same output convention as the schoolbook branches below: LAZYQ — product * 2^-(LOGM+1) in
            // [-q/2, 3q/2); canonical — product * 2^-32 in [0, q)
            constexpr uint32_t EXTRA = LAZYQ ? c_powmod((Q + 1) / 2, LOGM + 1, Q) : c_powmod(T::C::R_MODQ, Q - 2, Q);
            constexpr uint32_t FIX = Inner::fix_value(EXTRA);
            const TwPair fix = LAZYQ ? tw_signed_c(FIX, Q) : tw_unsigned_c(FIX, Q);
            uint32_t zz[32];
            Inner::product(x, y, zz, fix);
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) QT_NUSS_PUT(j, zz[j]);
        } else if (RING == 0) {
            // naive (NTT.cu:147-165) keeps two chains per output, A over j <= k and B over j > k, folding
            // after every term.  A chain's final value is a function of the INTEGER sum T of its products
            // only: it is congruent to T, lies in [0, 2^32-1], and is 0 exactly when T = 0 (a fold of a
            // non-zero value is never 0).  So each chain is summed exactly in three 32-bit limbs (one
            // multiply-add-with-carry per term) and folded once with the same end-around-carry adds — bit for
            // bit the reference's result, including which representation of zero comes out.
#ifndef QT_NUSS_RING_HALVES
#define QT_NUSS_RING_HALVES 0  // measured (run r02D): 39.0 vs 50.0 M polymul/s at n=1024, 88.9 vs 136.9 at n=512 — bit-exact, but two wide
#endif                          // multiplies per term cost more multiply-pipe time than one multiply + two carry additions; kept for A/B
            using O0 = NussOps<SET, 0>;
            if (QT_NUSS_RING_HALVES) {
                // A/B variant (off).  The same exact sums without a carry chain: y = ylo + 2^16 yhi, and a chain's T = L + 2^16 H with L = sum x ylo,
                // H = sum x yhi — at most 32 products below 2^48 each, so L and H fit 64-bit accumulators as they are (one
                // IMAD.WIDE per half and term instead of multiply + two carry additions).  In Z/(2^32-1) 2^32 = 1 and 2^16 is a
                // rotation by 16 bits, end-around-carry addition is associative and never turns a non-zero operand into 0, so
                // fold(L) (+) rotl16(fold(H)) is the same word as the fold of T's own limbs — including which zero comes out.
                uint32_t ylo[32], yhi[32];
#pragma unroll
                for (uint32_t j = 0; j < 32; j++) { ylo[j] = y[j] & 0xFFFFu; yhi[j] = y[j] >> 16; }
#pragma unroll
                for (uint32_t k = 0; k < 32; k++) {
                    uint64_t al = 0, ah = 0, bl = 0, bh = 0;
#pragma unroll
                    for (uint32_t j = 0; j < 32; j++) {
                        if (j <= k) { al += (uint64_t)x[j] * ylo[(k - j) & 31]; ah += (uint64_t)x[j] * yhi[(k - j) & 31]; }
                        else { bl += (uint64_t)x[j] * ylo[(32 + k - j) & 31]; bh += (uint64_t)x[j] * yhi[(32 + k - j) & 31]; }
                    }
                    const uint32_t fa = O0::fold(ah), fb = O0::fold(bh);
                    const uint32_t A = O0::add(O0::fold(al), __funnelshift_l(fa, fa, 16));
                    const uint32_t B = O0::add(O0::fold(bl), __funnelshift_l(fb, fb, 16));
                    QT_NUSS_PUT(k, O0::sub(A, B));
                }
            } else {
#pragma unroll
            for (uint32_t k = 0; k < 32; k++) {
                uint32_t a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
                for (uint32_t j = 0; j < 32; j++) {
                    if (j <= k)
                        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                            : "+r"(a0), "+r"(a1), "+r"(a2) : "r"(x[j]), "r"(y[(k - j) & 31]));
                    else
                        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                            : "+r"(b0), "+r"(b1), "+r"(b2) : "r"(x[j]), "r"(y[(32 + k - j) & 31]));
                }
                QT_NUSS_PUT(k, O0::sub(O0::add(O0::add(a0, a1), a2), O0::add(O0::add(b0, b1), b2)));
            }
            }
        } else if (LAZYQ) {
            // 2^32 (the Montgomery factor) times 2^-(LOGM+1) (every halving of the inverse stages), signed Shoup form
            constexpr uint32_t FIX = c_mulmod(T::C::R_MODQ, c_powmod((Q + 1) / 2, LOGM + 1, Q), Q);
            const TwPair fix = tw_signed_c(FIX, Q);
#pragma unroll
            for (uint32_t k = 0; k < 32; k++) {
                int64_t acc = 0;
#pragma unroll
                for (uint32_t j = 0; j < 32; j++) {
                    const int64_t pr = (int64_t)(int32_t)x[j] * (int64_t)(int32_t)((j <= k) ? y[(k - j) & 31] : y[(32 + k - j) & 31]);
                    acc = (j <= k) ? acc + pr : acc - pr;  // wrapped terms enter negated
                }
                const uint32_t m = (uint32_t)acc * (0u - T::C::QINV_NEG);                 // lo(acc) * q^-1
                const uint32_t red = (uint32_t)(acc >> 32) - (uint32_t)mulhi32s(m, Q);   // acc * 2^-32, |red| < 2^30
                QT_NUSS_PUT(k, T::smul_shoup(red, fix));                                  // [-q/2, 3q/2)
            }
        } else {
            uint32_t ny[32];
#pragma unroll
            for (uint32_t j = 0; j < 32; j++) ny[j] = Q - y[j];  // wrapped terms enter negated
#pragma unroll
            for (uint32_t k = 0; k < 32; k++) {
                uint32_t r = 0;
#pragma unroll
                for (uint32_t g = 0; g < 32 / TERMS; g++) {
                    uint64_t acc = 0;
#pragma unroll
                    for (uint32_t jj = 0; jj < TERMS; jj++) {
                        const uint32_t j = g * TERMS + jj;
                        acc += (uint64_t)x[j] * ((j <= k) ? y[(k - j) & 31] : ny[(32 + k - j) & 31]);
                    }
                    const uint32_t m = (uint32_t)acc * T::C::QINV_NEG;          // Montgomery: acc * 2^-32
                    const uint32_t red = (uint32_t)((acc + (uint64_t)m * Q) >> 32);  // in [0, 2q)
                    r = (g == 0) ? red : r + red;
                }
                // canonical: r < (32/TERMS)*2q
                if (32 / TERMS == 1) r = T::csub(r, Q);
                else r = T::csub(T::fold2q(r), Q);
                QT_NUSS_PUT(k, r);
            }
        }
#undef QT_NUSS_PUT
    }
};

#ifndef QT_NUSS_PREFETCH
#define QT_NUSS_PREFETCH 1
#endif
#ifndef QT_NUSS_F64_SHIFT_RED
#define QT_NUSS_F64_SHIFT_RED 1
#endif
#ifndef QT_NUSS_UNIFORM_WARP
#define QT_NUSS_UNIFORM_WARP 1  // the warp index as a provably warp-uniform value (see qt_kernels.cuh: warp_index); run r02F: ring
                                // 2^32-1 52.5 vs 49.7, n=512 FP64 rows 236.1 vs 232.7, n=1024 FP64 rows 96.4 vs 96.3 M polymul/s
#endif
template <int SET, int RING, int MODE = 0, bool LIFT = false>
__global__ void __launch_bounds__(NussWarp<SET, RING, MODE>::WARPS * 32)
k_nussbaumer_warp(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    static_assert(!LIFT || RING == 0, "the lift belongs to the ring 2^32-1");
    using W = NussWarp<SET, RING, MODE>;
    using K = NussCfg<SET>;
    using O = typename W::O;
    using T = Tile<SET>;
    static_assert(K::R == 32, "warp-resident Nussbaumer needs 32 columns");
    extern __shared__ uint4 nuss_smem_raw[];
    const uint32_t warp = QT_NUSS_UNIFORM_WARP ? __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) : threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* sx = reinterpret_cast<uint32_t*>(nuss_smem_raw) + warp * W::WARP_WORDS;
    uint32_t* sy = sx + W::ROWS * W::RS;
    // 2^32 mod q with its Shoup companion: removes the Montgomery factor of the products
    const TwPair rfix{T::C::R_MODQ, (uint32_t)(((uint64_t)T::C::R_MODQ << 32) / T::Q)};
    for (size_t p = (size_t)warp * gridDim.x + blockIdx.x; p < batch; p += (size_t)gridDim.x * W::WARPS) {  // SM-interleaved
        const uint32_t* gx = x + p * K::N;
        const uint32_t* gy = y + p * K::N;
#if QT_NUSS_PREFETCH
        if (p + (size_t)gridDim.x * W::WARPS < batch) {  // the next polynomial of this warp: its lines towards L2 now
#if QT_NUSS_PREFETCH == 2
            asm volatile("prefetch.global.L1 [%0];" ::"l"(gx + (size_t)gridDim.x * W::WARPS * K::N + K::M * lane));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(gy + (size_t)gridDim.x * W::WARPS * K::N + K::M * lane));
#else
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gx + (size_t)gridDim.x * W::WARPS * K::N + K::M * lane));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gy + (size_t)gridDim.x * W::WARPS * K::N + K::M * lane));
#endif
        }
#endif
        uint32_t v[W::ROWS];
#ifndef QT_NUSS_F64_PAIRED
#define QT_NUSS_F64_PAIRED 0  // measured (run r02G): 90.8 vs 96.6 M polymul/s at n=1024, 227.6 vs 236.7 at n=512 — rejected, kept for A/B
#endif
        constexpr bool PAIRED = QT_NUSS_F64_PAIRED && W::F64 && QT_NUSS_F64_SHIFT_RED && !LIFT;
        if constexpr (PAIRED) {
            uint32_t u[W::ROWS];
            const uint4* gxv = reinterpret_cast<const uint4*>(gx + K::M * lane);
            const uint4* gyv = reinterpret_cast<const uint4*>(gy + K::M * lane);
#pragma unroll
            for (uint32_t c = 0; c < K::M / 4; c++) {
                const uint4 a = gxv[c], b = gyv[c];
                v[4 * c] = a.x; v[4 * c + 1] = a.y; v[4 * c + 2] = a.z; v[4 * c + 3] = a.w;
                u[4 * c] = b.x; u[4 * c + 1] = b.y; u[4 * c + 2] = b.z; u[4 * c + 3] = b.w;
            }
#pragma unroll
            for (uint32_t i = 0; i < K::M; i++) { v[i + K::M] = v[i]; u[i + K::M] = u[i]; }
            W::forward2(v, u, lane);
#pragma unroll
            for (uint32_t r = 0; r < W::ROWS; r++) {
                v[r] -= (uint32_t)((int32_t)v[r] >> T::QS) * T::Q;
                u[r] -= (uint32_t)((int32_t)u[r] >> T::QS) * T::Q;
            }
#pragma unroll
            for (uint32_t r = 0; r < W::ROWS; r++) { sx[r * W::RS + lane] = v[r]; sy[r * W::RS + lane] = u[r]; }
        } else
#pragma unroll 1
        for (int op = 0; op < 2; op++) {
            const uint4* g = reinterpret_cast<const uint4*>((op ? gy : gx) + K::M * lane);  // X_i[a] = x[m*a + i]
#pragma unroll
            for (uint32_t c = 0; c < K::M / 4; c++) {
                const uint4 u = g[c];
                v[4 * c] = u.x; v[4 * c + 1] = u.y; v[4 * c + 2] = u.z; v[4 * c + 3] = u.w;
            }
            if (LIFT) {  // QT_RING_2P32M1_LIFT_Q: centred representatives of the Z_q operands
#pragma unroll
                for (uint32_t i = 0; i < K::M; i++) v[i] = NussOps<SET, 0>::lift_in(v[i]);
            }
            if (W::CENTRE_X) {
                const uint32_t thr = op ? 0xFFFFFFFFu : T::Q / 2;  // x only
#pragma unroll
                for (uint32_t i = 0; i < K::M; i++) v[i] -= (v[i] > thr) ? T::Q : 0u;
            }
#pragma unroll
            for (uint32_t i = 0; i < K::M; i++) v[i + K::M] = v[i];  // rows m..2m-1 are copies (NTT.cu:187-191)
            W::forward(v, lane);
            if constexpr (W::F64) {
                // FP64 row products take operands in [-q/2, 3q/2).
#ifndef QT_NUSS_F64_SHIFT_RED
#define QT_NUSS_F64_SHIFT_RED 1
#endif
                if (QT_NUSS_F64_SHIFT_RED) {
                    // q = 2^QS + d with a small d and |v| <= 2^LOGM q after the forward stages: k = v >> QS is within one of
                    // v / q, so v - k q = (v mod 2^QS) - k d lies in (-(2^LOGM + 2) d, 2^QS + (2^LOGM + 2) d) — a shift and ONE
                    // multiply-add per coefficient instead of the three multiplies of a Shoup reduction (W::F64_SHIFT_OK).  The
                    // halvings of the inverse stages, which the Shoup form carried on x, move to the final multiplication.
                    static_assert(W::F64_SHIFT_OK, "shift-based reduction: range");
#pragma unroll
                    for (uint32_t r = 0; r < W::ROWS; r++) v[r] -= (uint32_t)((int32_t)v[r] >> T::QS) * T::Q;
                } else {
                    // x also takes 2^-(LOGM+1), the halvings of the inverse stages
                    const TwPair one{1u, T::C::MU32}, halves = tw_signed_c(c_powmod((T::Q + 1) / 2, K::LOGM + 1, T::Q), T::Q);
                    const TwPair sw{op ? one.w : halves.w, op ? one.ws : halves.ws};
#pragma unroll
                    for (uint32_t r = 0; r < W::ROWS; r++) v[r] = T::smul_shoup(v[r], sw);
                }
            }
            uint32_t* s = op ? sy : sx;
#pragma unroll
            for (uint32_t r = 0; r < W::ROWS; r++) s[r * W::RS + lane] = v[r];
        }
        __syncwarp();
#pragma unroll 1
        for (uint32_t h = 0; h < W::RPL; h++) W::product_row(sx + (lane + 32 * h) * W::RS, sy + (lane + 32 * h) * W::RS);
        __syncwarp();
#pragma unroll
        for (uint32_t r = 0; r < W::ROWS; r++) v[r] = sx[r * W::RS + lane];
        __syncwarp();
        W::inverse(v, lane);
        // recombination (NTT.cu:271-276): z[m*a + i] = Z_i[a] + Z_{m+i}[a-1]; a = 0 wraps with a sign
        uint4* gz = reinterpret_cast<uint4*>(z + p * K::N + K::M * lane);
#pragma unroll
        for (uint32_t c = 0; c < K::M / 4; c++) {
            uint32_t o[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t i = 4 * c + k;
                const uint32_t up = __shfl_sync(0xffffffffu, v[K::M + i], (lane - 1) & 31u);
                uint32_t r = (lane == 0) ? O::sub(v[i], up) : O::add(v[i], up);
                if (W::F64 && QT_NUSS_F64_SHIFT_RED) r = T::scanon(T::smul_shoup(r, tw_signed_c(c_powmod((T::Q + 1) / 2, K::LOGM + 1, T::Q), T::Q)));  // the deferred halvings
                else if (W::LAZYQ) r = T::scanon(T::smul_shoup(r, TwPair{1u, T::C::MU32}));  // any |r| < 2^31 -> [0, q)
                else if (RING == 1) r = T::csub(T::mul_shoup(r, rfix), T::Q);
                else if (LIFT) r = NussOps<SET, 0>::lift_out(r);
                o[k] = r;
            }
            gz[c] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- block-pass warp kernel (m = 32: n = 1024, 2048) --------------------------------------------------------------
// One WARP per polynomial, like k_nussbaumer_warp, but the 2m = 64 rows are processed as two BLOCKS of 32: every forward
// stage j < log2 m and every inverse stage j < log2 m pairs rows inside one block (rows m..2m-1 start as copies of rows
// 0..m-1, NTT.cu:187-191, so the first stage is the copy itself), only the last inverse stage and the recombination join
// block 0 with block 1.  Per block: load, forward stages in registers (rotation = warp shuffles), rows to shared
// memory, row products (lane = row, all 32 lanes busy), rows back, inverse stages; block 0's result is parked in the
// warp's own shared-memory slots while block 1 runs.  What this buys over the whole-polynomial warp kernel:
//   * 32 * (r/32) row registers live instead of 64 * (r/32): n = 2048 (r = 64: two coefficients per lane) fits a warp at
//     all — it used to run on the shared-memory kernel k_nussbaumer with one thread per row and CTA-wide barriers;
//   * Z_q: block 1 is block 0's transform of the TWISTED rows (row i rotated by w^(i r/m) first: the standard split of a
//     length-2m transform into the even and the odd outputs), so the SAME stage code serves both blocks in a rolled loop —
//     half the straight-line code of the stage phases (the instruction-fetch stall of the whole-polynomial kernel).
//     The ring 2^32-1 keeps the reference's own rotation amounts and operation order (two compile-time copies), because
//     there the representation of zero depends on it.
template <int SET, int RING, int MODE = 0, bool LIFT = false> struct NussBlk {
    using K = NussCfg<SET>;
    using T = Tile<SET>;
    using W = NussWarp<SET, RING, MODE>;
    using O = typename W::O;
    static constexpr uint32_t M = K::M, R = K::R, LOGM = K::LOGM, EPL = R / 32, Q = K::Q, UNIT = K::ROT_UNIT;
    static_assert(M == 32 && (EPL == 1 || EPL == 2), "block-pass kernel: 32 rows per block, 32 or 64 columns");
    static constexpr bool LAZYQ = W::LAZYQ, F64 = W::F64, REC = W::REC;
    static_assert(R == 32 || (!LAZYQ && !F64), "64-column rows exist for the 30-bit modulus only");
    static_assert(R == 32 || RING == 0 || REC, "64-column Z_q rows: recursive row products");
    static constexpr bool TWIST = RING == 1;  // block 1 = twisted block 0 (exact in Z_q; see above)
    static constexpr uint32_t RS = R + 1;                 // row stride: conflict-free along a row and down a column
    static constexpr uint32_t BLK_WORDS = M * RS;
    static constexpr uint32_t WARP_WORDS = 3 * BLK_WORDS;  // X rows | Y rows | parked result of block 0
#ifndef QT_NUSS_BLK_WARPS
#define QT_NUSS_BLK_WARPS 16
#endif
#ifndef QT_NUSS_BLK_WARPS_F64
#define QT_NUSS_BLK_WARPS_F64 12
#endif
    // registers: the three-limb ring products and the FP64 rows want ~170 (12 warps), the recursive rows up to 255 (8 warps)
    static constexpr uint32_t WARPS = R == 64 ? 8 : ((F64 || RING == 0) ? QT_NUSS_BLK_WARPS_F64 : (REC ? 8 : QT_NUSS_BLK_WARPS));
    static constexpr size_t SMEM_BYTES = (size_t)WARPS * WARP_WORDS * sizeof(uint32_t);

    struct Row { uint32_t c[EPL]; };  // coefficients lane + 32 e of one row

    // row * w^s, s in [0, r): coefficient a takes (a - s) mod r, negated on wrap-around (NTT.cu:210-215)
    static __device__ __forceinline__ Row rot_fwd(Row t, uint32_t s, uint32_t lane) {
        const int d = (int)lane - (int)s;
        Row o;
        if (EPL == 1) {
            const uint32_t src = __shfl_sync(0xffffffffu, t.c[0], d & 31);
            o.c[0] = d >= 0 ? src : O::neg(src);
        } else {
            const uint32_t a0 = __shfl_sync(0xffffffffu, t.c[0], d & 31), a1 = __shfl_sync(0xffffffffu, t.c[EPL - 1], d & 31);
            const bool sw = (d >> 5) & 1;  // floor(d / 32) odd: the source sits in the other half
            const uint32_t r0 = sw ? a1 : a0, r1 = sw ? a0 : a1;
            o.c[0] = d >= 0 ? r0 : O::neg(r0);
            o.c[EPL - 1] = d >= -32 ? r1 : O::neg(r1);
        }
        return o;
    }
    // row * w^-s: coefficient a takes (a + s) mod r, negated on wrap-around (NTT.cu:262-267)
    static __device__ __forceinline__ Row rot_inv(Row t, uint32_t s, uint32_t lane) {
        const uint32_t d = lane + s;
        Row o;
        if (EPL == 1) {
            const uint32_t src = __shfl_sync(0xffffffffu, t.c[0], d & 31);
            o.c[0] = d < 32 ? src : O::neg(src);
        } else {
            const uint32_t a0 = __shfl_sync(0xffffffffu, t.c[0], d & 31), a1 = __shfl_sync(0xffffffffu, t.c[EPL - 1], d & 31);
            const bool sw = (d >> 5) & 1;
            const uint32_t r0 = sw ? a1 : a0, r1 = sw ? a0 : a1;
            o.c[0] = d < 64 ? r0 : O::neg(r0);
            o.c[EPL - 1] = d < 32 ? r1 : O::neg(r1);
        }
        return o;
    }
    // rotation exponent of stage j, LOCAL group i (of 2^(LOGM-1-j)) of block H:  rot(i + H 2^(LOGM-1-j), j) of the
    // whole-polynomial numbering = ((brev(i) + H) << j) * UNIT  (the block bit is the top bit of the group index)
    static __host__ __device__ constexpr uint32_t rot(uint32_t i, uint32_t j, uint32_t H) {
        return ((c_bitrev(i, LOGM - j) + H) << j) * UNIT;
    }

    template <uint32_t H> static __device__ __forceinline__ void forward(Row (&v)[M], uint32_t lane) {
#pragma unroll
        for (int j = (int)LOGM - 1; j >= 0; j--) {
#pragma unroll
            for (uint32_t bf = 0; bf < M / 2; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t I = (i << (j + 1)) + t, L = I + (1u << j), sr = rot(i, (uint32_t)j, H);
                Row tv = v[L];
                if (sr != 0) tv = rot_fwd(tv, sr, lane);
#pragma unroll
                for (uint32_t e = 0; e < EPL; e++) {
                    const uint32_t vi = v[I].c[e];
                    v[L].c[e] = O::sub(vi, tv.c[e]);
                    v[I].c[e] = O::add(vi, tv.c[e]);
                }
            }
        }
    }
    template <uint32_t H> static __device__ __forceinline__ void inverse(Row (&z)[M], uint32_t lane) {
#pragma unroll
        for (uint32_t j = 0; j < LOGM; j++) {
#pragma unroll
            for (uint32_t bf = 0; bf < M / 2; bf++) {
                const uint32_t i = bf >> j, t = bf & ((1u << j) - 1);
                const uint32_t A = (i << (j + 1)) + t, B = A + (1u << j), sr = rot(i, j, H);
                Row tv;
#pragma unroll
                for (uint32_t e = 0; e < EPL; e++) {
                    const uint32_t za = z[A].c[e], zb = z[B].c[e];
                    tv.c[e] = O::half(O::sub(za, zb));
                    z[A].c[e] = O::half(O::add(za, zb));
                }
                if (sr != 0) tv = rot_inv(tv, sr, lane);
                z[B] = tv;
            }
        }
    }

    // one row product, rows where they lie in shared memory (z overwrites x)
    static __device__ __forceinline__ void product_row(uint32_t* xr, uint32_t* yr) {
        if constexpr (R == 32) W::template product_row<false>(xr, yr);  // (row stride R + 1 here)
        else if constexpr (RING == 0) Nuss<SET, 0>::product(xr, yr);
        else Nuss<SET, 1>::product_recursive(xr, yr);
    }

#ifndef QT_NUSS_BLK_LOCKSTEP
#define QT_NUSS_BLK_LOCKSTEP 0
#endif
    // LOCKSTEP (64-column rows only): the warps of a CTA enter every phase together (CTA barriers), so that the straight-line code
    // of a phase — the row products alone are ~100 KiB — is fetched once per CTA and phase instead of once per drifting warp
    static constexpr bool LOCKSTEP = QT_NUSS_BLK_LOCKSTEP && R == 64;
    static __device__ __forceinline__ void phase_barrier() {
        if (LOCKSTEP) __syncthreads();
    }
    // everything of block H up to its inverse stages; result in v.  `act`: this warp has a polynomial (LOCKSTEP: warps without one
    // still walk through the barriers)
    template <uint32_t H> static __device__ __forceinline__ void block(Row (&v)[M], const uint32_t* gx, const uint32_t* gy,
                                                                     uint32_t* sx, uint32_t* sy, uint32_t lane, bool twist, bool act = true) {
        phase_barrier();
        if (act) {
#pragma unroll 1
        for (int op = 0; op < 2; op++) {
#pragma unroll
            for (uint32_t e = 0; e < EPL; e++) {
                const uint4* g = reinterpret_cast<const uint4*>((op ? gy : gx) + M * (lane + 32 * e));  // X_i[a] = x[m a + i]
#pragma unroll
                for (uint32_t c = 0; c < M / 4; c++) {
                    const uint4 u = g[c];
                    v[4 * c].c[e] = u.x; v[4 * c + 1].c[e] = u.y; v[4 * c + 2].c[e] = u.z; v[4 * c + 3].c[e] = u.w;
                }
            }
            if (LIFT) {
#pragma unroll
                for (uint32_t i = 0; i < M; i++)
#pragma unroll
                    for (uint32_t e = 0; e < EPL; e++) v[i].c[e] = NussOps<SET, 0>::lift_in(v[i].c[e]);
            }
            if (W::CENTRE_X) {
                const uint32_t thr = op ? 0xFFFFFFFFu : T::Q / 2;  // x only
#pragma unroll
                for (uint32_t i = 0; i < M; i++) v[i].c[0] -= (v[i].c[0] > thr) ? T::Q : 0u;
            }
            if (TWIST && twist) {  // block 1: row i * w^(i r/m), then block 0's stage code
#pragma unroll
                for (uint32_t i = 1; i < M; i++) v[i] = rot_fwd(v[i], i * UNIT, lane);
            }
            forward<H>(v, lane);
            if constexpr (F64) {
                // FP64 row products take operands in [-q/2, 3q/2); x also takes 2^-(LOGM+1), the halvings of the inverse stages
                const TwPair one{1u, T::C::MU32}, halves = tw_signed_c(c_powmod((T::Q + 1) / 2, LOGM + 1, T::Q), T::Q);
                const TwPair sw{op ? one.w : halves.w, op ? one.ws : halves.ws};
#pragma unroll
                for (uint32_t i = 0; i < M; i++) v[i].c[0] = T::smul_shoup(v[i].c[0], sw);
            }
            uint32_t* s = op ? sy : sx;
#pragma unroll
            for (uint32_t i = 0; i < M; i++)
#pragma unroll
                for (uint32_t e = 0; e < EPL; e++) s[i * RS + lane + 32 * e] = v[i].c[e];
        }
        __syncwarp();
        }
        phase_barrier();
        if (act) {
        product_row(sx + lane * RS, sy + lane * RS);
        __syncwarp();
        }
        phase_barrier();
        if (act) {
#pragma unroll
        for (uint32_t i = 0; i < M; i++)
#pragma unroll
            for (uint32_t e = 0; e < EPL; e++) v[i].c[e] = sx[i * RS + lane + 32 * e];
        __syncwarp();
        inverse<H>(v, lane);
        if (TWIST && twist) {
#pragma unroll
            for (uint32_t i = 1; i < M; i++) v[i] = rot_inv(v[i], i * UNIT, lane);
        }
        }
    }
};

template <int SET, int RING, int MODE = 0, bool LIFT = false>
__global__ void __launch_bounds__(NussBlk<SET, RING, MODE, LIFT>::WARPS * 32)
k_nussbaumer_blk(const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch) {
    using NB = NussBlk<SET, RING, MODE, LIFT>;
    using K = NussCfg<SET>;
    using O = typename NB::O;
    using T = Tile<SET>;
    using Row = typename NB::Row;
    static_assert(!LIFT || RING == 0, "the lift belongs to the ring 2^32-1");
    constexpr uint32_t M = NB::M, R = NB::R, EPL = NB::EPL, RS = NB::RS;
    extern __shared__ uint4 nuss_smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;  // (uniform index measured: 15.4 vs 15.5 M polymul/s)
    uint32_t* sx = reinterpret_cast<uint32_t*>(nuss_smem_raw) + warp * NB::WARP_WORDS;
    uint32_t* sy = sx + NB::BLK_WORDS;
    uint32_t* sp = sy + NB::BLK_WORDS;
    // 2^32 mod q with its Shoup companion: removes the Montgomery factor of the canonical 32-column row products
    const TwPair rfix{T::C::R_MODQ, (uint32_t)(((uint64_t)T::C::R_MODQ << 32) / T::Q)};
    const size_t stride = (size_t)gridDim.x * NB::WARPS;
    // LOCKSTEP: a CTA-uniform loop (warp 0's polynomial index decides), warps past the end of the batch only keep the barriers company
    for (size_t p0 = blockIdx.x; p0 < batch; p0 += stride) {
        const size_t p = p0 + (size_t)warp * gridDim.x;  // SM-interleaved
        const bool act = p < batch;
        if (!NB::LOCKSTEP && !act) break;
        const uint32_t* gx = x + (act ? p : 0) * K::N;
        const uint32_t* gy = y + (act ? p : 0) * K::N;
        if (p + stride < batch) {  // next polynomial of this warp: its lines towards L2 while this one is computed
#pragma unroll
            for (uint32_t e = 0; e < EPL; e++) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(gx + stride * K::N + M * (lane + 32 * e)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(gy + stride * K::N + M * (lane + 32 * e)));
            }
        }
        Row v[M];
        if constexpr (NB::TWIST) {
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {  // one copy of the stage code for both blocks
                NB::template block<0>(v, gx, gy, sx, sy, lane, h != 0, act);
                if (h == 0 && act) {
#pragma unroll
                    for (uint32_t i = 0; i < M; i++)
#pragma unroll
                        for (uint32_t e = 0; e < EPL; e++) sp[i * RS + lane + 32 * e] = v[i].c[e];  // parked in this lane's own slots
                }
            }
        } else {
            NB::template block<0>(v, gx, gy, sx, sy, lane, false, act);
            if (act) {
#pragma unroll
                for (uint32_t i = 0; i < M; i++)
#pragma unroll
                    for (uint32_t e = 0; e < EPL; e++) sp[i * RS + lane + 32 * e] = v[i].c[e];
            }
            NB::template block<1>(v, gx, gy, sx, sy, lane, false, act);
        }
        if (!act) continue;
        // last inverse stage (rows i and m+i, no rotation, NTT.cu:241-269 with j = log2 m) and the recombination
        // z[m a + i] = Z_i[a] + Z_{m+i}[a-1], a = 0 wraps with a sign (NTT.cu:271-276)
#pragma unroll
        for (uint32_t c = 0; c < M / 4; c++) {
            uint32_t o[EPL][4];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t i = 4 * c + k;
                Row za, zb, up;
#pragma unroll
                for (uint32_t e = 0; e < EPL; e++) {
                    const uint32_t a = sp[i * RS + lane + 32 * e], b = v[i].c[e];
                    zb.c[e] = O::half(O::sub(a, b));
                    za.c[e] = O::half(O::add(a, b));
                }
                // up[a] = Z_{m+i}[a - 1]
                const uint32_t u0 = __shfl_sync(0xffffffffu, zb.c[0], (lane - 1) & 31u);
                const uint32_t u1 = EPL == 2 ? __shfl_sync(0xffffffffu, zb.c[EPL - 1], (lane - 1) & 31u) : u0;
                up.c[0] = (EPL == 2 && lane == 0) ? u1 : u0;  // a = 0 takes coefficient r - 1 (subtracted)
                if (EPL == 2) up.c[EPL - 1] = lane == 0 ? u0 : u1;  // a = 32 takes coefficient 31
#pragma unroll
                for (uint32_t e = 0; e < EPL; e++) {
                    uint32_t r = (e == 0 && lane == 0) ? O::sub(za.c[e], up.c[e]) : O::add(za.c[e], up.c[e]);
                    if (NB::LAZYQ) r = T::scanon(T::smul_shoup(r, TwPair{1u, T::C::MU32}));  // any |r| < 2^31 -> [0, q)
                    else if (RING == 1 && R == 32) r = T::csub(T::mul_shoup(r, rfix), T::Q);
                    else if (LIFT) r = NussOps<SET, 0>::lift_out(r);
                    o[e][k] = r;
                }
            }
#pragma unroll
            for (uint32_t e = 0; e < EPL; e++)
                reinterpret_cast<uint4*>(z + p * K::N + M * (lane + 32 * e))[c] = make_uint4(o[e][0], o[e][1], o[e][2], o[e][3]);
        }
        __syncwarp();  // the parked slots and the row buffers are rewritten by the next polynomial
    }
}

// Row products of the Z_q kernels: schoolbook (the reference's structure) or split once more (NussInner).
// QT_NUSS_AUTO_RECURSIVE is what "automatic" picks.
#ifndef QT_NUSS_AUTO_RECURSIVE
#define QT_NUSS_AUTO_RECURSIVE 1
#endif
#ifndef QT_NUSS_AUTO_F64
#define QT_NUSS_AUTO_F64 1
#endif
enum : int { NUSS_AUTO = 0, NUSS_SCHOOLBOOK = 1, NUSS_RECURSIVE = 2, NUSS_FP64 = 3,
             NUSS_WHOLE = 16 };  // flag: the whole-polynomial kernels of round 1 instead of the block-pass kernel (A/B)
#ifndef QT_NUSS_BLK
#define QT_NUSS_BLK 1  // the block-pass warp kernel serves the m = 32 sets (n = 1024, 2048)
#endif
// measured (run r02f): n = 2048 15.6 vs 12.2 M polymul/s (Z_q, recursive rows) and 6.9 vs 5.8 (ring 2^32-1) against the shared-memory
// kernel; n = 1024 it LOSES to the whole-polynomial warp kernel (FP64 rows 78.8 vs 86.1, ring 41.0 vs 48.7: the second pass
// re-reads the operands and the twist costs 3 x 31 extra rotations), so it serves the 64-column split only
#ifndef QT_NUSS_BLK_ALL
#define QT_NUSS_BLK_ALL 0
#endif
template <int SET> constexpr bool nuss_has_blk() { return QT_NUSS_BLK && NussCfg<SET>::M == 32 && (QT_NUSS_BLK_ALL || NussCfg<SET>::R == 64); }
template <int SET> constexpr bool nuss_has_f64() { return NussCfg<SET>::R == 32 && NussRowF64<SET>::OK; }

template <class Kern> int nuss_prepare(Kern k, int threads, size_t smem, int* occ_min) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, threads, smem);
    if (e != cudaSuccess) return (int)e;
    *occ_min = o < *occ_min ? o : *occ_min;
    return 0;
}

template <int SET> int nuss_setup(int num_sms, int* grid) {
    using K = NussCfg<SET>;
    int occ = 64, rc = 0;
    if constexpr (nuss_has_blk<SET>()) {  // grid = #SMs (one CTA each); the occupancy of these is not part of `grid`
        int o1 = 64;
        if ((rc = nuss_prepare(k_nussbaumer_blk<SET, 0, 0>, NussBlk<SET, 0, 0>::WARPS * 32, NussBlk<SET, 0, 0>::SMEM_BYTES, &o1))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer_blk<SET, 0, 0, true>, NussBlk<SET, 0, 0, true>::WARPS * 32, NussBlk<SET, 0, 0, true>::SMEM_BYTES, &o1))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer_blk<SET, 1, 1>, NussBlk<SET, 1, 1>::WARPS * 32, NussBlk<SET, 1, 1>::SMEM_BYTES, &o1))) return rc;
        if constexpr (K::R == 32) {
            if ((rc = nuss_prepare(k_nussbaumer_blk<SET, 1, 0>, NussBlk<SET, 1, 0>::WARPS * 32, NussBlk<SET, 1, 0>::SMEM_BYTES, &o1))) return rc;
            if constexpr (nuss_has_f64<SET>())
                if ((rc = nuss_prepare(k_nussbaumer_blk<SET, 1, 2>, NussBlk<SET, 1, 2>::WARPS * 32, NussBlk<SET, 1, 2>::SMEM_BYTES, &o1))) return rc;
        }
        if (o1 < 1) return -4;
    }
    if constexpr (K::R == 32) {  // warp-resident kernels
        using W = NussWarp<SET, 0>;      // (each kernel with its OWN geometry: the row stride, and with it the shared-memory
        using WS = NussWarp<SET, 1, 0>;  //  size, differs between the ring and the Z_q kernels)
        using WR = NussWarp<SET, 1, 1>;
        if ((rc = nuss_prepare(k_nussbaumer_warp<SET, 0, 0>, W::WARPS * 32, W::SMEM_BYTES, &occ))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer_warp<SET, 0, 0, true>, W::WARPS * 32, W::SMEM_BYTES, &occ))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer_warp<SET, 1, 0>, WS::WARPS * 32, WS::SMEM_BYTES, &occ))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer_warp<SET, 1, 1>, WR::WARPS * 32, WR::SMEM_BYTES, &occ))) return rc;
        if constexpr (nuss_has_f64<SET>()) {
            using WF = NussWarp<SET, 1, 2>;
            if ((rc = nuss_prepare(k_nussbaumer_warp<SET, 1, 2>, WF::WARPS * 32, WF::SMEM_BYTES, &occ))) return rc;
        }
    } else {
        if ((rc = nuss_prepare(k_nussbaumer<SET, 0, false>, K::THREADS, K::SMEM_BYTES, &occ))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer<SET, 0, false, true>, K::THREADS, K::SMEM_BYTES, &occ))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer<SET, 1, false>, K::THREADS, K::SMEM_BYTES, &occ))) return rc;
        if ((rc = nuss_prepare(k_nussbaumer<SET, 1, true>, K::THREADS, K::SMEM_BYTES, &occ))) return rc;
    }
    *grid = (occ < 1 ? 1 : occ) * num_sms;
    return 0;
}

template <int SET>
int nuss_launch(int max_grid, const uint32_t* x, const uint32_t* y, uint32_t* z, size_t batch, int ring, int variant,
                cudaStream_t s, int grid_occ = 1) {
    using K = NussCfg<SET>;
    const bool whole = (variant & NUSS_WHOLE) != 0;
    variant &= ~NUSS_WHOLE;
    const uint32_t lift = ring == 2 ? 1u : 0u;  // QT_RING_2P32M1_LIFT_Q: the ring 2^32-1 kernels with the lift epilogue
    if (ring == 2) ring = 0;
    const bool f64 = ring == 1 && nuss_has_f64<SET>() && (variant == NUSS_FP64 || (variant == NUSS_AUTO && QT_NUSS_AUTO_F64));
    const bool rec = ring == 1 && !f64 && (variant == NUSS_RECURSIVE || (variant == NUSS_AUTO && QT_NUSS_AUTO_RECURSIVE));
    if (variant == NUSS_FP64 && ring == 1 && !nuss_has_f64<SET>()) return -4;  // QT_ERR_UNSUPPORTED
    if constexpr (nuss_has_blk<SET>()) {
        // block-pass warp kernel; 64-column Z_q rows exist in recursive form only (schoolbook on request: whole-polynomial kernel)
        const bool blk_ok = !whole && !(K::R == 64 && ring == 1 && !rec) && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 15) == 0;
        if (blk_ok) {
            const int sms = max_grid / (grid_occ < 1 ? 1 : grid_occ);
            const int g = (int)(batch < (size_t)sms ? batch : (size_t)sms);
#define QT_BLK_LAUNCH(RING_, MODE_, LIFT_)                                                                         \
    k_nussbaumer_blk<SET, RING_, MODE_, LIFT_><<<g, NussBlk<SET, RING_, MODE_, LIFT_>::WARPS * 32,                 \
                                                 NussBlk<SET, RING_, MODE_, LIFT_>::SMEM_BYTES, s>>>(x, y, z, batch)
            if (ring == 0 && lift) QT_BLK_LAUNCH(0, 0, true);
            else if (ring == 0) QT_BLK_LAUNCH(0, 0, false);
            else if (rec) QT_BLK_LAUNCH(1, 1, false);
            else if constexpr (K::R == 32) {
                if (f64) {
                    if constexpr (nuss_has_f64<SET>()) QT_BLK_LAUNCH(1, 2, false);
                } else QT_BLK_LAUNCH(1, 0, false);
            }
#undef QT_BLK_LAUNCH
            return (int)cudaGetLastError();
        }
    }
    if constexpr (K::R == 32) {
        if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 15) != 0) return -2;  // 128-bit accesses
        using W = NussWarp<SET, 0>;
        const int g = (int)(batch < (size_t)max_grid ? batch : (size_t)max_grid);  // small batches spread over all SMs
        if (ring == 0 && lift) k_nussbaumer_warp<SET, 0, 0, true><<<g, W::WARPS * 32, W::SMEM_BYTES, s>>>(x, y, z, batch);
        else if (ring == 0) k_nussbaumer_warp<SET, 0, 0><<<g, W::WARPS * 32, W::SMEM_BYTES, s>>>(x, y, z, batch);
        else if (f64) {
            if constexpr (nuss_has_f64<SET>())
                k_nussbaumer_warp<SET, 1, 2><<<g, NussWarp<SET, 1, 2>::WARPS * 32, NussWarp<SET, 1, 2>::SMEM_BYTES, s>>>(x, y, z, batch);
        } else if (rec) k_nussbaumer_warp<SET, 1, 1><<<g, NussWarp<SET, 1, 1>::WARPS * 32, NussWarp<SET, 1, 1>::SMEM_BYTES, s>>>(x, y, z, batch);
        else k_nussbaumer_warp<SET, 1, 0><<<g, NussWarp<SET, 1, 0>::WARPS * 32, NussWarp<SET, 1, 0>::SMEM_BYTES, s>>>(x, y, z, batch);
        return (int)cudaGetLastError();
    } else {
        const size_t groups = (batch + K::P - 1) / K::P;
        const int grid = (int)(groups < (size_t)max_grid ? groups : (size_t)max_grid);
        if (ring == 0 && lift) k_nussbaumer<SET, 0, false, true><<<grid, K::THREADS, K::SMEM_BYTES, s>>>(x, y, z, batch);
        else if (ring == 0) k_nussbaumer<SET, 0, false><<<grid, K::THREADS, K::SMEM_BYTES, s>>>(x, y, z, batch);
        else if (rec) k_nussbaumer<SET, 1, true><<<grid, K::THREADS, K::SMEM_BYTES, s>>>(x, y, z, batch);
        else k_nussbaumer<SET, 1, false><<<grid, K::THREADS, K::SMEM_BYTES, s>>>(x, y, z, batch);
        return (int)cudaGetLastError();
    }
}

#endif  // __CUDACC__

}  // namespace qt
