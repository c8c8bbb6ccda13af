// Native-mode arithmetic: counter-based Philox4x32-7 and the two normal generators.
#pragma once
#include <stdint.h>

namespace mcgp {

// Philox4x32-R (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// key = (seed_lo, seed_hi); counter = (sim_lo, sim_hi, lap<<8 | lane, race stream).
// R = MCGP_PHILOX_ROUNDS = 7: the paper's Table 2 lists Philox4x32-7 as the fastest variant that is Crush-resistant
// (passes SmallCrush, Crush and BigCrush); 10 is its conservative default.  A race consumes one call per lane per lap
// PAIR and the call is the largest single item of the lap budget (26 of 190 executed instructions per race-lap at
// R = 10), so the three safety-margin rounds are spent elsewhere; -DMCGP_PHILOX_ROUNDS=10 restores them.  The scalar
// mirror (oracle/native_mirror.c: MIRROR_PHILOX_ROUNDS) must use the same R; both are pinned to the Random123
// known-answer vectors for R = 7 and R = 10 (tests/test_native_mirror.py).
// The round keys depend only on the seed, so the host expands them once and passes them as a kernel
// parameter: they sit in the constant bank and feed the round's 3-input XOR (LOP3) as immediate-like operands
// instead of costing two uniform adds per round per call.
#ifndef MCGP_PHILOX_ROUNDS
#define MCGP_PHILOX_ROUNDS 7
#endif
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};

__host__ __device__ inline PhiloxKeys philox_expand_key(uint32_t seed_lo, uint32_t seed_hi) {
    PhiloxKeys k;
    for (int r = 0; r < 10; r++) {
        k.k0[r] = seed_lo + 0x9E3779B9u * (uint32_t)r;
        k.k1[r] = seed_hi + 0xBB67AE85u * (uint32_t)r;
    }
    return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < MCGP_PHILOX_ROUNDS; r++) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// ---- fast normals: Box-Muller on the MUFU unit (lg2 / sqrt / sin / cos approximations) --------
__device__ __forceinline__ float fast_radius(uint32_t w) {
    // u = ((w>>8)+1) * 2^-24 in (0,1];  r = sqrt(-2 ln u) = sqrt(-2 ln2 * (lg2(k) - 24))
    float lg;  // the argument is an integer in [1, 2^24]: no denormal guard needed around MUFU.LG2
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)((w >> 8) + 1u)));
    float r2 = fmaxf(__fmaf_rn(lg, -1.3862943611198906f, 33.27106466687737f), 0.0f);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
    return r;
}
__device__ __forceinline__ float fast_angle(uint32_t w) {
    return (float)(w >> 8) * 3.7450703e-07f;  // 2*pi / 2^24
}
__device__ __forceinline__ float fast_normal(uint32_t w1, uint32_t w2) {
    return fast_radius(w1) * __cosf(fast_angle(w2));
}
__device__ __forceinline__ void fast_normal2(uint32_t w1, uint32_t w2, float& za, float& zb) {
    float r = fast_radius(w1), a = fast_angle(w2);
    za = r * __cosf(a);
    zb = r * __sinf(a);
}

// ---- exact normals: only IEEE-754 add/mul/fma/sqrt and integer ops, so that the scalar CPU
// mirror (oracle/native_mirror.c) reproduces them bit for bit.  Polynomials: Cephes logf/sinf/cosf.
__device__ __forceinline__ float exact_log(float x) {  // x in [2^-24, 1]
    uint32_t b = __float_as_uint(x);
    int e = (int)(b >> 23) - 126;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f000000u);  // [0.5, 1)
    if (m < 0.70710678118654752440f) { e -= 1; m = __fadd_rn(__fadd_rn(m, m), -1.0f); }
    else m = __fadd_rn(m, -1.0f);
    float z = __fmul_rn(m, m);
    float y = 7.0376836292e-2f;
    y = __fmaf_rn(y, m, -1.1514610310e-1f);
    y = __fmaf_rn(y, m, 1.1676998740e-1f);
    y = __fmaf_rn(y, m, -1.2420140846e-1f);
    y = __fmaf_rn(y, m, 1.4249322787e-1f);
    y = __fmaf_rn(y, m, -1.6668057665e-1f);
    y = __fmaf_rn(y, m, 2.0000714765e-1f);
    y = __fmaf_rn(y, m, -2.4999993993e-1f);
    y = __fmaf_rn(y, m, 3.3333331174e-1f);
    y = __fmul_rn(__fmul_rn(y, m), z);
    float fe = (float)e;
    y = __fmaf_rn(-2.12194440e-4f, fe, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = __fadd_rn(m, y);
    return __fmaf_rn(0.693359375f, fe, r);
}
__device__ __forceinline__ void exact_normal2(uint32_t w1, uint32_t w2, float& za, float& zb) {
    float u = __fmul_rn((float)((w1 >> 8) + 1u), 5.9604644775390625e-08f);  // 2^-24, exact
    float rad = __fsqrt_rn(__fmul_rn(-2.0f, exact_log(u)));
    uint32_t k = w2 >> 8;                       // angle in 2^-24 turns
    uint32_t q = (k + (1u << 21)) >> 22;        // nearest quarter turn, 0..4
    int r = (int)k - (int)(q << 22);            // [-2^21, 2^21)
    float x = __fmul_rn((float)r, 3.7450703e-07f);
    float x2 = __fmul_rn(x, x);
    float pc = __fmaf_rn(2.443315711809948e-5f, x2, -1.388731625493765e-3f);
    pc = __fmaf_rn(pc, x2, 4.166664568298827e-2f);
    float cx = __fmaf_rn(__fmul_rn(x2, x2), pc, __fmaf_rn(-0.5f, x2, 1.0f));
    float ps = __fmaf_rn(-1.9515295891e-4f, x2, 8.3321608736e-3f);
    ps = __fmaf_rn(ps, x2, -1.6666654611e-1f);
    float sx = __fmaf_rn(__fmul_rn(x, x2), ps, x);
    float c, s;
    switch (q & 3u) {
        case 0: c = cx; s = sx; break;
        case 1: c = -sx; s = cx; break;
        case 2: c = -cx; s = -sx; break;
        default: c = sx; s = -cx; break;
    }
    za = __fmul_rn(rad, c);
    zb = __fmul_rn(rad, s);
}

}  // namespace mcgp
