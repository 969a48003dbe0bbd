// Philox4x32-7 counter-based RNG + normal transform, device side (7 rounds: the Crush-resistant minimum of
// Salmon et al., SC'11 - three rounds fewer than Random123's default is 3 of ~30 instructions per OU sample).
// Bit-for-bit the stream specified in oracle/philox.py (counter layout, 23-bit uniforms,
// Box-Muller with cos/sin ordering).  One Philox block feeds four consecutive samples.
#pragma once
#include <stdint.h>

namespace itjl {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

constexpr int kPhiloxRounds = 7;

constexpr uint32_t kStreamNoise = 0u;
constexpr uint32_t kStreamInit = 1u;
constexpr uint32_t kStreamPropose = 2u;    // PMC proposals (itjl_pmc_propose): counter = (attempt, 0, proposal index, c3)

__host__ __device__ __forceinline__ uint32_t philox_c3(uint32_t stream, uint64_t generation) {
    return ((stream & 0xFu) << 28) | (uint32_t)(generation & 0x0FFFFFFFull);
}

// The round keys (k0 + r W0, k1 + r W1) depend on the seed only: the host expands them once per
// launch into the kernel parameters, so the rounds read them as constant-bank operands instead of
// holding twenty registers (or re-adding the Weyl constants) in every thread.
struct PhiloxKeys {
    uint32_t k[2 * kPhiloxRounds];
};

__host__ __device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < kPhiloxRounds; ++r) {
        K.k[2 * r] = k0; K.k[2 * r + 1] = k1;
        k0 += kPhiloxW0; k1 += kPhiloxW1;
    }
    return K;
}

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& K, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ((r >> 9) + 0.5) * 2^-23 : exact in fp32, in (0, 1).  Built without an integer->float
// conversion: 0x3F800000 | (r >> 9) is the float 1 + (r >> 9) 2^-23 in [1, 2), and subtracting
// 1 - 2^-24 is exact (the result is an odd multiple of 2^-24 below 1).
__device__ __forceinline__ float u12(uint32_t r) {
    return __uint_as_float(0x3F800000u | (r >> 9));
}
__device__ __forceinline__ float u01(uint32_t r) {
    return u12(r) - 0.99999994f;
}

// Box-Muller on the SFU: lg2.approx / sqrt.approx / sin.approx / cos.approx (absolute error of
// a draw ~5e-7, far below the 1e-5 parity tolerance; the distribution is unaffected).
// z0 = R cos(2 pi v), z1 = R sin(2 pi v), R = sqrt(-2 ln u); the angle is folded to
// theta = 2 pi v - pi in (-pi, pi), where MUFU.SIN/COS are most accurate: cos(2 pi v) = -cos(theta).
// With w = u12(rb) = 1 + v - 2^-24:  theta = 2 pi w - 3 pi + 2 pi 2^-24.
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& z0, float& z1) {
    const float u = u01(ra);
    float l2, rad, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(l2 * -1.3862943611198906f));
    const float th = fmaf(u12(rb), 6.283185307179586f, -9.424777586262351f);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    z0 = -rad * c;
    z1 = -rad * s;
}

}  // namespace itjl
