// mcrt_numerics.h -- the numerical contract shared by the sm_100a kernels, the C++ host
// and the CPU oracle.
//
// Why this header exists: deterministic-mode parity requires bit-identical fp32/fp64 results on
// the CPU (oracle) and on the GPU (kernels).  IEEE add/sub/mul/div/sqrt are identical on both
// as long as no FMA contraction happens (device code is compiled with -fmad=false, host code
// with -ffp-contract=off), but libm transcendentals (glibc expf/logf/powf/sin/cos vs. CUDA's)
// are not.  Every transcendental the reference's hot path calls
//   std::exp(float)  ray.cpp:102, main.cpp:135      std::log(float)  ray.cpp:112
//   std::pow         ray.cpp:131,158,160,223        sin/cos (double) ray.cpp:181-182
// is therefore implemented here once, in plain IEEE double arithmetic (+,-,*,/, floor and the
// explicitly written IEEE fused multiply-add -- all correctly rounded, hence identical on both
// sides), and rounded to float where the reference rounds to float.  The double result has
// <= ~1 ulp(double) error, so the float rounding equals the correctly-rounded value except with
// probability ~2^-28 per call; tests/test_host_and_numerics.py pins these against glibc.
//
// Also here: the counter-based Philox4x32-10 generator that replaces the reference's
// per-call std::random_device + std::mt19937 (scene.cpp:132-135, ray.cpp:85-88,175-179,216-219).
#ifndef MCRT_NUMERICS_H
#define MCRT_NUMERICS_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define MCRT_HD __host__ __device__ __forceinline__
#else
#define MCRT_HD static inline
#endif

// ----------------------------------------------------------------------------------------------
// bit casts
// ----------------------------------------------------------------------------------------------
MCRT_HD uint64_t mc_d2u(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
MCRT_HD double mc_u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
MCRT_HD uint32_t mc_f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
MCRT_HD float mc_u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

#define MC_INF_D  (mc_u2d(0x7ff0000000000000ULL))
#define MC_NAN_D  (mc_u2d(0x7ff8000000000000ULL))

// glibc's M_PI; ray.cpp keeps it (it includes neither psf.h nor transducer.h, which redefine it).
#define MC_PI_D 3.14159265358979323846
// "#define M_PI 3.14159" of psf.h:9 and transducer.h:12
#define MC_PI_REDEFINED 3.14159

// 2^k as a double for -1022 <= k <= 1023
MCRT_HD double mc_pow2i(int k) { return mc_u2d((uint64_t)(k + 1023) << 52); }

// p * 2^k with k possibly outside the normal exponent range (two-step scaling, IEEE-exact ops)
MCRT_HD double mc_scale2(double p, int k)
{
    if (k > 1023) {
        p = p * mc_pow2i(1023); k -= 1023;
        if (k > 1023) k = 1023;
        return p * mc_pow2i(k);
    }
    if (k < -1022) {
        p = p * mc_pow2i(-1022); k += 1022;
        if (k < -1022) k = -1022;
        return p * mc_pow2i(k);
    }
    return p * mc_pow2i(k);
}

// ----------------------------------------------------------------------------------------------
// exp / log / pow in double, built from IEEE basic operations and the IEEE fused multiply-add
// ----------------------------------------------------------------------------------------------
// fma(a, b, c) is an IEEE-754 operation (one rounding): __fma_rn on the device and the hardware / libm fma on the host
// return the same bits, so using it EXPLICITLY keeps the contract (what is forbidden is the compiler contracting a
// separate multiply and add behind our back: -fmad=false / -ffp-contract=off stay).  Round 2: the polynomial cores
// below are near-minimax fits (scripts/gen_numerics_coeffs.py, mpmath) evaluated with FMAs in two interleaved
// Horner chains -- half the fp64 operations and a third of the dependent-chain depth of the round-1 Taylor/Horner
// forms, which made the shading phase of the trace kernels wait on the fp64 pipe (profiles/r02*_numerics*).
MCRT_HD double mc_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

MCRT_HD double mc_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return MC_INF_D;
    if (x < -745.2) return 0.0;
    const double LN2_HI = 6.93147180369123816490e-01;  // 0x3fe62e42fee00000, 32 significant bits
    const double LN2_LO = 1.90821492927058770002e-10;
    const double kf = floor(mc_fma(x, 1.44269504088896338700e+00, 0.5));
    const double r = mc_fma(-kf, LN2_LO, mc_fma(-kf, LN2_HI, x));   // |r| <= ln2/2 (+ rounding)
    // exp(r) = 1 + r + r^2 G(r); G = degree-9 near-minimax on |r| <= 0.34661: relative error of the sum < 1.7e-17
    const double G0 = 0x1.0000000000001p-1, G1 = 0x1.5555555555556p-3, G2 = 0x1.5555555553d63p-5, G3 = 0x1.11111111109b3p-7,
                 G4 = 0x1.6c16c1788bd90p-10, G5 = 0x1.a01a01a7c41d5p-13, G6 = 0x1.a019b90d2ae7ap-16, G7 = 0x1.71de0dae63bb3p-19,
                 G8 = 0x1.289185613a3d6p-22, G9 = 0x1.af38a9b0ec855p-26;
    const double w = r * r;
    double ge = mc_fma(w, G8, G6), go = mc_fma(w, G9, G7);          // G(r) = ge(w) + r go(w)
    ge = mc_fma(w, ge, G4); go = mc_fma(w, go, G5);
    ge = mc_fma(w, ge, G2); go = mc_fma(w, go, G3);
    ge = mc_fma(w, ge, G0); go = mc_fma(w, go, G1);
    const double g = mc_fma(r, go, ge);
    const double p = mc_fma(w, g, r) + 1.0;
    return mc_scale2(p, (int)kf);
}

MCRT_HD double mc_log(double x)
{
    if (x != x) return x;
    if (x < 0.0) return MC_NAN_D;
    if (x == 0.0) return -MC_INF_D;
    uint64_t u = mc_d2u(x);
    if (u == 0x7ff0000000000000ULL) return x;
    int e = 0;
    if ((u >> 52) == 0) {               // subnormal: scale by 2^54
        x = x * 18014398509481984.0;
        u = mc_d2u(x);
        e = -54;
    }
    e += (int)((u >> 52) & 0x7ff) - 1023;
    double m = mc_u2d((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);   // [1,2)
    if (m > 1.41421356237309514547) { m = m * 0.5; e += 1; }                   // [0.7071,1.4142]
    const double f = m - 1.0;
    const double s = f / (2.0 + f);     // |s| <= 0.1716
    const double z = s * s;
    // log(m) = 2 atanh(s) = 2s (1 + z Q(z)); Q = degree-6 near-minimax on z <= 0.029438: error of the bracket < 4.7e-18
    const double L0 = 0x1.5555555555558p-2, L1 = 0x1.99999999952d7p-3, L2 = 0x1.2492492df281ap-3, L3 = 0x1.c71c62e3f11e6p-4,
                 L4 = 0x1.7462b51cb66b1p-4, L5 = 0x1.39fe51a7c18f9p-4, L6 = 0x1.2b5900de53b32p-4;
    const double w = z * z;
    double qe = mc_fma(w, L6, L4), qo = mc_fma(w, L5, L3);          // Q(z) = qe(w) + z qo(w)
    qe = mc_fma(w, qe, L2); qo = mc_fma(w, qo, L1);
    qe = mc_fma(w, qe, L0);
    const double t = z * mc_fma(z, qo, qe);
    const double s2 = s + s;
    const double logm = mc_fma(s2, t, s2);
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double ef = (double)e;
    return mc_fma(ef, LN2_HI, mc_fma(ef, LN2_LO, logm));
}

// pow(x,y) with the C99 special cases the hot path can reach.
MCRT_HD double mc_pow(double x, double y)
{
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x != x || y != y) return MC_NAN_D;
    if (y == 1.0) return x;            // exact in IEEE pow; every example material has specularity 1 (ray.cpp:154-164)
    const double ay = fabs(y);
    const bool y_is_int = (ay >= 9007199254740992.0) || (floor(y) == y);
    bool y_is_odd = false;
    if (y_is_int && ay < 9007199254740992.0) {
        const double h = y * 0.5;
        y_is_odd = (floor(h) != h);
    }
    if (x == 0.0) {
        const bool neg0 = (mc_d2u(x) >> 63) != 0;
        if (y > 0.0) return (y_is_odd && neg0) ? -0.0 : 0.0;
        return (y_is_odd && neg0) ? -MC_INF_D : MC_INF_D;
    }
    double sign = 1.0;
    if (x < 0.0) {
        if (!y_is_int) return MC_NAN_D;
        if (y_is_odd) sign = -1.0;
        x = -x;
    }
    if (x == MC_INF_D) return sign * (y > 0.0 ? MC_INF_D : 0.0);
    if (ay == MC_INF_D) {
        if (x == 1.0) return 1.0;
        return ((x > 1.0) == (y > 0.0)) ? MC_INF_D : 0.0;
    }
    return sign * mc_exp(y * mc_log(x));
}

// float entry points: the reference's std::exp(float) / std::log(float) / std::pow(float,float)
MCRT_HD float mc_expf(float x) { return (float)mc_exp((double)x); }
MCRT_HD float mc_logf(float x) { return (float)mc_log((double)x); }
MCRT_HD float mc_powf(float x, float y) { return (float)mc_pow((double)x, (double)y); }

// ----------------------------------------------------------------------------------------------
// sin / cos in double for moderate arguments (the hot path only needs [0, 2*pi))
// ----------------------------------------------------------------------------------------------
MCRT_HD void mc_sincos(double a, double* sn, double* cs)
{
    const double PIO2_HI = 1.57079632673412561417e+00;   // first 33 bits of pi/2
    const double PIO2_LO = 6.07710050650619224932e-11;
    const double kf = floor(mc_fma(a, 6.36619772367581382433e-01, 0.5));
    const double r = mc_fma(-kf, PIO2_LO, mc_fma(-kf, PIO2_HI, a));  // |r| <= pi/4 (+ rounding)
    const double z = r * r, w = z * z;
    // sin(r) = r + r z S(z), cos(r) = 1 - z/2 + z^2 C(z); S, C = degree-5 near-minimax on z <= (pi/4)^2:
    // absolute errors < 1.4e-17 and < 9e-19 before rounding
    const double S0 = -0x1.5555555555555p-3, S1 = 0x1.1111111110bb1p-7, S2 = -0x1.a01a019e8357dp-13, S3 = 0x1.71de37961e4c6p-19,
                 S4 = -0x1.ae600a926c89ap-26, S5 = 0x1.5e0af186af739p-33;
    const double C0 = 0x1.5555555555555p-5, C1 = -0x1.6c16c16c16966p-10, C2 = 0x1.a01a019f4e867p-16, C3 = -0x1.27e4fa17a41b4p-22,
                 C4 = 0x1.1eeb68b109173p-29, C5 = -0x1.907d7aebd5e3dp-37;
    double se = mc_fma(w, S4, S2), so = mc_fma(w, S5, S3), ce = mc_fma(w, C4, C2), co = mc_fma(w, C5, C3);
    se = mc_fma(w, se, S0); so = mc_fma(w, so, S1); ce = mc_fma(w, ce, C0); co = mc_fma(w, co, C1);
    const double sr = mc_fma(r * z, mc_fma(z, so, se), r);
    const double cr = mc_fma(w, mc_fma(z, co, ce), mc_fma(-0.5, z, 1.0));
    // quadrant
    const double q4 = kf - 4.0 * floor(kf * 0.25);
    const int q = (int)q4;
    double s_, c_;
    if (q == 0)      { s_ = sr;  c_ = cr;  }
    else if (q == 1) { s_ = cr;  c_ = -sr; }
    else if (q == 2) { s_ = -sr; c_ = -cr; }
    else             { s_ = -cr; c_ = sr;  }
    *sn = s_; *cs = c_;
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1) -> 4 x u32
// ----------------------------------------------------------------------------------------------
struct mc_u32x4 { uint32_t v[4]; };

MCRT_HD uint32_t mc_mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MCRT_HD mc_u32x4 mc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                  uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mc_mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mc_mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n1 = lo1;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        const uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    mc_u32x4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// uniform double in (0,1): (w + 0.5) * 2^-32  -- never 0, never 1
MCRT_HD double mc_u01d(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386962890625e-10; }
// uniform float in (0,1): 24 bits, ((w>>8) + 0.5) * 2^-24 -- exactly representable, never 0 or 1
MCRT_HD float mc_u01f(uint32_t w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// RNG keying (DESIGN.md "RNG"): key = (seed lo, seed hi); counter = (frame, element, sample,
// bounce*MC_RNG_BLOCKS_PER_BOUNCE + block).  block 0 -> {thickness u1, thickness u2, shininess u,
// reflect/refract u}; blocks 1.. -> {azimuth u, radius u, -, -} for attempt (block-1) of the disk
// sampling loop of ray.cpp:171-184.
#define MC_RNG_BLOCKS_PER_BOUNCE 8u

MCRT_HD mc_u32x4 mc_rng_block(uint64_t seed, uint32_t frame, uint32_t element, uint32_t sample,
                              uint32_t bounce, uint32_t block)
{
    return mc_philox4x32_10(frame, element, sample, bounce * MC_RNG_BLOCKS_PER_BOUNCE + block,
                            (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
}

#endif  // MCRT_NUMERICS_H
