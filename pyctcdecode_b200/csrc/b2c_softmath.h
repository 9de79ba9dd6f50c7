// b200ctc -- float32 exp / log with a FIXED sequence of IEEE operations (fused multiply-add, multiply, add, integer
// arithmetic on the bit patterns): the same bits from nvcc device code and from g++ host code, which is what lets
// the oracle (its own copy of these two functions) and the kernels agree to the last bit while the log-softmax is
// computed in float32 like the reference does (decoder.py:180-197 on float32 logits).  Accuracy against the exact
// functions: exp <= 0.99 ulp on [-87, 0], log <= 0.74 ulp on [1, 1100] (tools/softmath_check.cpp).
#pragma once
#include <stdint.h>
#if defined(__CUDA_ARCH__)
#define B2C_SM_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define B2C_SM_MUL(a, b) __fmul_rn((a), (b))
#define B2C_SM_ADD(a, b) __fadd_rn((a), (b))
#define B2C_SM_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define B2C_SM_FMA(a, b, c) __builtin_fmaf((a), (b), (c))
#define B2C_SM_MUL(a, b) ((a) * (b))
#define B2C_SM_ADD(a, b) ((a) + (b))
#if defined(__CUDACC__)
#define B2C_SM_HD __host__ __device__ __forceinline__
#else
#define B2C_SM_HD inline
#endif
#endif
B2C_SM_HD float b2c_sm_from_bits(uint32_t u) { union { uint32_t u; float f; } c; c.u = u; return c.f; }
B2C_SM_HD uint32_t b2c_sm_bits(float f) { union { uint32_t u; float f; } c; c.f = f; return c.u; }

// exp(x), float32, for the log-softmax (x = logit - row max <= 0 in practice).  n = round(x log2 e) by the
// magic-number trick, r = x - n ln2 in two fused steps, degree-6 polynomial, 2^n by exponent arithmetic.
// Below -87 the result is defined as 0 (no subnormal results), above 88.7 as +inf.
B2C_SM_HD float b2c_sm_expf(float x) {
    if (!(x == x)) return x;
    if (x > 88.7f) return b2c_sm_from_bits(0x7F800000u);
    if (x < -87.0f) return 0.0f;
    const float t = B2C_SM_FMA(x, 1.44269504088896341f, 12582912.0f);
    const float n = B2C_SM_ADD(t, -12582912.0f);
    float r = B2C_SM_FMA(n, -0.693359375f, x);
    r = B2C_SM_FMA(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = B2C_SM_FMA(p, r, 1.3981999507e-3f);
    p = B2C_SM_FMA(p, r, 8.3334519073e-3f);
    p = B2C_SM_FMA(p, r, 4.1665795894e-2f);
    p = B2C_SM_FMA(p, r, 1.6666665459e-1f);
    p = B2C_SM_FMA(p, r, 5.0000001201e-1f);
    const float r2 = B2C_SM_MUL(r, r);
    p = B2C_SM_FMA(p, r2, r);
    p = B2C_SM_ADD(p, 1.0f);
    const int32_t ni = static_cast<int32_t>(n);
    return b2c_sm_from_bits(b2c_sm_bits(p) + (static_cast<uint32_t>(ni) << 23));
}

// rint(exp(d) * 2^32) as an integer -- the addend of one element in the softmax denominator -- for d <= 0 that is finite
// or -inf (rows without NaN / +inf; the caller has checked).  The SAME sequence of operations as b2c_sm_expf followed
// by the scaling and rounding of b2c_sm_quantum, without their special-value branches: d is clamped at -87 (there
// exp(d) 2^32 < 0.5, the same addend 0 as the definition's cut-off), n is read from the bits of the magic-number sum
// instead of converted, and the factor 2^32 is folded into the exponent arithmetic (an exact scaling either way).
// Bit-identical to b2c_sm_quantum(b2c_sm_expf(d)) on that domain (tools/softmath_check.cpp, tests/test_host_logic.py).
B2C_SM_HD uint64_t b2c_sm_quantum_fast(float d) {
#if defined(__CUDA_ARCH__)
    const float x = fmaxf(d, -87.0f);
#else
    const float x = d < -87.0f ? -87.0f : d;
#endif
    const float t = B2C_SM_FMA(x, 1.44269504088896341f, 12582912.0f);
    const float n = B2C_SM_ADD(t, -12582912.0f);
    float r = B2C_SM_FMA(n, -0.693359375f, x);
    r = B2C_SM_FMA(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = B2C_SM_FMA(p, r, 1.3981999507e-3f);
    p = B2C_SM_FMA(p, r, 8.3334519073e-3f);
    p = B2C_SM_FMA(p, r, 4.1665795894e-2f);
    p = B2C_SM_FMA(p, r, 1.6666665459e-1f);
    p = B2C_SM_FMA(p, r, 5.0000001201e-1f);
    const float r2 = B2C_SM_MUL(r, r);
    p = B2C_SM_FMA(p, r2, r);
    p = B2C_SM_ADD(p, 1.0f);
    // t = 1.5 * 2^23 + n exactly (|n| <= 126): n sits in the low mantissa bits of t
    const uint32_t ni32 = b2c_sm_bits(t) - 0x4B400000u + 32u;
    const float e32 = b2c_sm_from_bits(b2c_sm_bits(p) + (ni32 << 23));
#if defined(__CUDA_ARCH__)
    return __float2ull_rn(e32);
#else
    return static_cast<uint64_t>(llrintf(e32));
#endif
}

// log(x), float32, x > 0 normal (the softmax denominator is in [1, V]); Cephes-style: x = m 2^e with
// m in [sqrt(1/2), sqrt(2)), polynomial in f = m - 1, e ln2 added in two parts.
B2C_SM_HD float b2c_sm_logf(float x) {
    if (!(x == x) || x < 0.0f) return b2c_sm_from_bits(0x7FC00000u);
    if (x == 0.0f) return b2c_sm_from_bits(0xFF800000u);
    const uint32_t ux = b2c_sm_bits(x);
    if (ux >= 0x7F800000u) return x;
    int32_t e = static_cast<int32_t>(ux >> 23) - 126;
    float m = b2c_sm_from_bits((ux & 0x007FFFFFu) | 0x3F000000u);      // [0.5, 1)
    if (m < 0.70710678118654752440f) { e -= 1; m = B2C_SM_ADD(m, m); }
    const float f = B2C_SM_ADD(m, -1.0f);
    const float z = B2C_SM_MUL(f, f);
    float p = 7.0376836292e-2f;
    p = B2C_SM_FMA(p, f, -1.1514610310e-1f);
    p = B2C_SM_FMA(p, f, 1.1676998740e-1f);
    p = B2C_SM_FMA(p, f, -1.2420140846e-1f);
    p = B2C_SM_FMA(p, f, 1.4249322787e-1f);
    p = B2C_SM_FMA(p, f, -1.6668057665e-1f);
    p = B2C_SM_FMA(p, f, 2.0000714765e-1f);
    p = B2C_SM_FMA(p, f, -2.4999993993e-1f);
    p = B2C_SM_FMA(p, f, 3.3333331174e-1f);
    float y = B2C_SM_MUL(B2C_SM_MUL(f, z), p);
    const float fe = static_cast<float>(e);
    y = B2C_SM_FMA(fe, -2.12194440e-4f, y);
    y = B2C_SM_FMA(z, -0.5f, y);
    float r = B2C_SM_ADD(f, y);
    r = B2C_SM_FMA(fe, 0.693359375f, r);
    return r;
}
