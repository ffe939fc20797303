// cwq_spec.cuh -- canonical arithmetic of the engine, device side.
//
// Everything that decides an index, a group boundary or a candidate count is written with explicit
// IEEE round-to-nearest intrinsics (__fmaf_rn, __fmul_rn, __fadd_rn, __fdiv_rn): no contraction, no
// fast-math, so the CPU oracle reproduces it bit for bit.  The only non-IEEE pieces are the four
// MUFU approximations inside box_muller(); their outputs are what the parity harness injects into
// the oracle (BASELINE.json north_star: "injects the identical Philox candidate draws").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cwq {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t STREAM_KEY1 = 42u;             // reference misc.py:11  seed = [seed, 42]
constexpr uint32_t OUTLIER_STREAM_OFFSET = 0x9E3779B9u;

// Philox4x32-10, full generality (used by the KAT entry point).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned long long p0 = (unsigned long long)PHILOX_M0 * c0;
    unsigned long long p1 = (unsigned long long)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// The candidate stream's block: counter (k, q, 0, 0), key (stream_seed, 42).  With c2 = c3 = 0 the
// first round's second multiply is a constant and the second round's first multiply depends only on
// (q, stream_seed), so ptxas hoists it out of the candidate loop: 18 of the 20 multiplies remain.
__device__ __forceinline__ void stream_block(uint32_t sseed, uint32_t k, uint32_t q, uint32_t (&o)[4]) {
  philox4x32_10(k, q, 0u, 0u, sseed, STREAM_KEY1, o);
}

// Two N(0,1) values from two Philox words (Box-Muller, 23-bit uniforms).
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& ze, float& zo) {
  float f1 = __uint_as_float((xa & 0x007fffffu) | 0x3f800000u);  // [1,2)
  float f2 = __uint_as_float((xb & 0x007fffffu) | 0x3f800000u);
  float u1 = __fsub_rn(2.0f, f1);                                // (0,1], exact
  float th = __fmaf_rn(f2, 6.2831855f, -9.424778f);              // [-pi, pi)
  float l, r, c, s;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  float r2 = __fmul_rn(l, -1.3862944f);                          // -2 ln u1
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  ze = __fmul_rn(r, c);
  zo = __fmul_rn(r, s);
}

__device__ __forceinline__ void normal_quad(uint32_t sseed, uint32_t k, uint32_t q, float (&z)[4]) {
  uint32_t o[4];
  stream_block(sseed, k, q, o);
  box_muller(o[0], o[1], z[0], z[1]);
  box_muller(o[2], o[3], z[2], z[3]);
}

// Deterministic natural log / exp (Cephes-style minimax polynomials), bit-identical to
// oracle/cwq_oracle.c:cwqo_detlogf / cwqo_detexpf.
__device__ __forceinline__ float detlogf(float x) {
  uint32_t ix = __float_as_uint(x);
  if ((ix << 1) == 0u) return __uint_as_float(0xff800000u);
  if (ix >> 31) return __uint_as_float(0x7fc00000u);
  if (ix >= 0x7f800000u) return x;
  int e = 0;
  if (ix < 0x00800000u) { x = __fmul_rn(x, 8388608.0f); ix = __float_as_uint(x); e = -23; }
  e += (int)(ix >> 23) - 127;
  float m = __uint_as_float((ix & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421356f) { m = __fmul_rn(m, 0.5f); e += 1; }
  float f = __fsub_rn(m, 1.0f);
  float z = __fmul_rn(f, f);
  float p = 7.0376836292e-2f;
  p = __fmaf_rn(p, f, -1.1514610310e-1f);
  p = __fmaf_rn(p, f, 1.1676998740e-1f);
  p = __fmaf_rn(p, f, -1.2420140846e-1f);
  p = __fmaf_rn(p, f, 1.4249322787e-1f);
  p = __fmaf_rn(p, f, -1.6668057665e-1f);
  p = __fmaf_rn(p, f, 2.0000714765e-1f);
  p = __fmaf_rn(p, f, -2.4999993993e-1f);
  p = __fmaf_rn(p, f, 3.3333331174e-1f);
  float y = __fmul_rn(__fmul_rn(f, z), p);
  float fe = (float)e;
  y = __fmaf_rn(fe, -2.12194440e-4f, y);
  y = __fmaf_rn(-0.5f, z, y);
  float r = __fadd_rn(f, y);
  return __fmaf_rn(fe, 0.693359375f, r);
}

__device__ __forceinline__ float detexpf(float x) {
  if (x != x) return x;
  if (x > 88.72283f) return __uint_as_float(0x7f800000u);
  if (x < -103.0f) return 0.0f;
  float n = floorf(__fmaf_rn(x, 1.44269504f, 0.5f));
  float r = __fmaf_rn(n, -0.693359375f, x);
  r = __fmaf_rn(n, 2.12194440e-4f, r);
  float z = __fmul_rn(r, r);
  float p = 1.9875691500e-4f;
  p = __fmaf_rn(p, r, 1.3981999507e-3f);
  p = __fmaf_rn(p, r, 8.3334519073e-3f);
  p = __fmaf_rn(p, r, 4.1665795894e-2f);
  p = __fmaf_rn(p, r, 1.6666665459e-1f);
  p = __fmaf_rn(p, r, 5.0000001201e-1f);
  float y = __fadd_rn(__fmaf_rn(p, z, r), 1.0f);
  int ni = (int)n, h = ni / 2;
  float s1 = __uint_as_float((uint32_t)(h + 127) << 23), s2 = __uint_as_float((uint32_t)(ni - h + 127) << 23);
  return __fmul_rn(__fmul_rn(y, s1), s2);
}

// KL( N(mu, sigma) || N(0,1) ) in nats.
__device__ __forceinline__ float kl01(float mu, float sigma) {
  float r = __fmul_rn(sigma, sigma), m = __fmul_rn(mu, mu);
  float l = detlogf(r);
  return __fadd_rn(__fmul_rn(0.5f, m), __fmul_rn(0.5f, __fsub_rn(__fsub_rn(r, 1.0f), l)));
}

constexpr float LN2_F = 0.6931472f;

// candidate count of a group: K = ceil(exp(total KL)), clamped to [1, 2^31-1]
__device__ __forceinline__ int32_t candidates_from_kl(float total) {
  float c = ceilf(detexpf(total));
  if (!(c >= 1.0f)) return 1;
  if (c >= 2147483648.0f) return 2147483647;
  return (int32_t)c;
}

// (score, index) -> a 64-bit key whose unsigned maximum is "largest score, then lowest index".
__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_key(float score, uint32_t k) {
  return ((unsigned long long)orderable(score) << 32) | (unsigned long long)(~k);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long key) { return ~(uint32_t)key; }
// key of "candidate 0 with score -inf": what a sequential `if (s > best)` scan starts from
constexpr unsigned long long KEY_INIT = ((unsigned long long)0x007fffffu << 32) | 0xffffffffull;

__device__ __forceinline__ uint16_t quantize_u16(float x) {
  float c = x < -30.0f ? -30.0f : (x > 30.0f ? 30.0f : x);
  float v = floorf(__fmaf_rn(__fadd_rn(c, 30.0f), 1092.25f, 0.5f));
  if (!(v >= 0.0f)) v = 0.0f;
  if (v > 65535.0f) v = 65535.0f;
  return (uint16_t)v;
}
__device__ __forceinline__ float dequantize_u16(uint16_t q) {
  return __fmaf_rn((float)q, 60.0f / 65535.0f, -30.0f);
}

}  // namespace cwq
