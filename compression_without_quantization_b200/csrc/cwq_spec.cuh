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

__device__ __forceinline__ void mulwide_generic(uint32_t a, uint32_t m, uint32_t& hi, uint32_t& lo) {
  unsigned long long p = (unsigned long long)a * m;   // plain C: lets the compiler fold constant operands
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

// Philox4x32-10, full generality (KAT entry point, decode / draw kernels).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulwide_generic(c0, PHILOX_M0, h0, l0);
    mulwide_generic(c2, PHILOX_M1, h1, l1);
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// 32 x 32 -> 64 bit multiply as ONE IMAD.WIDE (the C form `(u64)a * m` makes ptxas emit two extra
// adds per product on sm_100a; measured in tools/microbench/is_loop.cu).
__device__ __forceinline__ void mulwide(uint32_t a, uint32_t m, uint32_t& hi, uint32_t& lo) {
  asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(m));
}

// ------------------------------------------------------------------------------------------------
// The candidate stream, replacing tf.random.stateless_normal (misc.py:10-11).
//
//   block(s, k, j) = Philox4x32-10(counter = (s, k, 0, j), key = (STREAM_KEY0, 42))
//
// s = stream seed (exactly what the reference passes as `seed`), k = candidate (row), j = block in the row.
// A row is cut into UNITS of 16 dims.  Unit u owns blocks 3u, 3u+1, 3u+2 and turns their 384 bits into
// sixteen 23-bit uniforms: quad b < 3 of the unit takes the low 23 bits of the four words of block 3u+b;
// quad 3 takes, for word i, the TOP BYTES of word i of the three blocks (block 3u in bits 0-7, 3u+1 in
// bits 8-15, 3u+2 in bits 16-22).  Low 23 bits and top byte of a word do not overlap, so the sixteen
// uniforms are independent.  Within a quad, words (0,1) -> dims 4q, 4q+1 and (2,3) -> dims 4q+2, 4q+3.
//
// Why this shape: IMAD.WIDE (4 cycles per warp on the FMA-heavy pipe) is the scarcest resource of the
// scoring loop.  (1) 3 instead of 4 Philox blocks per 16 dims.  (2) With c2 = 0, round 0 needs one
// multiply that depends on s only; in rounds 1-2 one product of each pair depends on (s, j) only --
// hoisted out of the candidate loop -- and the other on (s, k) only -- shared by all blocks of a candidate.
// Only rounds 3..9 cost two IMAD.WIDE per block.  (3) The key is a constant, so the round keys are
// immediates of the three-input XORs instead of per-group registers.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t STREAM_KEY0 = 0x6A09E667u;     // frac(sqrt 2): a nothing-up-my-sleeve constant

__device__ __forceinline__ void mulwide(uint32_t a, uint32_t m, uint32_t& hi, uint32_t& lo);

// one whole block straight from the definition (decode / replay / debug kernels), one IMAD.WIDE per product; round 0's
// second product is 0 * M1 = 0 (c2 = 0)
__device__ __forceinline__ void stream_block(uint32_t sseed, uint32_t k, uint32_t j, uint32_t (&o)[4]) {
  uint32_t c0 = sseed, c1 = k, c2 = 0u, c3 = j;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulwide(c0, PHILOX_M0, h0, l0);
    if (r == 0) { h1 = 0u; l1 = 0u; } else mulwide(c2, PHILOX_M1, h1, l1);
    const uint32_t n0 = h1 ^ c1 ^ (STREAM_KEY0 + (uint32_t)r * PHILOX_W0);
    const uint32_t n2 = h0 ^ c3 ^ (STREAM_KEY1 + (uint32_t)r * PHILOX_W1);
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// number of Philox blocks a row of nq quads (4 dims each) needs: 3 per full unit, min(rest, 3) for the tail
__host__ __device__ __forceinline__ int row_blocks(int nq) { return 3 * (nq >> 2) + ((nq & 3) < 3 ? (nq & 3) : 3); }

// -- the same block, factored for the hot loops --------------------------------------------------
// state after round 2 is (h ^ a, l, l0 ^ b, c3) with (a, b, c3) a function of (s, j) only and
// (h, l, l0) a function of (s, k) only.
struct BlockConst { uint32_t a, b, c3; };
struct LaneShared { uint32_t h, l, l0; };

__device__ __forceinline__ BlockConst block_const(uint32_t sseed, uint32_t j) {
  uint32_t hs, ls, h1, l1, h0, l0;
  mulwide(sseed, PHILOX_M0, hs, ls);                                        // round 0, p0 (s only)
  mulwide(hs ^ j ^ STREAM_KEY1, PHILOX_M1, h1, l1);                         // round 1, p1
  mulwide(h1 ^ (STREAM_KEY0 + PHILOX_W0), PHILOX_M0, h0, l0);               // round 2, p0
  BlockConst c;
  c.a = l1 ^ (STREAM_KEY0 + 2u * PHILOX_W0);
  c.b = h0 ^ (STREAM_KEY1 + 2u * PHILOX_W1);
  c.c3 = l0;
  return c;
}

__device__ __forceinline__ LaneShared lane_shared(uint32_t sseed, uint32_t k) {
  uint32_t hs, ls, h0, l0, h1, l1;
  mulwide(sseed, PHILOX_M0, hs, ls);                                        // round 0, p0 (loop invariant)
  mulwide(k ^ STREAM_KEY0, PHILOX_M0, h0, l0);                              // round 1, p0
  mulwide(h0 ^ ls ^ (STREAM_KEY1 + PHILOX_W1), PHILOX_M1, h1, l1);          // round 2, p1
  LaneShared s;
  s.h = h1; s.l = l1; s.l0 = l0;
  return s;
}

// rounds 3..9 from the factored state
__device__ __forceinline__ void stream_block_tail(const LaneShared& ls, const BlockConst& bc, uint32_t (&o)[4]) {
  uint32_t c0 = ls.h ^ bc.a, c1 = ls.l, c2 = ls.l0 ^ bc.b, c3 = bc.c3;
#pragma unroll
  for (int r = 3; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulwide(c2, PHILOX_M1, h1, l1);      // statement order matters to ptxas' schedule of the candidate loop: this one measured
    mulwide(c0, PHILOX_M0, h0, l0);      // 0.2 % faster than (c0, c2); eight orders tried span 1.3 % (tools/ab_variants.py)
    uint32_t n0 = h1 ^ c1 ^ (STREAM_KEY0 + (uint32_t)r * PHILOX_W0);
    uint32_t n2 = h0 ^ c3 ^ (STREAM_KEY1 + (uint32_t)r * PHILOX_W1);
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// quad 3 of a unit: top bytes of word i of blocks 0, 1, 2 -> bits 0-7, 8-15, 16-23 (two PRMT per word)
__device__ __forceinline__ uint32_t harvest01(uint32_t w0, uint32_t w1) { return __byte_perm(w0, w1, 0x0073); }
__device__ __forceinline__ uint32_t harvest2(uint32_t t01, uint32_t w2) { return __byte_perm(t01, w2, 0x0710); }

// Bits of 1.0f, read from constant memory so that ptxas keeps them in a REGISTER: LOP3 has one immediate
// slot, so `(x & 0x7fffff) | 0x3f800000` with two literals costs two LOP3 per Philox word; with the
// exponent in a register it is one (`lop3 d, x, 0x7fffff, one, (a&b)|c`).  Same bits either way.
static __constant__ uint32_t c_one_bits = 0x3f800000u;

__device__ __forceinline__ float mantissa_to_1_2(uint32_t x) {
  uint32_t r;
  asm("lop3.b32 %0, %1, 0x007fffff, %2, 0xEA;" : "=r"(r) : "r"(x), "r"(c_one_bits));
  return __uint_as_float(r);
}

// The stream's RAW draws y: Box-Muller of two 23-bit uniforms WITHOUT the factor sqrt(2 ln 2):
//   y = sqrt(-lg2 u1) * (cos t, sin t),   y ~ N(0, 1 / (2 ln 2));   the N(0,1) draw is  z = NORMAL_SCALE * y.
// The hot loops never form z: the factor is folded into what multiplies the draw (importance sampler: a' = a * 2 ln 2,
// b' = b * sqrt(2 ln 2); greedy: step scale s' = s * sqrt(2 ln 2)), which removes one multiply per Box-Muller pair from
// the candidate loop.  z itself is formed only where a latent is output (decode, outliers, the seam's N(0,1) view).
constexpr float NORMAL_SCALE = 1.1774100f;       // float32(sqrt(2 ln 2))
constexpr float NORMAL_SCALE_SQ = 1.3862944f;    // float32(2 ln 2)

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& ye, float& yo) {
  float l, r, c, s;
  float f1 = mantissa_to_1_2(xa);  // [1,2)
  float f2 = mantissa_to_1_2(xb);
  float u1 = __fsub_rn(2.0f, f1);                                // (0,1], exact
  float th = __fmaf_rn(f2, 6.2831855f, -9.424778f);              // [-pi, pi)
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));         // <= 0
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-l));        // the negation is a source modifier of MUFU.SQRT
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  ye = __fmul_rn(r, c);
  yo = __fmul_rn(r, s);
}

__device__ __forceinline__ void box_muller_quad(const uint32_t (&w)[4], float (&z)[4]) {
  box_muller(w[0], w[1], z[0], z[1]);
  box_muller(w[2], w[3], z[2], z[3]);
}

// The NQ (1..4) quads of one unit of candidate k: f(quad_in_unit, z[4]) is called for each, in order.
// bc(b) returns the BlockConst of the unit's block b.  Everything stays in registers.
template <int NQ, class GetBC, class F>
__device__ __forceinline__ void unit_quads(const LaneShared& ls, GetBC&& bc, F&& f) {
  constexpr int NB = NQ < 3 ? NQ : 3;
  uint32_t v[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    uint32_t o[4];
    stream_block_tail(ls, bc(b), o);
    if (NQ == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = b == 0 ? o[i] : (b == 1 ? harvest01(v[i], o[i]) : harvest2(v[i], o[i]));
    }
    float z[4];
    box_muller_quad(o, z);
    f(b, z);
  }
  if (NQ == 4) {
    float z[4];
    box_muller_quad(v, z);
    f(3, z);
  }
}

// The same for a unit whose LAST quad carries at most two dims (NP = 2 NQ - 1 Box-Muller pairs): the last quad's second
// pair -- and, for NQ == 4, the harvest of its two words -- is never formed; f(quad, z, pairs) sees pairs == 1 there and
// reads z[0], z[1] only.  The dims left out have zero coefficients in every caller, so the result is the same bits.
template <int NQ, class GetBC, class F>
__device__ __forceinline__ void unit_quads_half(const LaneShared& ls, GetBC&& bc, F&& f) {
  constexpr int NB = NQ < 3 ? NQ : 3;
  uint32_t v[2] = {0u, 0u};
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    uint32_t o[4];
    stream_block_tail(ls, bc(b), o);
    if (NQ == 4) {
#pragma unroll
      for (int i = 0; i < 2; ++i) v[i] = b == 0 ? o[i] : (b == 1 ? harvest01(v[i], o[i]) : harvest2(v[i], o[i]));
    }
    float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (NQ < 4 && b == NB - 1) {
      box_muller(o[0], o[1], z[0], z[1]);
      f(b, z, 1);
    } else {
      box_muller_quad(o, z);
      f(b, z, 2);
    }
  }
  if (NQ == 4) {
    float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    box_muller(v[0], v[1], z[0], z[1]);
    f(3, z, 1);
  }
}

// All nq quads of candidate k's row: f(quad, z[4]) in quad order; bc(j) = BlockConst of row block j.
template <class GetBC, class F>
__device__ __forceinline__ void row_quads(const LaneShared& ls, int nq, GetBC&& bc, F&& f) {
  int Q = 0, j = 0;
  for (; Q + 4 <= nq; Q += 4, j += 3)
    unit_quads<4>(ls, [&](int b) { return bc(j + b); }, [&](int qi, const float (&z)[4]) { f(Q + qi, z); });
  switch (nq - Q) {
    case 1: unit_quads<1>(ls, [&](int b) { return bc(j + b); }, [&](int qi, const float (&z)[4]) { f(Q + qi, z); }); break;
    case 2: unit_quads<2>(ls, [&](int b) { return bc(j + b); }, [&](int qi, const float (&z)[4]) { f(Q + qi, z); }); break;
    case 3: unit_quads<3>(ls, [&](int b) { return bc(j + b); }, [&](int qi, const float (&z)[4]) { f(Q + qi, z); }); break;
    default: break;
  }
}

// One quad of a row, straight from the definition (decode / replay / debug kernels: O(N) work).
__device__ __forceinline__ void quad_words(uint32_t sseed, uint32_t k, uint32_t quad, uint32_t (&w)[4]) {
  const uint32_t u = quad >> 2, qi = quad & 3u;
  if (qi < 3u) {
    stream_block(sseed, k, 3u * u + qi, w);
  } else {
    uint32_t o0[4], o1[4], o2[4];
    stream_block(sseed, k, 3u * u, o0);
    stream_block(sseed, k, 3u * u + 1u, o1);
    stream_block(sseed, k, 3u * u + 2u, o2);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = harvest2(harvest01(o0[i], o1[i]), o2[i]);
  }
}

// raw draws y of one quad (what the candidate loops score), and the N(0,1) draws z = NORMAL_SCALE * y (what is output)
__device__ __forceinline__ void raw_quad(uint32_t sseed, uint32_t k, uint32_t quad, float (&y)[4]) {
  uint32_t w[4];
  quad_words(sseed, k, quad, w);
  box_muller_quad(w, y);
}
__device__ __forceinline__ void normal_quad(uint32_t sseed, uint32_t k, uint32_t quad, float (&z)[4]) {
  raw_quad(sseed, k, quad, z);
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = __fmul_rn(NORMAL_SCALE, z[i]);
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

// Per-dim prologue of the importance sampler (CIS:133-163): rescale q by p, outlier test against
// `limit` bits, KL of what stays, score coefficients (a, b).  m bit0 = dim replaced by the prior, bit1 = outlier (side info).
struct PrologueDim { float kl; float2 coef; uint8_t m; };
__device__ __forceinline__ PrologueDim is_prologue_dim(float ql, float qs, float pl, float ps, float limit);

__device__ __forceinline__ PrologueDim is_prologue_dim(float ql, float qs, float pl, float ps, float limit) {
  const float tl = __fdiv_rn(__fsub_rn(ql, pl), ps);
  const float ts = __fdiv_rn(qs, ps);
  const float k0 = kl01(tl, ts);
  const float kl_bits = __fdiv_rn(k0, LN2_F);
  PrologueDim r;
  r.m = 0;
  if (!(kl_bits <= limit)) r.m |= 1;
  if (kl_bits > limit) r.m |= 2;
  // a dim replaced by the prior is N(0, 1) against N(0, 1): kl01(0, 1) and both coefficients are exactly +0, so the
  // replaced branch is a select, not a second evaluation (same bits as evaluating the formulas at (0, 1))
  const float iv = __fdiv_rn(1.0f, __fmul_rn(ts, ts));
  const bool rep = (r.m & 1) != 0;
  r.kl = rep ? 0.0f : k0;
  // log q(z) - log p(z) = a z^2 + b z + const with z = NORMAL_SCALE * y: the kernels score the raw draw y with
  // (a * 2 ln 2, b * sqrt(2 ln 2))
  r.coef = rep ? make_float2(0.0f, 0.0f)
               : make_float2(__fmul_rn(__fmaf_rn(-0.5f, iv, 0.5f), NORMAL_SCALE_SQ), __fmul_rn(__fmul_rn(tl, iv), NORMAL_SCALE));
  return r;
}

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
