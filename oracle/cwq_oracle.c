/*
 * cwq_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the MIRACLE relative-entropy-coding hot path of
 * gergely-flamich/compression-without-quantization, written against the canonical arithmetic
 * spec in DESIGN.md section 3.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object; the product (libcwq.so) never does.
 *
 * Parity status: the reference samplers need TensorFlow 1.x + TFP, neither of which is installed
 * here, and the reference ships no golden vectors for this path.  The pins
 * are therefore: Random123 Philox4x32-10 known-answer vectors, the reference's own
 * code/binary_io.py (imported as-is when generating tests/golden/), and golden vectors produced by
 * executing the UNMODIFIED reference sampler sources under a numpy stand-in for the TF ops they
 * call (oracle/tf_shim, generator tests/golden/gen_reference_goldens.py).  The TF/TFP kernels
 * themselves (stateless_normal, Eigen reduction order, TFP formula revision) stay unpinned.
 *
 * Reference lines restated (all under /root/reference/code):
 *   misc.py:3-17                         -> cwqo_normals        (the injection seam)
 *   coded_importance_sampler.py:48-51    -> cwqo_kl01, cwqo_group_totals (K = ceil(exp(sum KL)))
 *   coded_importance_sampler.py:55-63    -> cwqo_is_encode_group
 *   coded_importance_sampler.py:104-109  -> cwqo_is_decode_group
 *   coded_importance_sampler.py:133-163  -> cwqo_is_prologue    (rescale, outlier split, KL)
 *   coded_importance_sampler.py:164-204  -> cwqo_group          (strict = 1)
 *   coded_greedy_sampler.py:207-252      -> cwqo_group          (strict = 0)
 *   coded_greedy_sampler.py:38-78        -> cwqo_greedy_encode_group
 *   coded_greedy_sampler.py:134-165      -> cwqo_greedy_decode_group
 *   coded_greedy_sampler.py:194-201      -> cwqo_greedy_prologue
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -mfma -pthread).  Floating point is fp32
 * IEEE round-to-nearest with fused multiply-adds ONLY where fmaf() is written.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <stdatomic.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
#define STREAM_KEY1 42u /* misc.py:11  seed = [seed, 42] */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- Philox4x32-10 (Salmon et al. 2011; Random123).  ctr/out 4 words, key 2 words. ---- */
void cwqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- the normal stream z(stream_seed, k, d): DESIGN.md section 3.1 ----
 * block(s, k, j) = philox(ctr = (s, k, 0, j), key = (STREAM_KEY0, 42)).  A row is cut into units of 16 dims;
 * unit u owns blocks 3u, 3u+1, 3u+2.  Quad b < 3 of the unit uses the low 23 bits of the four words of block
 * 3u+b; quad 3 uses, for word i, the top bytes of word i of the three blocks (block 3u in bits 0-7, 3u+1 in
 * bits 8-15, 3u+2 in bits 16-23; low 23 bits taken).  Words (0,1) of a quad -> dims 4q, 4q+1; (2,3) -> dims
 * 4q+2, 4q+3.  The oracle evaluates the float transform with libm; the engine uses MUFU approximations, so
 * engine draws are injected when bit-exact indices are asserted. */
#define STREAM_KEY0 0x6A09E667u

static inline void stream_block(uint32_t sseed, uint32_t k, uint32_t j, uint32_t o[4]) {
  uint32_t ctr[4] = {sseed, k, 0u, j}, key[2] = {STREAM_KEY0, STREAM_KEY1};
  cwqo_philox4x32_10(ctr, key, o);
}

/* The stream's RAW draws: y = sqrt(-lg2 u1) (cos t, sin t), y ~ N(0, 1 / (2 ln 2)); the N(0,1) draw is
 * z = NORMAL_SCALE * y.  The candidate loops score y with the factor folded into the coefficients / the step scale. */
#define NORMAL_SCALE 1.1774100f    /* float32(sqrt(2 ln 2)) */
#define NORMAL_SCALE_SQ 1.3862944f /* float32(2 ln 2) */
static inline void box_muller(uint32_t xa, uint32_t xb, float* ye, float* yo) {
  float f1 = u2f((xa & 0x007fffffu) | 0x3f800000u); /* [1,2) */
  float f2 = u2f((xb & 0x007fffffu) | 0x3f800000u);
  float u1 = 2.0f - f1;                              /* (0,1], exact */
  float th = fmaf(f2, 6.2831855f, -9.424778f);       /* [-pi, pi) */
  float r = sqrtf(-log2f(u1));
  *ye = r * cosf(th);
  *yo = r * sinf(th);
}

static inline uint32_t harvest(uint32_t w0, uint32_t w1, uint32_t w2) {
  return (w0 >> 24) | ((w1 >> 24) << 8) | ((w2 >> 24) << 16);
}

/* the four uniform words of quad `quad` of row k, straight from the definition */
static inline void quad_words(uint32_t sseed, uint32_t k, uint32_t quad, uint32_t w[4]) {
  uint32_t u = quad >> 2, qi = quad & 3u;
  if (qi < 3u) { stream_block(sseed, k, 3u * u + qi, w); return; }
  uint32_t o0[4], o1[4], o2[4];
  stream_block(sseed, k, 3u * u, o0);
  stream_block(sseed, k, 3u * u + 1u, o1);
  stream_block(sseed, k, 3u * u + 2u, o2);
  for (int i = 0; i < 4; ++i) w[i] = harvest(o0[i], o1[i], o2[i]);
}

static inline void raw_quad(uint32_t sseed, uint32_t k, uint32_t quad, float z[4]) {
  uint32_t w[4];
  quad_words(sseed, k, quad, w);
  box_muller(w[0], w[1], &z[0], &z[1]);
  box_muller(w[2], w[3], &z[2], &z[3]);
}

/* Same values as raw_quad for callers that walk the quads of ONE row in order (0, 1, 2, ...): the three
 * blocks of the current unit are kept in `unit` so that quad 3 does not recompute them (the candidate loops
 * timed as the CPU baseline do the same work per candidate as the GPU kernels: 3 blocks per 16 dims). */
static inline void raw_quad_seq(uint32_t sseed, uint32_t k, uint32_t quad, uint32_t unit[3][4], float z[4]) {
  uint32_t qi = quad & 3u, w[4];
  if (qi < 3u) {
    stream_block(sseed, k, 3u * (quad >> 2) + qi, unit[qi]);
    memcpy(w, unit[qi], sizeof(w));
  } else {
    for (int i = 0; i < 4; ++i) w[i] = harvest(unit[0][i], unit[1][i], unit[2][i]);
  }
  box_muller(w[0], w[1], &z[0], &z[1]);
  box_muller(w[2], w[3], &z[2], &z[3]);
}

/* rows k0..k0+nk-1, columns 0..G-1, row-major [nk, G]: raw != 0 -> the raw draws y (what the candidate loops score and
 * the injection seam carries), raw == 0 -> N(0,1) draws z = NORMAL_SCALE * y (misc.py:10-15 with loc=0, scale=1) */
void cwqo_normals(uint32_t sseed, uint32_t k0, uint32_t nk, uint32_t G, int raw, float* out) {
  for (uint32_t i = 0; i < nk; ++i) {
    uint32_t unit[3][4];
    for (uint32_t q = 0; q * 4 < G; ++q) {
      float z[4];
      raw_quad_seq(sseed, k0 + i, q, unit, z);
      for (uint32_t j = 0; j < 4 && q * 4 + j < G; ++j) out[(size_t)i * G + q * 4 + j] = raw ? z[j] : NORMAL_SCALE * z[j];
    }
  }
}

/* raw Philox words of blocks j = 0..nb-1 of rows k0..k0+nk-1, for the integer KAT / engine-vs-oracle bit check */
void cwqo_stream_words(uint32_t sseed, uint32_t k0, uint32_t nk, uint32_t nb, uint32_t* out) {
  for (uint32_t i = 0; i < nk; ++i)
    for (uint32_t j = 0; j < nb; ++j) stream_block(sseed, k0 + i, j, out + ((size_t)i * nb + j) * 4);
}

/* the uniform words (before the 23-bit mask) of quads 0..nq-1 of rows k0..k0+nk-1 */
void cwqo_quad_words(uint32_t sseed, uint32_t k0, uint32_t nk, uint32_t nq, uint32_t* out) {
  for (uint32_t i = 0; i < nk; ++i)
    for (uint32_t q = 0; q < nq; ++q) quad_words(sseed, k0 + i, q, out + ((size_t)i * nq + q) * 4);
}

/* ---- deterministic log / exp (Cephes-style polynomials, explicit fmaf): spec section 3.3 ---- */
float cwqo_detlogf(float x) {
  uint32_t ix = f2u(x);
  if ((ix << 1) == 0) return -INFINITY;
  if (ix >> 31) return NAN;
  if (ix >= 0x7f800000u) return x;
  int e = 0;
  if (ix < 0x00800000u) { x = x * 8388608.0f; ix = f2u(x); e = -23; }
  e += (int)(ix >> 23) - 127;
  float m = u2f((ix & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
  float f = m - 1.0f;
  float z = f * f;
  float p = 7.0376836292e-2f;
  p = fmaf(p, f, -1.1514610310e-1f);
  p = fmaf(p, f, 1.1676998740e-1f);
  p = fmaf(p, f, -1.2420140846e-1f);
  p = fmaf(p, f, 1.4249322787e-1f);
  p = fmaf(p, f, -1.6668057665e-1f);
  p = fmaf(p, f, 2.0000714765e-1f);
  p = fmaf(p, f, -2.4999993993e-1f);
  p = fmaf(p, f, 3.3333331174e-1f);
  float y = (f * z) * p;
  float fe = (float)e;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(-0.5f, z, y);
  float r = f + y;
  return fmaf(fe, 0.693359375f, r);
}

float cwqo_detexpf(float x) {
  if (x != x) return x;
  if (x > 88.72283f) return INFINITY;
  if (x < -103.0f) return 0.0f;
  float n = floorf(fmaf(x, 1.44269504f, 0.5f));
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float z = r * r;
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  float y = fmaf(p, z, r) + 1.0f;
  int ni = (int)n, h = ni / 2;
  float s1 = u2f((uint32_t)(h + 127) << 23), s2 = u2f((uint32_t)(ni - h + 127) << 23);
  return (y * s1) * s2;
}

/* KL( N(mu, sigma) || N(0,1) ) in nats, spec section 3.4 (tfd.kl_divergence at CIS:48,142,160, CGS:201) */
float cwqo_kl01(float mu, float sigma) {
  float r = sigma * sigma, m = mu * mu;
  float l = cwqo_detlogf(r);
  return 0.5f * m + 0.5f * ((r - 1.0f) - l);
}

static const float LN2_F = 0.6931472f; /* float32(np.log(2)) */

/* ---- CIS:133-163.  Rescale q by p, split out high-KL dims, KL of what is left. ----
 * mask[i]: bit0 = "not (kl_bits <= limit)" (t replaced by p, CIS:144-145), bit1 = kl_bits > limit
 * (an outlier sent as side information, CIS:148). */
void cwqo_is_prologue(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                      int64_t n, float dim_kl_bit_limit, float* t_loc, float* t_scale, float* kl,
                      uint8_t* mask) {
  for (int64_t i = 0; i < n; ++i) {
    float tl = (q_loc[i] - p_loc[i]) / p_scale[i];
    float ts = q_scale[i] / p_scale[i];
    float kl_bits = cwqo_kl01(tl, ts) / LN2_F;
    uint8_t m = 0;
    if (!(kl_bits <= dim_kl_bit_limit)) { m |= 1; tl = 0.0f; ts = 1.0f; }
    if (kl_bits > dim_kl_bit_limit) m |= 2;
    t_loc[i] = tl; t_scale[i] = ts; mask[i] = m;
    kl[i] = cwqo_kl01(tl, ts);
  }
}

/* CGS:194-201: rescale and KL, no outlier handling */
void cwqo_greedy_prologue(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                          int64_t n, float* t_loc, float* t_scale, float* kl) {
  for (int64_t i = 0; i < n; ++i) {
    float tl = (q_loc[i] - p_loc[i]) / p_scale[i];
    float ts = q_scale[i] / p_scale[i];
    t_loc[i] = tl; t_scale[i] = ts;
    kl[i] = cwqo_kl01(tl, ts);
  }
}

/* ---- G1: sequential greedy packing, CIS:164-204 (strict=1: '>') / CGS:207-252 (strict=0: '>=').
 * size_limit: a boundary is forced once the running group holds size_limit dims, i.e. the reference's
 * test log2(size+1) >(=) max_group_size_bits restated on integers (host computes it, see
 * oracle.py:size_limit).  Accumulator float32, budget float64 (CIS:181).  starts must hold n+2 entries;
 * returns the number of groups (len(starts) - 1). */
int64_t cwqo_group(const float* kl, int64_t n, double n_nats_per_group, int64_t size_limit, int strict,
                   int64_t* starts) {
  int64_t ns = 0;
  starts[ns++] = 0;
  int64_t size = 0;
  float acc = 0.0f;
  for (int64_t idx = 0; idx < n; ++idx) {
    float cand = acc + kl[idx];
    int over_size = size >= size_limit;
    int over_kl = strict ? ((double)cand > n_nats_per_group) : ((double)cand >= n_nats_per_group);
    if (over_size || over_kl || idx == n - 1) {
      starts[ns++] = idx;
      size = 1;
      acc = kl[idx];
    } else {
      acc = cand;
      size += 1;
    }
  }
  starts[ns++] = n;
  return ns - 1;
}

/* per-group KL total (sequential fp32, the accumulator G1 carried) and K = ceil(exp(total)), CIS:49-51 */
void cwqo_group_totals(const float* kl, const int64_t* starts, int64_t n_groups, float* total, int32_t* K) {
  for (int64_t g = 0; g < n_groups; ++g) {
    float acc = 0.0f;
    for (int64_t i = starts[g]; i < starts[g + 1]; ++i) acc = acc + kl[i];
    total[g] = acc;
    float c = ceilf(cwqo_detexpf(acc));
    int32_t k;
    if (!(c >= 1.0f)) k = 1;
    else if (c >= 2147483648.0f) k = 2147483647;
    else k = (int32_t)c;
    K[g] = k;
  }
}

/* score coefficients, spec section 3.5: log q(z) - log p(z) = a z^2 + b z + const for p = N(0,1) */
void cwqo_is_coef(const float* t_loc, const float* t_scale, int64_t n, float* a, float* b) {
  for (int64_t i = 0; i < n; ++i) {
    float iv = 1.0f / (t_scale[i] * t_scale[i]);
    /* log q(z) - log p(z) = a z^2 + b z + const, z = NORMAL_SCALE * y: coefficients of the raw draw y */
    a[i] = fmaf(-0.5f, iv, 0.5f) * NORMAL_SCALE_SQ;
    b[i] = (t_loc[i] * iv) * NORMAL_SCALE;
  }
}

/* ---- CIS:55-63: K candidates, importance weight, first maximal index.  inj: optional [K,G] draws. */
uint32_t cwqo_is_encode_group(const float* a, const float* b, int32_t G, uint32_t K, uint32_t sseed,
                              const float* inj, float* best_score_out) {
  float best = -INFINITY;
  uint32_t best_k = 0;
  int32_t nq = (G + 3) / 4;
  for (uint32_t k = 0; k < K; ++k) {
    float acc0 = 0.0f, acc1 = 0.0f;
    uint32_t unit[3][4];
    for (int32_t q = 0; q < nq; ++q) {
      float z[4];
      if (inj) { for (int j = 0; j < 4; ++j) z[j] = (q * 4 + j < G) ? inj[(size_t)k * G + q * 4 + j] : 0.0f; }
      else raw_quad_seq(sseed, k, (uint32_t)q, unit, z);
      for (int j = 0; j < 4 && q * 4 + j < G; ++j) {
        int d = q * 4 + j;
        float w = fmaf(a[d], z[j], b[d]);
        if (j & 1) acc1 = fmaf(z[j], w, acc1); else acc0 = fmaf(z[j], w, acc0);
      }
    }
    float s = acc0 + acc1;
    if (s > best) { best = s; best_k = k; }
  }
  if (best_score_out) *best_score_out = best;
  return best_k;
}

/* CIS:104-109 with a counter-based stream: row `index` only */
void cwqo_is_decode_group(int32_t G, uint32_t index, uint32_t sseed, float* out) {
  cwqo_normals(sseed, index, 1, (uint32_t)G, 0, out);
}

/* ---- a tiny pthread parallel-for over groups (bench.py cpu_baseline / --impl reference) ---- */
typedef void (*group_fn)(int64_t g, void* arg);
typedef struct { atomic_llong next; int64_t g1; group_fn fn; void* arg; } pf_t;
static void* pf_worker(void* p) {
  pf_t* s = (pf_t*)p;
  for (;;) {
    int64_t g = (int64_t)atomic_fetch_add(&s->next, 1);
    if (g >= s->g1) break;
    s->fn(g, s->arg);
  }
  return NULL;
}
static void parallel_for(int64_t g0, int64_t g1, int n_threads, group_fn fn, void* arg) {
  pf_t s; atomic_init(&s.next, (long long)g0); s.g1 = g1; s.fn = fn; s.arg = arg;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < n_threads; ++t) if (pthread_create(&th[started], NULL, pf_worker, &s) == 0) ++started;
  pf_worker(&s);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

typedef struct { const float *a, *b; const int64_t* starts; const int32_t* K; uint32_t seed; uint32_t* index_out; float* z_out; } is_job_t;
static void is_job(int64_t g, void* p) {
  is_job_t* j = (is_job_t*)p;
  int64_t s = j->starts[g];
  int32_t G = (int32_t)(j->starts[g + 1] - s);
  uint32_t ss = j->seed + (uint32_t)g; /* CIS:243  seed + i */
  uint32_t idx = cwqo_is_encode_group(j->a + s, j->b + s, G, (uint32_t)j->K[g], ss, NULL, NULL);
  j->index_out[g] = idx;
  if (j->z_out && G > 0) cwqo_is_decode_group(G, idx, ss, j->z_out + s);
}
/* CIS:233-245 over groups [g0, g1) on n_threads host threads */
void cwqo_is_encode_groups(const float* a, const float* b, const int64_t* starts, const int32_t* K,
                           int64_t g0, int64_t g1, uint32_t seed, int n_threads, uint32_t* index_out, float* z_out) {
  is_job_t j = {a, b, starts, K, seed, index_out, z_out};
  parallel_for(g0, g1, n_threads, is_job, &j);
}

/* ---- greedy sampler, CGS:38-78.  shard scale s = f32(rho) / f32(sqrt(n_steps)) (CGS:42-45, p = N(0,1)).
 * inj: optional [n_steps, n_samples, G] draws.  indices_out[n_steps]; best_out[G]. */
float cwqo_greedy_shard_scale(float rho, int32_t n_steps) {
  return rho / (float)sqrt((double)n_steps);
}
/* what multiplies a RAW draw in one greedy step */
float cwqo_greedy_step_scale(float shard_scale) { return shard_scale * NORMAL_SCALE; }
float cwqo_normal_scale(void) { return NORMAL_SCALE; }

static inline uint32_t greedy_stream_seed(uint32_t group_seed, uint32_t step) {
  return 1000u * group_seed + step; /* CGS:55  seed = 1000 * seed + i, wrapping u32 */
}

/* shard_loc (may be NULL = 0): the per-dim loc of the proposal shard in rescaled coordinates, (p_loc / n_steps) / p_scale
 * (CGS:42 `loc=p_loc / n_steps`); every step adds it to the running sample before the candidates are scored (CGS:57). */
void cwqo_greedy_encode_group(const float* t_loc, const float* t_scale, const float* shard_loc, int32_t G, int32_t n_steps,
                              uint32_t n_samples, float shard_scale, uint32_t group_seed, const float* inj,
                              uint32_t* indices_out, float* best_out) {
  float* isig = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1));
  float* c = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1));
  float* e = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1));
  const float step_scale = shard_scale * NORMAL_SCALE;   /* the step's N(0,1) draw is NORMAL_SCALE * y */
  for (int32_t d = 0; d < G; ++d) { isig[d] = 1.0f / t_scale[d]; c[d] = step_scale * isig[d]; best_out[d] = 0.0f; }
  int32_t nq = (G + 3) / 4;
  for (int32_t step = 0; step < n_steps; ++step) {
    uint32_t ss = greedy_stream_seed(group_seed, (uint32_t)step);
    for (int32_t d = 0; d < G; ++d) {
      best_out[d] = best_out[d] + (shard_loc ? shard_loc[d] : 0.0f);   /* the step's loc; the chosen z is added below */
      e[d] = (best_out[d] - t_loc[d]) * isig[d];
    }
    float best = INFINITY;
    uint32_t best_k = 0;
    for (uint32_t k = 0; k < n_samples; ++k) {
      float acc0 = 0.0f, acc1 = 0.0f;
      uint32_t unit[3][4];
      for (int32_t q = 0; q < nq; ++q) {
        float z[4];
        if (inj) { for (int j = 0; j < 4; ++j) z[j] = (q * 4 + j < G) ? inj[((size_t)step * n_samples + k) * G + q * 4 + j] : 0.0f; }
        else raw_quad_seq(ss, k, (uint32_t)q, unit, z);
        for (int j = 0; j < 4 && q * 4 + j < G; ++j) {
          int d = q * 4 + j;
          float u = fmaf(z[j], c[d], e[d]);
          if (j & 1) acc1 = fmaf(u, u, acc1); else acc0 = fmaf(u, u, acc0);
        }
      }
      float s = acc0 + acc1;
      if (s < best) { best = s; best_k = k; }
    }
    indices_out[step] = best_k;
    for (int32_t q = 0; q < nq; ++q) {
      float z[4];
      if (inj) { for (int j = 0; j < 4; ++j) z[j] = (q * 4 + j < G) ? inj[((size_t)step * n_samples + best_k) * G + q * 4 + j] : 0.0f; }
      else raw_quad(ss, best_k, (uint32_t)q, z);
      for (int j = 0; j < 4 && q * 4 + j < G; ++j) best_out[q * 4 + j] = fmaf(step_scale, z[j], best_out[q * 4 + j]);
    }
  }
  free(isig); free(c); free(e);
}

/* CGS:134-165 with a counter-based stream */
void cwqo_greedy_decode_group(int32_t G, int32_t n_steps, float shard_scale, const float* shard_loc, uint32_t group_seed,
                              const uint32_t* indices, float* out) {
  for (int32_t d = 0; d < G; ++d) out[d] = 0.0f;
  int32_t nq = (G + 3) / 4;
  const float step_scale = shard_scale * NORMAL_SCALE;
  for (int32_t step = 0; step < n_steps; ++step) {
    uint32_t ss = greedy_stream_seed(group_seed, (uint32_t)step);
    for (int32_t q = 0; q < nq; ++q) {
      float z[4];
      raw_quad(ss, indices[step], (uint32_t)q, z);
      for (int j = 0; j < 4 && q * 4 + j < G; ++j) {
        int d = q * 4 + j;
        out[d] = fmaf(step_scale, z[j], out[d] + (shard_loc ? shard_loc[d] : 0.0f));
      }
    }
  }
}

/* shard loc of the per-group greedy entry points in rescaled coordinates: (p_loc / n_steps) / p_scale (CGS:42) */
void cwqo_greedy_shard_loc(const float* p_loc, const float* p_scale, int64_t n, int32_t n_steps, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = (p_loc[i] / (float)n_steps) / p_scale[i];
}

typedef struct { const float *t_loc, *t_scale; const int64_t* starts; int32_t n_steps; uint32_t n_samples; float shard_scale; uint32_t seed; uint32_t* indices_out; float* best_out; } gr_job_t;
static void gr_job(int64_t g, void* p) {
  gr_job_t* j = (gr_job_t*)p;
  int64_t s = j->starts[g];
  int32_t G = (int32_t)(j->starts[g + 1] - s);
  cwqo_greedy_encode_group(j->t_loc + s, j->t_scale + s, NULL, G, j->n_steps, j->n_samples, j->shard_scale,
                           j->seed + (uint32_t)g, NULL, j->indices_out + (size_t)g * j->n_steps, j->best_out + s);
}
/* CGS:273-284 over groups [g0, g1) on n_threads host threads */
void cwqo_greedy_encode_groups(const float* t_loc, const float* t_scale, const int64_t* starts, int64_t g0,
                               int64_t g1, int32_t n_steps, uint32_t n_samples, float shard_scale, uint32_t seed,
                               int n_threads, uint32_t* indices_out, float* best_out) {
  gr_job_t j = {t_loc, t_scale, starts, n_steps, n_samples, shard_scale, seed, indices_out, best_out};
  parallel_for(g0, g1, n_threads, gr_job, &j);
}

/* out[d] = fma(s, z[d], out[d]) -- the greedy running-sample update (CGS:57,63) on injected rows */
void cwqo_axpy_fma(float s, const float* z, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = fmaf(s, z[i], out[i]);
}

/* un-rescale (CIS:265, CGS:292): sample = p_scale * z + p_loc, two roundings */
void cwqo_unrescale(const float* z, const float* p_loc, const float* p_scale, int64_t n, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = p_scale[i] * z[i] + p_loc[i];
}

/* quint16 MIN_COMBINED over [-30, 30] (CIS:156, 351), spec section 3.8 */
uint16_t cwqo_quantize_u16(float x) {
  float c = x < -30.0f ? -30.0f : (x > 30.0f ? 30.0f : x);
  float v = floorf(fmaf(c + 30.0f, 1092.25f, 0.5f)); /* 65535 / 60 = 1092.25 exactly */
  if (!(v >= 0.0f)) v = 0.0f;
  if (v > 65535.0f) v = 65535.0f;
  return (uint16_t)v;
}
float cwqo_dequantize_u16(uint16_t q) { return fmaf((float)q, 60.0f / 65535.0f, -30.0f); }

/* Outlier dims (CIS:150): the reference draws an UNSEEDED target.sample(); the spec seeds it:
 * z = lane 0 of block(k = dim, q = 0) of stream (seed + 0x9E3779B9), value = q_scale * z + q_loc. */
float cwqo_outlier_normal(uint32_t seed, uint32_t dim) {
  float z[4];
  raw_quad(seed + 0x9E3779B9u, dim, 0u, z);
  return NORMAL_SCALE * z[0];
}

