// Inner-loop variants of the importance-sampler scoring kernel (k_is_score), timed in isolation:
// every lane scores `ncand` candidates of a 16-dim group (4 Philox blocks, 8 Box-Muller pairs,
// 16 score terms per candidate).  Used to pick the code shape that goes into cwq_samplers.cu.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false is_loop.cu -o is_loop
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

template <bool ASM>
__device__ __forceinline__ void mulhilo(uint32_t a, uint32_t m, uint32_t& hi, uint32_t& lo) {
  if (ASM) {
    asm("{ .reg .b64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0, %1}, p; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(m));
  } else {
    unsigned long long p = (unsigned long long)a * m; hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
  }
}

template <bool ASM>
__device__ __forceinline__ void block(uint32_t k, uint32_t q, uint32_t s, uint32_t (&o)[4]) {
  uint32_t c0 = k, c1 = q, c2 = 0, c3 = 0, k0 = s, k1 = 42u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo<ASM>(c0, M0, h0, l0); mulhilo<ASM>(c2, M1, h1, l1);
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0; k0 += W0; k1 += W1;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& ze, float& zo) {
  float f1 = __uint_as_float((xa & 0x007fffffu) | 0x3f800000u);
  float f2 = __uint_as_float((xb & 0x007fffffu) | 0x3f800000u);
  float u1 = __fsub_rn(2.0f, f1);
  float th = __fmaf_rn(f2, 6.2831855f, -9.424778f);
  float l, r, c, s;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  float r2 = __fmul_rn(l, -1.3862944f);
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  ze = __fmul_rn(r, c); zo = __fmul_rn(r, s);
}

// uniforms through I2FP (FMA-pipe conversion) instead of LOP3 mantissa masks (ALU pipe)
__device__ __forceinline__ void box_muller_cvt(uint32_t xa, uint32_t xb, float& ze, float& zo) {
  float fa = __uint2float_rn(xa), fb = __uint2float_rn(xb);
  float u1 = __fmaf_rn(fa, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (x + 0.5) * 2^-32 in (0, 1]
  float th = __fmaf_rn(fb, 1.4629180792671596e-9f, -3.1415927f);                 // 2 pi * x * 2^-32 - pi
  float l, r, c, s;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  float r2 = __fmul_rn(l, -1.3862944f);
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  ze = __fmul_rn(r, c); zo = __fmul_rn(r, s);
}

// ---- packed fp32x2 helpers (sm_100: FFMA2 / FMUL2) ----
typedef unsigned long long f2x;
__device__ __forceinline__ f2x pack2(float a, float b) { f2x r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f2x v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2x fma2(f2x a, f2x b, f2x c) { f2x r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2x mul2(f2x a, f2x b) { f2x r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// packed Box-Muller: returns (ze, zo) as one f32x2
__device__ __forceinline__ f2x box_muller2(uint32_t xa, uint32_t xb, f2x c_mul, f2x c_add) {
  float f1 = __uint_as_float((xa & 0x007fffffu) | 0x3f800000u);
  float f2 = __uint_as_float((xb & 0x007fffffu) | 0x3f800000u);
  f2x ut = fma2(pack2(f1, f2), c_mul, c_add);   // (2 - f1, 2pi f2 - 3pi)
  float u1, th; unpack2(ut, u1, th);
  float l, r, c, s;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  float r2 = __fmul_rn(l, -1.3862944f);
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  return mul2(pack2(r, r), pack2(c, s));
}

// VARIANT 0: C multiply, scalar FP, coefs in smem (what v0 of the product does)
// VARIANT 1: asm mul.wide
// VARIANT 2: asm mul.wide + packed f32x2 for Box-Muller affine, r*(c,s) and the score
// VARIANT 3: variant 1 with coefficients in registers
// VARIANT 4: variant 2 with coefficients in registers
template <int V, int MINB>
__global__ void __launch_bounds__(256, MINB) k_loop(uint32_t* out, uint32_t seed, int ncand, const float4* __restrict__ coef4,
                                                    long long* cyc) {
  __shared__ float4 sc[8];          // (a0,b0,a1,b1) per dim pair [V 0,1,3]  or (a0,a1,b0,b1) [V 2,4]
  if (threadIdx.x < 8) sc[threadIdx.x] = coef4[threadIdx.x];
  __syncthreads();
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  float best = -3.0e38f; uint32_t bestk = 0;
  float4 rc[8];
  if (V >= 3) {
#pragma unroll
    for (int i = 0; i < 8; ++i) rc[i] = sc[i];
  }
  const f2x c_mul = pack2(-1.0f, 6.2831855f), c_add = pack2(2.0f, -9.424778f);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ncand; ++it) {
    uint32_t k = tid + it * stride;
    if (V == 2 || V == 4) {
      f2x acc = pack2(0.0f, 0.0f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o[4];
        block<true>(k, q, seed, o);
        f2x z01 = box_muller2(o[0], o[1], c_mul, c_add);
        f2x z23 = box_muller2(o[2], o[3], c_mul, c_add);
        float4 c01 = V == 4 ? rc[2 * q] : sc[2 * q], c23 = V == 4 ? rc[2 * q + 1] : sc[2 * q + 1];
        acc = fma2(z01, fma2(pack2(c01.x, c01.y), z01, pack2(c01.z, c01.w)), acc);
        acc = fma2(z23, fma2(pack2(c23.x, c23.y), z23, pack2(c23.z, c23.w)), acc);
      }
      float a0, a1; unpack2(acc, a0, a1);
      float s = __fadd_rn(a0, a1);
      if (s > best) { best = s; bestk = k; }
    } else {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o[4];
        if (V == 0) block<false>(k, q, seed, o); else block<true>(k, q, seed, o);
        float z[4];
        if (V == 7) { box_muller_cvt(o[0], o[1], z[0], z[1]); box_muller_cvt(o[2], o[3], z[2], z[3]); }
        else { box_muller(o[0], o[1], z[0], z[1]); box_muller(o[2], o[3], z[2], z[3]); }
        float4 c01 = V == 3 ? rc[2 * q] : sc[2 * q], c23 = V == 3 ? rc[2 * q + 1] : sc[2 * q + 1];
        acc0 = __fmaf_rn(z[0], __fmaf_rn(c01.x, z[0], c01.y), acc0);
        acc1 = __fmaf_rn(z[1], __fmaf_rn(c01.z, z[1], c01.w), acc1);
        acc0 = __fmaf_rn(z[2], __fmaf_rn(c23.x, z[2], c23.y), acc0);
        acc1 = __fmaf_rn(z[3], __fmaf_rn(c23.z, z[3], c23.w), acc1);
      }
      float s = __fadd_rn(acc0, acc1);
      if (s > best) { best = s; bestk = k; }
    }
  }
  long long t1 = clock64();
  if ((bestk ^ __float_as_uint(best)) == 0x12345u) out[tid & 1023] = bestk;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ---- software-pipelined variants: the instruction stream of every warp is a uniform mix ----
struct PhState { uint32_t c0, c1, c2, c3; };
#define MULWIDE_V(lo, hi, a, m) asm volatile("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(m))
#define MULWIDE(lo, hi, a, m) asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(m))
// rounds 0 and 1 of counter (k, q, 0, 0): M1*0 folds away and M0*(q ^ key0) is loop-invariant (uh, ul)
__device__ __forceinline__ void ph_round01(PhState& p, uint32_t k, uint32_t uh, uint32_t ul, uint32_t k1_0, uint32_t k1_1) {
  uint32_t h0, l0, h1, l1;
  MULWIDE_V(l0, h0, k, M0);                       // round 0: c0 = k
  uint32_t c2 = h0 ^ k1_0, c3 = l0;               // c0' = q ^ k0 (uniform), c1' = 0
  MULWIDE_V(l1, h1, c2, M1);                      // round 1
  p.c0 = h1 ^ 0u ^ 0u; p.c0 = h1;                 // n0 = h1 ^ c1'(0) ^ k0_1  -> key folded by caller through uh? (kept exact below)
  p.c1 = l1; p.c2 = uh ^ c3 ^ k1_1; p.c3 = ul;
}
__device__ __forceinline__ void ph_round(PhState& p, uint32_t k0, uint32_t k1) {
  uint32_t h0, l0, h1, l1;
  MULWIDE_V(l0, h0, p.c0, M0);
  MULWIDE_V(l1, h1, p.c2, M1);
  uint32_t n0 = h1 ^ p.c1 ^ k0, n2 = h0 ^ p.c3 ^ k1;
  p.c0 = n0; p.c1 = l1; p.c2 = n2; p.c3 = l0;
}
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sqrt(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_cos(float x) { float r; asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sin(float x) { float r; asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int V, int MINB>
__global__ void __launch_bounds__(256, MINB) k_pipe5(uint32_t* out, uint32_t seed, int ncand, const float4* __restrict__ coef4,
                                                     long long* cyc) {
  __shared__ float4 sc[8];
  if (threadIdx.x < 8) sc[threadIdx.x] = coef4[threadIdx.x];
  __syncthreads();
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  float best = -3.0e38f; uint32_t bestk = 0;
  uint32_t rk[10];
#pragma unroll
  for (int r = 0; r < 10; ++r) rk[r] = seed + (uint32_t)r * W0;
  uint32_t uh[4], ul[4];   // M0 * (q ^ key0): round-1 product, invariant over candidates
#pragma unroll
  for (int q = 0; q < 4; ++q) MULWIDE(ul[q], uh[q], (uint32_t)q ^ rk[0], M0);
  uint32_t w[16];   // Philox words of the candidate being scored
  {
    PhState p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { ph_round01(p[q], tid, uh[q], ul[q], 42u, 42u + W1); p[q].c0 ^= rk[1]; }
#pragma unroll
    for (int r = 2; r < 10; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) ph_round(p[q], rk[r], 42u + (uint32_t)r * W1);
#pragma unroll
    for (int q = 0; q < 4; ++q) { w[4 * q] = p[q].c0; w[4 * q + 1] = p[q].c1; w[4 * q + 2] = p[q].c2; w[4 * q + 3] = p[q].c3; }
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ncand; ++it) {
    const uint32_t k = tid + it * stride;
    PhState p[4];
    float acc0 = 0.f, acc1 = 0.f;
    // tick t: Philox round(s) of the NEXT candidate (4 blocks) + Box-Muller pair t of THIS candidate (t < 8)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (t == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { ph_round01(p[q], k + stride, uh[q], ul[q], 42u, 42u + W1); p[q].c0 ^= rk[1]; }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) ph_round(p[q], rk[t + 1], 42u + (uint32_t)(t + 1) * W1);
      }
      if (t < 8) {
        float f1 = __uint_as_float((w[2 * t] & 0x007fffffu) | 0x3f800000u);
        float f2 = __uint_as_float((w[2 * t + 1] & 0x007fffffu) | 0x3f800000u);
        float u1 = __fsub_rn(2.0f, f1);
        float th = __fmaf_rn(f2, 6.2831855f, -9.424778f);
        float l = mufu_lg2(u1);
        float r = mufu_sqrt(__fmul_rn(l, -1.3862944f));
        float c = mufu_cos(th), sn = mufu_sin(th);
        float ze = __fmul_rn(r, c), zo = __fmul_rn(r, sn);
        float4 cf = sc[t];
        acc0 = __fmaf_rn(ze, __fmaf_rn(cf.x, ze, cf.y), acc0);
        acc1 = __fmaf_rn(zo, __fmaf_rn(cf.z, zo, cf.w), acc1);
      }
    }
    float s = __fadd_rn(acc0, acc1);
    if (s > best) { best = s; bestk = k; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { w[4 * q] = p[q].c0; w[4 * q + 1] = p[q].c1; w[4 * q + 2] = p[q].c2; w[4 * q + 3] = p[q].c3; }
  }
  long long t1 = clock64();
  if ((bestk ^ __float_as_uint(best)) == 0x12345u) out[tid & 1023] = bestk;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V, int MINB>
void run5(const char* name, int sms, int bps, uint32_t* d_out, long long* d_cyc, const float4* d_coef) {
  int grid = sms * bps, threads = 256, ncand = 2048 / bps * 2;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_pipe5<V, MINB>));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pipe5<V, MINB>, threads, 0));
  if (occ < bps) { printf("%-34s regs=%3d occ=%d < bps=%d, skipped\n", name, fa.numRegs, occ, bps); return; }
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_pipe5<V, MINB><<<grid, threads>>>(d_out, 7, ncand, d_coef, d_cyc);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    k_pipe5<V, MINB><<<grid, threads>>>(d_out, 7 + rep, ncand, d_coef, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best_ms) best_ms = ms;
  }
  double cd = (double)grid * threads * ncand * 16.0;
  printf("%-34s regs=%3d bps=%d  %8.3f ms  %.3e cand-dims/s\n", name, fa.numRegs, bps, best_ms, cd / (best_ms * 1e-3));
}

template <int V, int MINB>
void run(const char* name, int sms, int bps, uint32_t* d_out, long long* d_cyc, const float4* d_coef) {
  int grid = sms * bps, threads = 256, ncand = 2048 / bps * 2;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_loop<V, MINB>));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_loop<V, MINB>, threads, 0));
  if (occ < bps) { printf("%-34s regs=%3d occ=%d < bps=%d, skipped\n", name, fa.numRegs, occ, bps); return; }
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_loop<V, MINB><<<grid, threads>>>(d_out, 7, ncand, d_coef, d_cyc);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    k_loop<V, MINB><<<grid, threads>>>(d_out, 7 + rep, ncand, d_coef, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best_ms) best_ms = ms;
  }
  long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double cd = (double)grid * threads * ncand * 16.0;
  printf("%-34s regs=%3d bps=%d  %8.3f ms  ~%4.0f MHz  %6.2f cd/clk/SM  %.3e cand-dims/s\n", name, fa.numRegs, bps, best_ms,
         cyc / (best_ms * 1e3), cd / ((double)cyc * sms), cd / (best_ms * 1e-3));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device: %s SMs=%d\n", p.name, sms);
  uint32_t* d_out; long long* d_cyc; float4* d_coef;
  CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_coef, 8 * sizeof(float4)));
  float4 h[8]; for (int i = 0; i < 8; ++i) { h[i] = make_float4(-0.5f - 0.1f * i, 0.3f * (i - 4), -0.7f + 0.05f * i, 0.2f * (i - 3)); }
  CK(cudaMemcpy(d_coef, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int bps : {2, 3, 4, 6, 8}) {
    run<0, 1>("V0 C-mul scalar smem", sms, bps, d_out, d_cyc, d_coef);
    run<1, 1>("V1 asm-mul scalar smem", sms, bps, d_out, d_cyc, d_coef);
    run<2, 1>("V2 asm-mul f32x2 smem", sms, bps, d_out, d_cyc, d_coef);
    run<3, 1>("V3 asm-mul scalar regs", sms, bps, d_out, d_cyc, d_coef);
    run<4, 1>("V4 asm-mul f32x2 regs", sms, bps, d_out, d_cyc, d_coef);
    run<7, 1>("V7 asm-mul I2FP uniforms", sms, bps, d_out, d_cyc, d_coef);
  }
  run5<5, 1>("V5 sw-pipelined scalar", sms, 2, d_out, d_cyc, d_coef);
  run5<5, 3>("V5 sw-pipelined scalar minb3", sms, 3, d_out, d_cyc, d_coef);
  run5<5, 4>("V5 sw-pipelined scalar minb4", sms, 4, d_out, d_cyc, d_coef);
  // register-capped builds
  run<1, 4>("V1 minb4", sms, 4, d_out, d_cyc, d_coef);
  run<7, 4>("V7 minb4", sms, 4, d_out, d_cyc, d_coef);
  run<7, 3>("V7 minb3", sms, 3, d_out, d_cyc, d_coef);
  run<2, 4>("V2 minb4", sms, 4, d_out, d_cyc, d_coef);
  run<1, 6>("V1 minb6", sms, 6, d_out, d_cyc, d_coef);
  run<2, 6>("V2 minb6", sms, 6, d_out, d_cyc, d_coef);
  run<1, 8>("V1 minb8", sms, 8, d_out, d_cyc, d_coef);
  run<2, 8>("V2 minb8", sms, 8, d_out, d_cyc, d_coef);
  return 0;
}
