// Which pipes overlap?  Each thread runs independent dependency chains of IMAD.WIDE (W), MUFU (M),
// LOP3 (L) and FFMA (F) in a chosen ratio; comparing the time of a mix with the times of its parts
// shows which instruction classes share an issue/dispatch resource on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ITERS = 8192;

template <int NW, int NM, int NL, int NF, int NP = 0>
__global__ void __launch_bounds__(256) k_mix(uint32_t* out, uint32_t seed) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t w[NW > 0 ? NW : 1], l[NL > 0 ? NL : 1], pw[NP > 0 ? NP : 1];
  for (int i = 0; i < NP; ++i) pw[i] = tid * 747796405u + i * 3u + seed;
  float m[NM > 0 ? NM : 1], f[NF > 0 ? NF : 1];
  for (int i = 0; i < NW; ++i) w[i] = tid * 2654435761u + i + seed;
  for (int i = 0; i < NL; ++i) l[i] = tid * 40503u + i * 77u + seed;
  for (int i = 0; i < NM; ++i) m[i] = 1.0f + 1e-3f * (float)((tid + i) & 1023);
  for (int i = 0; i < NF; ++i) f[i] = 1.0f + 1e-4f * (float)((tid + 3 * i) & 1023);
  float c1 = 1.0f + 1e-7f * (float)seed, c2 = 1e-9f * (float)seed;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < NW; ++i) { uint32_t lo, hi; asm volatile("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(w[i]), "r"(0xD2511F53u)); w[i] = hi; l[NL > 0 ? i % (NL > 0 ? NL : 1) : 0] ^= (NL > 0 ? 0u : lo); if (NL == 0) w[i] ^= lo; }
#pragma unroll
      for (int i = 0; i < NP; ++i) {   // 32x32->64 product through the FP64 pipe: hi = lo32(fma.rz(2^52 + a, M * 2^-32, 2^52 - M * 2^20)), lo = a * M
        double a = __hiloint2double(0x43300000, (int)pw[i]);
        double h = __fma_rz(a, 3528531795.0 * 2.3283064365386963e-10, 4503599627370496.0 - 3528531795.0 * 1048576.0);
        uint32_t hi = (uint32_t)__double2loint(h), lo = pw[i] * 0xD2511F53u;
        pw[i] = hi; if (NL > 0) l[i % (NL > 0 ? NL : 1)] ^= lo; else pw[i] ^= lo;
      }
#pragma unroll
      for (int i = 0; i < NM; ++i) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m[i])); m[i] = r; }
#pragma unroll
      for (int i = 0; i < NL; ++i) { l[i] = (l[i] ^ seed ^ (uint32_t)it) & 0x7fffffffu | (l[i] >> 3); }
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i] = __fmaf_rn(f[i], c1, c2);
    }
  }
  uint32_t acc = 0;
  for (int i = 0; i < NW; ++i) acc ^= w[i];
  for (int i = 0; i < NP; ++i) acc ^= pw[i];
  for (int i = 0; i < NL; ++i) acc ^= l[i];
  for (int i = 0; i < NM; ++i) acc ^= __float_as_uint(m[i]);
  for (int i = 0; i < NF; ++i) acc ^= __float_as_uint(f[i]);
  if (acc == 0x12345u) out[tid & 1023] = acc;
}

template <int NW, int NM, int NL, int NF, int NP = 0>
float run(const char* name, int sms, uint32_t* d_out) {
  int grid = sms * 4, threads = 256;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_mix<NW, NM, NL, NF, NP><<<grid, threads>>>(d_out, 1); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(a)); k_mix<NW, NM, NL, NF, NP><<<grid, threads>>>(d_out, 2 + r); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  double groups = (double)grid * threads * ITERS * 4;   // one "group" = NW W + NM M + NL L(x~3 lop) + NF F
  printf("%-26s P=%d W=%d M=%d L=%d F=%d  %8.3f ms   %.3f ns per 1e6 thread-groups  (cycles/warp-group/SMSP @1.96GHz: %.2f)\n", name, NP, NW, NM, NL, NF, best,
         best * 1e6 / (groups / 1e6), best * 1e-3 * 1.96e9 / (groups / 32.0 / (sms * 4)));
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
  uint32_t* d_out; CK(cudaMalloc(&d_out, 4096));
  run<4, 0, 0, 0>("W only", sms, d_out);
  run<8, 0, 0, 0>("W only x8 chains", sms, d_out);
  run<0, 2, 0, 0>("M only", sms, d_out);
  run<0, 4, 0, 0>("M only x4 chains", sms, d_out);
  run<0, 0, 6, 0>("L only", sms, d_out);
  run<0, 0, 0, 5>("F only", sms, d_out);
  run<4, 2, 0, 0>("W+M", sms, d_out);
  run<4, 0, 6, 0>("W+L", sms, d_out);
  run<4, 0, 0, 5>("W+F", sms, d_out);
  run<0, 2, 6, 0>("M+L", sms, d_out);
  run<0, 2, 0, 5>("M+F", sms, d_out);
  run<0, 0, 6, 5>("L+F", sms, d_out);
  run<4, 2, 6, 0>("W+M+L", sms, d_out);
  run<4, 2, 6, 5>("W+M+L+F", sms, d_out);
  run<8, 4, 12, 10>("2x(W+M+L+F)", sms, d_out);
  // the score kernel's own mix per 5 candidate-dims: 18 IMAD.WIDE, 10 MUFU, ~28 ALU (9 L = 27), 25 FP32
  run<18, 10, 9, 25>("kernel mix / 5 cd", sms, d_out);
  run<18, 10, 0, 25>("kernel mix / 5 cd, no ALU", sms, d_out);
  run<18, 0, 0, 25>("kernel mix / 5 cd, W+F", sms, d_out);
  run<18, 10, 0, 0>("kernel mix / 5 cd, W+M", sms, d_out);
  run<9, 5, 4, 12>("kernel mix / 2.5 cd", sms, d_out);
  // 3 Philox blocks per 16 dims instead of 4 (radius words + shared angle word): 44 instead of 58 IMAD.WIDE per candidate
  run<0, 0, 0, 0, 4>("P only (DFMA hi + IMAD lo)", sms, d_out);
  run<0, 0, 0, 0, 8>("P only x8", sms, d_out);
  run<0, 10, 9, 25, 18>("kernel mix, products via FP64", sms, d_out);
  run<9, 10, 9, 25, 9>("kernel mix, half via FP64", sms, d_out);
  run<12, 10, 9, 25, 6>("kernel mix, third via FP64", sms, d_out);
  // stream v2 (3 blocks + harvested quad per 16 dims): 44 IMAD.WIDE, 32 MUFU, ~74 ALU (6 L = 18 + 11 xor), 81 FP32 per 16 cd
  run<11, 8, 6, 20>("v2 mix / 4 cd", sms, d_out);
  run<11, 8, 0, 20>("v2 mix / 4 cd, no ALU", sms, d_out);
  run<11, 0, 0, 20>("v2 mix / 4 cd, W+F", sms, d_out);
  run<11, 8, 0, 0>("v2 mix / 4 cd, W+M", sms, d_out);
  run<0, 8, 0, 20>("v2 mix / 4 cd, M+F", sms, d_out);
  run<22, 16, 12, 40>("v2 mix / 8 cd", sms, d_out);
  run<14, 10, 9, 25>("3-block mix / 5 cd", sms, d_out);
  run<14, 10, 10, 25>("3-block mix / 5 cd, +3 ALU", sms, d_out);
  run<14, 10, 9, 23>("3-block mix / 5 cd, -2 FP", sms, d_out);
  return 0;
}
