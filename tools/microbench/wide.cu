// IMAD.WIDE operand-form check: immediate vs uniform-register vs per-lane register multiplier.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ITERS = 8192;
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t m_param, const uint32_t* mptr) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t w[8];
  for (int i = 0; i < 8; ++i) w[i] = tid * 2654435761u + i;
  uint32_t m = MODE == 0 ? 0xD2511F53u : (MODE == 1 ? m_param : mptr[tid & 31]);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int i = 0; i < 8; ++i) { uint32_t lo, hi; asm volatile("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(w[i]), "r"(m)); w[i] = hi ^ lo; }
  }
  uint32_t acc = 0; for (int i = 0; i < 8; ++i) acc ^= w[i];
  if (acc == 0x12345u) out[tid & 1023] = acc;
}
template <int MODE> void run(const char* name, int sms, uint32_t* d_out, uint32_t* d_m) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k<MODE><<<sms * 4, 256>>>(d_out, 0xD2511F53u, d_m); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); k<MODE><<<sms * 4, 256>>>(d_out, 0xD2511F53u, d_m); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  double ops = (double)sms * 4 * 256 * ITERS * 32;
  printf("%-28s %8.3f ms  %.3e IMAD.WIDE/s  cycles/warp-instr/SMSP @1.96GHz: %.2f\n", name, ms, ops / (ms * 1e-3), ms * 1e-3 * 1.96e9 / (ops / 32 / (sms * 4)));
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
  uint32_t *d_out, *d_m; CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_m, 128));
  uint32_t h[32]; for (int i = 0; i < 32; ++i) h[i] = 0xD2511F53u; CK(cudaMemcpy(d_m, h, 128, cudaMemcpyHostToDevice));
  run<0>("immediate multiplier (+LOP)", sms, d_out, d_m);
  run<1>("uniform-register multiplier", sms, d_out, d_m);
  run<2>("per-lane register multiplier", sms, d_out, d_m);
  return 0;
}
