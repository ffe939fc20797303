// cwq_p2p.cu -- the index exchange of a sharded run over NVLink peer memory instead of a collective call.
//
// Groups are independent (reference CIS:233-245), so every rank scores a range of groups and all ranks need all
// indices before the decode / packing stages.  With peer-mapped buffers the rank that found an index stores it straight
// into every rank's index array (4 bytes per group and peer, through NVSwitch), then raises a flag on every peer; a
// one-block kernel on each rank waits for the flags of all ranks before the finishing stages run.  Nothing in this
// exchange involves the host: no sizes to agree on, no collective launch.  Buffers are double-buffered by the parity
// of the exchange counter: a rank can only be one exchange ahead of the slowest peer (it needs the peer's flag to pass).
#include "cwq_common.cuh"
#include "cwq_spec.cuh"

namespace cwq {

using ull = unsigned long long;

struct PeerTable { void* p[CWQ_P2P_MAX_WORLD]; };

__global__ void __launch_bounds__(256) k_p2p_publish(const ull* __restrict__ keys, const long long* __restrict__ range,
                                                     long long capacity, PeerTable peers, int world, size_t byte_offset) {
  const long long g = range[0] + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= range[1] || g >= capacity) return;
  const uint32_t v = key_index(keys[g]);
  for (int r = 0; r < world; ++r) reinterpret_cast<uint32_t*>((char*)peers.p[r] + byte_offset)[g] = v;
}

// flags live at the start of every rank's exchange buffer: flags[parity * CWQ_P2P_MAX_WORLD + source rank] = epoch
__global__ void k_p2p_signal(PeerTable peers, int world, int rank, int parity, ull epoch) {
  __threadfence_system();   // the index stores of the preceding kernel are ordered before the flags, system wide
  const int r = threadIdx.x;
  if (r < world) {
    ull* flag = reinterpret_cast<ull*>(peers.p[r]) + parity * CWQ_P2P_MAX_WORLD + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
  }
}

__global__ void k_p2p_wait(const ull* __restrict__ local_flags, int world, int parity, ull epoch, long long timeout_cycles,
                           int* __restrict__ status) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const ull* flag = local_flags + parity * CWQ_P2P_MAX_WORLD + r;
  const long long t0 = clock64();
  for (;;) {
    ull v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > timeout_cycles) { atomicExch(status, 1 + r); break; }   // never hang the GPU: report and go on
    __nanosleep(200);
  }
}

}  // namespace cwq

using namespace cwq;

extern "C" size_t cwq_p2p_buffer_bytes(int64_t capacity_groups) {
  if (capacity_groups <= 0) return 0;
  return CWQ_P2P_HEADER_BYTES + 2 * align_up((size_t)capacity_groups * sizeof(uint32_t));
}

extern "C" int cwq_p2p_alloc(size_t bytes, void** buffer, uint8_t* h_handle64) {
  CWQ_REQUIRE(buffer && h_handle64 && bytes >= CWQ_P2P_HEADER_BYTES, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  CWQ_CUDA(cudaMalloc(&p, bytes));
  CWQ_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  CWQ_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(h_handle64, &h, 64);
  *buffer = p;
  return CWQ_OK;
}

extern "C" int cwq_p2p_open(const uint8_t* h_handle64, void** peer_buffer) {
  CWQ_REQUIRE(h_handle64 && peer_buffer, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, 64);
  CWQ_CUDA(cudaIpcOpenMemHandle(peer_buffer, h, cudaIpcMemLazyEnablePeerAccess));
  return CWQ_OK;
}

extern "C" int cwq_p2p_close(void* peer_buffer) {
  if (peer_buffer) CWQ_CUDA(cudaIpcCloseMemHandle(peer_buffer));
  return CWQ_OK;
}

extern "C" int cwq_p2p_free(void* buffer) {
  if (buffer) CWQ_CUDA(cudaFree(buffer));
  return CWQ_OK;
}

extern "C" int cwq_p2p_exchange_indices(const void* score_workspace, int64_t n_groups, const int64_t* range_dev,
                                        void* const* h_peer_buffers, int32_t world, int32_t rank, uint64_t epoch,
                                        int64_t capacity_groups, double timeout_seconds, uint32_t* index, void* stream) {
  CWQ_REQUIRE(score_workspace && range_dev && h_peer_buffers && index, "bad arguments");
  CWQ_REQUIRE(world >= 1 && world <= CWQ_P2P_MAX_WORLD && rank >= 0 && rank < world, "world must be 1..CWQ_P2P_MAX_WORLD");
  CWQ_REQUIRE(n_groups > 0 && n_groups <= capacity_groups && epoch > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  // keys of the scoring stage: second array of its workspace (see is_encode_impl: unit_off[n_groups + 1], keys[n_groups])
  Workspace ws(const_cast<void*>(score_workspace), (size_t)-1);
  ws.take<ull>((size_t)n_groups + 1);
  const ull* keys = ws.take<ull>((size_t)n_groups);
  PeerTable peers;
  for (int r = 0; r < CWQ_P2P_MAX_WORLD; ++r) peers.p[r] = r < world ? h_peer_buffers[r] : nullptr;
  const int parity = (int)(epoch & 1u);
  const size_t byte_offset = CWQ_P2P_HEADER_BYTES + (size_t)parity * align_up((size_t)capacity_groups * sizeof(uint32_t));
  k_p2p_publish<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(keys, (const long long*)range_dev, n_groups, peers, world,
                                                                 byte_offset);
  CWQ_LAUNCH_CHECK();
  k_p2p_signal<<<1, 32, 0, st>>>(peers, world, rank, parity, (ull)epoch);
  CWQ_LAUNCH_CHECK();
  int clock_khz = 1965000, device = 0;
  CWQ_CUDA(cudaGetDevice(&device));
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
  const long long timeout_cycles = (long long)(timeout_seconds * 1e3 * (double)clock_khz);
  int* status = reinterpret_cast<int*>((char*)h_peer_buffers[rank] + CWQ_P2P_STATUS_OFFSET);
  k_p2p_wait<<<1, 32, 0, st>>>(reinterpret_cast<const ull*>(h_peer_buffers[rank]), world, parity, (ull)epoch, timeout_cycles, status);
  CWQ_LAUNCH_CHECK();
  // out of the exchange buffer at once: a peer may overwrite this parity two exchanges from now, while the caller may
  // still be reading its index array (downloads on another stream)
  CWQ_CUDA(cudaMemcpyAsync(index, (const char*)h_peer_buffers[rank] + byte_offset, (size_t)n_groups * sizeof(uint32_t),
                           cudaMemcpyDeviceToDevice, st));
  return CWQ_OK;
}
