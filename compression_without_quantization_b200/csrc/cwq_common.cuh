// cwq_common.cuh -- error plumbing, segment lookup and a small device-wide scan used by the O(N)
// stages (grouping compaction, unit offsets, bit offsets, outlier compaction).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda_runtime.h>
#include "../../include/cwq.h"

namespace cwq {

void set_error(const std::string& msg);

#define CWQ_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      ::cwq::set_error(std::string(__func__) + ": " + (msg));    \
      return CWQ_ERR_ARG;                                        \
    }                                                            \
  } while (0)

#define CWQ_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      ::cwq::set_error(std::string(__func__) + ": " + #expr + ": " + cudaGetErrorString(e_));  \
      return CWQ_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// every kernel launch of the library goes through this: it also feeds cwq_launch_count()
void count_launches(int n);
#define CWQ_LAUNCH_CHECK()            \
  do {                                \
    ::cwq::count_launches(1);         \
    CWQ_CUDA(cudaGetLastError());     \
  } while (0)

// optional CUDA-event timing of kernels on their launching stream (cwq_profile_*): level 1 records the scoring kernel
// only, level 2 every kernel of the fused pipeline.  trace_begin returns a slot (or -1), trace_end closes it.
enum TraceKind { TR_SCORE = 0, TR_FRONT_TILES = 1, TR_FRONT_CHAIN = 2, TR_FRONT_GROUPS = 3, TR_FRONT_PREFIX = 4, TR_INDEX = 5,
                 TR_FINISH_DIMS = 6, TR_PACK = 7 };
int trace_begin(int kind, cudaStream_t st);
void trace_end(int slot, cudaStream_t st);

// Programmatic dependent launch along the kernel chain of the fused pipeline (tiles -> chain -> groups -> prefix -> score ->
// index -> finish -> pack): a kernel launched with the attribute may become resident while its predecessor on the stream is
// still running; it must call pdl_wait() before it touches anything the predecessor (or anything earlier on the stream)
// wrote, and the predecessor lets it in early with pdl_trigger().  Launch latency and grid ramp-up then overlap the tail of
// the previous kernel -- what a step of eight 10-30 us kernels is made of.  CWQ_PDL=0 in the environment turns it off.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t size, used;
  Workspace(void* p, size_t n) : base((char*)p), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (used + bytes > size) return nullptr;
    T* r = (T*)(base + used);
    used += bytes;
    return r;
  }
};

__device__ __forceinline__ int find_segment(const cwq_segment* __restrict__ segs, int n_seg, long long i) {
  int lo = 0, hi = n_seg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs[mid].start <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// Exclusive scan of in(i), i in [0, n), delivered to out(i, exclusive_prefix, value).
// Three launches: per-tile sums, one-block scan of the sums, per-tile scan + delivery.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem /*[33]*/, T& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? smem[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < nwarps) smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  T r = smem[warp] + incl - v;
  block_total = smem[32];
  __syncthreads();
  return r;
}

template <typename T, class InF>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(InF in, long long n, T* __restrict__ tile_sums) {
  __shared__ T sm[33];
  long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  T s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (base + j < n) s += in(base + j);
  T total;
  block_exclusive_scan<T>(s, sm, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(1024) k_scan_of_sums(T* __restrict__ tile_sums, long long nt, T* __restrict__ total_out) {
  __shared__ T sm[33];
  long long per = (nt + blockDim.x - 1) / blockDim.x;
  long long b = (long long)threadIdx.x * per, e = b + per < nt ? b + per : nt;
  T s = 0;
  for (long long i = b; i < e; ++i) s += tile_sums[i];
  T total;
  T off = block_exclusive_scan<T>(s, sm, total);
  for (long long i = b; i < e; ++i) { T v = tile_sums[i]; tile_sums[i] = off; off += v; }
  if (threadIdx.x == 0 && total_out) *total_out = total;
}

template <typename T, class InF, class OutF>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_deliver(InF in, OutF out, long long n, const T* __restrict__ tile_offsets) {
  __shared__ T sm[33];
  long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) { v[j] = (base + j < n) ? in(base + j) : T(0); s += v[j]; }
  T total;
  T pre = block_exclusive_scan<T>(s, sm, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out(base + j, pre, v[j]);
    pre += v[j];
  }
}

inline long long scan_tiles(long long n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p);
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v);

// ---------------------------------------------------------------------------------------------
// Exclusive prefix of per-tile aggregates across the CTAs of ONE kernel, two levels, no serial chain.
// A decoupled look-back degenerates when every tile is resident and publishes at the same moment (tile t then walks
// t / 32 windows back); here tiles are grouped in blocks of TP_BLOCK: every tile publishes its aggregate, the tile that
// publishes last in its block adds the block up, and a tile's prefix is [all earlier blocks] + [earlier tiles of its own
// block], summed in parallel by the whole CTA.  Tiles must take their ids from a ticket counter (every predecessor of a
// running tile is then running or done); all words start at zero; values stay below 2^62.  NV values per tile.
// ---------------------------------------------------------------------------------------------
constexpr int TP_BLOCK = 32;
constexpr unsigned long long TP_FLAG = 1ull << 63, TP_VAL = TP_FLAG - 1ull;

template <int NV>
struct TilePrefix {
  unsigned long long* tile_words;    // [NV][n_tiles]
  unsigned long long* block_words;   // [NV][n_blocks]
  unsigned int* block_count;         // [n_blocks]
  long long n_tiles, n_blocks;
  __host__ __device__ static size_t words(long long n_tiles) { return (size_t)NV * (size_t)(n_tiles + (n_tiles + TP_BLOCK - 1) / TP_BLOCK) + (size_t)((n_tiles + TP_BLOCK - 1) / TP_BLOCK + 1) / 2 + 1; }
  __host__ __device__ static TilePrefix carve(unsigned long long* base, long long n_tiles) {
    TilePrefix p;
    p.n_tiles = n_tiles;
    p.n_blocks = (n_tiles + TP_BLOCK - 1) / TP_BLOCK;
    p.tile_words = base;
    p.block_words = base + (size_t)NV * n_tiles;
    p.block_count = reinterpret_cast<unsigned int*>(base + (size_t)NV * (n_tiles + p.n_blocks));
    return p;
  }
  // ONE thread of tile t
  __device__ void publish(long long t, const unsigned long long (&agg)[NV]) const {
#pragma unroll
    for (int v = 0; v < NV; ++v) st_volatile_u64(&tile_words[(size_t)v * n_tiles + t], TP_FLAG | agg[v]);
    __threadfence();
    const long long b = t / TP_BLOCK;
    const long long first = b * TP_BLOCK;
    const int members = (int)(first + TP_BLOCK <= n_tiles ? TP_BLOCK : n_tiles - first);
    if (atomicAdd(&block_count[b], 1u) == (unsigned)(members - 1)) {
      __threadfence();
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        unsigned long long s = 0;
        for (int k = 0; k < members; ++k) s += ld_volatile_u64(&tile_words[(size_t)v * n_tiles + first + k]) & TP_VAL;
        st_volatile_u64(&block_words[(size_t)v * n_blocks + b], TP_FLAG | s);
      }
    }
  }
  // ALL threads of tile t (after publish); sm: [NV * 33] shared words.  Returns the exclusive prefixes to every thread.
  __device__ void exclusive(long long t, unsigned long long* sm, unsigned long long (&out)[NV]) const {
    const long long b = t / TP_BLOCK;
    const int in_block = (int)(t - b * TP_BLOCK);
    unsigned long long part[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) part[v] = 0;
    for (long long j = threadIdx.x; j < b + in_block; j += blockDim.x) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const unsigned long long* w = j < b ? &block_words[(size_t)v * n_blocks + j] : &tile_words[(size_t)v * n_tiles + b * TP_BLOCK + (j - b)];
        unsigned long long x;
        do { x = ld_volatile_u64(w); } while (!(x & TP_FLAG));
        part[v] += x & TP_VAL;
      }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part[v] += __shfl_xor_sync(0xffffffffu, part[v], o);
      if (lane == 0) sm[v * 33 + warp] = part[v];
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      unsigned long long s = 0;
      for (int k = 0; k < nwarps; ++k) s += sm[v * 33 + k];
      out[v] = s;
    }
    __syncthreads();
  }
};

// ---------------------------------------------------------------------------------------------
// Single-pass scan support: decoupled look-back over per-tile aggregates (one status word per tile).
// Tiles must take their ids from a ticket counter (atomicAdd), so that every predecessor of a running tile is itself
// running or done.  status[] must be zero before the kernel.  Called by ALL 32 lanes of ONE warp of the CTA;
// returns the exclusive prefix of `aggregate` over tiles [0, tile) to every lane and publishes the inclusive one.
// Values must stay below 2^62.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_VAL = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* status, long long tile,
                                                                 unsigned long long aggregate) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(&status[0], LB_INC | aggregate);
    return 0ull;
  }
  if (lane == 0) st_volatile_u64(&status[tile], LB_AGG | aggregate);
  unsigned long long excl = 0ull;
  long long look = tile - 1;
  for (;;) {
    const long long idx = look - lane;
    unsigned long long sv = LB_INC;                        // before tile 0: an inclusive prefix of 0
    if (idx >= 0) {
      do { sv = ld_volatile_u64(&status[idx]); } while ((sv >> 62) == 0ull);
    }
    const unsigned inc = __ballot_sync(0xffffffffu, (sv >> 62) == 2ull);
    const int first = inc ? __ffs(inc) - 1 : 31;           // nearest predecessor holding an inclusive prefix
    unsigned long long v = lane <= first ? (sv & LB_VAL) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (inc) break;
    look -= 32;
  }
  if (lane == 0) st_volatile_u64(&status[tile], LB_INC | (excl + aggregate));
  return excl;
}

// tile_sums must hold scan_tiles(n) entries; total_out (device) may be null
template <typename T, class InF, class OutF>
cudaError_t device_exclusive_scan(InF in, OutF out, long long n, T* tile_sums, T* total_out, cudaStream_t st) {
  if (n <= 0) {
    if (total_out) return cudaMemsetAsync(total_out, 0, sizeof(T), st);
    return cudaSuccess;
  }
  long long nt = scan_tiles(n);
  k_scan_tile_sums<T, InF><<<(unsigned)nt, SCAN_THREADS, 0, st>>>(in, n, tile_sums);
  k_scan_of_sums<T><<<1, 1024, 0, st>>>(tile_sums, nt, total_out);
  k_scan_deliver<T, InF, OutF><<<(unsigned)nt, SCAN_THREADS, 0, st>>>(in, out, n, tile_sums);
  count_launches(3);
  return cudaGetLastError();
}

}  // namespace cwq
