// cwq_group.cu -- per-dim prologues and the KL-budget grouping (reference CIS:133-204, CGS:194-252).
//
// The reference packs dims into groups with a sequential fp32 accumulator; that recurrence is not
// associative, so a prefix sum would move boundaries.  Here it is evaluated EXACTLY but in parallel:
//   1. k_next:       next[i] = where the following boundary falls if a group starts at dim i
//                    (a bounded sequential fp32 walk per dim, all dims in parallel);
//   2. k_tile_exit:  pointer jumping inside 4096-dim tiles -> exit1[i] = first chain node past the tile;
//   3. k_super_exit: the same over 32-tile super-tiles, only for the few possible entry nodes;
//   4. k_chain_top / k_tile_entries / k_mark: follow the one true chain 0 -> ... -> n top-down;
//   5. a device scan compacts the marked boundaries into starts[], gid[], per-segment offsets.
#include "cwq_common.cuh"
#include "cwq_spec.cuh"

namespace cwq {

// ------------------------------------------------------------------------------------------
// prologues
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_is_prologue(const float* __restrict__ q_loc, const float* __restrict__ q_scale,
                                                     const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                                                     long long n, const cwq_segment* __restrict__ segs, int n_seg,
                                                     float* __restrict__ kl, float2* __restrict__ coef,
                                                     uint8_t* __restrict__ mask) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cwq_segment& sg = segs[find_segment(segs, n_seg, i)];
  const PrologueDim r = is_prologue_dim(q_loc[i], q_scale[i], p_loc[i], p_scale[i], sg.dim_kl_bit_limit);
  kl[i] = r.kl;
  coef[i] = r.coef;
  mask[i] = r.m | (i == sg.start ? 4 : 0);   // bit 2: first dim of a segment (outlier compaction reads its per-segment offsets here)
}

__global__ void __launch_bounds__(256) k_greedy_prologue(const float* __restrict__ q_loc, const float* __restrict__ q_scale,
                                                         const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                                                         long long n, float* __restrict__ kl, float2* __restrict__ tparam) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ps = p_scale[i];
  float tl = __fdiv_rn(__fsub_rn(q_loc[i], p_loc[i]), ps);
  float ts = __fdiv_rn(q_scale[i], ps);
  kl[i] = kl01(tl, ts);
  tparam[i] = make_float2(tl, __fdiv_rn(1.0f, ts));
}

// ------------------------------------------------------------------------------------------
// grouping
// ------------------------------------------------------------------------------------------
constexpr int TILE1 = 4096;          // nodes per tile (its next[] slice lives in shared memory: 16 KB)
constexpr int TILES_PER_SUPER = 32;  // tiles per super-tile
constexpr long long TILE2 = (long long)TILE1 * TILES_PER_SUPER;

// the reference's loop body from state (size, acc) at dim idx0, until the next boundary
__device__ __forceinline__ long long walk(const float* __restrict__ kl, long long idx0, long long size, float acc,
                                          const cwq_segment& sg) {
  const long long end = sg.end;
  const double budget = sg.n_nats_per_group;
  const long long limit = sg.size_limit;
  const bool strict = sg.strict != 0;
  for (long long idx = idx0; idx < end; ++idx) {
    float cand = __fadd_rn(acc, kl[idx]);
    bool over = strict ? ((double)cand > budget) : ((double)cand >= budget);
    if (size >= limit || over || idx == end - 1) return idx;
    acc = cand;
    size += 1;
  }
  return end;
}

// thread 0: virtual start of segment 0 -> *first; thread i+1: node i -> next[i]
__global__ void __launch_bounds__(256) k_next(const float* __restrict__ kl, long long n, const cwq_segment* __restrict__ segs,
                                              int n_seg, int* __restrict__ next, int* __restrict__ first) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n) return;
  if (t == 0) { *first = (int)walk(kl, segs[0].start, 0, 0.0f, segs[0]); return; }
  long long i = t - 1;
  int s = find_segment(segs, n_seg, i);
  cwq_segment sg = segs[s];
  long long r;
  if (i == sg.end - 1) {
    // last dim of a segment: the chain continues with the next segment's first boundary
    r = (s + 1 < n_seg) ? walk(kl, segs[s + 1].start, 0, 0.0f, segs[s + 1]) : n;
  } else {
    r = walk(kl, i + 1, 1, kl[i], sg);
  }
  next[i] = (int)r;
}

__global__ void __launch_bounds__(256) k_tile_exit(const int* __restrict__ next, long long n, int* __restrict__ exit1) {
  __shared__ int ptr[TILE1];
  const long long tile0 = (long long)blockIdx.x * TILE1;
  const long long tile_end = tile0 + TILE1 < n ? tile0 + TILE1 : n;
  for (int j = threadIdx.x; j < TILE1; j += blockDim.x) ptr[j] = (tile0 + j < n) ? next[tile0 + j] : (int)n;
  __syncthreads();
  for (;;) {
    int nv[TILE1 / 256];
    int changed = 0;
#pragma unroll
    for (int r = 0; r < TILE1 / 256; ++r) {
      int j = threadIdx.x + r * 256;
      int v = ptr[j];
      if (v < tile_end) { v = ptr[v - tile0]; changed = 1; }
      nv[r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TILE1 / 256; ++r) ptr[threadIdx.x + r * 256] = nv[r];
    if (!__syncthreads_or(changed)) break;
  }
  for (int j = threadIdx.x; j < TILE1; j += blockDim.x)
    if (tile0 + j < n) exit1[tile0 + j] = ptr[j];
}

// exit2[u * E2 + e] = first chain node at or past the end of super-tile u, entering at its e-th node
__global__ void __launch_bounds__(256) k_super_exit(const int* __restrict__ exit1, long long n, int E2, long long n_super,
                                                    int* __restrict__ exit2) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_super * E2) return;
  long long u = t / E2;
  int e = (int)(t - u * E2);
  long long u0 = u * TILE2, u_end = u0 + TILE2 < n ? u0 + TILE2 : n;
  long long v = u0 + e;
  if (v >= n) { exit2[t] = (int)n; return; }
  while (v < u_end) v = exit1[v];
  exit2[t] = (int)v;
}

__global__ void k_chain_top(const int* __restrict__ first, const int* __restrict__ exit2, long long n, int E2,
                            int* __restrict__ entry2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long v = *first;
  while (v < n) {
    long long u = v / TILE2;
    entry2[u] = (int)v;
    v = exit2[u * E2 + (v - u * TILE2)];
  }
}

__global__ void __launch_bounds__(128) k_tile_entries(const int* __restrict__ entry2, const int* __restrict__ exit1,
                                                      long long n, long long n_super, int* __restrict__ entry1) {
  long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_super) return;
  long long v = entry2[u];
  if (v < 0) return;
  long long u_end = (u + 1) * TILE2 < n ? (u + 1) * TILE2 : n;
  while (v < u_end) {
    entry1[v / TILE1] = (int)v;
    v = exit1[v];
  }
}

// one CTA per tile: the tile's next[] slice is staged in shared memory (coalesced), one thread follows the
// chain through it setting flags in shared memory, then the whole tile's flags are written out coalesced
// (every byte of mark[] is written here, so no memset is needed)
__global__ void __launch_bounds__(256) k_mark(const int* __restrict__ entry1, const int* __restrict__ next, long long n,
                                              long long n_tiles, uint8_t* __restrict__ mark) {
  __shared__ int snext[TILE1];
  __shared__ uint8_t sflag[TILE1];
  const long long t = blockIdx.x;
  if (t >= n_tiles) return;
  const long long v0 = entry1[t];
  const long long tile0 = t * TILE1;
  const long long t_end = tile0 + TILE1 < n ? tile0 + TILE1 : n;
  for (int j = threadIdx.x; j < TILE1; j += blockDim.x) {
    sflag[j] = 0;
    if (v0 >= 0) snext[j] = (tile0 + j < n) ? next[tile0 + j] : (int)n;
  }
  __syncthreads();
  if (threadIdx.x == 0 && v0 >= 0) {
    long long v = v0;
    while (v < t_end) {
      sflag[v - tile0] = 1;
      v = snext[v - tile0];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < TILE1; j += blockDim.x)
    if (tile0 + j < n) mark[tile0 + j] = sflag[j];
}

__global__ void k_mark_segment_starts(const cwq_segment* __restrict__ segs, int n_seg, uint8_t* __restrict__ mark) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg) mark[segs[s].start] |= 2;   // runs after k_mark on the same stream; one writer per byte
}

struct MarkCount {
  const uint8_t* mark;
  __device__ uint32_t operator()(long long i) const { uint8_t m = mark[i]; return (uint32_t)(m & 1) + (uint32_t)((m >> 1) & 1); }
};

struct GroupDeliver {
  const uint8_t* mark;
  const cwq_segment* segs;
  int n_seg;
  long long* starts;
  int* gid;
  long long* seg_group_off;
  __device__ void operator()(long long i, uint32_t base, uint32_t cnt) const {
    uint8_t m = mark[i];
    if (m & 2) {  // first group of a segment, [start, first boundary)
      starts[base] = i;
      seg_group_off[find_segment(segs, n_seg, i)] = (long long)base;
    }
    if (m & 1) starts[base + ((m >> 1) & 1)] = i;
    gid[i] = (int)(base + cnt) - 1;
  }
};

__global__ void k_group_finish(const uint32_t* __restrict__ n_groups, long long n, int n_seg, long long* __restrict__ starts,
                               long long* __restrict__ seg_group_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    starts[*n_groups] = n;
    seg_group_off[n_seg] = (long long)*n_groups;
  }
}

__global__ void __launch_bounds__(256) k_group_seed(const long long* __restrict__ starts, const uint32_t* __restrict__ n_groups_p,
                                                    const cwq_segment* __restrict__ segs, int n_seg,
                                                    const long long* __restrict__ seg_group_off, uint32_t* __restrict__ group_seed) {
  long long ng = *n_groups_p;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (long long)gridDim.x * blockDim.x) {
    int s = find_segment(segs, n_seg, starts[g]);
    group_seed[g] = segs[s].seed + (uint32_t)(g - seg_group_off[s]);
  }
}

__global__ void __launch_bounds__(256) k_group_totals(const float* __restrict__ kl, const long long* __restrict__ starts,
                                                      long long n_groups, const long long* __restrict__ ng_dev,
                                                      float* __restrict__ total_kl, int* __restrict__ K) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ng_dev) n_groups = *ng_dev;
  if (g >= n_groups) return;
  float acc = 0.0f;
  for (long long i = starts[g], e = starts[g + 1]; i < e; ++i) acc = __fadd_rn(acc, kl[i]);
  total_kl[g] = acc;
  K[g] = candidates_from_kl(acc);
}

// decoder side: group tables from given starts.  One thread per group paints gid over its dims.
__global__ void __launch_bounds__(256) k_expand_groups(const long long* __restrict__ starts, long long n_groups, long long n,
                                                       const cwq_segment* __restrict__ segs, int n_seg,
                                                       const long long* __restrict__ seg_group_off, int* __restrict__ gid,
                                                       uint32_t* __restrict__ group_seed) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  int lo = 0, hi = n_seg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (seg_group_off[mid] <= g) lo = mid; else hi = mid - 1;
  }
  group_seed[g] = segs[lo].seed + (uint32_t)(g - seg_group_off[lo]);
  long long b = starts[g], e = starts[g + 1];
  if (b < 0) b = 0;
  if (e > n) e = n;
  for (long long i = b; i < e; ++i) gid[i] = (int)g;
}

static int max_entry_span(const cwq_segment* h_segs, int n_seg, long long* out) {
  long long m = 0;
  for (int s = 0; s < n_seg; ++s) m = h_segs[s].size_limit > m ? h_segs[s].size_limit : m;
  *out = m + 2;  // hop length <= size_limit + 1
  return 0;
}

}  // namespace cwq

using namespace cwq;

extern "C" int cwq_is_prologue(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                               int64_t n, const cwq_segment* segments, int32_t n_segments, float* kl, float* coef,
                               uint8_t* mask, void* stream) {
  CWQ_REQUIRE(n > 0 && n < 2147483000LL, "n out of range");
  CWQ_REQUIRE(n_segments > 0 && segments, "need at least one segment");
  k_is_prologue<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_loc, q_scale, p_loc, p_scale, n, segments,
                                                                              n_segments, kl, (float2*)coef, mask);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_greedy_prologue(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                                   int64_t n, float* kl, float* tparam, void* stream) {
  CWQ_REQUIRE(n > 0 && n < 2147483000LL, "n out of range");
  k_greedy_prologue<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_loc, q_scale, p_loc, p_scale, n, kl,
                                                                                  (float2*)tparam);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" size_t cwq_group_workspace_bytes(int64_t n, int32_t n_segments) {
  (void)n_segments;
  if (n <= 0) return 4096;
  size_t n_tiles = (size_t)((n + TILE1 - 1) / TILE1), n_super = (size_t)((n + TILE2 - 1) / TILE2);
  size_t b = 0;
  b += align_up((size_t)n * 4) * 2;          // next, exit1
  b += align_up((size_t)n);                  // mark
  b += align_up(n_tiles * 4);                // entry1
  b += align_up(n_super * 4);                // entry2
  b += align_up(n_super * (size_t)TILE2 * 4);  // exit2 (worst case: every node of a super-tile is an entry)
  b += align_up((size_t)scan_tiles(n) * 4);  // scan tile sums
  return b + 4096;                           // scalars + slack
}

// h_segments is a HOST copy of the same table: the entry tables are sized from its size limits.
extern "C" int cwq_group(const float* kl, int64_t n, const cwq_segment* segments, const cwq_segment* h_segments,
                         int32_t n_segments, int64_t* starts, uint32_t* group_seed, int32_t* gid, int64_t* seg_group_off,
                         void* workspace, size_t workspace_bytes, void* stream) {
  CWQ_REQUIRE(n > 0 && n < 2147483000LL, "n out of range");
  CWQ_REQUIRE(n_segments > 0 && n_segments <= 1 << 20 && segments && h_segments, "bad segment table");
  cudaStream_t st = (cudaStream_t)stream;
  const cwq_segment* h_segs = h_segments;
  bool ok = h_segs[0].start == 0 && h_segs[n_segments - 1].end == n;
  for (int s = 0; ok && s < n_segments; ++s) {
    ok = h_segs[s].end > h_segs[s].start && h_segs[s].size_limit >= 0 && (s == 0 || h_segs[s].start == h_segs[s - 1].end);
  }
  CWQ_REQUIRE(ok, "segments must be non-empty, contiguous, ascending and cover [0, n)");
  long long span = 0;
  max_entry_span(h_segs, n_segments, &span);

  const long long n_tiles = (n + TILE1 - 1) / TILE1, n_super = (n + TILE2 - 1) / TILE2;
  const int E2 = (int)(span < TILE2 ? span : TILE2);
  Workspace ws(workspace, workspace_bytes);
  int* next = ws.take<int>((size_t)n);
  int* exit1 = ws.take<int>((size_t)n);
  uint8_t* mark = ws.take<uint8_t>((size_t)n);
  int* entry1 = ws.take<int>((size_t)n_tiles);
  int* entry2 = ws.take<int>((size_t)n_super);
  int* exit2 = ws.take<int>((size_t)n_super * (size_t)E2);
  uint32_t* tile_sums = ws.take<uint32_t>((size_t)scan_tiles(n));
  int* first = ws.take<int>(1);
  uint32_t* n_groups = ws.take<uint32_t>(1);
  if (!next || !exit1 || !mark || !entry1 || !entry2 || !exit2 || !tile_sums || !first || !n_groups) {
    set_error("cwq_group: workspace too small");
    return CWQ_ERR_WORKSPACE;
  }
  CWQ_CUDA(cudaMemsetAsync(entry1, 0xff, (size_t)n_tiles * 4, st));
  CWQ_CUDA(cudaMemsetAsync(entry2, 0xff, (size_t)n_super * 4, st));
  k_next<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(kl, n, segments, n_segments, next, first);
  CWQ_LAUNCH_CHECK();
  k_tile_exit<<<(unsigned)n_tiles, 256, 0, st>>>(next, n, exit1);
  CWQ_LAUNCH_CHECK();
  k_super_exit<<<(unsigned)((n_super * E2 + 255) / 256), 256, 0, st>>>(exit1, n, E2, n_super, exit2);
  CWQ_LAUNCH_CHECK();
  k_chain_top<<<1, 32, 0, st>>>(first, exit2, n, E2, entry2);
  CWQ_LAUNCH_CHECK();
  k_tile_entries<<<(unsigned)((n_super + 127) / 128), 128, 0, st>>>(entry2, exit1, n, n_super, entry1);
  CWQ_LAUNCH_CHECK();
  k_mark<<<(unsigned)n_tiles, 256, 0, st>>>(entry1, next, n, n_tiles, mark);
  CWQ_LAUNCH_CHECK();
  k_mark_segment_starts<<<(n_segments + 127) / 128, 128, 0, st>>>(segments, n_segments, mark);
  CWQ_LAUNCH_CHECK();
  MarkCount in{mark};
  GroupDeliver out{mark, segments, n_segments, (long long*)starts, gid, (long long*)seg_group_off};
  CWQ_CUDA((device_exclusive_scan<uint32_t, MarkCount, GroupDeliver>(in, out, n, tile_sums, n_groups, st)));
  k_group_finish<<<1, 32, 0, st>>>(n_groups, n, n_segments, (long long*)starts, (long long*)seg_group_off);
  CWQ_LAUNCH_CHECK();
  k_group_seed<<<1184, 256, 0, st>>>((const long long*)starts, n_groups, segments, n_segments,
                                     (const long long*)seg_group_off, group_seed);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_expand_groups(const int64_t* starts, int64_t n_groups, int64_t n, const cwq_segment* segments,
                                 int32_t n_segments, const int64_t* seg_group_off, int32_t* gid, uint32_t* group_seed,
                                 void* stream) {
  CWQ_REQUIRE(n > 0 && n_groups > 0 && n_segments > 0, "bad sizes");
  k_expand_groups<<<(unsigned)((n_groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)starts, n_groups, n, segments, n_segments, (const long long*)seg_group_off, gid, group_seed);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_group_totals(const float* kl, const int64_t* starts, int64_t n_groups, const int64_t* n_groups_dev,
                                float* total_kl, int32_t* K, void* stream) {
  CWQ_REQUIRE(n_groups > 0, "n_groups must be positive");
  k_group_totals<<<(unsigned)((n_groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kl, (const long long*)starts, n_groups,
                                                                                      (const long long*)n_groups_dev, total_kl, K);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}
