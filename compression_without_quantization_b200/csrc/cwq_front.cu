// cwq_front.cu -- the O(N) stages of the grouped importance-sampler encode (CIS:133-204, 262-267, 258-260) as SIX
// kernels instead of thirty launches: per-dim prologue, KL-budget grouping, candidate counts, unit offsets before the
// scoring kernel; regeneration of the chosen latents, outliers and the Elias-delta bitstream after it.
//
// Grouping.  The reference packs dims into groups with a sequential fp32 accumulator (CIS:178-204); from a group that
// starts at dim i the next boundary next(i) is a function of kl[i .. i + L] only (L = size limit), so the groups are the
// chain  first -> next(first) -> ...  and a hop is at most L + 1 dims long.  Dims are cut into tiles of T:
//   k_front_tiles  (one CTA per tile)  prologue of the tile's dims (+ a halo of L + 1), next() of every dim from shared
//                  memory, then for each of the E = L + 1 possible ENTRY offsets of the chain into the tile one thread
//                  follows the chain through the tile: exit offset into the next tile, number of nodes, and the visited
//                  dims as a bit row.  Nothing here depends on other tiles.
//   k_front_chain  (one CTA)  composes the per-tile entry -> exit maps (two levels, shared memory) to the true entry of
//                  every tile, prefix-sums node counts and outlier counts over tiles, per-segment group offsets.
//   k_front_groups (one CTA per tile)  reads the bit row of the tile's true entry: group starts, group of every dim,
//                  seeds, sequential fp32 KL totals, candidate counts K, and the exclusive prefix of scoring units over
//                  all groups (decoupled look-back across tiles).
// The result is bit-identical to cwq_is_prologue + cwq_group + cwq_group_totals (cwq_group.cu), which stay as the general
// path (greedy sampler: L up to 4095) and as the cross-check in tests/test_gpu_parity.py.
#include "cwq_common.cuh"
#include "cwq_spec.cuh"
#include "cwq_front.cuh"
#include <cstdlib>

namespace cwq {

using ull = unsigned long long;

__device__ __forceinline__ int find_segment_between(const cwq_segment* __restrict__ segs, int lo, int hi, long long i) {
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs[mid].start <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Segment of dim x.  tile_segment (may be null): segment of every dim 1024 k, built by the host next to the segment table;
// with it the lookup is one load plus a short advance instead of a binary search of dependent loads at kernel start.
__device__ __forceinline__ int segment_of(const cwq_segment* __restrict__ segs, int n_seg, const int* __restrict__ tile_segment,
                                          long long x) {
  if (!tile_segment) return find_segment(segs, n_seg, x);
  int s = tile_segment[x >> 10];
  while (s + 1 < n_seg && segs[s + 1].start <= x) ++s;
  return s;
}

// The budget test of the reference compares the float32 running sum, promoted to double, with the float64 budget
// n_bits * ln 2 - 1 (CIS:181,188 / CGS:226,233).  (double)x > b  <=>  x >= [smallest float32 above b], and
// (double)x >= b  <=>  x >= [smallest float32 at or above b]: one float32 compare per step, same decisions (NaN: never over).
__device__ __forceinline__ float kl_threshold(double budget, bool strict) {
  float t = __double2float_ru(budget);
  if (strict && (double)t == budget) {
    uint32_t u = __float_as_uint(t);
    if ((u << 1) == 0u) u = 1u;                       // +-0 -> smallest positive subnormal
    else u = (u >> 31) ? u - 1u : u + 1u;
    t = __uint_as_float(u);
  }
  return t;
}

struct SegLocal {      // a segment seen from a tile: positions relative to the tile's first dim
  int start, end;      // clipped to int range (n < 2^31)
  int limit;
  float thr;
};
__device__ __forceinline__ SegLocal seg_local(const cwq_segment& sg, long long tile0) {
  SegLocal r;
  const long long a = sg.start - tile0, b = sg.end - tile0;
  r.start = a < -2147483000LL ? -2147483000 : (int)a;
  r.end = b > 2147483000LL ? 2147483000 : (int)b;
  r.limit = sg.size_limit > 2147483000LL ? 2147483000 : (int)sg.size_limit;
  r.thr = kl_threshold(sg.n_nats_per_group, sg.strict != 0);
  return r;
}

// the reference's loop body from state (size, acc) at local dim j0 until the next boundary (cwq_group.cu:walk restated on
// the tile's shared copy of kl, 32-bit indices, float32 threshold)
__device__ __forceinline__ int walk_tile(const float* __restrict__ kl_s, int j0, int size, float acc, const SegLocal& sg) {
  if (j0 >= sg.end) return sg.end;
  // the size test fires at j0 + (limit - size), the end-of-segment test at end - 1: whichever comes first closes the walk;
  // before it only the budget test can
  const int room = sg.limit > size ? sg.limit - size : 0;
  const int jstop = j0 + room < sg.end - 1 ? j0 + room : sg.end - 1;
  for (int j = j0; j < jstop; ++j) {
    acc = __fadd_rn(acc, kl_s[j]);
    if (acc >= sg.thr) return j;
  }
  return jstop;
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 1: tiles
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FRONT_THREADS) k_front_tiles(
    const float* __restrict__ q_loc, const float* __restrict__ q_scale, const float* __restrict__ p_loc,
    const float* __restrict__ p_scale, long long n, const cwq_segment* __restrict__ segs, int n_seg, int T, int E,
    float* __restrict__ kl, float2* __restrict__ coef, uint8_t* __restrict__ mask, uint8_t* __restrict__ hop,
    int* __restrict__ seg_first, uint8_t* __restrict__ maps, uint16_t* __restrict__ counts, uint32_t* __restrict__ rows,
    uint32_t* __restrict__ tile_out, const int* __restrict__ tile_segment) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = T >> 5;
  float* kl_s = reinterpret_cast<float*>(smem_raw);                                   // [T + E]
  uint32_t* rows_s = reinterpret_cast<uint32_t*>(smem_raw + (size_t)(T + E) * 4);     // [E * W]
  uint8_t* hop_s = smem_raw + (size_t)(T + E) * 4 + (size_t)E * W * 4;                // [T]
  __shared__ uint32_t s_out;
  const int tid = threadIdx.x;
  const long long t = blockIdx.x;
  const long long tile0 = t * T;
  const int lim = (int)(tile0 + T < n ? T : n - tile0);                 // dims of the tile
  const int hlim = (int)(tile0 + T + E < n ? T + E : n - tile0);        // ... and of the halo behind it
  // every thread looks the two segments up itself (same addresses: broadcast loads), so nothing waits on a barrier for it
  const int slo = segment_of(segs, n_seg, tile_segment, tile0), shi = segment_of(segs, n_seg, tile_segment, tile0 + hlim - 1);
  if (tid == 0) s_out = 0;
  for (int j = tid; j < E * W; j += FRONT_THREADS) rows_s[j] = 0u;
  const bool one_seg = slo == shi;                // the common case: tile and halo inside one segment
  const cwq_segment seg0 = segs[slo];
  const SegLocal sl0 = seg_local(seg0, tile0);

  // ---- prologue of the tile's dims and of the halo the boundary walks look at ----
  uint32_t my_out = 0;
  for (int li = tid; li < hlim; li += FRONT_THREADS) {
    const long long i = tile0 + li;
    float limit = seg0.dim_kl_bit_limit;
    bool first = li == sl0.start;
    if (!one_seg) {
      const int s = find_segment_between(segs, slo, shi, i);
      limit = segs[s].dim_kl_bit_limit;
      first = i == segs[s].start;
    }
    const PrologueDim r = is_prologue_dim(q_loc[i], q_scale[i], p_loc[i], p_scale[i], limit);
    kl_s[li] = r.kl;
    if (li < T) {
      kl[i] = r.kl;
      coef[i] = r.coef;
      mask[i] = r.m | (first ? 4 : 0);
      my_out += (r.m >> 1) & 1u;
    }
  }
  __syncthreads();

  // ---- next boundary from every dim of the tile ----
  for (int li = tid; li < lim; li += FRONT_THREADS) {
    int r;
    if (one_seg) {
      if (li == sl0.end - 1) r = (int)(n - tile0);        // last dim of the last segment (one_seg: nothing follows in the halo)
      else r = walk_tile(kl_s, li + 1, 1, kl_s[li], sl0);
      if (li == sl0.start) seg_first[slo] = (int)(tile0 + walk_tile(kl_s, li, 0, 0.0f, sl0));
    } else {
      const long long i = tile0 + li;
      const int s = find_segment_between(segs, slo, shi, i);
      const cwq_segment sg = segs[s];
      const SegLocal sl = seg_local(sg, tile0);
      if (li == sl.end - 1) {
        if (s + 1 < n_seg) {
          const SegLocal nx = seg_local(segs[s + 1], tile0);
          r = walk_tile(kl_s, nx.start, 0, 0.0f, nx);
        } else {
          r = (int)(n - tile0);
        }
      } else {
        r = walk_tile(kl_s, li + 1, 1, kl_s[li], sl);
      }
      if (li == sl.start) seg_first[s] = (int)(tile0 + walk_tile(kl_s, li, 0, 0.0f, sl));   // end of the segment's first group
    }
    const uint8_t h = (uint8_t)(r - li);
    hop_s[li] = h;
    hop[tile0 + li] = h;
  }
  __syncthreads();

  // ---- the chain through the tile from each possible entry offset ----
  if (tid < E) {
    int v = tid;
    int cnt = 0;
    uint32_t* row = rows_s + (size_t)tid * W;
    int cur_w = -1;
    uint32_t acc = 0u;
    while (v < lim) {
      const int w = v >> 5;
      if (w != cur_w) {
        if (cur_w >= 0) row[cur_w] = acc;
        cur_w = w;
        acc = 0u;
      }
      acc |= 1u << (v & 31);
      ++cnt;
      v += hop_s[v];
    }
    if (cur_w >= 0) row[cur_w] = acc;
    maps[t * E + tid] = (uint8_t)(v >= T ? v - T : 0);
    counts[t * E + tid] = (uint16_t)cnt;
  }
  // outliers of the tile
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_out += __shfl_xor_sync(0xffffffffu, my_out, o);
  if ((tid & 31) == 0 && my_out) atomicAdd(&s_out, my_out);
  __syncthreads();
  for (int j = tid; j < E * W; j += FRONT_THREADS) rows[(size_t)t * E * W + j] = rows_s[j];
  if (tid == 0) tile_out[t] = s_out;
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 2: chain through the tiles
// ------------------------------------------------------------------------------------------------------------------
constexpr int CHAIN_THREADS = 1024;
constexpr int CHAIN_BLOCK = 32;     // tiles per composed block

template <typename T>
__device__ __forceinline__ T chain_block_scan(T v, T* sm /*[33]*/, T& total) {
  return block_exclusive_scan<T>(v, sm, total);
}

__global__ void __launch_bounds__(CHAIN_THREADS) k_front_chain(
    long long n, const cwq_segment* __restrict__ segs, int n_seg, int T, int E, long long n_tiles,
    const int* __restrict__ seg_first, const uint8_t* __restrict__ maps, const uint16_t* __restrict__ counts,
    const uint32_t* __restrict__ rows, const uint32_t* __restrict__ tile_out, uint8_t* __restrict__ entry,
    uint32_t* __restrict__ base, uint32_t* __restrict__ tile_out_off, long long* __restrict__ starts,
    long long* __restrict__ seg_group_off, long long* __restrict__ seg_outlier_off, ull* __restrict__ status,
    ull* __restrict__ scalars) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* maps_s = smem_raw;                                        // [n_tiles * E]
  const long long nb = (n_tiles + CHAIN_BLOCK - 1) / CHAIN_BLOCK;
  uint8_t* comp_s = smem_raw + (size_t)n_tiles * E;                  // [nb * E]  composed maps of blocks of tiles
  uint8_t* bentry_s = comp_s + (size_t)nb * E;                       // [nb]      entry of each block
  uint8_t* entry_s = bentry_s + nb;                                  // [n_tiles]
  __shared__ uint32_t sm_scan[33];
  const int tid = threadIdx.x;
  const int W = T >> 5;

  for (long long j = tid; j < n_tiles * E; j += CHAIN_THREADS) maps_s[j] = maps[j];
  for (long long j = tid; j < (long long)TilePrefix<2>::words(n_tiles); j += CHAIN_THREADS) status[j] = 0ull;   // tile-prefix words of k_front_groups
  if (tid < 8) scalars[tid] = 0ull;                                                  // tickets / counters of later kernels
  __syncthreads();
  // level 1: entry -> exit of each block of CHAIN_BLOCK tiles
  for (long long task = tid; task < nb * E; task += CHAIN_THREADS) {
    const long long b = task / E;
    int x = (int)(task - b * E);
    const long long t1 = (b + 1) * CHAIN_BLOCK < n_tiles ? (b + 1) * CHAIN_BLOCK : n_tiles;
    for (long long tt = b * CHAIN_BLOCK; tt < t1; ++tt) x = maps_s[tt * E + x];
    comp_s[task] = (uint8_t)x;
  }
  __syncthreads();
  // level 2: the entry of every block, serially (<= 256 blocks for 8192 tiles)
  if (tid == 0) {
    int x = seg_first[0];          // the first chain node: first boundary of segment 0 (walk from state (0, 0))
    for (long long b = 0; b < nb; ++b) {
      bentry_s[b] = (uint8_t)x;
      x = comp_s[b * E + x];
    }
  }
  __syncthreads();
  for (long long b = tid; b < nb; b += CHAIN_THREADS) {
    int x = bentry_s[b];
    const long long t1 = (b + 1) * CHAIN_BLOCK < n_tiles ? (b + 1) * CHAIN_BLOCK : n_tiles;
    for (long long tt = b * CHAIN_BLOCK; tt < t1; ++tt) {
      entry_s[tt] = (uint8_t)x;
      x = maps_s[tt * E + x];
    }
  }
  __syncthreads();

  // groups before each tile: chain nodes on the true path + segment starts; outliers before each tile
  const long long per = (n_tiles + CHAIN_THREADS - 1) / CHAIN_THREADS;
  const long long a = (long long)tid * per, bnd = a + per < n_tiles ? a + per : n_tiles;
  uint32_t my_groups = 0, my_out = 0;
  for (long long tt = a; tt < bnd; ++tt) {
    my_groups += counts[tt * E + entry_s[tt]];
    my_out += tile_out[tt];
  }
  // segment starts of my tiles: segments are sorted, so they form a contiguous run
  if (a < bnd) {
    const long long d0 = a * T, d1 = bnd * T < n ? bnd * T : n;
    int s = find_segment(segs, n_seg, d0);
    if (segs[s].start < d0) ++s;
    for (; s < n_seg && segs[s].start < d1; ++s) ++my_groups;
  }
  uint32_t total_groups, total_out;
  uint32_t gpre = chain_block_scan<uint32_t>(my_groups, sm_scan, total_groups);
  uint32_t opre = chain_block_scan<uint32_t>(my_out, sm_scan, total_out);
  {
    int s = 0;
    if (a < bnd) {
      s = find_segment(segs, n_seg, a * T);
      if (segs[s].start < a * T) ++s;
    }
    for (long long tt = a; tt < bnd; ++tt) {
      entry[tt] = entry_s[tt];
      base[tt] = gpre;
      tile_out_off[tt] = opre;
      gpre += counts[tt * E + entry_s[tt]];
      opre += tile_out[tt];
      const long long d1 = (tt + 1) * T;
      for (; s < n_seg && segs[s].start < d1; ++s) ++gpre;
    }
  }
  if (tid == 0) {
    starts[total_groups] = n;
    seg_group_off[n_seg] = (long long)total_groups;
    seg_outlier_off[n_seg] = (long long)total_out;
  }
  __syncthreads();
  // first group of every segment: groups before its tile + chain nodes of the tile before its start + earlier segment
  // starts in the same tile (a segment's forced first group precedes a chain node at the same dim)
  for (int s = tid; s < n_seg; s += CHAIN_THREADS) {
    const long long pos = segs[s].start;
    const long long tt = pos / T;
    const int li = (int)(pos - tt * T);
    const uint32_t* row = rows + ((size_t)tt * E + entry_s[tt]) * W;
    uint32_t c = 0;
    for (int w = 0; w < (li >> 5); ++w) c += __popc(row[w]);
    if (li & 31) c += __popc(row[li >> 5] & ((1u << (li & 31)) - 1u));
    int k = 0;
    for (int sp = s - 1; sp >= 0 && segs[sp].start >= tt * T; --sp) ++k;
    seg_group_off[s] = (long long)base[tt] + c + k;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 3: groups of each tile
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ ull units_of(int K) { return (ull)((K + FRONT_UNIT_CAND - 1) / FRONT_UNIT_CAND); }
__device__ __forceinline__ ull work_of(int K, int dims) {
  const int blocks = row_blocks((dims + 3) >> 2);
  return (ull)K * (ull)(blocks > 0 ? blocks : 1);
}

__global__ void __launch_bounds__(FRONT_THREADS) k_front_groups(
    long long n, const cwq_segment* __restrict__ segs, int n_seg, int T, int E, long long n_tiles,
    const float* __restrict__ kl, const uint8_t* __restrict__ hop, const int* __restrict__ seg_first,
    const uint32_t* __restrict__ rows, const uint8_t* __restrict__ entry, const uint32_t* __restrict__ base,
    const long long* __restrict__ seg_group_off, long long* __restrict__ starts, int* __restrict__ gid,
    uint32_t* __restrict__ group_seed, float* __restrict__ total_kl, int* __restrict__ K, ull* __restrict__ keys,
    ull* __restrict__ unit_off, ull* __restrict__ work_off, ull* __restrict__ prefix_words, const int* __restrict__ tile_segment) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = T >> 5;
  float* kl_s = reinterpret_cast<float*>(smem_raw);                        // [T + E] the tile's kl and the halo behind it
  uint32_t* node_w = reinterpret_cast<uint32_t*>(kl_s + T + E);            // [W] chain nodes of the tile
  uint32_t* seg_w = node_w + W;                                            // [W] segment starts of the tile
  uint32_t* node_pre = seg_w + W;                                          // [W] nodes before each word
  uint32_t* seg_pre = node_pre + W;                                        // [W]
  uint8_t* hop_s = reinterpret_cast<uint8_t*>(seg_pre + W);                // [T]
  uint16_t* slot_s = reinterpret_cast<uint16_t*>(hop_s + T);               // [2 T] the tile's groups in order: start dim | 0x8000 for a chain node
  __shared__ ull sm_scan[2 * 33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long t = blockIdx.x;
  const long long tile0 = t * T;
  const int lim = (int)(tile0 + T < n ? T : n - tile0);
  const int hlim = (int)(tile0 + T + E < n ? T + E : n - tile0);
  const uint32_t B = base[t];
  const uint32_t* __restrict__ row = rows + ((size_t)t * E + entry[t]) * W;
  for (int w = tid; w < W; w += FRONT_THREADS) { node_w[w] = row[w]; seg_w[w] = 0u; }
  for (int j = tid; j < hlim; j += FRONT_THREADS) kl_s[j] = kl[tile0 + j];
  for (int j = tid; j < lim; j += FRONT_THREADS) hop_s[j] = hop[tile0 + j];
  const int slo = segment_of(segs, n_seg, tile_segment, tile0), shi = segment_of(segs, n_seg, tile_segment, tile0 + lim - 1);
  __syncthreads();
  if (tid == 0) {
    for (int s = slo; s <= shi; ++s) {
      const long long p = segs[s].start;
      if (p >= tile0 && p < tile0 + lim) seg_w[(p - tile0) >> 5] |= 1u << ((p - tile0) & 31);
    }
  }
  __syncthreads();
  if (warp == 0) {   // exclusive prefix of the word popcounts, 32 words at a time
    uint32_t carry_n = 0, carry_s = 0;
    for (int w0 = 0; w0 < W; w0 += 32) {
      const int w = w0 + lane;
      uint32_t cn = w < W ? __popc(node_w[w]) : 0u, cs = w < W ? __popc(seg_w[w]) : 0u;
      uint32_t in = cn, is = cs;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, in, o), b = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) { in += a; is += b; }
      }
      if (w < W) { node_pre[w] = carry_n + in - cn; seg_pre[w] = carry_s + is - cs; }
      carry_n += __shfl_sync(0xffffffffu, in, 31);
      carry_s += __shfl_sync(0xffffffffu, is, 31);
    }
  }
  __syncthreads();

  // Pass 1, one thread per D consecutive dims (cheap, uniform): the group of every dim, and the tile's groups laid out in
  // order as a slot list -- slot = chain nodes + segment starts before the dim, which the popcount prefixes give without
  // atomics; a dim that is both (a segment whose first dim alone is over the budget: empty forced group, CIS:183-204)
  // takes two slots, the forced group first.
  const int D = T / FRONT_THREADS;
  const int li0 = tid * D;
  const bool one_seg = slo == shi;
  if (li0 < lim) {
    const int w0 = li0 >> 5, b0 = li0 & 31;
    const uint32_t below = (1u << b0) - 1u;
    uint32_t nodes = node_pre[w0] + __popc(node_w[w0] & below);     // chain nodes / segment starts strictly before li0
    uint32_t sstarts = seg_pre[w0] + __popc(seg_w[w0] & below);
    for (int r = 0; r < D && li0 + r < lim; ++r) {
      const int li = li0 + r;
      const uint32_t is_node = (node_w[li >> 5] >> (li & 31)) & 1u, is_seg = (seg_w[li >> 5] >> (li & 31)) & 1u;
      if (is_seg) slot_s[nodes + sstarts] = (uint16_t)li;
      if (is_node) slot_s[nodes + sstarts + is_seg] = (uint16_t)(li | 0x8000);
      nodes += is_node;
      sstarts += is_seg;
      gid[tile0 + li] = (int)(B + nodes + sstarts) - 1;
    }
  }
  __syncthreads();
  // Pass 2, one thread per GROUP (the sequential fp32 total, the candidate count and five stores are ~10x the work of a
  // dim, and only one dim in ~8-12 starts a group: done per dim it ran with most lanes idle): rounds of FRONT_THREADS slots,
  // the running unit / work prefix carried from round to round.
  const int n_slots = (int)(node_pre[W - 1] + __popc(node_w[W - 1]) + seg_pre[W - 1] + __popc(seg_w[W - 1]));
  ull tot_u = 0, tot_w = 0;
  for (int base_slot = 0; base_slot < n_slots; base_slot += FRONT_THREADS) {
    const int slot = base_slot + tid;
    ull my_units = 0, my_work = 0;
    long long g = -1;
    if (slot < n_slots) {
      const uint32_t e = slot_s[slot];
      const int li = (int)(e & 0x7fffu);
      const bool forced = (e & 0x8000u) == 0u;          // the segment's forced first group; else a chain node's group
      const int s = one_seg ? slo : find_segment_between(segs, slo, shi, tile0 + li);
      const cwq_segment& sg = segs[s];
      const long long seg_end_l = sg.end - tile0;
      g = (long long)B + slot;
      // a node's group ends at the next node -- or at the end of its segment (the last dim of a segment hops into the next one)
      long long end_l = forced ? (long long)seg_first[s] - tile0 : (long long)li + hop_s[li];
      if (end_l > seg_end_l) end_l = seg_end_l;
      float acc = 0.0f;
      for (int k = li; k < (int)end_l; ++k) acc = __fadd_rn(acc, kl_s[k]);
      const int Kg = candidates_from_kl(acc);
      starts[g] = tile0 + li;
      total_kl[g] = acc;
      K[g] = Kg;
      group_seed[g] = sg.seed + (uint32_t)(g - seg_group_off[s]);
      keys[g] = KEY_INIT;
      my_units = units_of(Kg);
      my_work = work_of(Kg, (int)end_l - li);
    }
    ull ru, rw;
    const ull upre = block_exclusive_scan<ull>(my_units, sm_scan, ru);
    const ull wpre = block_exclusive_scan<ull>(my_work, sm_scan, rw);
    // prefixes relative to the tile; the tile's totals are published for k_front_prefix, which completes them
    if (g >= 0) {
      unit_off[g] = tot_u + upre;
      work_off[g] = tot_w + wpre;
    }
    tot_u += ru;
    tot_w += rw;
  }
  if (tid == 0) {
    const TilePrefix<2> tp = TilePrefix<2>::carve(prefix_words, n_tiles);
    const ull agg[2] = {tot_u, tot_w};
    tp.publish(t, agg);
  }
}

// kernel 3b: add to every group the units / work of the tiles before its own (all tiles have published: nothing waits)
constexpr int PREFIX_THREADS = 64;
__global__ void __launch_bounds__(PREFIX_THREADS) k_front_prefix(long long n_tiles, const uint32_t* __restrict__ base,
                                                                 const long long* __restrict__ ng_dev, ull* __restrict__ unit_off,
                                                                 ull* __restrict__ work_off, ull* __restrict__ prefix_words) {
  pdl_wait();
  pdl_trigger();
  __shared__ ull sm[2 * 33];
  const long long t = blockIdx.x;
  const TilePrefix<2> tp = TilePrefix<2>::carve(prefix_words, n_tiles);
  ull excl[2];
  tp.exclusive(t, sm, excl);
  const long long ng = *ng_dev;
  const long long g0 = base[t], g1 = t + 1 < n_tiles ? (long long)base[t + 1] : ng;
  for (long long g = g0 + threadIdx.x; g < g1; g += PREFIX_THREADS) {
    unit_off[g] += excl[0];
    work_off[g] += excl[1];
  }
  if (t == n_tiles - 1 && threadIdx.x == 0) {
    unit_off[ng] = excl[0] + (ld_volatile_u64(&tp.tile_words[t]) & TP_VAL);
    work_off[ng] = excl[1] + (ld_volatile_u64(&tp.tile_words[n_tiles + t]) & TP_VAL);
  }
}

// Elias-delta code of index + 1 (BIO:7-20)
__device__ __forceinline__ void elias_delta_bits(uint32_t x, ull& pattern, int& len) {
  const int nb = 31 - __clz((int)x);
  const int l = 31 - __clz(nb + 1);
  pattern = ((ull)(nb + 1) << nb) | (ull)(x & ((nb == 0) ? 0u : (0xffffffffu >> (32 - nb))));
  len = 2 * l + nb + 1;
}

// OR `len` bits (right-aligned in pattern) into an MSB-first BYTE stream at bit `pos`, through 32-bit words
__device__ __forceinline__ void put_bits_bytes(uint32_t* words, ull pos, ull pattern, int len) {
  ull w = pos >> 5;
  int o = (int)(pos & 31);
  int remaining = len;
  while (remaining > 0) {
    const int take = 32 - o < remaining ? 32 - o : remaining;
    const uint32_t chunk = (uint32_t)((pattern >> (remaining - take)) & ((1ull << take) - 1ull));
    atomicOr(&words[w], __byte_perm(chunk << (32 - o - take), 0u, 0x0123));
    remaining -= take;
    ++w;
    o = 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// after the scoring kernel: regenerate the chosen rows, un-rescale, outliers (CIS:104-109, 148-158, 262-267)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FRONT_THREADS) k_finish_dims(
    long long n, const cwq_segment* __restrict__ segs, int n_seg, int T, const float* __restrict__ q_loc,
    const float* __restrict__ q_scale, const float* __restrict__ p_loc, const float* __restrict__ p_scale,
    const uint8_t* __restrict__ mask, const long long* __restrict__ starts, const uint32_t* __restrict__ gseed,
    const int* __restrict__ gid, const uint32_t* __restrict__ index, const uint32_t* __restrict__ tile_out_off,
    float* __restrict__ sample, long long* __restrict__ outlier_idx, uint16_t* __restrict__ outlier_q,
    long long* __restrict__ seg_outlier_off, const uint32_t* __restrict__ base, const long long* __restrict__ ng_dev,
    long long n_tiles, ull* __restrict__ pack_words, const int* __restrict__ tile_segment) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t sm_scan[33];
  __shared__ ull sm_bits[33];
  const int tid = threadIdx.x;
  const long long t = blockIdx.x;
  {
    // code bits of the groups that START in this tile, published for the packing kernel that follows on the stream
    const long long g0 = base[t], g1 = t + 1 < n_tiles ? (long long)base[t + 1] : *ng_dev;
    ull bits = 0;
    for (long long g = g0 + tid; g < g1; g += FRONT_THREADS) {
      ull pat; int len;
      elias_delta_bits(index[g] + 1u, pat, len);
      bits += (ull)len;
    }
    ull tot_bits;
    block_exclusive_scan<ull>(bits, sm_bits, tot_bits);
    if (tid == 0) {
      const TilePrefix<1> tp = TilePrefix<1>::carve(pack_words, n_tiles);
      const ull agg[1] = {tot_bits};
      tp.publish(t, agg);
    }
  }
  const long long tile0 = t * T;
  const long long tile_end = tile0 + T < n ? tile0 + T : n;
  const int D = T / FRONT_THREADS;
  const long long i0 = tile0 + (long long)tid * D;
  // D consecutive dims per thread, in chunks of 4 (vector loads / stores); the Philox block of a quad is computed once for
  // the dims of the thread that share it
  const bool aligned = (((uintptr_t)p_loc | (uintptr_t)p_scale | (uintptr_t)sample | (uintptr_t)gid) & 15u) == 0 &&
                       ((uintptr_t)mask & 3u) == 0;     // caller-owned buffers: vector access only when it is legal
  uint32_t my_out = 0;
  for (int c = 0; c < D; c += 4) {
    const long long i = i0 + c;
    if (aligned && i + 4 <= tile_end) {
      const uint32_t m4 = *reinterpret_cast<const uint32_t*>(mask + i);
      my_out += __popc(m4 & 0x02020202u);
    } else {
      for (int r = 0; r < 4; ++r)
        if (i + r < tile_end) my_out += (mask[i + r] >> 1) & 1u;
    }
  }
  uint32_t tot;
  uint32_t pre = block_exclusive_scan<uint32_t>(my_out, sm_scan, tot) + tile_out_off[t];
  int last_g = -1;
  uint32_t last_q = 0;
  float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int c = 0; c < D; c += 4) {
    const long long ic = i0 + c;
    if (ic >= tile_end) break;
    const bool vec = aligned && ic + 4 <= tile_end;
    uint32_t m4 = 0;
    int4 gv = make_int4(0, 0, 0, 0);
    float4 plv = make_float4(0.f, 0.f, 0.f, 0.f), psv = plv;
    if (vec) {
      m4 = *reinterpret_cast<const uint32_t*>(mask + ic);
      gv = *reinterpret_cast<const int4*>(gid + ic);
      plv = *reinterpret_cast<const float4*>(p_loc + ic);
      psv = *reinterpret_cast<const float4*>(p_scale + ic);
    }
    float out4[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long i = ic + r;
      if (!vec && i >= tile_end) { out4[r] = 0.0f; continue; }
      const uint8_t mm = vec ? (uint8_t)(m4 >> (8 * r)) : mask[i];
      int s = -1;
      if (mm & 6) s = segment_of(segs, n_seg, tile_segment, i);
      if (mm & 4) seg_outlier_off[s] = (long long)pre;       // outliers before this segment
      if (mm & 2) {
        const long long local = i - segs[s].start;
        float zo[4];
        normal_quad(segs[s].seed + OUTLIER_STREAM_OFFSET, (uint32_t)local, 0u, zo);
        const float v = __fadd_rn(__fmul_rn(q_scale[i], zo[0]), q_loc[i]);
        const uint16_t qv = quantize_u16(v);
        out4[r] = dequantize_u16(qv);
        outlier_idx[pre] = local;
        outlier_q[pre] = qv;
        ++pre;
      } else {
        const int g = vec ? (r == 0 ? gv.x : r == 1 ? gv.y : r == 2 ? gv.z : gv.w) : gid[i];
        const uint32_t d = (uint32_t)(i - starts[g]);
        if (g != last_g || (d >> 2) != last_q) {
          normal_quad(gseed[g], index[g], d >> 2, z);
          last_g = g;
          last_q = d >> 2;
        }
        const float zz = (d & 2) ? ((d & 1) ? z[3] : z[2]) : ((d & 1) ? z[1] : z[0]);
        const float ps = vec ? (r == 0 ? psv.x : r == 1 ? psv.y : r == 2 ? psv.z : psv.w) : p_scale[i];
        const float pl = vec ? (r == 0 ? plv.x : r == 1 ? plv.y : r == 2 ? plv.z : plv.w) : p_loc[i];
        out4[r] = __fadd_rn(__fmul_rn(ps, zz), pl);
      }
    }
    if (vec) {
      *reinterpret_cast<float4*>(sample + ic) = make_float4(out4[0], out4[1], out4[2], out4[3]);
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (ic + r < tile_end) sample[ic + r] = out4[r];
    }
  }
}

// Elias-delta bitstream (CIS:258-260): every dim tile packs the codes of the groups that start in it; bit offsets from the
// tile aggregates k_finish_dims published (nothing waits), atomicOr of byte-swapped words into the zeroed stream
__global__ void __launch_bounds__(FRONT_THREADS) k_pack_elias(const uint32_t* __restrict__ index, const uint32_t* __restrict__ base,
                                                              const long long* __restrict__ ng_dev, long long n_tiles,
                                                              uint32_t* __restrict__ words, long long* __restrict__ bit_off,
                                                              ull* __restrict__ pack_words, long long cap) {
  pdl_wait();
  pdl_trigger();
  __shared__ ull sm_scan[33];
  const int tid = threadIdx.x;
  const long long t = blockIdx.x;
  const TilePrefix<1> tp = TilePrefix<1>::carve(pack_words, n_tiles);
  ull excl[1];
  tp.exclusive(t, sm_scan, excl);
  ull running = excl[0];
  const long long ng = *ng_dev;
  const long long g0 = base[t], g1 = t + 1 < n_tiles ? (long long)base[t + 1] : ng;
  for (long long c = g0; c < g1; c += FRONT_THREADS) {
    const long long g = c + tid;
    ull pat = 0;
    int len = 0;
    if (g < g1) elias_delta_bits(index[g] + 1u, pat, len);
    ull tot;
    const ull pre = running + block_exclusive_scan<ull>((ull)len, sm_scan, tot);
    if (g < g1) {
      bit_off[g] = (long long)pre;
      put_bits_bytes(words, pre, pat, len);
    }
    running += tot;
  }
  if (t == n_tiles - 1 && tid == 0) {
    bit_off[cap] = (long long)running;      // the total also at a place the host knows before it knows the group count
    bit_off[ng] = (long long)running;
  }
}

}  // namespace cwq

using namespace cwq;

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
namespace cwq {

bool front_plan(long long n, long long max_group_size, FrontPlan* p) {
  if (n <= 0 || max_group_size < 0 || max_group_size > 128) return false;
  const int E = (int)max_group_size + 1;
  // 1024-dim tiles keep the per-tile chain walk short (latency of small inputs); above 2M dims the single-CTA chain
  // kernel and the per-tile fixed costs weigh more and 2048 is faster (measured: profiles/r02_front_tile_size.log)
  int t_min = n > (2LL << 20) ? 2048 : 1024;
  if (const char* e = getenv("CWQ_FRONT_TILE")) {      // tuning knob: smallest tile size tried
    const int v = atoi(e);
    if (v == 2048 || v == 4096 || v == 8192 || v == 16384) t_min = v;
  }
  for (int T = t_min; T <= 16384; T <<= 1) {
    const long long nt = (n + T - 1) / T;
    const long long nb = (nt + CHAIN_BLOCK - 1) / CHAIN_BLOCK;
    const size_t smem_chain = (size_t)nt * E + (size_t)nb * E + (size_t)nb + (size_t)nt + 64;
    const size_t smem_tiles = (size_t)(T + E) * 4 + (size_t)E * (T / 32) * 4 + (size_t)T + 64;
    if (smem_chain <= 200 * 1024 && smem_tiles <= 160 * 1024) {
      p->T = T; p->E = E; p->W = T / 32; p->n_tiles = nt;
      p->smem_tiles = smem_tiles; p->smem_chain = smem_chain; p->smem_groups = (size_t)(T + E) * 4 + (size_t)4 * (T / 32) * 4 + (size_t)T + (size_t)4 * T;
      return true;
    }
  }
  return false;
}

size_t front_workspace_bytes(long long n, int n_seg, long long cap, const FrontPlan& p) {
  size_t b = 0;
  b += align_up(((size_t)cap + 1) * 8) * 2;                 // unit_off, work_off
  b += align_up((size_t)cap * 8);                           // keys
  b += align_up((size_t)p.n_tiles * 4);                     // tile_out_off
  b += align_up(64);                                        // scalars
  b += align_up(TilePrefix<1>::words(p.n_tiles) * 8);       // tile-prefix words of the packing kernel
  b += align_up((size_t)n);                                 // hop
  b += align_up((size_t)n_seg * 4);                         // seg_first
  b += align_up((size_t)p.n_tiles * p.E);                   // maps
  b += align_up((size_t)p.n_tiles * p.E * 2);               // counts
  b += align_up((size_t)p.n_tiles * p.E * p.W * 4);         // rows
  b += align_up((size_t)p.n_tiles * 4) * 2;                 // tile_out, base
  b += align_up((size_t)p.n_tiles);                         // entry
  b += align_up(TilePrefix<2>::words(p.n_tiles) * 8);       // tile-prefix words of the group kernel (units, work)
  return b + 1024;
}

bool front_carve(void* workspace, size_t bytes, long long n, int n_seg, long long cap, const FrontPlan& p, FrontWorkspace* w) {
  Workspace ws(workspace, bytes);
  w->unit_off = ws.take<ull>((size_t)cap + 1);
  w->work_off = ws.take<ull>((size_t)cap + 1);
  w->keys = ws.take<ull>((size_t)cap);
  w->tile_out_off = ws.take<uint32_t>((size_t)p.n_tiles);
  w->scalars = ws.take<ull>(8);
  w->pack_status = ws.take<ull>(TilePrefix<1>::words(p.n_tiles));
  w->hop = ws.take<uint8_t>((size_t)n);
  w->seg_first = ws.take<int>((size_t)n_seg);
  w->maps = ws.take<uint8_t>((size_t)p.n_tiles * p.E);
  w->counts = ws.take<uint16_t>((size_t)p.n_tiles * p.E);
  w->rows = ws.take<uint32_t>((size_t)p.n_tiles * p.E * p.W);
  w->tile_out = ws.take<uint32_t>((size_t)p.n_tiles);
  w->base = ws.take<uint32_t>((size_t)p.n_tiles);
  w->entry = ws.take<uint8_t>((size_t)p.n_tiles);
  w->status = ws.take<ull>(TilePrefix<2>::words(p.n_tiles));
  return w->status != nullptr;
}

int front_group(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, long long n,
                const cwq_segment* segs, int n_seg, const FrontPlan& p, const FrontWorkspace& w, const cwq_is_buffers* b,
                cudaStream_t st) {
  if (p.smem_tiles > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_front_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_tiles));
  if (p.smem_chain > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_front_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_chain));
  if (p.smem_groups > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_front_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_groups));
  const int* ts = b->tile_segment;
  const long long* ng_dev = (const long long*)b->seg_group_off + n_seg;
  int slot = trace_begin(TR_FRONT_TILES, st);
  CWQ_CUDA(launch_chain(k_front_tiles, dim3((unsigned)p.n_tiles), dim3(FRONT_THREADS), p.smem_tiles, st, q_loc, q_scale, p_loc,
                        p_scale, (long long)n, segs, n_seg, p.T, p.E, b->kl, (float2*)b->coef, b->mask, w.hop, w.seg_first, w.maps,
                        w.counts, w.rows, w.tile_out, ts));
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  slot = trace_begin(TR_FRONT_CHAIN, st);
  CWQ_CUDA(launch_chain(k_front_chain, dim3(1), dim3(CHAIN_THREADS), p.smem_chain, st, (long long)n, segs, n_seg, p.T, p.E, p.n_tiles,
                        w.seg_first, w.maps, w.counts, w.rows, w.tile_out, w.entry, w.base, w.tile_out_off, (long long*)b->starts,
                        (long long*)b->seg_group_off, (long long*)b->seg_outlier_off, w.status, w.scalars));
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  slot = trace_begin(TR_FRONT_GROUPS, st);
  CWQ_CUDA(launch_chain(k_front_groups, dim3((unsigned)p.n_tiles), dim3(FRONT_THREADS), p.smem_groups, st, (long long)n, segs, n_seg,
                        p.T, p.E, p.n_tiles, b->kl, w.hop, w.seg_first, w.rows, w.entry, w.base, (const long long*)b->seg_group_off,
                        (long long*)b->starts, b->gid, b->group_seed, b->total_kl, b->K, w.keys, w.unit_off, w.work_off, w.status,
                        ts));
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  slot = trace_begin(TR_FRONT_PREFIX, st);
  CWQ_CUDA(launch_chain(k_front_prefix, dim3((unsigned)p.n_tiles), dim3(PREFIX_THREADS), 0, st, p.n_tiles, w.base, ng_dev, w.unit_off,
                        w.work_off, w.status));
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  return CWQ_OK;
}

// zeroes what the FINISH kernels OR / publish into (the bitstream, the packing kernel's tile words).  Nothing between the
// GROUP stage and FINISH touches them, so a call that runs all stages clears them first and keeps its kernel chain unbroken
// (a memset between two kernels ends a programmatic dependent launch chain).
int front_finish_clear(const FrontPlan& p, const FrontWorkspace& w, const cwq_is_buffers* b, cudaStream_t st) {
  CWQ_CUDA(cudaMemsetAsync(b->bits, 0, b->bits_capacity_bytes, st));
  CWQ_CUDA(cudaMemsetAsync(w.pack_status, 0, TilePrefix<1>::words(p.n_tiles) * 8, st));
  return CWQ_OK;
}

int front_finish(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, long long n,
                 const cwq_segment* segs, int n_seg, long long cap, const FrontPlan& p, const FrontWorkspace& w,
                 const cwq_is_buffers* b, cudaStream_t st) {
  const long long* ng_dev = (const long long*)b->seg_group_off + n_seg;
  int slot = trace_begin(TR_FINISH_DIMS, st);
  CWQ_CUDA(launch_chain(k_finish_dims, dim3((unsigned)p.n_tiles), dim3(FRONT_THREADS), 0, st, (long long)n, segs, n_seg, p.T, q_loc,
                        q_scale, p_loc, p_scale, b->mask, (const long long*)b->starts, b->group_seed, b->gid, b->index,
                        w.tile_out_off, b->sample, (long long*)b->outlier_idx, b->outlier_q, (long long*)b->seg_outlier_off, w.base,
                        ng_dev, p.n_tiles, w.pack_status, b->tile_segment));
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  slot = trace_begin(TR_PACK, st);
  CWQ_CUDA(launch_chain(k_pack_elias, dim3((unsigned)p.n_tiles), dim3(FRONT_THREADS), 0, st, b->index, w.base, ng_dev, p.n_tiles,
                        (uint32_t*)b->bits, (long long*)b->bit_off, w.pack_status, (long long)cap));
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  return CWQ_OK;
}

}  // namespace cwq
