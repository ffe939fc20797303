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

namespace cwq {

using ull = unsigned long long;

__device__ __forceinline__ int find_segment_between(const cwq_segment* __restrict__ segs, int lo, int hi, long long i) {
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs[mid].start <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// the reference's loop body from state (size, acc) at dim idx0 until the next boundary (same as cwq_group.cu:walk), with
// kl read from the tile's shared copy (kl_s[0] is dim tile0)
__device__ __forceinline__ long long walk_tile(const float* __restrict__ kl_s, long long tile0, long long idx0, long long size,
                                               float acc, const cwq_segment& sg) {
  const long long end = sg.end;
  const double budget = sg.n_nats_per_group;
  const long long limit = sg.size_limit;
  const bool strict = sg.strict != 0;
  for (long long idx = idx0; idx < end; ++idx) {
    const float cand = __fadd_rn(acc, kl_s[idx - tile0]);
    const bool over = strict ? ((double)cand > budget) : ((double)cand >= budget);
    if (size >= limit || over || idx == end - 1) return idx;
    acc = cand;
    size += 1;
  }
  return end;
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 1: tiles
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FRONT_THREADS) k_front_tiles(
    const float* __restrict__ q_loc, const float* __restrict__ q_scale, const float* __restrict__ p_loc,
    const float* __restrict__ p_scale, long long n, const cwq_segment* __restrict__ segs, int n_seg, int T, int E,
    float* __restrict__ kl, float2* __restrict__ coef, uint8_t* __restrict__ mask, uint8_t* __restrict__ hop,
    int* __restrict__ seg_first, uint8_t* __restrict__ maps, uint16_t* __restrict__ counts, uint32_t* __restrict__ rows,
    uint32_t* __restrict__ tile_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = T >> 5;
  float* kl_s = reinterpret_cast<float*>(smem_raw);                                   // [T + E]
  uint32_t* rows_s = reinterpret_cast<uint32_t*>(smem_raw + (size_t)(T + E) * 4);     // [E * W]
  uint8_t* hop_s = smem_raw + (size_t)(T + E) * 4 + (size_t)E * W * 4;                // [T]
  __shared__ int s_seg[2];
  __shared__ uint32_t s_out;
  const int tid = threadIdx.x;
  const long long t = blockIdx.x;
  const long long tile0 = t * T;
  const long long tile_end = tile0 + T < n ? tile0 + T : n;
  const long long halo_end = tile0 + T + E < n ? tile0 + T + E : n;
  if (tid == 0) {
    s_seg[0] = find_segment(segs, n_seg, tile0);
    s_seg[1] = find_segment(segs, n_seg, halo_end - 1);
    s_out = 0;
  }
  for (int j = tid; j < E * W; j += FRONT_THREADS) rows_s[j] = 0u;
  __syncthreads();
  const int slo = s_seg[0], shi = s_seg[1];

  // ---- prologue of the tile's dims and of the halo the boundary walks look at ----
  uint32_t my_out = 0;
  for (long long li = tid; li < halo_end - tile0; li += FRONT_THREADS) {
    const long long i = tile0 + li;
    const int s = find_segment_between(segs, slo, shi, i);
    const PrologueDim r = is_prologue_dim(q_loc[i], q_scale[i], p_loc[i], p_scale[i], segs[s].dim_kl_bit_limit);
    kl_s[li] = r.kl;
    if (li < T) {
      kl[i] = r.kl;
      coef[i] = r.coef;
      mask[i] = r.m | (i == segs[s].start ? 4 : 0);
      my_out += (r.m >> 1) & 1u;
    }
  }
  __syncthreads();

  // ---- next boundary from every dim of the tile ----
  for (long long li = tid; li < tile_end - tile0; li += FRONT_THREADS) {
    const long long i = tile0 + li;
    const int s = find_segment_between(segs, slo, shi, i);
    const cwq_segment sg = segs[s];
    long long r;
    if (i == sg.end - 1) r = (s + 1 < n_seg) ? walk_tile(kl_s, tile0, segs[s + 1].start, 0, 0.0f, segs[s + 1]) : n;
    else r = walk_tile(kl_s, tile0, i + 1, 1, kl_s[li], sg);
    const uint8_t h = (uint8_t)(r - i);
    hop_s[li] = h;
    hop[i] = h;
    if (i == sg.start) seg_first[s] = (int)walk_tile(kl_s, tile0, i, 0, 0.0f, sg);   // end of the segment's first group
  }
  __syncthreads();

  // ---- the chain through the tile from each possible entry offset ----
  if (tid < E) {
    const long long lim = tile_end - tile0;
    long long v = tid;
    int cnt = 0;
    uint32_t* row = rows_s + (size_t)tid * W;
    int cur_w = -1;
    uint32_t acc = 0u;
    while (v < lim) {
      const int w = (int)(v >> 5);
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
  for (long long j = tid; j < 2 * n_tiles; j += CHAIN_THREADS) status[j] = 0ull;   // look-back words of k_front_groups
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

__global__ void __launch_bounds__(FRONT_THREADS) k_front_groups(
    long long n, const cwq_segment* __restrict__ segs, int n_seg, int T, int E, long long n_tiles,
    const float* __restrict__ kl, const uint8_t* __restrict__ hop, const int* __restrict__ seg_first,
    const uint32_t* __restrict__ rows, const uint8_t* __restrict__ entry, const uint32_t* __restrict__ base,
    const long long* __restrict__ seg_group_off, long long* __restrict__ starts, int* __restrict__ gid,
    uint32_t* __restrict__ group_seed, float* __restrict__ total_kl, int* __restrict__ K, ull* __restrict__ keys,
    ull* __restrict__ unit_off, ull* __restrict__ work_off, ull* __restrict__ status, ull* __restrict__ scalars) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = T >> 5;
  uint32_t* node_w = reinterpret_cast<uint32_t*>(smem_raw);     // [W] chain nodes of the tile
  uint32_t* seg_w = node_w + W;                                 // [W] segment starts of the tile
  uint32_t* node_pre = seg_w + W;                               // [W] nodes before each word
  uint32_t* seg_pre = node_pre + W;                             // [W]
  __shared__ ull sm_scan[33];
  __shared__ long long s_tile;
  __shared__ int s_seg[2];
  __shared__ ull s_excl[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = (long long)atomicAdd(&scalars[FRONT_TICKET_GROUPS], 1ull);   // ticket order: look-back needs it
  __syncthreads();
  const long long t = s_tile;
  const long long tile0 = t * T;
  const long long tile_end = tile0 + T < n ? tile0 + T : n;
  const uint32_t B = base[t];
  for (int w = tid; w < W; w += FRONT_THREADS) {
    node_w[w] = rows[((size_t)t * E + entry[t]) * W + w];
    seg_w[w] = 0u;
  }
  if (tid == 0) {
    s_seg[0] = find_segment(segs, n_seg, tile0);
    s_seg[1] = find_segment(segs, n_seg, tile_end - 1);
  }
  __syncthreads();
  const int slo = s_seg[0], shi = s_seg[1];
  if (tid == 0) {
    for (int s = slo; s <= shi; ++s) {
      const long long p = segs[s].start;
      if (p >= tile0 && p < tile_end) seg_w[(p - tile0) >> 5] |= 1u << ((p - tile0) & 31);
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

  // each thread owns D consecutive dims: thread order == group order
  const int D = T / FRONT_THREADS;
  const long long li0 = (long long)tid * D;
  ull my_units = 0, my_work = 0;
  for (int r = 0; r < D; ++r) {
    const long long li = li0 + r, i = tile0 + li;
    if (i >= tile_end) break;
    const int w = (int)(li >> 5), b = (int)(li & 31);
    const uint32_t le = b == 31 ? 0xffffffffu : ((2u << b) - 1u);
    const uint32_t nodes_le = node_pre[w] + __popc(node_w[w] & le), segs_le = seg_pre[w] + __popc(seg_w[w] & le);
    const uint32_t is_node = (node_w[w] >> b) & 1u, is_seg = (seg_w[w] >> b) & 1u;
    gid[i] = (int)(B + nodes_le + segs_le) - 1;
    if (!(is_node | is_seg)) continue;
    const int s = find_segment_between(segs, slo, shi, i);
    const uint32_t seed0 = segs[s].seed;
    const long long seg_end = segs[s].end;
    const long long goff = seg_group_off[s];
    for (int which = 0; which < 2; ++which) {      // 0: the segment's forced first group, 1: the chain node's group
      if (which == 0 ? !is_seg : !is_node) continue;
      const long long g = which == 0 ? (long long)(B + (nodes_le - is_node) + (segs_le - 1)) : (long long)(B + nodes_le - 1 + segs_le);
      // a node's group ends at the next node -- or at the end of its segment (the last dim of a segment hops into the next one)
      const long long end = which == 0 ? (long long)seg_first[s] : (i + hop[i] < seg_end ? i + hop[i] : seg_end);
      float acc = 0.0f;
      for (long long k = i; k < end; ++k) acc = __fadd_rn(acc, kl[k]);
      const int Kg = candidates_from_kl(acc);
      starts[g] = i;
      total_kl[g] = acc;
      K[g] = Kg;
      group_seed[g] = seed0 + (uint32_t)(g - goff);
      keys[g] = KEY_INIT;
      my_units += units_of(Kg);
      const int nq = (int)((end - i + 3) >> 2);
      const int blocks = row_blocks(nq);
      my_work += (ull)Kg * (ull)(blocks > 0 ? blocks : 1);
    }
  }
  ull tot_units, tot_work;
  ull upre = block_exclusive_scan<ull>(my_units, sm_scan, tot_units);
  ull wpre = block_exclusive_scan<ull>(my_work, sm_scan, tot_work);
  if (warp == 0) {
    const ull eu = lookback_exclusive(status, t, tot_units);
    const ull ew = lookback_exclusive(status + n_tiles, t, tot_work);
    if (lane == 0) { s_excl[0] = eu; s_excl[1] = ew; }
  }
  __syncthreads();
  upre += s_excl[0];
  wpre += s_excl[1];
  for (int r = 0; r < D; ++r) {
    const long long li = li0 + r, i = tile0 + li;
    if (i >= tile_end) break;
    const int w = (int)(li >> 5), b = (int)(li & 31);
    const uint32_t is_node = (node_w[w] >> b) & 1u, is_seg = (seg_w[w] >> b) & 1u;
    if (!(is_node | is_seg)) continue;
    const uint32_t le = b == 31 ? 0xffffffffu : ((2u << b) - 1u);
    const uint32_t nodes_le = node_pre[w] + __popc(node_w[w] & le), segs_le = seg_pre[w] + __popc(seg_w[w] & le);
    const int s = find_segment_between(segs, slo, shi, i);
    const long long seg_end = segs[s].end;
    for (int which = 0; which < 2; ++which) {
      if (which == 0 ? !is_seg : !is_node) continue;
      const long long g = which == 0 ? (long long)(B + (nodes_le - is_node) + (segs_le - 1)) : (long long)(B + nodes_le - 1 + segs_le);
      // a node's group ends at the next node -- or at the end of its segment (the last dim of a segment hops into the next one)
      const long long end = which == 0 ? (long long)seg_first[s] : (i + hop[i] < seg_end ? i + hop[i] : seg_end);
      const int Kg = K[g];
      unit_off[g] = upre;
      work_off[g] = wpre;
      upre += units_of(Kg);
      const int nq = (int)((end - i + 3) >> 2);
      const int blocks = row_blocks(nq);
      wpre += (ull)Kg * (ull)(blocks > 0 ? blocks : 1);
    }
  }
  if (t == n_tiles - 1 && tid == FRONT_THREADS - 1) {   // the thread holding the last groups closes both prefixes
    const long long ng = seg_group_off[n_seg];
    unit_off[ng] = upre;
    work_off[ng] = wpre;
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
    long long* __restrict__ seg_outlier_off, ull* __restrict__ pack_status, long long pack_tiles, ull* __restrict__ scalars) {
  __shared__ uint32_t sm_scan[33];
  const int tid = threadIdx.x;
  const long long t = blockIdx.x;
  // the packing kernel that follows on the stream starts from clean look-back words and a clean ticket
  for (long long j = t * FRONT_THREADS + tid; j < pack_tiles; j += (long long)gridDim.x * FRONT_THREADS) pack_status[j] = 0ull;
  if (t == 0 && tid == 0) scalars[FRONT_TICKET_PACK] = 0ull;
  const long long tile0 = t * T;
  const long long tile_end = tile0 + T < n ? tile0 + T : n;
  const int D = T / FRONT_THREADS;
  const long long i0 = tile0 + (long long)tid * D;
  uint32_t my_out = 0;
  for (int r = 0; r < D; ++r) {
    const long long i = i0 + r;
    if (i >= tile_end) break;
    my_out += (mask[i] >> 1) & 1u;
  }
  uint32_t tot;
  uint32_t pre = block_exclusive_scan<uint32_t>(my_out, sm_scan, tot) + tile_out_off[t];
  for (int r = 0; r < D; ++r) {
    const long long i = i0 + r;
    if (i >= tile_end) break;
    const uint8_t m = mask[i];
    int s = -1;
    if (m & 6) s = find_segment(segs, n_seg, i);
    if (m & 4) seg_outlier_off[s] = (long long)pre;       // outliers before this segment
    if (m & 2) {
      const long long local = i - segs[s].start;
      float z[4];
      normal_quad(segs[s].seed + OUTLIER_STREAM_OFFSET, (uint32_t)local, 0u, z);
      const float v = __fadd_rn(__fmul_rn(q_scale[i], z[0]), q_loc[i]);
      const uint16_t qv = quantize_u16(v);
      sample[i] = dequantize_u16(qv);
      outlier_idx[pre] = local;
      outlier_q[pre] = qv;
      ++pre;
    } else {
      const int g = gid[i];
      const uint32_t d = (uint32_t)(i - starts[g]);
      float z[4];
      normal_quad(gseed[g], index[g], d >> 2, z);
      const float zz = (d & 2) ? ((d & 1) ? z[3] : z[2]) : ((d & 1) ? z[1] : z[0]);
      sample[i] = __fadd_rn(__fmul_rn(p_scale[i], zz), p_loc[i]);
    }
  }
}

// Elias-delta packing of index + 1 (BIO:7-20), one pass: per-tile bit counts, look-back, atomicOr of byte-swapped words
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

constexpr int PACK_ITEMS = 8;
constexpr int PACK_TILE = FRONT_THREADS * PACK_ITEMS;

__global__ void __launch_bounds__(FRONT_THREADS) k_pack_elias(const uint32_t* __restrict__ index, long long n_groups_cap,
                                                              const long long* __restrict__ ng_dev, uint32_t* __restrict__ words,
                                                              long long* __restrict__ bit_off, ull* __restrict__ status,
                                                              ull* __restrict__ scalars) {
  __shared__ ull sm_scan[33];
  __shared__ long long s_tile;
  __shared__ ull s_excl;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = (long long)atomicAdd(&scalars[FRONT_TICKET_PACK], 1ull);
  __syncthreads();
  const long long t = s_tile;
  const long long ng = ng_dev ? *ng_dev : n_groups_cap;
  const long long g0 = t * PACK_TILE + (long long)tid * PACK_ITEMS;
  ull pat[PACK_ITEMS];
  int len[PACK_ITEMS];
  ull mine = 0;
#pragma unroll
  for (int r = 0; r < PACK_ITEMS; ++r) {
    len[r] = 0;
    pat[r] = 0;
    if (g0 + r < ng) {
      elias_delta_bits(index[g0 + r] + 1u, pat[r], len[r]);
      mine += (ull)len[r];
    }
  }
  ull tot;
  ull pre = block_exclusive_scan<ull>(mine, sm_scan, tot);
  if (tid < 32) {
    const ull e = lookback_exclusive(status, t, tot);
    if (tid == 0) s_excl = e;
  }
  __syncthreads();
  pre += s_excl;
#pragma unroll
  for (int r = 0; r < PACK_ITEMS; ++r) {
    if (g0 + r < ng) {
      bit_off[g0 + r] = (long long)pre;
      put_bits_bytes(words, pre, pat[r], len[r]);
      pre += (ull)len[r];
      if (g0 + r == ng - 1) bit_off[ng] = (long long)pre;
    }
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
  for (int T = 1024; T <= 16384; T <<= 1) {
    const long long nt = (n + T - 1) / T;
    const long long nb = (nt + CHAIN_BLOCK - 1) / CHAIN_BLOCK;
    const size_t smem_chain = (size_t)nt * E + (size_t)nb * E + (size_t)nb + (size_t)nt + 64;
    const size_t smem_tiles = (size_t)(T + E) * 4 + (size_t)E * (T / 32) * 4 + (size_t)T + 64;
    if (smem_chain <= 200 * 1024 && smem_tiles <= 160 * 1024) {
      p->T = T; p->E = E; p->W = T / 32; p->n_tiles = nt;
      p->smem_tiles = smem_tiles; p->smem_chain = smem_chain; p->smem_groups = (size_t)4 * (T / 32) * 4;
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
  b += align_up((size_t)((cap + PACK_TILE - 1) / PACK_TILE) * 8);   // look-back words of the packing kernel
  b += align_up((size_t)n);                                 // hop
  b += align_up((size_t)n_seg * 4);                         // seg_first
  b += align_up((size_t)p.n_tiles * p.E);                   // maps
  b += align_up((size_t)p.n_tiles * p.E * 2);               // counts
  b += align_up((size_t)p.n_tiles * p.E * p.W * 4);         // rows
  b += align_up((size_t)p.n_tiles * 4) * 2;                 // tile_out, base
  b += align_up((size_t)p.n_tiles);                         // entry
  b += align_up((size_t)p.n_tiles * 8 * 2);                 // look-back words of the group kernel (units, work)
  return b + 1024;
}

bool front_carve(void* workspace, size_t bytes, long long n, int n_seg, long long cap, const FrontPlan& p, FrontWorkspace* w) {
  Workspace ws(workspace, bytes);
  w->unit_off = ws.take<ull>((size_t)cap + 1);
  w->work_off = ws.take<ull>((size_t)cap + 1);
  w->keys = ws.take<ull>((size_t)cap);
  w->tile_out_off = ws.take<uint32_t>((size_t)p.n_tiles);
  w->scalars = ws.take<ull>(8);
  w->pack_status = ws.take<ull>((size_t)((cap + PACK_TILE - 1) / PACK_TILE));
  w->hop = ws.take<uint8_t>((size_t)n);
  w->seg_first = ws.take<int>((size_t)n_seg);
  w->maps = ws.take<uint8_t>((size_t)p.n_tiles * p.E);
  w->counts = ws.take<uint16_t>((size_t)p.n_tiles * p.E);
  w->rows = ws.take<uint32_t>((size_t)p.n_tiles * p.E * p.W);
  w->tile_out = ws.take<uint32_t>((size_t)p.n_tiles);
  w->base = ws.take<uint32_t>((size_t)p.n_tiles);
  w->entry = ws.take<uint8_t>((size_t)p.n_tiles);
  w->status = ws.take<ull>((size_t)p.n_tiles * 2);
  return w->status != nullptr;
}

int front_group(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, long long n,
                const cwq_segment* segs, int n_seg, const FrontPlan& p, const FrontWorkspace& w, const cwq_is_buffers* b,
                cudaStream_t st) {
  if (p.smem_tiles > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_front_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_tiles));
  if (p.smem_chain > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_front_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_chain));
  k_front_tiles<<<(unsigned)p.n_tiles, FRONT_THREADS, p.smem_tiles, st>>>(
      q_loc, q_scale, p_loc, p_scale, n, segs, n_seg, p.T, p.E, b->kl, (float2*)b->coef, b->mask, w.hop, w.seg_first, w.maps,
      w.counts, w.rows, w.tile_out);
  CWQ_LAUNCH_CHECK();
  k_front_chain<<<1, CHAIN_THREADS, p.smem_chain, st>>>(n, segs, n_seg, p.T, p.E, p.n_tiles, w.seg_first, w.maps, w.counts, w.rows,
                                                       w.tile_out, w.entry, w.base, w.tile_out_off, (long long*)b->starts,
                                                       (long long*)b->seg_group_off, (long long*)b->seg_outlier_off, w.status,
                                                       w.scalars);
  CWQ_LAUNCH_CHECK();
  k_front_groups<<<(unsigned)p.n_tiles, FRONT_THREADS, p.smem_groups, st>>>(
      n, segs, n_seg, p.T, p.E, p.n_tiles, b->kl, w.hop, w.seg_first, w.rows, w.entry, w.base, (const long long*)b->seg_group_off,
      (long long*)b->starts, b->gid, b->group_seed, b->total_kl, b->K, w.keys, w.unit_off, w.work_off, w.status, w.scalars);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

int front_finish(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, long long n,
                 const cwq_segment* segs, int n_seg, long long cap, const FrontPlan& p, const FrontWorkspace& w,
                 const cwq_is_buffers* b, cudaStream_t st) {
  const long long pack_tiles = (cap + PACK_TILE - 1) / PACK_TILE;
  CWQ_CUDA(cudaMemsetAsync(b->bits, 0, b->bits_capacity_bytes, st));
  k_finish_dims<<<(unsigned)p.n_tiles, FRONT_THREADS, 0, st>>>(n, segs, n_seg, p.T, q_loc, q_scale, p_loc, p_scale, b->mask,
                                                              (const long long*)b->starts, b->group_seed, b->gid, b->index,
                                                              w.tile_out_off, b->sample, (long long*)b->outlier_idx, b->outlier_q,
                                                              (long long*)b->seg_outlier_off, w.pack_status, pack_tiles, w.scalars);
  CWQ_LAUNCH_CHECK();
  k_pack_elias<<<(unsigned)pack_tiles, FRONT_THREADS, 0, st>>>(b->index, cap, (const long long*)b->seg_group_off + n_seg,
                                                              (uint32_t*)b->bits, (long long*)b->bit_off, w.pack_status, w.scalars);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

}  // namespace cwq
