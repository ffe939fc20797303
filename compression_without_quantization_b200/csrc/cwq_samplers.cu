// cwq_samplers.cu -- the candidate loops: importance-sampler scoring/argmax, greedy persistent
// step loop, the decode (regenerate-from-index) kernels, outliers, index->bit packing and the
// debug view of the candidate stream.  Reference lines are cited at each extern "C" entry in cwq.h.
//
// Work decomposition of the importance sampler (the dominant kernel, k_is_score):
//   a group's K candidates are cut into UNITs of 1024; a WARP claims a run of consecutive
//   units from a global counter (guided: long runs first, single units at the end), scores 32 candidates at a time (one candidate per lane, its
//   Philox blocks (3 per 16 dims) and Box-Muller pairs generated in registers, never stored), keeps a
//   per-lane (score, index) best, shuffle-reduces it and publishes one 64-bit atomicMax per run.
//   No __syncthreads, no HBM traffic beyond the 8 B/dim coefficients and 8 B/group keys.
#include "cwq_common.cuh"
#include "cwq_spec.cuh"
#include "cwq_front.cuh"
#include <atomic>
#include <cooperative_groups.h>
#include <cstdlib>

namespace cwq {

using ull = unsigned long long;

// ------------------------------------------------------------------------------------------
// importance sampler: encode
// ------------------------------------------------------------------------------------------
constexpr int IS_THREADS = 256;
constexpr int IS_WARPS = IS_THREADS / 32;
constexpr int UNIT_CAND = 1024;  // candidates per unit
#ifndef CWQ_UNIT_BATCH
#define CWQ_UNIT_BATCH 4
#endif
// build-time knobs of tools/ab_variants.py (same results, different code generation)
#ifndef CWQ_IS_MIN_BLOCKS
#define CWQ_IS_MIN_BLOCKS 3
#endif

constexpr int UNIT_BATCH = CWQ_UNIT_BATCH;    // units claimed per atomic (low-K instantiation; the default one is guided)
#ifndef CWQ_GUIDED_MAX
#define CWQ_GUIDED_MAX 64
#endif
#ifndef CWQ_GUIDED_DIV
#define CWQ_GUIDED_DIV 4
#endif
constexpr int GUIDED_MAX = CWQ_GUIDED_MAX, GUIDED_DIV = CWQ_GUIDED_DIV;

// The scored range of groups is [g0, g1).  Either end may live on the device (g0_dev / g1_dev non-null): g1_dev is the
// group count of a synchronisation-free encode, (g0_dev, g1_dev) the range a rank computed for itself on the device;
// the host-side values are then only capacities and every scan runs over the capacity.
struct GroupRange {
  long long g0, g1;
  const long long* g0_dev;
  const long long* g1_dev;
  __device__ long long begin() const { return g0_dev ? *g0_dev : g0; }
  __device__ long long end() const { return g1_dev ? *g1_dev : g1; }
};
struct UnitCount {
  const int* K;
  GroupRange r;
  __device__ ull operator()(long long j) const {
    const long long g = r.begin() + j;
    if (g >= r.end()) return 0ull;
    return (ull)((K[g] + UNIT_CAND - 1) / UNIT_CAND);
  }
};
struct UnitDeliver {
  ull* unit_off;
  ull* keys;
  GroupRange r;
  __device__ void operator()(long long j, ull pre, ull) const {
    const long long g = r.begin() + j;
    if (g >= r.end()) return;
    unit_off[j] = pre;
    keys[g] = KEY_INIT;
  }
};
__global__ void k_units_finish(ull* unit_off, GroupRange r, const ull* total, ull* counter) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { unit_off[r.end() - r.begin()] = *total; *counter = 0; }
}

// ---- contiguous equal-work ranges of groups for the ranks of a sharded run, computed on the device ----
// work of a group = candidates x Philox blocks per candidate; cut r (1 <= r < world) closes range r-1 behind the first
// group whose inclusive work prefix reaches r/world of the total (integer arithmetic: every rank gets the same cuts).
struct ShardWork {
  const int* K;
  const long long* starts;
  const long long* ng_dev;
  __device__ ull operator()(long long j) const {
    if (j >= *ng_dev) return 0ull;
    const int nq = (int)((starts[j + 1] - starts[j] + 3) >> 2);
    const int blocks = row_blocks(nq);
    return (ull)K[j] * (ull)(blocks > 0 ? blocks : 1);
  }
};
struct ShardCuts {
  const ull* total;
  int world;
  long long* bounds;
  __device__ void operator()(long long j, ull pre, ull cnt) const {
    if (cnt == 0ull) return;
    const ull tot = *total;
    for (int r = 1; r < world; ++r) {
      const ull target = tot * (ull)r;
      if (pre * (ull)world < target && (pre + cnt) * (ull)world >= target) bounds[r] = j + 1;
    }
  }
};
__global__ void k_shard_bounds_init(long long* bounds, int world, const long long* ng_dev) {
  const int r = threadIdx.x;
  if (r <= world) bounds[r] = r == 0 ? 0 : *ng_dev;
}

__device__ __forceinline__ ull warp_max_key(ull key) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ull other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  return key;
}

// score of one quad: two accumulators (even / odd dims), spec section 3.7
__device__ __forceinline__ void is_score_quad(const float4* __restrict__ sc, int q, const float (&z)[4], float& acc0, float& acc1) {
  const float4 c01 = sc[2 * q], c23 = sc[2 * q + 1];
  acc0 = __fmaf_rn(z[0], __fmaf_rn(c01.x, z[0], c01.y), acc0);
  acc1 = __fmaf_rn(z[1], __fmaf_rn(c01.z, z[1], c01.w), acc1);
  acc0 = __fmaf_rn(z[2], __fmaf_rn(c23.x, z[2], c23.y), acc0);
  acc1 = __fmaf_rn(z[3], __fmaf_rn(c23.z, z[3], c23.w), acc1);
}

// the first pair of a quad only (its last two dims do not exist: unit_quads_half)
__device__ __forceinline__ void is_score_pair(const float4* __restrict__ sc, int q, const float (&z)[4], float& acc0, float& acc1) {
  const float4 c01 = sc[2 * q];
  acc0 = __fmaf_rn(z[0], __fmaf_rn(c01.x, z[0], c01.y), acc0);
  acc1 = __fmaf_rn(z[1], __fmaf_rn(c01.z, z[1], c01.w), acc1);
}

// candidates [kb, ke) of one group of NQ <= 4 quads (one unit), coefficients staged in this warp's shared slice.
// Per candidate: 2 shared IMAD.WIDE (lane_shared) + 14 per block (rounds 3..9), min(NQ, 3) blocks; the
// (seed, block) parts of rounds 0-2 live in registers (bc).  HALF: the last quad holds one or two dims, its second
// Box-Muller pair is skipped (a third of the 13-15-dim groups of the default configuration end that way).
template <int NQ, bool HALF>
__device__ __forceinline__ ull is_score_range_impl(const float4* __restrict__ sc, uint32_t sseed, uint32_t kb, uint32_t ke, int lane) {
  float best = __uint_as_float(0xff800000u);
  uint32_t bestk = 0xffffffffu;
  constexpr int NB = NQ < 3 ? NQ : 3;
  BlockConst bc[NB > 0 ? NB : 1];
#pragma unroll
  for (int j = 0; j < NB; ++j) bc[j] = block_const(sseed, (uint32_t)j);
  for (uint32_t k = kb + lane; k < ke; k += 32) {
    float acc0 = 0.0f, acc1 = 0.0f;
    if (NQ > 0) {
      const LaneShared ls = lane_shared(sseed, k);
      if (HALF) {
        unit_quads_half<(NQ > 0 ? NQ : 1)>(ls, [&](int b) { return bc[b]; },
                                           [&](int q, const float (&z)[4], int pairs) {
                                             if (pairs == 2) is_score_quad(sc, q, z, acc0, acc1); else is_score_pair(sc, q, z, acc0, acc1);
                                           });
      } else {
        unit_quads<(NQ > 0 ? NQ : 1)>(ls, [&](int b) { return bc[b]; },
                                      [&](int q, const float (&z)[4]) { is_score_quad(sc, q, z, acc0, acc1); });
      }
    }
    float s = __fadd_rn(acc0, acc1);
    if (s > best) { best = s; bestk = k; }
  }
  return ((ull)__float_as_uint(best) << 32) | (ull)bestk;
}
template <int NQ, bool HALF>
__device__ __forceinline__ void is_score_range(const float4* __restrict__ sc, uint32_t sseed, uint32_t kb, uint32_t ke,
                                               int lane, float& best, uint32_t& bestk) {
  const ull r = is_score_range_impl<NQ, HALF>(sc, sseed, kb, ke, lane);
  best = __uint_as_float((uint32_t)(r >> 32));
  bestk = (uint32_t)r;
}

// 16 < G <= 64: coefficients and the per-(seed, block) Philox constants staged in this warp's shared slice
__device__ __forceinline__ void is_score_range_mid(const float4* __restrict__ sc, const uint4* __restrict__ sb, int nq, uint32_t sseed,
                                                   uint32_t kb, uint32_t ke, int lane, float& best, uint32_t& bestk) {
  for (uint32_t k = kb + lane; k < ke; k += 32) {
    const LaneShared ls = lane_shared(sseed, k);
    float acc0 = 0.0f, acc1 = 0.0f;
    row_quads(ls, nq,
              [&](int j) { const uint4 t = sb[j]; BlockConst c; c.a = t.x; c.b = t.y; c.c3 = t.z; return c; },
              [&](int q, const float (&z)[4]) { is_score_quad(sc, q, z, acc0, acc1); });
    float s = __fadd_rn(acc0, acc1);
    if (s > best) { best = s; bestk = k; }
  }
}

// any group size: coefficients read through the read-only path (warp-uniform addresses)
__device__ __forceinline__ void is_score_range_generic(const float2* __restrict__ coef, int G, uint32_t sseed, uint32_t kb,
                                                       uint32_t ke, int lane, float& best, uint32_t& bestk) {
  const int nq = (G + 3) >> 2;
  for (uint32_t k = kb + lane; k < ke; k += 32) {
    const LaneShared ls = lane_shared(sseed, k);
    float acc0 = 0.0f, acc1 = 0.0f;
    row_quads(ls, nq, [&](int j) { return block_const(sseed, (uint32_t)j); },
              [&](int q, const float (&z)[4]) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  int d = 4 * q + jj;
                  float2 c = d < G ? __ldg(&coef[d]) : make_float2(0.0f, 0.0f);
                  float w = __fmaf_rn(c.x, z[jj], c.y);
                  if (jj & 1) acc1 = __fmaf_rn(z[jj], w, acc1); else acc0 = __fmaf_rn(z[jj], w, acc0);
                }
              });
    float s = __fadd_rn(acc0, acc1);
    if (s > best) { best = s; bestk = k; }
  }
}

constexpr double LOW_K_MAX_NATS = 12.0 * 0.6931471805599453 - 1.0;   // budgets up to 12 bits per group (CIS:181)

// SMALL: every group has <= 16 dims (the reference defaults): only the register-resident path is compiled in,
// which keeps the kernel at 80 registers without spills / 3 resident CTAs per SM (measured best: 2 CTAs at 112 registers
// and 4 CTAs at 64 registers with spills are 4 % and 0.5 % slower, profiles/r02_s2_ab_inline_occupancy.log).
// LOWK: the launch for small KL budgets (<= 12 bits per group: tens to hundreds of candidates, one unit per group).  The
// per-group bookkeeping weighs as much as the scoring there, so this instantiation drops what it can: when every group is
// exactly one unit (checked on the device: #units == #groups) the group of a unit needs no search and no unit prefix, and
// the warp argmax is two REDUX instead of five 64-bit shuffle rounds.  Kept apart from the default instantiation on
// purpose: ptxas schedules the 231-instruction candidate loop differently whenever the code around it changes, and the
// default's schedule is the one measured fastest on the headline batch (tools/ab_variants.py).
template <bool SMALL, bool LOWK>
__global__ void __launch_bounds__(IS_THREADS, CWQ_IS_MIN_BLOCKS) k_is_score(const float2* __restrict__ coef, const long long* __restrict__ starts,
                                                         const int* __restrict__ K, const uint32_t* __restrict__ gseed,
                                                         GroupRange range, const ull* __restrict__ unit_off_any,
                                                         int unit_off_is_global, ull* __restrict__ counter,
                                                         ull* __restrict__ keys) {
  pdl_wait();      // no early trigger: the kernel fills the SMs for most of a step, what follows moves in as its CTAs retire
  __shared__ float4 s_coef[IS_WARPS][SMALL ? 8 : 32];   // 16 or 64 (a, b) pairs per warp
  __shared__ uint4 s_quad[IS_WARPS][SMALL ? 1 : 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long g0 = range.begin(), ng = range.end() - g0;
  // unit_off[j] = units before group g0 + j: either the range's own scan (starts at 0) or a slice of the prefix over all
  // groups that the fused front end wrote (then rebased on the first unit of the range)
  const ull* __restrict__ unit_off = unit_off_any + (unit_off_is_global ? g0 : 0);
  const ull u_first = unit_off[0];
  const ull n_units = unit_off[ng];
  float2* my_coef = reinterpret_cast<float2*>(s_coef[warp]);
  const bool one_each = LOWK && (n_units - u_first) == (ull)ng;      // every group of the range is exactly one unit
  // Guided self-scheduling (default instantiation): a claim takes 1/GUIDED_DIV of a warp's fair share of the units still
  // unclaimed, at most GUIDED_MAX and at least one -- long claims while plenty is left (claim, search and staging paid
  // once per run), single units at the end (the tail of the kernel is one unit long instead of four).  The share comes
  // from a relaxed read of the counter: a heuristic only, the atomicAdd decides.  Measured against fixed claims of 4 units
  // (tools/ab_variants.py, same candidate-loop code): 24 images 107.15 -> 106.05 ms, one image 4.644 -> 4.579 ms;
  // dividing by 2 or 1 instead of 4 loses on one image (units differ in cost), 8 gains less.
  const ull span_units = n_units - u_first;
  const ull per_round = (ull)gridDim.x * IS_WARPS * GUIDED_DIV;
  for (;;) {
    ull u = 0;
    ull batch = (ull)UNIT_BATCH;
    if (!LOWK) {
      if (lane == 0) {
        const ull seen = ld_volatile_u64(counter);
        const ull left = seen < span_units ? span_units - seen : 0ull;
        ull b = left / per_round;
        b = b < 1ull ? 1ull : (b > (ull)GUIDED_MAX ? (ull)GUIDED_MAX : b);
        batch = b;
        u = atomicAdd(counter, b);
      }
      batch = __shfl_sync(0xffffffffu, batch, 0);
    } else if (lane == 0) u = atomicAdd(counter, (ull)UNIT_BATCH);
    u = __shfl_sync(0xffffffffu, u, 0) + u_first;
    if (u >= n_units) break;
    const ull u_end = u + batch < n_units ? u + batch : n_units;
    // group of unit u: the largest j with unit_off[j] <= u (unit_off is strictly increasing: every group has a unit).
    // 32-ary search, one probe per lane and round: ~4 rounds of one L2 access instead of 18 dependent ones.
    long long lo = 0, hi = ng - 1;
    if (LOWK && one_each) lo = hi = (long long)(u - u_first);
    while (lo < hi) {
      const long long step = (hi - lo + 31) >> 5;
      long long probe = lo + (long long)(lane + 1) * step;
      probe = probe < hi ? probe : hi;
      const int c = __popc(__ballot_sync(0xffffffffu, unit_off[probe] <= u));   // lanes 0..c-1 probe at or below the answer
      const long long below = lo + (long long)c * step, above = lo + (long long)(c + 1) * step;
      const long long new_lo = c == 0 ? lo : (below < hi ? below : hi);
      const long long new_hi = c == 32 ? hi : (above < hi ? above : hi) - 1;
      lo = new_lo;
      hi = new_hi > new_lo ? new_hi : new_lo;
    }
    long long j = lo;
    while (u < u_end) {
      ull off_j, off_n;
      if (LOWK && one_each) { off_j = u_first + (ull)j; off_n = off_j + 1ull; }
      else { off_j = unit_off[j]; off_n = unit_off[j + 1]; }
      const ull ue = u_end < off_n ? u_end : off_n;
      const long long g = g0 + j;
      const uint32_t Kg = (uint32_t)K[g];
      const uint32_t kb = (uint32_t)(u - off_j) * UNIT_CAND;
      const ull ke64 = (ue - off_j) * (ull)UNIT_CAND;
      const uint32_t ke = ke64 < Kg ? (uint32_t)ke64 : Kg;
      const long long s = starts[g];
      const int G = (int)(starts[g + 1] - s);
      const uint32_t sseed = gseed[g];
      float best = __uint_as_float(0xff800000u);
      uint32_t bestk = 0xffffffffu;
      if (SMALL || G <= 16) {
        __syncwarp();
        if (lane < 16) my_coef[lane] = lane < G ? coef[s + lane] : make_float2(0.0f, 0.0f);
        __syncwarp();
        const float4* sc = s_coef[warp];
        switch ((G + 1) >> 1) {      // Box-Muller pairs of the row: an odd count leaves the last quad half empty
          case 0: is_score_range<0, false>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 1: is_score_range<1, true>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 2: is_score_range<1, false>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 3: is_score_range<2, true>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 4: is_score_range<2, false>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 5: is_score_range<3, true>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 6: is_score_range<3, false>(sc, sseed, kb, ke, lane, best, bestk); break;
          case 7: is_score_range<4, true>(sc, sseed, kb, ke, lane, best, bestk); break;
          default: is_score_range<4, false>(sc, sseed, kb, ke, lane, best, bestk); break;
        }
      } else if (!SMALL && G <= 64) {
        const int nq = (G + 3) >> 2;
        __syncwarp();
        my_coef[lane] = lane < G ? coef[s + lane] : make_float2(0.0f, 0.0f);
        my_coef[lane + 32] = lane + 32 < G ? coef[s + lane + 32] : make_float2(0.0f, 0.0f);
        if (lane < row_blocks(nq)) {
          BlockConst c = block_const(sseed, (uint32_t)lane);
          s_quad[warp][lane] = make_uint4(c.a, c.b, c.c3, 0u);
        }
        __syncwarp();
        is_score_range_mid(s_coef[warp], s_quad[warp], nq, sseed, kb, ke, lane, best, bestk);
      } else if (!SMALL) {
        is_score_range_generic(coef + s, G, sseed, kb, ke, lane, best, bestk);
      }
      // warp argmax; low-K: as two REDUX -- the largest orderable score, then the lowest index among the lanes that hold it,
      // i.e. the maximum of pack_key() over the lanes (0 = a lane without a candidate; orderable() of a real score is never 0)
      if (LOWK) {
        const uint32_t sb = bestk == 0xffffffffu ? 0u : orderable(best);
        const uint32_t top = __reduce_max_sync(0xffffffffu, sb);
        const uint32_t kmin = __reduce_min_sync(0xffffffffu, sb == top ? bestk : 0xffffffffu);
        if (lane == 0 && top != 0u) atomicMax(&keys[g], ((ull)top << 32) | (ull)(~kmin));
      } else {
        ull key = bestk == 0xffffffffu ? 0ull : pack_key(best, bestk);
        key = warp_max_key(key);
        if (lane == 0 && key != 0ull) atomicMax(&keys[g], key);
      }
      u = ue;
      ++j;
    }
  }
}

// also hands the unit counter back at zero for the next scoring launch on these buffers
__global__ void __launch_bounds__(256) k_is_index(const ull* __restrict__ keys, GroupRange range, uint32_t* __restrict__ index,
                                                  ull* __restrict__ counter) {
  pdl_wait();
  pdl_trigger();
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ull;
  // grid-stride: the host often knows only the capacity of the range (its end lives on the device)
  const long long end = range.end(), stride = (long long)gridDim.x * blockDim.x;
  for (long long g = range.begin() + (long long)blockIdx.x * blockDim.x + threadIdx.x; g < end; g += stride) index[g] = key_index(keys[g]);
}

// ------------------------------------------------------------------------------------------
// importance sampler: decode (row index[g] of each group's table), un-rescaled
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_is_decode(const long long* __restrict__ starts, const uint32_t* __restrict__ gseed,
                                                   const int* __restrict__ gid, const uint32_t* __restrict__ index,
                                                   long long n, const float* __restrict__ p_loc,
                                                   const float* __restrict__ p_scale, float* __restrict__ sample) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int g = gid[i];
  uint32_t d = (uint32_t)(i - starts[g]);
  float z[4];
  normal_quad(gseed[g], index[g], d >> 2, z);
  float zz = (d & 2) ? ((d & 1) ? z[3] : z[2]) : ((d & 1) ? z[1] : z[0]);
  sample[i] = __fadd_rn(__fmul_rn(p_scale[i], zz), p_loc[i]);
}

// ------------------------------------------------------------------------------------------
// outliers
// ------------------------------------------------------------------------------------------
struct OutlierCount {
  const uint8_t* mask;
  __device__ uint32_t operator()(long long i) const { return (uint32_t)((mask[i] >> 1) & 1); }
};
struct OutlierDeliver {
  const uint8_t* mask;
  const float* q_loc;
  const float* q_scale;
  const cwq_segment* segs;
  int n_seg;
  float* sample;
  long long* outlier_idx;
  uint16_t* outlier_q;
  long long* seg_outlier_off;
  __device__ void operator()(long long i, uint32_t pre, uint32_t cnt) const {
    const uint8_t m = mask[i];
    if ((m & 6) == 0) return;
    int s = find_segment(segs, n_seg, i);
    if (m & 4) seg_outlier_off[s] = (long long)pre;   // outliers before this segment
    if (cnt == 0) return;
    long long local = i - segs[s].start;
    float z[4];
    normal_quad(segs[s].seed + OUTLIER_STREAM_OFFSET, (uint32_t)local, 0u, z);
    float v = __fadd_rn(__fmul_rn(q_scale[i], z[0]), q_loc[i]);
    const uint16_t qv = quantize_u16(v);
    // The encoder's sample carries what the decoder will reconstruct (the reference returns the unquantised draw,
    // CIS:267, so its level-1 prior differs between encoder and decoder whenever level 2 has outliers).
    sample[i] = dequantize_u16(qv);
    outlier_idx[pre] = local;
    outlier_q[pre] = qv;
  }
};
__global__ void k_outlier_finish(const uint32_t* __restrict__ total, int n_seg, long long* __restrict__ off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) off[n_seg] = (long long)*total;
}

__global__ void __launch_bounds__(256) k_outlier_decode(float* __restrict__ sample, const long long* __restrict__ pos,
                                                        const uint16_t* __restrict__ q, long long m) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  // The listed dims are overwritten whatever the value: the reference scatters into a dense zero tensor and keeps the
  // sample where `updates == 0` (CIS:351-361), a sentinel the explicit index list makes unnecessary.
  sample[pos[j]] = dequantize_u16(q[j]);
}

// ------------------------------------------------------------------------------------------
// greedy sampler
// ------------------------------------------------------------------------------------------
constexpr int GREEDY_THREADS = 512;

// one quad of the greedy score: u = z * c + e, acc += u * u (even / odd dims), spec section 3.8
__device__ __forceinline__ void greedy_score_quad(const float4* __restrict__ ce4, int q, const float (&z)[4], float& acc0, float& acc1) {
  const float4 c01 = ce4[2 * q], c23 = ce4[2 * q + 1];
  float u0 = __fmaf_rn(z[0], c01.x, c01.y);
  float u1 = __fmaf_rn(z[1], c01.z, c01.w);
  float u2 = __fmaf_rn(z[2], c23.x, c23.y);
  float u3 = __fmaf_rn(z[3], c23.z, c23.w);
  acc0 = __fmaf_rn(u0, u0, acc0);
  acc1 = __fmaf_rn(u1, u1, acc1);
  acc0 = __fmaf_rn(u2, u2, acc0);
  acc1 = __fmaf_rn(u3, u3, acc1);
}


__global__ void __launch_bounds__(GREEDY_THREADS) k_greedy_encode(const float2* __restrict__ tparam, const long long* __restrict__ starts,
                                                                  const uint32_t* __restrict__ gseed, long long g0, int n_steps,
                                                                  uint32_t n_samples, float shard_scale,
                                                                  const float* __restrict__ shard_loc,
                                                                  uint32_t* __restrict__ indices, float* __restrict__ best_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ ull s_warp_key[GREEDY_THREADS / 32];
  __shared__ uint32_t s_choice;
  const long long g = g0 + blockIdx.x;
  const long long s = starts[g];
  const int G = (int)(starts[g + 1] - s);
  const int nq = (G + 3) >> 2, Gp = nq * 4;
  float2* ce = reinterpret_cast<float2*>(smem_raw);            // (c_d, e_d)
  uint4* qtab = reinterpret_cast<uint4*>(smem_raw + (size_t)Gp * sizeof(float2));   // (a, b, c3, -) per quad, this step's stream
  float* best = reinterpret_cast<float*>(smem_raw + (size_t)Gp * sizeof(float2) + (size_t)nq * sizeof(uint4));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t group_seed = gseed[g];
  const float step_scale = __fmul_rn(shard_scale, NORMAL_SCALE);   // the step's N(0,1) draw is NORMAL_SCALE * y: folded into the scale
  for (int d = tid; d < Gp; d += GREEDY_THREADS) {
    best[d] = 0.0f;
    ce[d] = make_float2(d < G ? __fmul_rn(step_scale, tparam[s + d].y) : 0.0f, 0.0f);
  }
  __syncthreads();
  for (int step = 0; step < n_steps; ++step) {
    const uint32_t sseed = 1000u * group_seed + (uint32_t)step;  // CGS:55
    for (int d = tid; d < G; d += GREEDY_THREADS) {
      float2 t = tparam[s + d];
      const float b = __fadd_rn(best[d], shard_loc ? shard_loc[s + d] : 0.0f);   // the shard's loc, CGS:42,57
      best[d] = b;
      ce[d].y = __fmul_rn(__fsub_rn(b, t.x), t.y);
    }
    for (int j = tid; j < row_blocks(nq); j += GREEDY_THREADS) {
      BlockConst c = block_const(sseed, (uint32_t)j);
      qtab[j] = make_uint4(c.a, c.b, c.c3, 0u);
    }
    __syncthreads();
    float bs = __uint_as_float(0x7f800000u);
    uint32_t bk = 0xffffffffu;
    const float4* ce4 = reinterpret_cast<const float4*>(ce);
    for (uint32_t k = tid; k < n_samples; k += GREEDY_THREADS) {
      const LaneShared ls = lane_shared(sseed, k);
      float acc0 = 0.0f, acc1 = 0.0f;
      row_quads(ls, nq,
                [&](int j) { const uint4 t = qtab[j]; BlockConst c; c.a = t.x; c.b = t.y; c.c3 = t.z; return c; },
                [&](int q, const float (&z)[4]) { greedy_score_quad(ce4, q, z, acc0, acc1); });
      float sc = __fadd_rn(acc0, acc1);
      if (sc < bs) { bs = sc; bk = k; }
    }
    // argmin, lowest index on ties == argmax of the negated score
    ull key = bk == 0xffffffffu ? 0ull : pack_key(-bs, bk);
    key = warp_max_key(key);
    if (lane == 0) s_warp_key[warp] = key;
    __syncthreads();
    if (warp == 0) {
      ull k2 = lane < GREEDY_THREADS / 32 ? s_warp_key[lane] : 0ull;
      k2 = warp_max_key(k2);
      if (lane == 0) {
        uint32_t choice = k2 == 0ull ? 0u : key_index(k2);
        s_choice = choice;
        indices[g * n_steps + step] = choice;
      }
    }
    __syncthreads();
    const uint32_t choice = s_choice;
    for (int q = tid; q < nq; q += GREEDY_THREADS) {
      float y[4];
      raw_quad(sseed, choice, (uint32_t)q, y);
#pragma unroll
      for (int j = 0; j < 4; ++j) best[4 * q + j] = __fmaf_rn(step_scale, y[j], best[4 * q + j]);
    }
    __syncthreads();
  }
  for (int d = tid; d < G; d += GREEDY_THREADS) best_out[s + d] = best[d];
}

// ------------------------------------------------------------------------------------------
// greedy sampler, lockstep variant for FEW groups (e.g. one image = 360 groups on 148 SMs): all groups
// advance one step at a time; inside a step the work is cut into (group, 1024-candidate chunk) units that
// CTAs claim dynamically, so the load is balanced however few groups there are.  Two grid-wide barriers
// per step (cooperative launch); step state (best, (c,e)) lives in global memory between steps.
// Same arithmetic and same argmin (order-independent 64-bit atomicMax) as k_greedy_encode.
// ------------------------------------------------------------------------------------------
constexpr int GREEDY_CHUNK = 1024;

__global__ void __launch_bounds__(GREEDY_THREADS) k_greedy_lockstep(const float2* __restrict__ tparam, const long long* __restrict__ starts,
                                                                    const uint32_t* __restrict__ gseed, long long g0, long long ng,
                                                                    int n_steps, uint32_t n_samples, float shard_scale,
                                                                    const float* __restrict__ shard_loc,
                                                                    float* __restrict__ best, float2* __restrict__ ce,
                                                                    ull* __restrict__ keys, unsigned int* __restrict__ counters,
                                                                    uint32_t* __restrict__ indices) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ ull s_warp_key[GREEDY_THREADS / 32];
  __shared__ unsigned int s_unit;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long d_begin = starts[g0], d_end = starts[g0 + ng];
  const unsigned int chunks = (n_samples + GREEDY_CHUNK - 1) / GREEDY_CHUNK;
  const unsigned long long n_units = (unsigned long long)ng * chunks;
  const float step_scale = __fmul_rn(shard_scale, NORMAL_SCALE);

  for (long long i = d_begin + (long long)blockIdx.x * blockDim.x + tid; i < d_end; i += (long long)gridDim.x * blockDim.x) {
    float2 t = tparam[i];
    const float b = __fadd_rn(0.0f, shard_loc ? shard_loc[i] : 0.0f);   // best holds "running sample + this step's shard loc"
    best[i] = b;
    ce[i] = make_float2(__fmul_rn(step_scale, t.y), __fmul_rn(__fsub_rn(b, t.x), t.y));
  }
  for (long long j = (long long)blockIdx.x * blockDim.x + tid; j < ng; j += (long long)gridDim.x * blockDim.x) keys[j] = 0ull;
  grid.sync();

  for (int step = 0; step < n_steps; ++step) {
    // ---- phase A: score every (group, chunk) unit ----
    for (;;) {
      __syncthreads();
      if (tid == 0) s_unit = atomicAdd(&counters[step], 1u);
      __syncthreads();
      const unsigned long long unit = s_unit;
      if (unit >= n_units) break;
      const long long j = (long long)(unit / chunks);
      const uint32_t chunk = (uint32_t)(unit - (unsigned long long)j * chunks);
      const long long g = g0 + j;
      const long long s = starts[g];
      const int G = (int)(starts[g + 1] - s);
      const int nq = (G + 3) >> 2, Gp = nq * 4;
      float2* s_ce = reinterpret_cast<float2*>(smem_raw);
      uint4* s_q = reinterpret_cast<uint4*>(smem_raw + (size_t)Gp * sizeof(float2));
      const uint32_t sseed = 1000u * gseed[g] + (uint32_t)step;
      for (int d = tid; d < Gp; d += GREEDY_THREADS) s_ce[d] = d < G ? ce[s + d] : make_float2(0.0f, 0.0f);
      for (int jb = tid; jb < row_blocks(nq); jb += GREEDY_THREADS) {
        BlockConst c = block_const(sseed, (uint32_t)jb);
        s_q[jb] = make_uint4(c.a, c.b, c.c3, 0u);
      }
      __syncthreads();
      float bs = __uint_as_float(0x7f800000u);
      uint32_t bk = 0xffffffffu;
      const float4* ce4 = reinterpret_cast<const float4*>(s_ce);
      const uint32_t k_end = (chunk + 1) * GREEDY_CHUNK < n_samples ? (chunk + 1) * GREEDY_CHUNK : n_samples;
      for (uint32_t k = chunk * GREEDY_CHUNK + tid; k < k_end; k += GREEDY_THREADS) {
        const LaneShared ls = lane_shared(sseed, k);
        float acc0 = 0.0f, acc1 = 0.0f;
        row_quads(ls, nq,
                  [&](int jb) { const uint4 t = s_q[jb]; BlockConst c; c.a = t.x; c.b = t.y; c.c3 = t.z; return c; },
                  [&](int q, const float (&z)[4]) { greedy_score_quad(ce4, q, z, acc0, acc1); });
        float sc = __fadd_rn(acc0, acc1);
        if (sc < bs) { bs = sc; bk = k; }
      }
      ull key = bk == 0xffffffffu ? 0ull : pack_key(-bs, bk);
      key = warp_max_key(key);
      if (lane == 0) s_warp_key[warp] = key;
      __syncthreads();
      if (warp == 0) {
        ull k2 = lane < GREEDY_THREADS / 32 ? s_warp_key[lane] : 0ull;
        k2 = warp_max_key(k2);
        if (lane == 0 && k2 != 0ull) atomicMax(&keys[j], k2);
      }
    }
    grid.sync();
    // ---- phase B: apply the chosen row of every group ----
    for (long long j = blockIdx.x; j < ng; j += gridDim.x) {
      const long long g = g0 + j;
      const long long s = starts[g];
      const int G = (int)(starts[g + 1] - s);
      const int nq = (G + 3) >> 2;
      const ull k2 = keys[j];
      const uint32_t choice = k2 == 0ull ? 0u : key_index(k2);
      const uint32_t sseed = 1000u * gseed[g] + (uint32_t)step;
      for (int q = tid; q < nq; q += GREEDY_THREADS) {
        float z[4];
        raw_quad(sseed, choice, (uint32_t)q, z);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int d = 4 * q + jj;
          if (d < G) {
            const float chosen = __fmaf_rn(step_scale, z[jj], best[s + d]);
            const float2 t = tparam[s + d];
            if (step + 1 == n_steps) { best[s + d] = chosen; continue; }
            const float b = __fadd_rn(chosen, shard_loc ? shard_loc[s + d] : 0.0f);
            best[s + d] = b;
            ce[s + d].y = __fmul_rn(__fsub_rn(b, t.x), t.y);
          }
        }
      }
      __syncthreads();   // every thread has read keys[j] before it is cleared
      if (tid == 0) { indices[g * n_steps + step] = choice; keys[j] = 0ull; }
    }
    grid.sync();
  }
}

__global__ void __launch_bounds__(256) k_greedy_decode(const long long* __restrict__ starts, const uint32_t* __restrict__ gseed,
                                                       const int* __restrict__ gid, const uint32_t* __restrict__ indices,
                                                       long long n, int n_steps, float shard_scale,
                                                       const float* __restrict__ shard_loc,
                                                       const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                                                       float* __restrict__ sample) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int g = gid[i];
  uint32_t d = (uint32_t)(i - starts[g]);
  uint32_t group_seed = gseed[g];
  const float l = shard_loc ? shard_loc[i] : 0.0f;
  const float step_scale = __fmul_rn(shard_scale, NORMAL_SCALE);
  float b = 0.0f;
  for (int step = 0; step < n_steps; ++step) {
    float z[4];
    raw_quad(1000u * group_seed + (uint32_t)step, indices[(long long)g * n_steps + step], d >> 2, z);
    float zz = (d & 2) ? ((d & 1) ? z[3] : z[2]) : ((d & 1) ? z[1] : z[0]);
    b = __fmaf_rn(step_scale, zz, __fadd_rn(b, l));
  }
  sample[i] = __fadd_rn(__fmul_rn(p_scale[i], b), p_loc[i]);
}

// loc of the proposal shard of the per-group greedy entry points, rescaled coordinates: (p_loc / n_steps) / p_scale (CGS:42)
__global__ void __launch_bounds__(256) k_greedy_shard_loc(const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                                                          long long n, float n_steps, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fdiv_rn(__fdiv_rn(p_loc[i], n_steps), p_scale[i]);
}

__global__ void __launch_bounds__(256) k_unrescale(const float* __restrict__ best, const float* __restrict__ p_loc,
                                                   const float* __restrict__ p_scale, long long n, float* __restrict__ sample) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sample[i] = __fadd_rn(__fmul_rn(p_scale[i], best[i]), p_loc[i]);
}

// ------------------------------------------------------------------------------------------
// index -> bits
// ------------------------------------------------------------------------------------------
// OR `len` bits (right-aligned in pattern) into a big-endian-bit-ordered word stream at bit `pos`
__device__ __forceinline__ void put_bits(uint32_t* words, ull pos, ull pattern, int len) {
  ull w = pos >> 5;
  int o = (int)(pos & 31);
  int remaining = len;
  while (remaining > 0) {
    int take = 32 - o < remaining ? 32 - o : remaining;
    uint32_t chunk = (uint32_t)((pattern >> (remaining - take)) & ((1ull << take) - 1ull));
    atomicOr(&words[w], chunk << (32 - o - take));
    remaining -= take;
    ++w;
    o = 0;
  }
}

__device__ __forceinline__ void elias_delta(uint32_t x, ull& pattern, int& len) {
  int n = 31 - __clz((int)x);          // floor(log2 x), BIO:11
  int l = 31 - __clz(n + 1);           // floor(log2 (n+1)), BIO:12
  pattern = ((ull)(n + 1) << n) | (ull)(x & ((n == 0) ? 0u : (0xffffffffu >> (32 - n))));
  len = 2 * l + n + 1;
}

struct EliasLen {
  const uint32_t* index;
  const long long* ng_dev;
  __device__ ull operator()(long long g) const {
    if (ng_dev && g >= *ng_dev) return 0ull;
    ull p; int len;
    elias_delta(index[g] + 1u, p, len);
    return (ull)len;
  }
};
struct EliasDeliver {
  const uint32_t* index;
  long long* bit_off;
  uint32_t* words;
  const long long* ng_dev;
  __device__ void operator()(long long g, ull pre, ull) const {
    if (ng_dev && g >= *ng_dev) return;
    ull p; int len;
    elias_delta(index[g] + 1u, p, len);
    bit_off[g] = (long long)pre;
    put_bits(words, pre, p, len);
  }
};
__global__ void k_bits_finish(long long* bit_off, long long ng, const long long* ng_dev, const ull* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    bit_off[ng] = (long long)*total;        // capacity slot: where the host finds the total before it knows the group count
    bit_off[ng_dev ? *ng_dev : ng] = (long long)*total;
  }
}
__global__ void __launch_bounds__(256) k_bswap_words(uint32_t* __restrict__ words, long long nw) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) words[i] = __byte_perm(words[i], 0u, 0x0123);
}

__global__ void __launch_bounds__(256) k_fixed_pack(const uint32_t* __restrict__ index, long long count, int b,
                                                    uint32_t* __restrict__ words) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  ull pattern = (ull)(__brev(index[i]) >> (32 - b));  // LSB first (BIO:41-53)
  put_bits(words, (ull)i * (ull)b, pattern, b);
}

__global__ void __launch_bounds__(256) k_fixed_unpack(const uint8_t* __restrict__ bits, long long count, int b,
                                                      uint32_t* __restrict__ index) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  ull p = (ull)i * (ull)b;
  uint32_t v = 0;
  for (int t = 0; t < b; ++t, ++p) v |= (uint32_t)((bits[p >> 3] >> (7 - (int)(p & 7))) & 1u) << t;
  index[i] = v;
}

// packed MSB-first bits -> the reference's '0'/'1' characters (what BIO:76-79 consumes), 8 characters per thread.
// `ascii` holds 8 * ceil(n_bits / 8) bytes: the last thread may write up to 7 characters past n_bits.
// 64-thread CTAs (18 registers per thread): one fits beside the three resident k_is_score CTAs of an SM (61 440 of
// 65 536 registers taken), so the expansion of one chunk can run on the download stream while the next chunk is scored.
constexpr int ASCII_THREADS = 64;
__global__ void __launch_bounds__(ASCII_THREADS) k_bits_to_ascii(const uint8_t* __restrict__ bits, long long n_bytes,
                                                       const long long* __restrict__ n_bits_dev, uint2* __restrict__ ascii) {
  pdl_wait();
  pdl_trigger();
  if (n_bits_dev) { const long long nb = (*n_bits_dev + 7) >> 3; n_bytes = nb < n_bytes ? nb : n_bytes; }
  // grid-stride: with the count on the device the host sizes the launch by the capacity, usually many times the code
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += stride) {
    const uint32_t b = bits[i];
    uint32_t lo = 0x30303030u, hi = 0x30303030u;   // characters 0-3 (bits 7..4) and 4-7 (bits 3..0), little-endian words
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      lo |= ((b >> (7 - c)) & 1u) << (8 * c);
      hi |= ((b >> (3 - c)) & 1u) << (8 * c);
    }
    ascii[i] = make_uint2(lo, hi);
  }
}

// ------------------------------------------------------------------------------------------
// the candidate stream, materialised (debug / oracle injection)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_draw(uint32_t sseed, uint32_t k0, uint32_t nk, uint32_t G, int raw, float* __restrict__ out) {
  const uint32_t nq = (G + 3) >> 2;
  ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (ull)nk * nq) return;
  uint32_t r = (uint32_t)(t / nq), q = (uint32_t)(t - (ull)r * nq);
  float z[4];
  if (raw) raw_quad(sseed, k0 + r, q, z); else normal_quad(sseed, k0 + r, q, z);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (4 * q + j < G) out[(ull)r * G + 4 * q + j] = z[j];
}

__global__ void __launch_bounds__(256) k_stream_words(uint32_t sseed, uint32_t k0, uint32_t nk, uint32_t nq, uint32_t* __restrict__ out) {
  ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (ull)nk * nq) return;
  uint32_t r = (uint32_t)(t / nq), q = (uint32_t)(t - (ull)r * nq);
  uint32_t o[4];
  stream_block(sseed, k0 + r, q, o);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[t * 4 + j] = o[j];
}

__global__ void __launch_bounds__(256) k_philox(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key,
                                                uint32_t* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t o[4];
  philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[4 * i + j] = o[j];
}

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// optional timing of the dominant kernel (k_is_score) with CUDA events on the launching stream
constexpr int PROFILE_RING = 1024;
static int g_profile = 0;        // 0 off, 1 scoring kernel only, 2 every kernel of the fused pipeline
static cudaEvent_t g_ev0[PROFILE_RING], g_ev1[PROFILE_RING];
static int g_kind[PROFILE_RING];
static int g_profile_count = 0;
static bool g_profile_events_ready = false;

int trace_begin(int kind, cudaStream_t st) {
  if (!g_profile || (g_profile == 1 && kind != TR_SCORE) || g_profile_count >= PROFILE_RING) return -1;
  const int slot = g_profile_count++;
  g_kind[slot] = kind;
  cudaEventRecord(g_ev0[slot], st);
  return slot;
}
void trace_end(int slot, cudaStream_t st) {
  if (slot >= 0) cudaEventRecord(g_ev1[slot], st);
}

static int g_sm_count = 0, g_is_blocks_per_sm = 0, g_is_blocks_per_sm_any = 0;
static int ensure_device_props() {
  if (g_sm_count > 0) return CWQ_OK;
  int dev = 0;
  CWQ_CUDA(cudaGetDevice(&dev));
  int sms = 0, bps = 0;
  CWQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CWQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_is_score<true, false>, IS_THREADS, 0));
  g_is_blocks_per_sm = bps > 0 ? bps : 1;
  CWQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_is_score<false, false>, IS_THREADS, 0));
  g_is_blocks_per_sm_any = bps > 0 ? bps : 1;
  if (const char* e = getenv("CWQ_IS_BLOCKS_PER_SM")) {   // tuning knob: fewer resident CTAs than the occupancy limit
    int v = atoi(e);
    if (v >= 1 && v <= g_is_blocks_per_sm) g_is_blocks_per_sm = v;
  }
  g_sm_count = sms;
  return CWQ_OK;
}

}  // namespace cwq

using namespace cwq;

extern "C" int cwq_abi_version(void) { return CWQ_ABI_VERSION; }
extern "C" const char* cwq_last_error(void) { return g_last_error.c_str(); }

extern "C" int cwq_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* clock_khz) {
  int v = 0;
  if (sm_count) { CWQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm_count = v; }
  if (cc_major) { CWQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device)); *cc_major = v; }
  if (cc_minor) { CWQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device)); *cc_minor = v; }
  if (clock_khz) { CWQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); *clock_khz = v; }
  return CWQ_OK;
}

extern "C" long long cwq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int cwq_profile_enable(int on) {
  if (on && !g_profile_events_ready) {
    for (int i = 0; i < PROFILE_RING; ++i) { CWQ_CUDA(cudaEventCreate(&g_ev0[i])); CWQ_CUDA(cudaEventCreate(&g_ev1[i])); }
    g_profile_events_ready = true;
  }
  g_profile = on < 0 ? 0 : (on > 2 ? 2 : on);
  g_profile_count = 0;
  return CWQ_OK;
}
extern "C" int cwq_profile_count(void) { return g_profile_count; }
extern "C" int cwq_profile_kind(int i) { return (i >= 0 && i < g_profile_count) ? g_kind[i] : -1; }
extern "C" int cwq_profile_get_ms(int i, float* ms) {
  CWQ_REQUIRE(i >= 0 && i < g_profile_count && ms, "no such profile record");
  CWQ_CUDA(cudaEventElapsedTime(ms, g_ev0[i], g_ev1[i]));
  return CWQ_OK;
}

extern "C" size_t cwq_is_encode_workspace_bytes(int64_t n_groups) {
  if (n_groups < 0) n_groups = 0;
  size_t b = 0;
  b += align_up(((size_t)n_groups + 1) * 8);             // unit_off
  b += align_up((size_t)n_groups * 8);                   // keys
  b += align_up((size_t)scan_tiles(n_groups + 1) * 8);   // scan tile sums
  return b + 1024;
}

// groups [g0, g1) of arrays holding n_groups entries; range ends on the device where given (see GroupRange)
// `front` (may be null): the fused front end already holds the unit prefix over all groups, initialised keys and a counter
static int is_encode_impl(const float* coef, const int64_t* starts, const int32_t* K, const uint32_t* group_seed, int64_t n_groups,
                          GroupRange range, int64_t max_group_size, uint32_t* index, void* workspace, size_t workspace_bytes,
                          void* stream, const FrontWorkspace* front = nullptr, bool low_k = false) {
  int rc = ensure_device_props();
  if (rc != CWQ_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // scans and the index kernel cover the host-known length of the range, or the whole capacity when an end is on the device
  const long long span = (range.g0_dev || range.g1_dev) ? n_groups : range.g1 - range.g0;
  ull *unit_off, *keys, *counter;
  if (front) {
    unit_off = front->unit_off;
    keys = front->keys;
    counter = front->scalars + FRONT_SCORE_COUNTER;      // zero: k_front_chain cleared it, k_is_index clears it after every scoring
  } else {
    Workspace ws(workspace, workspace_bytes);
    unit_off = ws.take<ull>((size_t)span + 1);
    keys = ws.take<ull>((size_t)n_groups);
    ull* tile_sums = ws.take<ull>((size_t)scan_tiles(span));
    ull* total = ws.take<ull>(1);
    counter = ws.take<ull>(1);
    if (!unit_off || !keys || !tile_sums || !total || !counter) { set_error("cwq_is_encode: workspace too small"); return CWQ_ERR_WORKSPACE; }
    UnitCount in{K, range};
    UnitDeliver out{unit_off, keys, range};
    CWQ_CUDA((device_exclusive_scan<ull, UnitCount, UnitDeliver>(in, out, span, tile_sums, total, st)));
    k_units_finish<<<1, 32, 0, st>>>(unit_off, range, total, counter);
    CWQ_LAUNCH_CHECK();
  }
  const int is_global = front ? 1 : 0;
  const int slot = trace_begin(TR_SCORE, st);
  const float2* cf = (const float2*)coef;
  const long long* stt = (const long long*)starts;
  const dim3 blk(IS_THREADS);
  const int ug = is_global;
  cudaError_t le;
  if (max_group_size <= 16) {
    const dim3 grid(g_sm_count * g_is_blocks_per_sm);
    le = low_k ? launch_chain(k_is_score<true, true>, grid, blk, 0, st, cf, stt, K, group_seed, range, (const ull*)unit_off, ug, counter, keys)
               : launch_chain(k_is_score<true, false>, grid, blk, 0, st, cf, stt, K, group_seed, range, (const ull*)unit_off, ug, counter, keys);
  } else {
    const dim3 grid(g_sm_count * g_is_blocks_per_sm_any);
    le = low_k ? launch_chain(k_is_score<false, true>, grid, blk, 0, st, cf, stt, K, group_seed, range, (const ull*)unit_off, ug, counter, keys)
               : launch_chain(k_is_score<false, false>, grid, blk, 0, st, cf, stt, K, group_seed, range, (const ull*)unit_off, ug, counter, keys);
  }
  CWQ_CUDA(le);
  CWQ_LAUNCH_CHECK();
  trace_end(slot, st);
  const int slot_i = trace_begin(TR_INDEX, st);
  const long long index_blocks = (span + 255) / 256, index_cap = 4LL * g_sm_count;
  CWQ_CUDA(launch_chain(k_is_index, dim3((unsigned)(index_blocks < index_cap ? index_blocks : index_cap)), dim3(256), 0, st,
                        (const ull*)keys, range, index, counter));
  CWQ_LAUNCH_CHECK();
  trace_end(slot_i, st);
  return CWQ_OK;
}

extern "C" int cwq_is_encode(const float* coef, const int64_t* starts, const int32_t* K, const uint32_t* group_seed,
                             int64_t n_groups, const int64_t* n_groups_dev, int64_t g0, int64_t g1, int64_t max_group_size,
                             uint32_t* index, void* workspace, size_t workspace_bytes, void* stream) {
  CWQ_REQUIRE(max_group_size >= 0, "max_group_size must be an upper bound of the group sizes");
  CWQ_REQUIRE(n_groups > 0 && g0 >= 0 && g1 <= n_groups && g0 <= g1, "bad group range");
  CWQ_REQUIRE(!n_groups_dev || (g0 == 0 && g1 == n_groups), "a device-side group count needs the full range [0, capacity)");
  if (g0 == g1) return CWQ_OK;
  GroupRange range{g0, g1, nullptr, (const long long*)n_groups_dev};
  return is_encode_impl(coef, starts, K, group_seed, n_groups, range, max_group_size, index, workspace, workspace_bytes, stream);
}

extern "C" int cwq_is_encode_device_range(const float* coef, const int64_t* starts, const int32_t* K, const uint32_t* group_seed,
                                          int64_t n_groups, const int64_t* range_dev, int64_t max_group_size, uint32_t* index,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  CWQ_REQUIRE(max_group_size >= 0, "max_group_size must be an upper bound of the group sizes");
  CWQ_REQUIRE(n_groups > 0 && range_dev, "bad arguments");
  GroupRange range{0, n_groups, (const long long*)range_dev, (const long long*)range_dev + 1};
  return is_encode_impl(coef, starts, K, group_seed, n_groups, range, max_group_size, index, workspace, workspace_bytes, stream);
}

extern "C" size_t cwq_shard_ranges_workspace_bytes(int64_t n_groups) {
  if (n_groups <= 0) return 0;
  return align_up((size_t)scan_tiles(n_groups) * sizeof(ull)) + align_up(sizeof(ull)) + 256;
}

extern "C" int cwq_is_shard_ranges(const int32_t* K, const int64_t* starts, int64_t n_groups, const int64_t* n_groups_dev,
                                   int32_t world, int64_t* bounds, void* workspace, size_t workspace_bytes, void* stream) {
  CWQ_REQUIRE(K && starts && n_groups_dev && bounds && n_groups > 0 && world >= 1 && world <= 1024, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  ull* tile_sums = ws.take<ull>((size_t)scan_tiles(n_groups));
  ull* total = ws.take<ull>(1);
  if (!tile_sums || !total) { set_error("cwq_is_shard_ranges: workspace too small"); return CWQ_ERR_WORKSPACE; }
  k_shard_bounds_init<<<1, 1024, 0, st>>>((long long*)bounds, world, (const long long*)n_groups_dev);
  CWQ_LAUNCH_CHECK();
  ShardWork in{K, (const long long*)starts, (const long long*)n_groups_dev};
  ShardCuts out{total, world, (long long*)bounds};
  CWQ_CUDA((device_exclusive_scan<ull, ShardWork, ShardCuts>(in, out, n_groups, tile_sums, total, st)));
  return CWQ_OK;
}

extern "C" int cwq_is_decode(const int64_t* starts, const uint32_t* group_seed, const int32_t* gid, const uint32_t* index,
                             int64_t n, const float* p_loc, const float* p_scale, float* sample, void* stream) {
  CWQ_REQUIRE(n > 0, "n must be positive");
  k_is_decode<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const long long*)starts, group_seed, gid, index, n,
                                                                            p_loc, p_scale, sample);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" size_t cwq_outlier_workspace_bytes(int64_t n) {
  if (n < 0) n = 0;
  return align_up((size_t)scan_tiles(n) * 4) + 4096;  // tile sums + scalars
}

extern "C" int cwq_outlier_encode(const float* q_loc, const float* q_scale, const uint8_t* mask, int64_t n,
                                  const cwq_segment* segments, int32_t n_segments, float* sample, int64_t* outlier_idx,
                                  uint16_t* outlier_q, int64_t* seg_outlier_off, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  CWQ_REQUIRE(n > 0 && n_segments > 0 && n_segments <= 1 << 20, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  uint32_t* tile_sums = ws.take<uint32_t>((size_t)scan_tiles(n));
  uint32_t* total = ws.take<uint32_t>(1);
  if (!tile_sums || !total) { set_error("cwq_outlier_encode: workspace too small"); return CWQ_ERR_WORKSPACE; }
  OutlierCount in{mask};
  OutlierDeliver out{mask, q_loc, q_scale, segments, n_segments, sample, (long long*)outlier_idx, outlier_q,
                     (long long*)seg_outlier_off};
  CWQ_CUDA((device_exclusive_scan<uint32_t, OutlierCount, OutlierDeliver>(in, out, n, tile_sums, total, st)));
  k_outlier_finish<<<1, 32, 0, st>>>(total, n_segments, (long long*)seg_outlier_off);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_outlier_decode(float* sample, const int64_t* outlier_pos, const uint16_t* outlier_q, int64_t m,
                                  void* stream) {
  if (m <= 0) return CWQ_OK;
  k_outlier_decode<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sample, (const long long*)outlier_pos, outlier_q, m);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" size_t cwq_greedy_workspace_bytes(int64_t n, int64_t n_groups, int32_t n_steps) {
  if (n < 0) n = 0;
  if (n_groups < 0) n_groups = 0;
  if (n_steps < 0) n_steps = 0;
  return align_up((size_t)n * 8) + align_up((size_t)n_groups * 8) + align_up((size_t)n_steps * 4) + 1024;
}

// mode: 0 = automatic, 1 = one persistent CTA per group, 2 = lockstep units (few groups)
extern "C" int cwq_greedy_encode(const float* tparam, const int64_t* starts, const uint32_t* group_seed, int64_t n_groups,
                                 int64_t g0, int64_t g1, int32_t n_steps, int32_t n_bits_per_step, float shard_scale,
                                 const float* shard_loc, int64_t max_group_size, int64_t n, int32_t mode, uint32_t* indices, float* best,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  CWQ_REQUIRE(n_groups > 0 && g0 >= 0 && g1 <= n_groups && g0 <= g1, "bad group range");
  CWQ_REQUIRE(n_steps > 0 && n_bits_per_step >= 0 && n_bits_per_step <= 30, "bad step parameters");
  CWQ_REQUIRE(max_group_size >= 0 && mode >= 0 && mode <= 2, "bad max_group_size or mode");
  if (g0 == g1) return CWQ_OK;
  int rc = ensure_device_props();
  if (rc != CWQ_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  size_t gp = (size_t)((max_group_size + 3) / 4 * 4);
  const long long ng = g1 - g0;
  if (mode == 0) mode = ng < 40LL * g_sm_count ? 2 : 1;   // measured: lockstep units are never slower up to 7.5k groups, 23 % faster at 360
  if (mode == 1) {
    size_t smem = gp * (sizeof(float2) + sizeof(float)) + gp / 4 * sizeof(uint4);
    CWQ_REQUIRE(smem <= 200 * 1024, "group too large for shared memory (max_group_size_bits too high)");
    if (smem > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_greedy_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_greedy_encode<<<(unsigned)ng, GREEDY_THREADS, smem, st>>>(
        (const float2*)tparam, (const long long*)starts, group_seed, g0, n_steps, 1u << n_bits_per_step, shard_scale, shard_loc,
        indices, best);
    CWQ_LAUNCH_CHECK();
    return CWQ_OK;
  }
  size_t smem = gp * sizeof(float2) + gp / 4 * sizeof(uint4);
  CWQ_REQUIRE(smem <= 200 * 1024, "group too large for shared memory (max_group_size_bits too high)");
  Workspace ws(workspace, workspace_bytes);
  float2* ce = ws.take<float2>((size_t)n);
  ull* keys = ws.take<ull>((size_t)ng);
  unsigned int* counters = ws.take<unsigned int>((size_t)n_steps);
  if (!ce || !keys || !counters) { set_error("cwq_greedy_encode: workspace too small"); return CWQ_ERR_WORKSPACE; }
  if (smem > 48 * 1024) CWQ_CUDA(cudaFuncSetAttribute(k_greedy_lockstep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bps = 0;
  CWQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_greedy_lockstep, GREEDY_THREADS, smem));
  CWQ_REQUIRE(bps >= 1, "lockstep greedy kernel does not fit on an SM");
  CWQ_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * (size_t)n_steps, st));
  const float2* a_tparam = (const float2*)tparam;
  const long long* a_starts = (const long long*)starts;
  long long a_g0 = g0, a_ng = ng;
  uint32_t a_ns = 1u << n_bits_per_step;
  void* args[] = {&a_tparam, &a_starts, &group_seed, &a_g0, &a_ng, &n_steps, &a_ns, &shard_scale, &shard_loc, &best, &ce, &keys,
                  &counters, &indices};
  CWQ_CUDA(cudaLaunchCooperativeKernel((const void*)k_greedy_lockstep, dim3((unsigned)(g_sm_count * bps)), dim3(GREEDY_THREADS), args, smem, st));
  count_launches(1);
  return CWQ_OK;
}

extern "C" int cwq_greedy_decode(const int64_t* starts, const uint32_t* group_seed, const int32_t* gid,
                                 const uint32_t* indices, int64_t n, int32_t n_steps, float shard_scale,
                                 const float* shard_loc, const float* p_loc, const float* p_scale, float* sample, void* stream) {
  CWQ_REQUIRE(n > 0 && n_steps > 0, "bad sizes");
  k_greedy_decode<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const long long*)starts, group_seed, gid, indices,
                                                                                n, n_steps, shard_scale, shard_loc, p_loc, p_scale,
                                                                                sample);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_greedy_shard_loc(const float* p_loc, const float* p_scale, int64_t n, int32_t n_steps, float* shard_loc,
                                    void* stream) {
  CWQ_REQUIRE(n > 0 && n_steps > 0, "bad sizes");
  k_greedy_shard_loc<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_loc, p_scale, n, (float)n_steps, shard_loc);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_unrescale(const float* best, const float* p_loc, const float* p_scale, int64_t n, float* sample,
                             void* stream) {
  CWQ_REQUIRE(n > 0, "n must be positive");
  k_unrescale<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(best, p_loc, p_scale, n, sample);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" size_t cwq_pack_workspace_bytes(int64_t n_groups) {
  if (n_groups < 0) n_groups = 0;
  return align_up((size_t)scan_tiles(n_groups) * 8) + 1024;
}

extern "C" int cwq_elias_delta_pack(const uint32_t* index, int64_t n_groups, const int64_t* n_groups_dev, uint8_t* bits,
                                    size_t bits_capacity_bytes, int64_t* bit_off, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  CWQ_REQUIRE(n_groups > 0, "n_groups must be positive");
  CWQ_REQUIRE(bits_capacity_bytes % 4 == 0 && bits_capacity_bytes >= ((size_t)n_groups * 42 + 31) / 32 * 4,
              "bits capacity must be a multiple of 4 and hold 42 bits per group");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  ull* tile_sums = ws.take<ull>((size_t)scan_tiles(n_groups));
  ull* total = ws.take<ull>(1);
  if (!tile_sums || !total) { set_error("cwq_elias_delta_pack: workspace too small"); return CWQ_ERR_WORKSPACE; }
  CWQ_CUDA(cudaMemsetAsync(bits, 0, bits_capacity_bytes, st));
  const long long* ngd = (const long long*)n_groups_dev;
  EliasLen in{index, ngd};
  EliasDeliver out{index, (long long*)bit_off, (uint32_t*)bits, ngd};
  CWQ_CUDA((device_exclusive_scan<ull, EliasLen, EliasDeliver>(in, out, n_groups, tile_sums, total, st)));
  k_bits_finish<<<1, 32, 0, st>>>((long long*)bit_off, n_groups, ngd, total);
  CWQ_LAUNCH_CHECK();
  long long nw = (long long)(bits_capacity_bytes / 4);
  k_bswap_words<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>((uint32_t*)bits, nw);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" size_t cwq_is_pipeline_workspace_bytes(int64_t n, int32_t n_segments, int64_t max_group_size) {
  if (n <= 0 || n_segments <= 0) return 0;
  const int64_t cap = n + n_segments;
  size_t w = cwq_group_workspace_bytes(n, n_segments);
  const size_t others[4] = {cwq_is_encode_workspace_bytes(cap), cwq_outlier_workspace_bytes(n), cwq_pack_workspace_bytes(cap),
                            cwq_shard_ranges_workspace_bytes(cap)};
  for (size_t o : others) w = o > w ? o : w;
  FrontPlan plan;
  if (front_plan(n, max_group_size, &plan)) {
    const size_t f = front_workspace_bytes(n, n_segments, cap, plan);
    w = f > w ? f : w;
  }
  return w;
}

namespace cwq {
// bounds[r] = one past the first group whose inclusive work prefix reaches r / world of the total (same cut rule as
// ShardCuts), from the exclusive prefix the fused front end wrote: one binary search per rank
__global__ void k_shard_bounds_from_prefix(const ull* __restrict__ work_off, const long long* __restrict__ ng_dev, int world,
                                           long long* __restrict__ bounds) {
  const int r = threadIdx.x;
  if (r > world) return;
  const long long ng = *ng_dev;
  if (r == 0) { bounds[0] = 0; return; }
  if (r == world) { bounds[world] = ng; return; }
  const ull tot = work_off[ng];
  const ull target = tot * (ull)r;
  long long lo = 0, hi = ng - 1;           // smallest j with work_off[j + 1] * world >= target
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (work_off[mid + 1] * (ull)world >= target) hi = mid; else lo = mid + 1;
  }
  bounds[r] = lo + 1;
}

bool pdl_enabled() {
  const char* e = getenv("CWQ_PDL");
  return !e || atoi(e) != 0;
}

static bool g_front_enabled = true;
}  // namespace cwq

extern "C" int cwq_set_fused_front(int on) {
  const int was = g_front_enabled ? 1 : 0;
  g_front_enabled = on != 0;
  return was;
}

extern "C" int cwq_is_pipeline_shard_ranges(int64_t n, int32_t n_segments, int64_t max_group_size, int32_t world, int64_t* bounds,
                                            const cwq_is_buffers* b, void* stream) {
  CWQ_REQUIRE(b && bounds && n > 0 && n_segments > 0 && world >= 1 && world <= 1024, "bad arguments");
  const int64_t cap = n + n_segments;
  const int64_t* ng_dev = b->seg_group_off + n_segments;
  FrontPlan plan;
  if (g_front_enabled && front_plan(n, max_group_size, &plan)) {
    FrontWorkspace fw;
    if (!front_carve(b->workspace, b->workspace_bytes, n, n_segments, cap, plan, &fw)) {
      set_error("cwq_is_pipeline_shard_ranges: workspace too small");
      return CWQ_ERR_WORKSPACE;
    }
    k_shard_bounds_from_prefix<<<1, 1056, 0, (cudaStream_t)stream>>>(fw.work_off, (const long long*)ng_dev, world, (long long*)bounds);
    CWQ_LAUNCH_CHECK();
    return CWQ_OK;
  }
  return CWQ_ERR_ARG;
}

extern "C" int cwq_is_encode_pipeline(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                                      int64_t n, const cwq_segment* segments, const cwq_segment* h_segments, int32_t n_segments,
                                      int64_t max_group_size, int32_t stages, const cwq_is_buffers* b, void* stream) {
  CWQ_REQUIRE(b && n > 0 && n_segments > 0, "bad arguments");
  CWQ_REQUIRE((stages & ~(CWQ_STAGE_GROUP | CWQ_STAGE_SCORE | CWQ_STAGE_FINISH)) == 0 && stages != 0, "unknown stage");
  const int64_t cap = n + n_segments;
  const int64_t* ng_dev = b->seg_group_off + n_segments;
  int rc = CWQ_OK;
  FrontPlan plan;
  if (g_front_enabled && front_plan(n, max_group_size, &plan)) {
    // the fused front end (cwq_front.cu): 3 kernels for GROUP, the scoring kernel on the prefix they leave, 2 for FINISH
    FrontWorkspace fw;
    if (!front_carve(b->workspace, b->workspace_bytes, n, n_segments, cap, plan, &fw)) {
      set_error("cwq_is_encode_pipeline: workspace too small");
      return CWQ_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool clear_first = (stages & CWQ_STAGE_GROUP) && (stages & CWQ_STAGE_FINISH);
    if (clear_first) {
      rc = front_finish_clear(plan, fw, b, st);
      if (rc != CWQ_OK) return rc;
    }
    if (stages & CWQ_STAGE_GROUP) {
      rc = front_group(q_loc, q_scale, p_loc, p_scale, n, segments, n_segments, plan, fw, b, st);
      if (rc != CWQ_OK) return rc;
    }
    if (stages & CWQ_STAGE_SCORE) {
      GroupRange range{0, cap, nullptr, (const long long*)ng_dev};
      if (b->score_range_dev) range = GroupRange{0, cap, (const long long*)b->score_range_dev, (const long long*)b->score_range_dev + 1};
      // small budgets everywhere in the batch -> the low-K instantiation of the scoring kernel (same results)
      bool low_k = h_segments != nullptr;
      for (int32_t i = 0; low_k && i < n_segments; ++i) low_k = h_segments[i].n_nats_per_group <= LOW_K_MAX_NATS;
      rc = is_encode_impl(b->coef, b->starts, b->K, b->group_seed, cap, range, max_group_size, b->index, nullptr, 0, stream, &fw,
                          low_k);
      if (rc != CWQ_OK) return rc;
    }
    if (stages & CWQ_STAGE_FINISH) {
      if (!clear_first) {
        rc = front_finish_clear(plan, fw, b, st);
        if (rc != CWQ_OK) return rc;
      }
      rc = front_finish(q_loc, q_scale, p_loc, p_scale, n, segments, n_segments, cap, plan, fw, b, st);
      if (rc != CWQ_OK) return rc;
    }
    return CWQ_OK;
  }
  if (stages & CWQ_STAGE_GROUP) {
    rc = cwq_is_prologue(q_loc, q_scale, p_loc, p_scale, n, segments, n_segments, b->kl, b->coef, b->mask, stream);
    if (rc != CWQ_OK) return rc;
    rc = cwq_group(b->kl, n, segments, h_segments, n_segments, b->starts, b->group_seed, b->gid, b->seg_group_off,
                   b->workspace, b->workspace_bytes, stream);
    if (rc != CWQ_OK) return rc;
    rc = cwq_group_totals(b->kl, b->starts, cap, ng_dev, b->total_kl, b->K, stream);
    if (rc != CWQ_OK) return rc;
  }
  if (stages & CWQ_STAGE_SCORE) {
    if (b->score_range_dev)
      rc = cwq_is_encode_device_range(b->coef, b->starts, b->K, b->group_seed, cap, b->score_range_dev, max_group_size, b->index,
                                      b->workspace, b->workspace_bytes, stream);
    else
      rc = cwq_is_encode(b->coef, b->starts, b->K, b->group_seed, cap, ng_dev, 0, cap, max_group_size, b->index, b->workspace,
                         b->workspace_bytes, stream);
    if (rc != CWQ_OK) return rc;
  }
  if (stages & CWQ_STAGE_FINISH) {
    rc = cwq_is_decode(b->starts, b->group_seed, b->gid, b->index, n, p_loc, p_scale, b->sample, stream);
    if (rc != CWQ_OK) return rc;
    rc = cwq_outlier_encode(q_loc, q_scale, b->mask, n, segments, n_segments, b->sample, b->outlier_idx, b->outlier_q,
                            b->seg_outlier_off, b->workspace, b->workspace_bytes, stream);
    if (rc != CWQ_OK) return rc;
    rc = cwq_elias_delta_pack(b->index, cap, ng_dev, b->bits, b->bits_capacity_bytes, b->bit_off, b->workspace,
                              b->workspace_bytes, stream);
    if (rc != CWQ_OK) return rc;
  }
  return CWQ_OK;
}

extern "C" int cwq_fixed_pack(const uint32_t* index, int64_t count, int32_t bits_per_index, uint8_t* bits,
                              size_t bits_capacity_bytes, void* stream) {
  CWQ_REQUIRE(count > 0 && bits_per_index >= 1 && bits_per_index <= 32, "bad sizes");
  CWQ_REQUIRE(bits_capacity_bytes % 4 == 0 && bits_capacity_bytes * 8 >= (size_t)count * (size_t)bits_per_index,
              "bits capacity must be a multiple of 4 and hold count * bits_per_index bits");
  cudaStream_t st = (cudaStream_t)stream;
  CWQ_CUDA(cudaMemsetAsync(bits, 0, bits_capacity_bytes, st));
  k_fixed_pack<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(index, count, bits_per_index, (uint32_t*)bits);
  CWQ_LAUNCH_CHECK();
  long long nw = (long long)(bits_capacity_bytes / 4);
  k_bswap_words<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>((uint32_t*)bits, nw);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_fixed_unpack(const uint8_t* bits, int64_t count, int32_t bits_per_index, uint32_t* index, void* stream) {
  CWQ_REQUIRE(count > 0 && bits_per_index >= 1 && bits_per_index <= 32, "bad sizes");
  k_fixed_unpack<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, count, bits_per_index, index);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_bits_to_ascii(const uint8_t* bits, int64_t n_bits, const int64_t* n_bits_dev, uint8_t* ascii, void* stream) {
  CWQ_REQUIRE(n_bits >= 0 && (n_bits == 0 || (bits && ascii)), "bad arguments");
  CWQ_REQUIRE((reinterpret_cast<uintptr_t>(ascii) & 7u) == 0, "ascii must be 8-byte aligned");
  if (n_bits == 0) return CWQ_OK;
  const long long n_bytes = (n_bits + 7) / 8;
  int rc = ensure_device_props();
  if (rc != CWQ_OK) return rc;
  const long long blocks = (n_bytes + ASCII_THREADS - 1) / ASCII_THREADS, most = 32LL * g_sm_count;
  CWQ_CUDA(launch_chain(k_bits_to_ascii, dim3((unsigned)(blocks < most ? blocks : most)), dim3(ASCII_THREADS), 0, (cudaStream_t)stream,
                        bits, n_bytes, (const long long*)n_bits_dev, reinterpret_cast<uint2*>(ascii)));
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_elias_delta_parse_host(const uint8_t* h_ascii_bits, int64_t n_bits, int64_t n_codes, uint32_t* h_index,
                                          int64_t* h_bits_consumed) {
  CWQ_REQUIRE(h_ascii_bits && h_index && n_bits >= 0 && n_codes >= 0, "bad arguments");
  int64_t p = 0;
  for (int64_t c = 0; c < n_codes; ++c) {
    int l = 0;
    while (p < n_bits && h_ascii_bits[p] == '0') { ++l; ++p; }   // BIO:28-29
    CWQ_REQUIRE(l <= 5 && p + l + 1 <= n_bits, "truncated or malformed Elias-delta stream");
    uint32_t n_plus_one = 0;
    for (int t = 0; t <= l; ++t) n_plus_one = (n_plus_one << 1) | (uint32_t)(h_ascii_bits[p++] == '1');   // BIO:33
    CWQ_REQUIRE(n_plus_one >= 1 && n_plus_one <= 32 && p + (int64_t)n_plus_one - 1 <= n_bits, "truncated or malformed Elias-delta stream");
    uint64_t x = 1;
    for (uint32_t t = 0; t + 1 < n_plus_one; ++t) x = (x << 1) | (uint64_t)(h_ascii_bits[p++] == '1');     // BIO:37
    h_index[c] = (uint32_t)(x - 1);
  }
  if (h_bits_consumed) *h_bits_consumed = p;
  return CWQ_OK;
}

extern "C" int cwq_draw_candidates(uint32_t stream_seed, uint32_t k0, uint32_t nk, uint32_t G, int32_t raw, float* out, void* stream) {
  if (nk == 0 || G == 0) return CWQ_OK;
  ull t = (ull)nk * ((G + 3) / 4);
  CWQ_REQUIRE(t < (1ull << 39), "too many draws for one call");
  k_draw<<<(unsigned)((t + 255) / 256), 256, 0, (cudaStream_t)stream>>>(stream_seed, k0, nk, G, raw, out);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_stream_words(uint32_t stream_seed, uint32_t k0, uint32_t nk, uint32_t nq, uint32_t* out, void* stream) {
  if (nk == 0 || nq == 0) return CWQ_OK;
  ull t = (ull)nk * nq;
  CWQ_REQUIRE(t < (1ull << 39), "too many blocks for one call");
  k_stream_words<<<(unsigned)((t + 255) / 256), 256, 0, (cudaStream_t)stream>>>(stream_seed, k0, nk, nq, out);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}

extern "C" int cwq_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int64_t n, void* stream) {
  if (n <= 0) return CWQ_OK;
  k_philox<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctr, key, out, n);
  CWQ_LAUNCH_CHECK();
  return CWQ_OK;
}
