/*
 * cwq.h -- C ABI of libcwq.so: the B200-native MIRACLE relative-entropy coding engine.
 *
 * This is the drop-in boundary for the hot path of gergely-flamich/compression-without-quantization
 * (code/coded_importance_sampler.py, code/coded_greedy_sampler.py, code/misc.py and the bit helpers
 * of code/binary_io.py).  The reference has no FFI of its own: its boundary is eight Python
 * callables.  The Python package compression_without_quantization_b200 keeps
 * those names and signatures and is a thin ctypes layer over the functions below; INTEGRATION.md
 * shows the binding.  Each entry point cites the reference lines it replaces
 * (CIS = code/coded_importance_sampler.py, CGS = code/coded_greedy_sampler.py, BIO = code/binary_io.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (caller-owned, e.g. obtained through DLPack) unless its name
 *     starts with h_; `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - return value: 0 = ok, negative = error (cwq_last_error() has the text).  There is no CPU
 *     fallback: without a CUDA device every compute entry point fails with CWQ_ERR_CUDA.
 *   - no hidden allocation: scratch is passed in as `workspace` (size from the *_workspace_bytes
 *     queries, 256-byte aligned).
 *   - a "segment" is one reference call: one flat latent vector with its own grouping parameters
 *     and seed.  Passing many segments batches many reference calls (e.g. both PLN levels of 24
 *     images) into one launch sequence; results per segment are identical to separate calls.
 */
#ifndef CWQ_H_
#define CWQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CWQ_ABI_VERSION 3

#if defined(__GNUC__)
#define CWQ_API __attribute__((visibility("default")))
#else
#define CWQ_API
#endif

#define CWQ_OK 0
#define CWQ_ERR_ARG (-1)
#define CWQ_ERR_CUDA (-2)
#define CWQ_ERR_WORKSPACE (-3)

typedef struct cwq_segment {
  int64_t start;            /* first dim of the segment in the flat arrays */
  int64_t end;              /* one past the last dim; segments are contiguous and ascending */
  double n_nats_per_group;  /* n_bits_per_group * ln 2 - 1            (CIS:181, CGS:226) */
  int64_t size_limit;       /* a boundary is forced once the running group holds this many dims:
                               the integer restatement of log2(size+1) >(=) max_group_size_bits (CIS:185-187) */
  float dim_kl_bit_limit;   /* IS only: dims above it are outliers    (CIS:142-148) */
  uint32_t seed;            /* group i of the segment uses seed + i   (CIS:243, CGS:282) */
  int32_t strict;           /* 1: '>' comparisons (CIS:187-188); 0: '>=' (CGS:232-233) */
  int32_t reserved;
} cwq_segment;

CWQ_API int cwq_abi_version(void);
CWQ_API const char* cwq_last_error(void);
/* device 0..; any out pointer may be NULL */
CWQ_API int cwq_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* clock_khz);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
CWQ_API long long cwq_launch_count(void);
/* CUDA-event timing of the dominant kernel (the importance-sampler scoring kernel) on its launching
 * stream: enable, run, synchronise, then read up to 256 records. */
CWQ_API int cwq_profile_enable(int on);   /* 0 off, 1 the scoring kernel, 2 every kernel of cwq_is_encode_pipeline */
CWQ_API int cwq_profile_count(void);
/* what record i timed: 0 scoring kernel, 1 tiles (prologue + boundaries), 2 chain, 3 groups, 4 unit/work prefix, 5 index
 * extraction, 6 decode + outliers, 7 Elias-delta packing; -1 no such record */
CWQ_API int cwq_profile_kind(int i);
CWQ_API int cwq_profile_get_ms(int i, float* ms);

/* ---- per-dim prologue ------------------------------------------------------------------------
 * CIS:133-163: rescale q by p, split out dims whose KL exceeds dim_kl_bit_limit, KL of the rest.
 * kl[n]; coef[2n] = (a, b) with log q(z) - log p(z) = a z^2 + b z + const; mask[n] bit0 = dim was
 * replaced by the prior (CIS:144-145), bit1 = outlier sent as side information (CIS:148). */
CWQ_API int cwq_is_prologue(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, int64_t n,
                    const cwq_segment* segments, int32_t n_segments, float* kl, float* coef, uint8_t* mask,
                    void* stream);
/* CGS:194-201: rescale and KL.  tparam[2n] = (mu, 1/sigma) of the rescaled target. */
CWQ_API int cwq_greedy_prologue(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                        int64_t n, float* kl, float* tparam, void* stream);

/* ---- grouping by KL budget -------------------------------------------------------------------
 * CIS:164-204 / CGS:207-252, the sequential fp32 packing, evaluated exactly but in parallel.
 * starts[n_groups + 1] (capacity n + n_segments + 1): group g covers [starts[g], starts[g+1]);
 * group_seed[g] = segment seed + index of g inside its segment; gid[n] = group of each dim;
 * seg_group_off[n_segments + 1] = first group of each segment (last entry = n_groups).
 * `segments` is the device table, `h_segments` a host copy of it (scratch is sized from it). */
CWQ_API size_t cwq_group_workspace_bytes(int64_t n, int32_t n_segments);
CWQ_API int cwq_group(const float* kl, int64_t n, const cwq_segment* segments, const cwq_segment* h_segments,
              int32_t n_segments, int64_t* starts, uint32_t* group_seed, int32_t* gid, int64_t* seg_group_off,
              void* workspace, size_t workspace_bytes, void* stream);
/* The same per-group tables from caller-supplied group starts (the decoder's view, CIS:294, CGS:323):
 * gid[n] and group_seed[n_groups] for starts[n_groups + 1]; seg_group_off as above (device). */
CWQ_API int cwq_expand_groups(const int64_t* starts, int64_t n_groups, int64_t n, const cwq_segment* segments,
                      int32_t n_segments, const int64_t* seg_group_off, int32_t* gid, uint32_t* group_seed,
                      void* stream);
/* CIS:48-51: total KL of each group (sequential fp32) and K = ceil(exp(total)).
 * n_groups_dev (here and below, may be NULL): when given, the group count is read on the DEVICE (it is
 * seg_group_off + n_segments as cwq_group wrote it) and n_groups is only the capacity of the arrays: the
 * whole encode then runs without a host synchronisation between grouping and scoring. */
CWQ_API int cwq_group_totals(const float* kl, const int64_t* starts, int64_t n_groups, const int64_t* n_groups_dev,
                     float* total_kl, int32_t* K, void* stream);

/* ---- importance sampler ----------------------------------------------------------------------
 * CIS:55-63 for groups [g0, g1): index[g] = first candidate maximising the importance weight
 * (0-based; the reference codes index + 1).  max_group_size: an upper bound of the group sizes (the largest
 * size_limit of the segments); <= 16 selects the kernel specialised for the reference defaults. */
CWQ_API size_t cwq_is_encode_workspace_bytes(int64_t n_groups);
CWQ_API int cwq_is_encode(const float* coef, const int64_t* starts, const int32_t* K, const uint32_t* group_seed,
                  int64_t n_groups, const int64_t* n_groups_dev, int64_t g0, int64_t g1, int64_t max_group_size,
                  uint32_t* index, void* workspace, size_t workspace_bytes, void* stream);
/* The same for the range [range_dev[0], range_dev[1]) read on the DEVICE (two int64): a rank of a sharded run
 * computes its own range behind the grouping and starts scoring without telling the host first.  n_groups is the
 * capacity of the arrays; the workspace is the one of cwq_is_encode_workspace_bytes(n_groups). */
CWQ_API int cwq_is_encode_device_range(const float* coef, const int64_t* starts, const int32_t* K, const uint32_t* group_seed,
                               int64_t n_groups, const int64_t* range_dev, int64_t max_group_size, uint32_t* index,
                               void* workspace, size_t workspace_bytes, void* stream);
/* Ranges of groups for the `world` ranks of a sharded run (groups are independent: CIS:233-245): bounds[world + 1],
 * ascending, bounds[0] = 0, bounds[world] = group count; rank r scores [bounds[r], bounds[r+1]).  Contiguous and
 * balanced on candidates x Philox blocks per candidate.  Computed on the device from the device-side group count, so
 * that cwq_is_encode_device_range(..., bounds + rank, ...) can follow without a host synchronisation. */
CWQ_API size_t cwq_shard_ranges_workspace_bytes(int64_t n_groups);
CWQ_API int cwq_is_shard_ranges(const int32_t* K, const int64_t* starts, int64_t n_groups, const int64_t* n_groups_dev,
                        int32_t world, int64_t* bounds, void* workspace, size_t workspace_bytes, void* stream);
/* CIS:104-109 + CIS:265/347: regenerate row index[g] of every group and un-rescale:
 * sample[i] = p_scale[i] * z + p_loc[i]. */
CWQ_API int cwq_is_decode(const int64_t* starts, const uint32_t* group_seed, const int32_t* gid, const uint32_t* index,
                  int64_t n, const float* p_loc, const float* p_scale, float* sample, void* stream);

/* CIS:148-158, 267: outliers.  sample[i] is overwritten at outlier dims with the (seeded) posterior draw AS THE DECODER
 * WILL SEE IT, dequantise(quantise(draw)) -- the reference keeps the unquantised draw there, an encoder/decoder
 * asymmetry; outlier_idx (index inside the segment) and outlier_q (quint16 over [-30,30]) are the side
 * information, concatenated per segment with seg_outlier_off[n_segments + 1] offsets. */
CWQ_API size_t cwq_outlier_workspace_bytes(int64_t n);
CWQ_API int cwq_outlier_encode(const float* q_loc, const float* q_scale, const uint8_t* mask, int64_t n,
                       const cwq_segment* segments, int32_t n_segments, float* sample, int64_t* outlier_idx,
                       uint16_t* outlier_q, int64_t* seg_outlier_off, void* workspace, size_t workspace_bytes,
                       void* stream);
/* CIS:351-361: dequantise and overwrite the listed dims (no `== 0` sentinel: the index list says which dims are
 * outliers).  outlier_pos holds positions in the flat array. */
CWQ_API int cwq_outlier_decode(float* sample, const int64_t* outlier_pos, const uint16_t* outlier_q, int64_t m,
                       void* stream);

/* ---- the grouped importance-sampler encode (CIS:112-274) as ONE call -----------------------------
 * Runs the stages above in order on `stream` with no host synchronisation and nothing between the launches
 * but this function (the Python layer spends ~0.5 ms per call sequencing them one by one): GROUP = prologue +
 * grouping + candidate counts, SCORE = cwq_is_encode over all groups, FINISH = decode + outliers + Elias-delta
 * packing.  Any subset of stages may be requested (a sharded run calls GROUP, scores its own range with
 * cwq_is_encode, all-gathers, then calls FINISH).  The group count stays on the device at
 * seg_group_off[n_segments]; every per-group array has capacity n + n_segments (starts and bit_off one more). */
#define CWQ_STAGE_GROUP 1
#define CWQ_STAGE_SCORE 2
#define CWQ_STAGE_FINISH 4
typedef struct cwq_is_buffers {
  float* kl;                /* [n]      */
  float* coef;              /* [2n]     */
  uint8_t* mask;            /* [n]      */
  int32_t* gid;             /* [n]      */
  float* sample;            /* [n]      */
  int64_t* outlier_idx;     /* [n]      */
  uint16_t* outlier_q;      /* [n]      */
  int64_t* starts;          /* [n + n_segments + 1] */
  uint32_t* group_seed;     /* [n + n_segments]     */
  float* total_kl;          /* [n + n_segments]     */
  int32_t* K;               /* [n + n_segments]     */
  uint32_t* index;          /* [n + n_segments]     */
  int64_t* bit_off;         /* [n + n_segments + 1] */
  int64_t* seg_group_off;   /* [n_segments + 1]     */
  int64_t* seg_outlier_off; /* [n_segments + 1]     */
  uint8_t* bits;            /* bits_capacity_bytes: a multiple of 4, >= 42 bits per group of capacity */
  size_t bits_capacity_bytes;
  void* workspace;          /* cwq_is_pipeline_workspace_bytes(n, n_segments, max_group_size), shared by the stages: its
                               contents must survive from the GROUP stage to the FINISH stage of one encode */
  size_t workspace_bytes;
  const int64_t* score_range_dev; /* NULL: the SCORE stage covers all groups; else two int64 on the device, the range
                                     [g0, g1) of groups this rank scores (sharded runs: cwq_is_pipeline_shard_ranges) */
  const int32_t* tile_segment;    /* NULL, or [ceil(n / 1024)] on the device: index of the segment that holds dim 1024 k.
                                     A hint only (results do not depend on it): the kernels then find a dim's segment with
                                     one load instead of a binary search over the segment table */
} cwq_is_buffers;
CWQ_API size_t cwq_is_pipeline_workspace_bytes(int64_t n, int32_t n_segments, int64_t max_group_size);
/* Contiguous work-balanced ranges of groups for `world` ranks (same cut rule as cwq_is_shard_ranges), computed on the
 * device behind the GROUP stage from the work prefix that stage leaves in the workspace: bounds[world + 1] (device). */
CWQ_API int cwq_is_pipeline_shard_ranges(int64_t n, int32_t n_segments, int64_t max_group_size, int32_t world, int64_t* bounds,
                                 const cwq_is_buffers* h_buffers, void* stream);
/* The pipeline runs its O(N) stages as six fused kernels (csrc/cwq_front.cu) whenever the largest size limit is <= 128;
 * cwq_set_fused_front(0) forces the stage-by-stage kernels instead (same results; used by tests as a cross-check).
 * Returns the previous setting. */
CWQ_API int cwq_set_fused_front(int on);
CWQ_API int cwq_is_encode_pipeline(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale,
                           int64_t n, const cwq_segment* segments, const cwq_segment* h_segments, int32_t n_segments,
                           int64_t max_group_size, int32_t stages, const cwq_is_buffers* h_buffers, void* stream);

/* ---- greedy sampler --------------------------------------------------------------------------
 * CGS:38-78 for groups [g0, g1): n_steps sequential steps of 2^n_bits_per_step candidates each.
 * mode 1: one persistent CTA per group, step state on chip (many groups); mode 2: all groups in lockstep,
 * (group, 1024-candidate) units claimed dynamically, two grid barriers per step (few groups, e.g. one
 * image); mode 0 picks.  Identical results.  indices[g * n_steps + j]; best[n] in rescaled coordinates.
 * shard_loc (may be NULL = all zero, what every grouped call of the reference passes: CGS:278-281): loc of the proposal
 * shard per dim in rescaled coordinates, (p_loc / n_steps) / p_scale (CGS:42 `loc=p_loc / n_steps`); it is added to the
 * running sample at the start of every step, before the candidates are scored (CGS:57).  The per-group entry points
 * (CGS:29-36, 93-99) take any proposal loc; they rescale by p_scale only and finish with sample = p_scale * best. */
CWQ_API size_t cwq_greedy_workspace_bytes(int64_t n, int64_t n_groups, int32_t n_steps);
CWQ_API int cwq_greedy_encode(const float* tparam, const int64_t* starts, const uint32_t* group_seed, int64_t n_groups,
                      int64_t g0, int64_t g1, int32_t n_steps, int32_t n_bits_per_step, float shard_scale,
                      const float* shard_loc, int64_t max_group_size, int64_t n, int32_t mode, uint32_t* indices,
                      float* best, void* workspace, size_t workspace_bytes, void* stream);
/* CGS:134-165 + CGS:292/362: replay the chosen rows and un-rescale. */
CWQ_API int cwq_greedy_decode(const int64_t* starts, const uint32_t* group_seed, const int32_t* gid,
                      const uint32_t* indices, int64_t n, int32_t n_steps, float shard_scale, const float* shard_loc,
                      const float* p_loc, const float* p_scale, float* sample, void* stream);
/* CGS:42: shard_loc[i] = (p_loc[i] / n_steps) / p_scale[i] (IEEE float32 divisions). */
CWQ_API int cwq_greedy_shard_loc(const float* p_loc, const float* p_scale, int64_t n, int32_t n_steps, float* shard_loc,
                         void* stream);
/* sample = p_scale * best + p_loc  (CGS:292) */
CWQ_API int cwq_unrescale(const float* best, const float* p_loc, const float* p_scale, int64_t n, float* sample,
                  void* stream);

/* ---- index -> bitstream ----------------------------------------------------------------------
 * BIO:7-20 Elias-delta of index + 1, codes concatenated MSB-first into bytes (the order the '0'/'1'
 * string has in the reference, BIO:76-79).  bit_off[n_groups + 1] = bit offset of each code; with a device-side group
 * count the total is ALSO written to bit_off[n_groups] of the capacity (the last slot), where a host that does not know
 * the count yet can read it. */
CWQ_API size_t cwq_pack_workspace_bytes(int64_t n_groups);
CWQ_API int cwq_elias_delta_pack(const uint32_t* index, int64_t n_groups, const int64_t* n_groups_dev, uint8_t* bits,
                         size_t bits_capacity_bytes, int64_t* bit_off, void* workspace, size_t workspace_bytes,
                         void* stream);
/* BIO:41-53 / CGS:81-87: fixed-width LSB-first codes, concatenated MSB-first into bytes. */
CWQ_API int cwq_fixed_pack(const uint32_t* index, int64_t count, int32_t bits_per_index, uint8_t* bits,
                   size_t bits_capacity_bytes, void* stream);
CWQ_API int cwq_fixed_unpack(const uint8_t* bits, int64_t count, int32_t bits_per_index, uint32_t* index, void* stream);
/* Packed MSB-first bits -> '0'/'1' characters, the form in which the reference carries every code
 * (CIS:258-260, CGS:286-287, consumed by BIO:76-79).  ascii: device buffer of 8 * ceil(n_bits / 8) bytes, 8-byte
 * aligned.  n_bits_dev (may be NULL): the number of bits is read on the device and n_bits is only the capacity
 * (e.g. bit_off + n_groups of cwq_elias_delta_pack with 42 bits per group as the bound). */
CWQ_API int cwq_bits_to_ascii(const uint8_t* bits, int64_t n_bits, const int64_t* n_bits_dev, uint8_t* ascii, void* stream);
/* BIO:22-39 over a whole stream of n_codes codes; HOST function on a HOST '0'/'1' byte string
 * (the sequential parse the decoder does, CIS:325-336).  h_index receives x - 1. */
CWQ_API int cwq_elias_delta_parse_host(const uint8_t* h_ascii_bits, int64_t n_bits, int64_t n_codes, uint32_t* h_index,
                               int64_t* h_bits_consumed);

/* ---- arithmetic coder of the side information (HOST functions; reference code/coding.pyx) ------
 * counts[n_symbols] > 0 are the empirical histogram (coding.pyx:29-55); messages end with symbol 0
 * (coding.pyx:180).  Bits are '0'/'1' bytes as everywhere at the reference boundary.
 * encode: coding.pyx:59-125.  decode: fast = 0 -> coding.pyx:129-216, fast = 1 -> coding.pyx:220-310. */
CWQ_API int cwq_ac_encode_host(const int64_t* h_counts, int32_t n_symbols, const int64_t* h_message, int64_t n_message,
                       int32_t precision, uint8_t* h_ascii_bits, int64_t capacity, int64_t* h_n_bits);
CWQ_API int cwq_ac_decode_host(const int64_t* h_counts, int32_t n_symbols, const uint8_t* h_ascii_bits, int64_t n_bits,
                       int32_t precision, int32_t fast, int64_t* h_message, int64_t capacity, int64_t* h_n_message,
                       int64_t* h_bits_used);

/* ---- the candidate stream (debug / oracle injection seam, misc.py:3-17) ---------------------- */
/* rows k0..k0+nk-1, G columns, of the candidate table of stream `stream_seed`: out[nk * G].  raw = 0: the N(0,1) draws z
 * that stand in for tf.random.stateless_normal; raw = 1: the stream's raw draws y, z = float32(sqrt(2 ln 2)) * y, which are
 * what the candidate loops score (the factor is folded into their coefficients) and what the parity harness injects into
 * the oracle. */
CWQ_API int cwq_draw_candidates(uint32_t stream_seed, uint32_t k0, uint32_t nk, uint32_t G, int32_t raw, float* out,
                        void* stream);
/* the raw Philox words of blocks j = 0..nq-1 of rows k0..k0+nk-1 (block(s,k,j)): out[nk * nq * 4] */
CWQ_API int cwq_stream_words(uint32_t stream_seed, uint32_t k0, uint32_t nk, uint32_t nq, uint32_t* out, void* stream);
/* plain Philox4x32-10 on n (counter, key) pairs: ctr[4n], key[2n], out[4n] */
CWQ_API int cwq_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CWQ_H_ */
