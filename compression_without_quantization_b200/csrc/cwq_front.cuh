// cwq_front.cuh -- interface between cwq_front.cu (the fused O(N) stages of the importance-sampler encode) and the
// pipeline entry point in cwq_samplers.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cwq.h"

namespace cwq {

constexpr int FRONT_THREADS = 256;
constexpr int FRONT_UNIT_CAND = 1024;     // candidates per scoring unit (== UNIT_CAND of the scoring kernel)
// slots of the small scalar block in the workspace (zeroed by k_front_chain)
constexpr int FRONT_TICKET_GROUPS = 0, FRONT_TICKET_PACK = 1, FRONT_SCORE_COUNTER = 2;

struct FrontPlan {
  int T, E, W;              // dims per tile, entry offsets per tile (size limit + 1), 32-bit words per bit row
  long long n_tiles;
  size_t smem_tiles, smem_chain, smem_groups;
};

struct FrontWorkspace {
  // alive from the GROUP stage to the FINISH stage
  unsigned long long* unit_off;     // [cap + 1] exclusive prefix of scoring units over groups
  unsigned long long* work_off;     // [cap + 1] exclusive prefix of K x Philox blocks (sharding)
  unsigned long long* keys;         // [cap]     (score, index) argmax keys
  uint32_t* tile_out_off;           // [n_tiles] outliers before each tile
  unsigned long long* scalars;      // [8]
  unsigned long long* pack_status;  // look-back words of the packing kernel
  // scratch of the GROUP stage
  uint8_t* hop;
  int* seg_first;
  uint8_t* maps;
  uint16_t* counts;
  uint32_t* rows;
  uint32_t* tile_out;
  uint32_t* base;
  uint8_t* entry;
  unsigned long long* status;
};

// false: the fused path does not apply (size limit above 128, or shapes that do not fit shared memory)
bool front_plan(long long n, long long max_group_size, FrontPlan* plan);
size_t front_workspace_bytes(long long n, int n_seg, long long cap, const FrontPlan& plan);
bool front_carve(void* workspace, size_t bytes, long long n, int n_seg, long long cap, const FrontPlan& plan, FrontWorkspace* w);
int front_group(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, long long n,
                const cwq_segment* segs, int n_seg, const FrontPlan& plan, const FrontWorkspace& w, const cwq_is_buffers* b,
                cudaStream_t st);
int front_finish_clear(const FrontPlan& plan, const FrontWorkspace& w, const cwq_is_buffers* b, cudaStream_t st);
int front_finish(const float* q_loc, const float* q_scale, const float* p_loc, const float* p_scale, long long n,
                 const cwq_segment* segs, int n_seg, long long cap, const FrontPlan& plan, const FrontWorkspace& w,
                 const cwq_is_buffers* b, cudaStream_t st);

}  // namespace cwq
