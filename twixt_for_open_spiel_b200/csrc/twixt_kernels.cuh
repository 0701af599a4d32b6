// twixt_kernels.cuh -- launch wrappers of the sm_100a kernels (defined in
// twixt_kernels_api.cu and twixt_kernel_playout.cu).  All env state is an
// array of records in global memory: records + env * record_words(n).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/twixt_b200.h"

namespace twixt {

// device-side counters, one per batch (twixt_stats in the C ABI)
struct DeviceStats {
  unsigned long long plies;
  unsigned long long games;
  unsigned long long red_wins;
  unsigned long long blue_wins;
  unsigned long long draws;
  unsigned long long swaps;
  unsigned long long max_length;
  // lowest index (within the call's range) of an illegal action seen by the
  // apply kernel, 0xFFFFFFFF = none
  unsigned int illegal_index;
  unsigned int pad;
  // ticket counter of the persistent playout kernel (zeroed before every launch)
  unsigned long long tickets;
  // validate kernel: (lowest invalid env << 8) | reason, all ones = every record passed
  unsigned long long invalid_code;
  // replay kernel: (lowest env that met an illegal action << 32) | that action, all ones = none
  unsigned long long replay_illegal;
  // clone kernel: lowest position whose gathered source id was out of range / inside the destination
  unsigned int bad_clone_index;
  unsigned int pad2;
  // only written by the bounds-instrumented test variant of the playout kernel (TW_PLAYOUT_BOUNDS_CHECK)
  unsigned long long bounds_violations;
};

struct PlayoutArgs {
  uint32_t* records;           // first env of the range
  int64_t count;
  int n;
  int max_plies;
  uint64_t seed;
  uint64_t stream_base;        // stream id of range element 0 when stream_ids == nullptr
  const uint64_t* stream_ids;  // [count] or nullptr
  float* out_returns;          // [count,2] or nullptr
  int32_t* out_lengths;        // [count] or nullptr
  uint16_t* out_actions;       // [trace_plies, count] or nullptr
  int trace_plies;
  DeviceStats* stats;
  unsigned long long* tickets;  // &stats->tickets
};

cudaError_t launch_reset(uint32_t* records, int64_t count, int n, cudaStream_t s);
cudaError_t launch_clone(uint32_t* dst, const uint32_t* src, const int64_t* src_ids, int64_t count, int n,
                         int64_t num_envs, int64_t dst_first, DeviceStats* stats, cudaStream_t s);
cudaError_t launch_replay(uint32_t* records, int64_t count, int n, const int32_t* actions, int64_t stride,
                          const int32_t* lengths, int32_t* out_applied, DeviceStats* stats, cudaStream_t s);
cudaError_t launch_step(uint32_t* record, int n, int action, twixt_step_result* out, int64_t* out_legal, cudaStream_t s);
cudaError_t launch_validate(const uint32_t* records, int64_t count, int n, DeviceStats* stats, cudaStream_t s);
const char* invalid_reason_text(unsigned code);
cudaError_t launch_legal_actions(const uint32_t* records, int64_t count, int n, void* out_actions, int elem_bytes,
                                 int64_t stride, int32_t* out_counts, cudaStream_t s);
cudaError_t launch_legal_mask(const uint32_t* records, int64_t count, int n, uint8_t* out, cudaStream_t s);
cudaError_t launch_apply(uint32_t* records, int64_t count, int n, const int32_t* actions, int32_t* out_status,
                         DeviceStats* stats, cudaStream_t s);
cudaError_t launch_query(const uint32_t* records, int64_t count, int n, int8_t* out_player, uint8_t* out_terminal,
                         float* out_returns, cudaStream_t s);
cudaError_t launch_observation(const uint32_t* records, int64_t count, int n, float* out, uint8_t* out_mask,
                               cudaStream_t s);
cudaError_t launch_playout(const PlayoutArgs& a, cudaStream_t s);
// one-time per-process setup of the playout kernels (opt-in shared memory)
cudaError_t playout_setup();
// the five size groups of twixt_kernel_playout.cu (board sizes 5+g, 10+g, 15+g, 20+g)
cudaError_t playout_setup_g0();
cudaError_t playout_setup_g1();
cudaError_t playout_setup_g2();
cudaError_t playout_setup_g3();
cudaError_t playout_setup_g4();
cudaError_t launch_playout_g0(const PlayoutArgs& a, cudaStream_t s);
cudaError_t launch_playout_g1(const PlayoutArgs& a, cudaStream_t s);
cudaError_t launch_playout_g2(const PlayoutArgs& a, cudaStream_t s);
cudaError_t launch_playout_g3(const PlayoutArgs& a, cudaStream_t s);
cudaError_t launch_playout_g4(const PlayoutArgs& a, cudaStream_t s);

}  // namespace twixt
