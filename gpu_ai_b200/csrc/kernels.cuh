// gpu_ai_b200/csrc/kernels.cuh -- launch interface between the C ABI (api.cu) and the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2p {

struct PlayoutParams {
  const uint4 *states;          // n packed leaf states (16 B each, one LDG.128 per refill)
  uint32_t n;                   // leaves
  uint32_t total;               // playouts = n * reps  (< 2^31 per launch)
  uint64_t rep_stride;          // playout id = pid_base + rep * rep_stride + leaf (global leaf count when sharded)
  uint64_t key;                 // Philox key
  uint64_t pid_base;            // playout id of work item 0
  int max_plies;                // < 0: to the end
  int8_t *winners;              // [total] or null
  uint32_t *plies;              // [total] or null
  uint4 *final_states;          // [total] or null
  unsigned int *leaf_wins;      // [n][2] PLAYER_1 / PLAYER_2 wins per leaf over all reps (atomically accumulated) or null
  unsigned long long *counters; // [4] draws, p1, p2, plies (atomically accumulated)
  unsigned int *next;           // work-queue head (zeroed by the launcher)
};

enum KernelMode { kRandomCanonical = 0, kRandomFast = 1, kHeuristic = 2, kLeafGen = 3 };

struct LaunchInfo {
  int grid, block, regs, blocks_per_sm;
};

// thread-per-playout, persistent lanes.  Returns cudaSuccess or the launch error.
cudaError_t launch_playout_lanes(const PlayoutParams &prm, KernelMode mode, int sm_count, cudaStream_t stream,
                                 LaunchInfo *info);
// warp-per-playout (small batches)
cudaError_t launch_playout_warp(const PlayoutParams &prm, KernelMode mode, int sm_count, cudaStream_t stream,
                                LaunchInfo *info);
cudaError_t launch_genmoves(const uint4 *states, uint32_t n, int max_moves, unsigned long long *moves, uint8_t *counts,
                            cudaStream_t stream);
cudaError_t launch_microbench(int which, int iters, int sm_count, uint32_t *sink, cudaStream_t stream,
                              double *thread_ops);

}  // namespace b2p
