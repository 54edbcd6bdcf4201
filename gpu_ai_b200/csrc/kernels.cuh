// gpu_ai_b200/csrc/kernels.cuh -- launch interface between the C ABI (api.cu) and the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2p {

struct PlayoutParams {
  const uint4 *states;          // n packed leaf states (16 B each, one LDG.128 per refill)
  uint32_t n;                   // leaves
  uint32_t total;               // playouts = n * reps  (< 2^31 per launch)
  uint32_t div_magic, div_shift; // w / n == umulhi(w, div_magic) >> div_shift for every w < 2^31 (set_divider)
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

// Exact division of a 31-bit work index by n without a DIV sequence in the refill path:
// m = floor(2^(31+s) / n) + 1 with s = ceil(log2 n) (Granlund-Montgomery round-up method, exact for 31-bit
// dividends because m*n - 2^(31+s) <= n <= 2^s).  n == 1 keeps the plain division (div_shift = 0xFFFFFFFF).
inline void set_divider(PlayoutParams &prm) {
  const uint32_t n = prm.n;
  if (n <= 1) { prm.div_magic = 0; prm.div_shift = 0xFFFFFFFFu; return; }
  uint32_t s = 0;
  while ((1ull << s) < n) s++;
  prm.div_magic = (uint32_t)((1ull << (31 + s)) / n + 1ull);
  prm.div_shift = s - 1;
}

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
