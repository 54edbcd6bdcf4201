// gpu_ai_b200/csrc/kernels.cu -- hand-written sm_100a kernels of the playout hot path.
//
// Replaces singlePlayoutKernel (src/singlePlayout.cu:16-69), playoutKernel
// (src/multiplePlayout.cu:14-51), coarsePlayoutKernel (src/coarsePlayout.cu:18-89),
// heuristicPlayoutKernel (src/heuristicPlayout.cu:16-100) and genMovesKernel
// (src/genMovesTest.cu:10-24) of the reference.
//
// Design (DESIGN.md has the full account):
//  * a playout lives entirely in registers: 3 board words, turn, draw counter, ply counter;
//    no shared memory on the random path, no recursion; the only local memory is an 80-byte
//    sequence buffer touched by the rare king multi-jump enumeration (~1 % of plies).
//  * persistent lanes: every lane of every resident warp owns one playout at a time; finished
//    lanes are refilled together, once every 4 plies, with ONE warp-aggregated atomicAdd on the
//    work-queue head (__ballot_sync + __popc prefix) -- the same point where all lanes draw their
//    next Philox4x32-10 block, so both the refill and the RNG run converged.
//  * the only global traffic is one coalesced LDG.128 per playout and one STG.8 per result.
//  * integer work only (LOP3 / SHF / IADD3 / POPC / IMAD); no tensor cores: nothing here is a
//    contraction.
//  * SIMT discipline (measured): nothing lane-dependent branches BEFORE the common
//    selection path -- rare work (multi-jump enumeration, publishing a result) comes after it or is
//    written branch-free; the heuristic ply is staged with explicit rejoin points.
#include "kernels.cuh"

#include "bitboard.cuh"
#define B2P_GAUSS_QUAL __device__ const
#include "gauss_table_bits.h"
#include "philox.cuh"
#include "playout_core.cuh"

namespace b2p {

namespace {

#ifndef B2P_LANE_BLOCK
#define B2P_LANE_BLOCK 128  // 64 and 256 threads per block measured within noise of 128 (profiles/r02q_ab.txt)
#endif
constexpr int kLaneBlock = B2P_LANE_BLOCK;
constexpr int kRatioA = 52, kRatioB = 49;  // material quotient table: numerator 0..51, denominator 0..48
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t pick4(const Philox4 &b, int q) {
  uint32_t r = b.v[0];
  r = q == 1 ? b.v[1] : r;
  r = q == 2 ? b.v[2] : r;
  r = q == 3 ? b.v[3] : r;
  return r;
}

// 16-bit draw -> N(0, 0.11^2): 10-bit quantile bucket + 6-bit linear interpolation (one fma).
// The shared table holds {T[i], (T[i+1] - T[i]) / 64}: fma(d / 64, k, lo) is bit-identical to the protocol's
// fma(d, k / 64, lo) (power-of-two scaling is exact), and costs one LDS.64 instead of two loads and a subtract.
__device__ __forceinline__ float gauss_lookup(const float2 *tab, uint32_t h) {
  const float2 v = tab[(h >> 6) & 1023u];
  return __fmaf_rn(v.y, (float)(h & 63u), v.x);
}

__device__ __forceinline__ void fill_gauss_table(float2 *tab) {
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const float lo = __uint_as_float(b2p_gauss_table_bits[i]), hi = __uint_as_float(b2p_gauss_table_bits[i + 1]);
    tab[i] = make_float2(lo, (hi - lo) * (1.0f / 64.0f));
  }
}

// ---------------------------------------------------------------------------------------------
// thread-per-playout, persistent lanes
// ---------------------------------------------------------------------------------------------
#ifndef B2P_HEUR_MIN_BLOCKS
#define B2P_HEUR_MIN_BLOCKS 7  // 72 registers, 7 blocks/SM: +5 % over uncapped (99 regs), 8 blocks spills and loses 8 % (profiles/r02g_ab.txt)
#endif
#ifndef B2P_RAND_MIN_BLOCKS
#define B2P_RAND_MIN_BLOCKS 1
#endif
template <int MODE, bool LIMITED>
__global__ void __launch_bounds__(kLaneBlock, MODE == kHeuristic ? (B2P_HEUR_MIN_BLOCKS * 128) / kLaneBlock : B2P_RAND_MIN_BLOCKS) playout_lanes_kernel(const PlayoutParams prm) {
  constexpr bool kHeur = MODE == kHeuristic;
  constexpr bool kLeaf = MODE == kLeafGen;
  constexpr int kOrder = MODE == kRandomFast ? kOrderFast : kOrderCanonical;
  constexpr uint32_t kDomain = kLeaf ? kDomainLeaf : kDomainRandom;

  // heuristic mode: Gaussian quantile table + table of all material quotients a/b (a <= 51, b <= 48: 12 kings a
  // side plus one crowning), both filled once per block; a lookup replaces an IEEE division per weight class
  __shared__ float2 s_gauss[kHeur ? 1024 : 1];
  __shared__ float s_ratio[kHeur ? kRatioA * kRatioB : 1];
  if (kHeur) {
    fill_gauss_table(s_gauss);
    for (int i = threadIdx.x; i < kRatioA * kRatioB; i += blockDim.x) s_ratio[i] = (float)(i / kRatioB) / (float)(i % kRatioB);
    __syncthreads();
  }

  const unsigned lane = threadIdx.x & 31u;
  const unsigned below = (1u << lane) - 1u;

  Game g;
  g.pos.own = g.pos.opp = g.pos.kings = 0;
  g.turn = g.msc = 0;
  uint32_t w = 0, ply = 0, blk = 0, my_leaf = 0;
  uint64_t pid = 0;
  int limit = prm.max_plies;
  bool busy = false;
  bool more = true;  // warp-uniform: the queue may still hold work
  uint32_t c_none = 0, c_p1 = 0, c_p2 = 0, c_plies = 0;

  for (;;) {
    // ---- refill: all idle lanes of the warp take consecutive work items --------------------
    const unsigned idle = __ballot_sync(kFull, !busy);
    if (idle != 0u && more) {
      const uint32_t cnt = (uint32_t)__popc(idle);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(prm.next, cnt);
      base = __shfl_sync(kFull, base, 0);
      more = base + cnt < prm.total;
      if (!busy) {
        w = base + (uint32_t)__popc(idle & below);
        if (w < prm.total) {
          uint32_t leaf = w, rep = 0;
          if (prm.total != prm.n) {
            rep = prm.div_shift != 0xFFFFFFFFu ? (__umulhi(w, prm.div_magic) >> prm.div_shift) : w / prm.n;
            leaf = w - rep * prm.n;
          }
          pid = prm.pid_base + (uint64_t)rep * prm.rep_stride + leaf;
          if (LIMITED) my_leaf = leaf;
          if (kLeaf) {
            g = load_game(0x00000FFFu, 0xFFF00000u, 0u, 0u);  // getStartingState, src/state.cpp:25-40
          } else {
            const uint4 s = __ldg(prm.states + leaf);
            g = load_game(s.x, s.y, s.z, s.w);
          }
          busy = true;
          ply = 0;
          blk = 0;
          limit = prm.max_plies;
        }
      }
    }
    if (__ballot_sync(kFull, busy) == 0u) break;

    // ---- one Philox block = the draws of the next 4 plies (converged across the warp) -------
    Philox4 rnd;
    if (!kHeur) rnd = philox_block(prm.key, pid, kDomain, blk);
    blk++;

#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      // lanes that will play a heuristic ply in this slot (all 32 lanes are together at the loop top)
      unsigned heur_lanes = 0;
      if (kHeur) heur_lanes = __ballot_sync(kFull, busy && !((LIMITED || kLeaf) && limit >= 0 && (int)ply >= limit));
      if (!busy) continue;
      int res = kRunning;
      if (kLeaf && blk == 1 && q == 0) {
        // draw 0 of the leaf stream: prefix length U{1..100} (src/driver.cpp:80,86)
        limit = 1 + (int)mulhi(rnd.v[0], 100u);
        continue;
      }
      if ((LIMITED || kLeaf) && limit >= 0 && (int)ply >= limit) {
        // stopped by the ply limit: report the winner if the game happens to be over here
        Game probe = g;
        res = random_ply<kOrderFast>(probe, 0u);
        if (res == kRunning) res = 3;  // marker: unfinished
      } else if (kHeur) {
        res = heuristic_ply(g, heur_lanes, [&](int b) { return philox_block(prm.key, pid, kDomainNoise | ((uint32_t)b << 8), ply); },
                            [&](uint32_t r) { return gauss_lookup(s_gauss, r); },
                            [&](uint32_t a, uint32_t b) {
                              return (a < (uint32_t)kRatioA && b < (uint32_t)kRatioB) ? s_ratio[a * kRatioB + b] : (float)a / (float)b;
                            });
      } else {
        res = random_ply<kOrder>(g, pick4(rnd, q));
      }
      if (res == kRunning) {
        ply++;
        continue;
      }
      // ---- playout finished: publish --------------------------------------------------------
      if (res == 3) res = kRunning;
      if (prm.winners) prm.winners[w] = (int8_t)res;
      if (LIMITED || kLeaf) {
        if (prm.plies) prm.plies[w] = ply;
        if (prm.final_states) {
          uint32_t o[4];
          store_game(g, o);
          prm.final_states[w] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (prm.leaf_wins && (res == 0 || res == 1)) atomicAdd(prm.leaf_wins + 2 * my_leaf + res, 1u);
      }
      c_none += res == -1;
      c_p1 += res == 0;
      c_p2 += res == 1;
      c_plies += ply;
      busy = false;
    }
  }

  // ---- per-warp reduction of the win counters, one atomic per counter per warp ---------------
  if (prm.counters) {
    c_none = __reduce_add_sync(kFull, c_none);
    c_p1 = __reduce_add_sync(kFull, c_p1);
    c_p2 = __reduce_add_sync(kFull, c_p2);
    unsigned long long pl = c_plies;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pl += __shfl_xor_sync(kFull, pl, o);
    if (lane == 0) {
      if (c_none) atomicAdd(prm.counters + 0, (unsigned long long)c_none);
      if (c_p1) atomicAdd(prm.counters + 1, (unsigned long long)c_p1);
      if (c_p2) atomicAdd(prm.counters + 2, (unsigned long long)c_p2);
      if (pl) atomicAdd(prm.counters + 3, pl);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// batched canonical move lists (parity kernel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) genmoves_kernel(const uint4 *__restrict__ states, uint32_t n, int max_moves,
                                                       unsigned long long *__restrict__ moves,
                                                       uint8_t *__restrict__ counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 s = __ldg(states + i);
  const int cnt = gen_moves_canonical(s.x, s.y, s.z & (s.x | s.y), s.w & 1u,
                                      reinterpret_cast<uint64_t *>(moves) + (size_t)i * max_moves, max_moves);
  counts[i] = (uint8_t)(cnt > 255 ? 255 : cnt);
}

template <int MODE, bool LIMITED>
cudaError_t launch_lanes_t(const PlayoutParams &prm_in, int sm_count, cudaStream_t stream, LaunchInfo *info) {
  PlayoutParams prm = prm_in;
  set_divider(prm);
  auto kern = playout_lanes_kernel<MODE, LIMITED>;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLaneBlock, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long long want = ((long long)prm.total + kLaneBlock - 1) / kLaneBlock;
  long long cap = (long long)sm_count * per_sm;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  if (info) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    info->grid = grid;
    info->block = kLaneBlock;
    info->regs = fa.numRegs;
    info->blocks_per_sm = per_sm;
  }
  e = cudaMemsetAsync(prm.next, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) return e;
  kern<<<grid, kLaneBlock, 0, stream>>>(prm);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_playout_lanes(const PlayoutParams &prm, KernelMode mode, int sm_count, cudaStream_t stream,
                                 LaunchInfo *info) {
  const bool limited = prm.max_plies >= 0 || prm.plies != nullptr || prm.final_states != nullptr || prm.leaf_wins != nullptr;
  switch (mode) {
    case kRandomCanonical:
      return limited ? launch_lanes_t<kRandomCanonical, true>(prm, sm_count, stream, info)
                     : launch_lanes_t<kRandomCanonical, false>(prm, sm_count, stream, info);
    case kRandomFast:
      return limited ? launch_lanes_t<kRandomFast, true>(prm, sm_count, stream, info)
                     : launch_lanes_t<kRandomFast, false>(prm, sm_count, stream, info);
    case kHeuristic:
      return limited ? launch_lanes_t<kHeuristic, true>(prm, sm_count, stream, info)
                     : launch_lanes_t<kHeuristic, false>(prm, sm_count, stream, info);
    case kLeafGen:
      return launch_lanes_t<kLeafGen, true>(prm, sm_count, stream, info);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_genmoves(const uint4 *states, uint32_t n, int max_moves, unsigned long long *moves, uint8_t *counts,
                            cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int block = 128;
  genmoves_kernel<<<(n + block - 1) / block, block, 0, stream>>>(states, n, max_moves, moves, counts);
  return cudaGetLastError();
}

}  // namespace b2p
