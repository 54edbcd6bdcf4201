// gpu_ai_b200/csrc/warp_kernel.cu -- warp-per-playout scheduling (B2P_SCHED_WARP).
//
// Replaces the scheduling of the reference's playoutKernel (src/multiplePlayout.cu:14-51) and
// heuristicPlayoutKernel (src/heuristicPlayout.cu:16-100): one group of 32 threads per playout.
// The reference gives each thread one board square and builds the move list with a shared-
// memory Blelloch scan and ~12 __syncthreads per move type (src/state.cu:182-237), then lets
// thread 0 pick and apply the move.  On a bitboard the 32 squares already ARE one machine word,
// so here the warp keeps the position uniform in registers (every lane the same three words) and
// splits across lanes only the work that is serial in the thread-per-playout kernel:
//   * multi-hop captures: lane i enumerates the sequences that start on square i (its own register
//     DFS), counts are combined with a shuffle scan, the owning lane walks to the chosen sequence
//     and broadcasts it;
//   * heuristic playouts: lane i scores the candidate moves of square i (its own Philox noise
//     draws); the best (weight, canonical index) pair is found with a 5-step shuffle reduction
//     instead of the reference's shared-memory max reduction (src/heuristicPlayout.cu:66-85).
// No shared memory (except the noise table), no __syncthreads in the ply loop, no data race
// (the reference's kernels rely on implicit warp lock-step).
//
// Throughput is ~10x below the thread-per-playout kernel (one playout per warp instead of 32);
// what it buys is latency for tiny batches (an MCTS batch of 50 leaves fills 50 warps on 50 SMs
// instead of 2 warps on one SM).  B2P_SCHED_AUTO picks it below a measured batch size.
#include "kernels.cuh"

#include "bitboard.cuh"
#define B2P_GAUSS_QUAL __device__ const
#include "gauss_table_bits.h"
#include "philox.cuh"
#include "playout_core.cuh"

namespace b2p {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kWarpBlock = 128;

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

// exclusive prefix sum of c over the lanes + total
__device__ __forceinline__ int warp_scan(int c, unsigned lane, int &total) {
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if ((int)lane >= o) incl += t;
  }
  total = __shfl_sync(kFull, incl, 31);
  return incl - c;
}

// ---- random ply, warp-cooperative ---------------------------------------------------------------
template <int ORDER>
__device__ __forceinline__ int warp_random_ply(Game &g, uint32_t r, unsigned lane) {
  if (g.msc >= kDrawPlies) return -1;
  const Pos p = g.pos;
  const PlyMasks m = ply_masks(p);
  uint32_t from = 0, to = 0, captured = 0;
  const int shape = capture_shape(p, m.jm, m.cap);
  if (shape == 2) {
    // lane = origin square: every capturing piece enumerates its own sequences in parallel
    const uint32_t origins = m.cap[0] | m.cap[1] | m.cap[2] | m.cap[3];
    const bool mine = (origins >> lane) & 1u;
    bool stopped = false;
    int c = 0;
    if (mine) c = for_each_capture_from(p, m.jm, (int)lane, [](const CaptureMove &) { return false; }, stopped);
    int n;
    const int excl = warp_scan(c, lane, n);
    int k = (int)mulhi(r, (uint32_t)n);
    if (ORDER == kOrderCanonical && g.turn) k = n - 1 - k;
    const bool owner = mine && k >= excl && k < excl + c;
    if (owner) {
      int left = k - excl;
      for_each_capture_from(p, m.jm, (int)lane, [&](const CaptureMove &cm) {
        if (left-- != 0) return false;
        from = 1u << cm.from; to = 1u << cm.to; captured = cm.captured;
        return true;
      }, stopped);
    }
    const int src = __ffs(__ballot_sync(kFull, owner)) - 1;
    from = __shfl_sync(kFull, from, src);
    to = __shfl_sync(kFull, to, src);
    captured = __shfl_sync(kFull, captured, src);
  } else {
    // one entry per first hop: the bitboard word is already the 32-lane vector; uniform across the warp
    if (pick_first_hop_and_chain<ORDER>(p, m, shape, g.turn, r, from, to, captured) == 0) return (int)(g.turn ^ 1u);
  }
  finish_ply(g, m.capture, from, to, captured);
  return kRunning;
}

// ---- heuristic ply, warp-cooperative --------------------------------------------------------------
// Candidate with canonical list index idx gets weight + noise(idx); the winner is the largest
// weight, ties to the smallest canonical index (= the reference's strict '>' scan,
// src/heuristicPlayout.cpp:27-37).
struct Best {
  float w;
  int idx;
  uint32_t from, to, captured;
};

__device__ __forceinline__ void consider(Best &b, float w, int idx, uint32_t from, uint32_t to, uint32_t captured) {
  if (w > b.w || (w == b.w && idx < b.idx)) { b.w = w; b.idx = idx; b.from = from; b.to = to; b.captured = captured; }
}

__device__ __forceinline__ int warp_heuristic_ply(Game &g, uint64_t key, uint64_t pid, uint32_t ply, const float2 *gauss,
                                                  unsigned lane) {
  if (g.msc >= kDrawPlies) return -1;
  const Pos p = g.pos;
  const PlyMasks m = ply_masks(p);
  const uint32_t my = material(p.own, p.kings), his = material(p.opp, p.kings);
  const uint32_t ownMen = p.own & ~p.kings;
  const bool rev = g.turn != 0;
  int cached = -1;
  Philox4 nb;
  nb.v[0] = nb.v[1] = nb.v[2] = nb.v[3] = 0;
  auto noise = [&](int idx) {
    const int b = idx >> 3;
    if (b != cached) {
      nb = philox_block(key, pid, kDomainNoise | ((uint32_t)b << 8), ply);
      cached = b;
    }
    const uint32_t word = pick4(nb, (idx >> 1) & 3);
    return gauss_lookup(gauss, (idx & 1) ? word >> 16 : word & 0xFFFFu);
  };
  Best best;
  best.w = -__int_as_float(0x7f800000);
  best.idx = 0x7fffffff;
  best.from = best.to = best.captured = 0;
  int n;
  if (m.capture) {
    const uint32_t origins = m.cap[0] | m.cap[1] | m.cap[2] | m.cap[3];
    const bool mine = (origins >> lane) & 1u;
    bool stopped = false;
    int c = 0;
    if (mine) c = for_each_capture_from(p, m.jm, (int)lane, [](const CaptureMove &) { return false; }, stopped);
    const int excl = warp_scan(c, lane, n);
    if (mine) {
      int li = 0;
      for_each_capture_from(p, m.jm, (int)lane, [&](const CaptureMove &cm) {
        const int idx = rev ? n - 1 - (excl + li) : excl + li;
        const uint32_t promo = (((ownMen >> cm.from) & 1u) && cm.to >= 28) ? 3u : 0u;
        const uint32_t loss = (uint32_t)(popc(cm.captured) + 3 * popc(cm.captured & p.kings));
        consider(best, (float)(my + promo) / (float)(his - loss) + noise(idx), idx, 1u << cm.from, 1u << cm.to, cm.captured);
        li++;
        return false;
      }, stopped);
    }
  } else {
    const uint32_t ownK = p.own & p.kings;
    const uint32_t a[4] = {p.own & m.e[0], p.own & m.e[1], ownK & m.e[2], ownK & m.e[3]};
    n = popc(a[0]) + popc(a[1]) + popc(a[2]) + popc(a[3]);
    if (n == 0) return (int)(g.turn ^ 1u);
    const uint32_t below = (1u << lane) - 1u;
    int li = popc(a[0] & below) + popc(a[1] & below) + popc(a[2] & below) + popc(a[3] & below);
    const float plain = (float)my / (float)his, crowned = (float)(my + 3u) / (float)his;
    const bool man = (ownMen >> lane) & 1u;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (!((a[d] >> lane) & 1u)) continue;
      const int t = step_target((int)lane, d);
      const int idx = rev ? n - 1 - li : li;
      consider(best, ((man && t >= 28) ? crowned : plain) + noise(idx), idx, 1u << lane, 1u << t, 0u);
      li++;
    }
  }
  // arg-max over the warp: (weight, then smallest canonical index)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ow = __shfl_xor_sync(kFull, best.w, o);
    const int oi = __shfl_xor_sync(kFull, best.idx, o);
    if (ow > best.w || (ow == best.w && oi < best.idx)) { best.w = ow; best.idx = oi; best.from = 0; }
  }
  // the lane that still holds the winning move (from != 0 and its idx is the winner) broadcasts it
  const unsigned holder = __ballot_sync(kFull, best.from != 0u);
  const int src = __ffs(holder) - 1;
  const uint32_t from = __shfl_sync(kFull, best.from, src);
  const uint32_t to = __shfl_sync(kFull, best.to, src);
  const uint32_t captured = __shfl_sync(kFull, best.captured, src);
  finish_ply(g, m.capture, from, to, captured);
  return kRunning;
}

template <int MODE>
__global__ void __launch_bounds__(kWarpBlock) playout_warp_kernel(const PlayoutParams prm) {
  constexpr bool kHeur = MODE == kHeuristic;
  constexpr int kOrder = MODE == kRandomFast ? kOrderFast : kOrderCanonical;
  __shared__ float2 s_gauss[kHeur ? 1024 : 1];
  if (kHeur) {
    fill_gauss_table(s_gauss);
    __syncthreads();
  }
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  uint32_t c_none = 0, c_p1 = 0, c_p2 = 0;
  unsigned long long c_plies = 0;

  for (uint32_t w = gwarp; w < prm.total; w += nwarps) {
    uint32_t leaf = w, rep = 0;
    if (prm.total != prm.n) {
      rep = prm.div_shift != 0xFFFFFFFFu ? (__umulhi(w, prm.div_magic) >> prm.div_shift) : w / prm.n;
      leaf = w - rep * prm.n;
    }
    const uint64_t pid = prm.pid_base + (uint64_t)rep * prm.rep_stride + leaf;
    const uint4 s = __ldg(prm.states + leaf);  // same address in all lanes: one broadcast transaction
    Game g = load_game(s.x, s.y, s.z, s.w);
    uint32_t ply = 0;
    int res;
    Philox4 rnd;
    rnd.v[0] = rnd.v[1] = rnd.v[2] = rnd.v[3] = 0;
    for (;;) {
      if (prm.max_plies >= 0 && (int)ply >= prm.max_plies) {
        Game probe = g;
        res = random_ply<kOrderFast>(probe, 0u);
        break;
      }
      if (kHeur) {
        res = warp_heuristic_ply(g, prm.key, pid, ply, s_gauss, lane);
      } else {
        if ((ply & 3u) == 0u) rnd = philox_block(prm.key, pid, kDomainRandom, ply >> 2);
        res = warp_random_ply<kOrder>(g, pick4(rnd, ply & 3), lane);
      }
      if (res != kRunning) break;
      ply++;
    }
    if (lane == 0) {
      if (prm.winners) prm.winners[w] = (int8_t)res;
      if (prm.plies) prm.plies[w] = ply;
      if (prm.final_states) {
        uint32_t o[4];
        store_game(g, o);
        prm.final_states[w] = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (prm.leaf_wins && (res == 0 || res == 1)) atomicAdd(prm.leaf_wins + 2 * leaf + res, 1u);
      c_none += res == -1;
      c_p1 += res == 0;
      c_p2 += res == 1;
      c_plies += ply;
    }
  }
  if (prm.counters && lane == 0) {
    if (c_none) atomicAdd(prm.counters + 0, (unsigned long long)c_none);
    if (c_p1) atomicAdd(prm.counters + 1, (unsigned long long)c_p1);
    if (c_p2) atomicAdd(prm.counters + 2, (unsigned long long)c_p2);
    if (c_plies) atomicAdd(prm.counters + 3, c_plies);
  }
}

template <int MODE>
cudaError_t launch_warp_t(const PlayoutParams &prm_in, int sm_count, cudaStream_t stream, LaunchInfo *info) {
  PlayoutParams prm = prm_in;
  set_divider(prm);
  auto kern = playout_warp_kernel<MODE>;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpBlock, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const int warps_per_block = kWarpBlock / 32;
  long long want = ((long long)prm.total + warps_per_block - 1) / warps_per_block;
  long long cap = (long long)sm_count * per_sm;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  if (info) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    info->grid = grid;
    info->block = kWarpBlock;
    info->regs = fa.numRegs;
    info->blocks_per_sm = per_sm;
  }
  kern<<<grid, kWarpBlock, 0, stream>>>(prm);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_playout_warp(const PlayoutParams &prm, KernelMode mode, int sm_count, cudaStream_t stream,
                                LaunchInfo *info) {
  switch (mode) {
    case kRandomCanonical: return launch_warp_t<kRandomCanonical>(prm, sm_count, stream, info);
    case kRandomFast: return launch_warp_t<kRandomFast>(prm, sm_count, stream, info);
    case kHeuristic: return launch_warp_t<kHeuristic>(prm, sm_count, stream, info);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace b2p
