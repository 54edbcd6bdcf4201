// gpu_ai_b200/csrc/philox.cuh
//
// Counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
// as 1, 2, 3", SC'11) -- one independent stream per playout, no RNG state in memory.  Replaces
// the reference's per-thread XORWOW states (`curand_init(SEED, tid, 0, ...)`,
// src/singlePlayout.cu:26-27, src/multiplePlayout.cu:24-25) and, on the host side, glibc
// rand() (src/player.cpp:15).
//
// Draw protocol (mirrored, independently, by oracle/chooser.h for the parity tests):
//   block(key, pid, domain, b) = philox4x32_10(ctr = {pid_lo, pid_hi, b, domain}, key = {key_lo, key_hi})
//   draw t of a stream         = block(key, pid, domain, t >> 2)[t & 3]
//   random playout, ply p      : move index = mulhi32(draw p of domain 0, n_moves)
//   heuristic playout          : candidate i of ply p gets noise from the 16-bit half (i & 1) of word
//                                (i >> 1) & 3 of block(key, pid, kDomainNoise | (i >> 3) << 8, p):
//                                one block serves 8 candidates
//   leaf generation            : domain 2; draw 0 -> prefix length 1 + mulhi32(r, 100),
//                                draw 1 + p -> move index of prefix ply p
#pragma once

#include <stdint.h>

#include "bitboard.cuh"

namespace b2p {

enum : uint32_t { kDomainRandom = 0u, kDomainNoise = 1u, kDomainLeaf = 2u };

struct Philox4 {
  uint32_t v[4];
};

B2P_HD void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
  const uint64_t p = (uint64_t)a * b;  // one IMAD.WIDE.U32 on sm_100a
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

B2P_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    uint32_t h0, l0, h1, l1;
    mulhilo(0xD2511F53u, c0, h0, l0);
    mulhilo(0xCD9E8D57u, c2, h1, l1);
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

B2P_HD Philox4 philox_block(uint64_t key, uint64_t pid, uint32_t domain, uint32_t block) {
  return philox4x32_10((uint32_t)pid, (uint32_t)(pid >> 32), block, domain, (uint32_t)key, (uint32_t)(key >> 32));
}

}  // namespace b2p
