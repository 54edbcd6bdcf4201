// tests/host_build/hostlib.cpp -- TEST-ONLY host compilation of the product's bit logic.
//
// Compiles gpu_ai_b200/csrc/{bitboard,philox,playout_core}.cuh with g++ (B2P_HD expands to
// `inline`) so that the CPU test-suite (`-m "not gpu"`) can check the bitboard move
// generator, the ply step and the Philox protocol against the oracle without a GPU.
// This library is NOT part of the product: libb2p.so has no host execution path and
// nothing in gpu_ai_b200/ loads this file.
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "../../gpu_ai_b200/csrc/playout_core.cuh"
#include "../../gpu_ai_b200/csrc/gauss_table_bits.h"

using namespace b2p;

static float gauss_sigma(uint32_t h) {
  uint32_t i = (h >> 6) & 1023u;
  float frac = (float)(h & 63u) * (1.0f / 64.0f);
  float lo, hi;
  std::memcpy(&lo, &b2p_gauss_table_bits[i], 4);
  std::memcpy(&hi, &b2p_gauss_table_bits[i + 1], 4);
  return __builtin_fmaf(hi - lo, frac, lo);
}

extern "C" {

void hb_genmoves_batch(const uint32_t *packed, size_t n, int max_moves, uint64_t *moves_out, uint8_t *counts_out) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const uint32_t *w = packed + 4 * i;
    int cnt = gen_moves_canonical(w[0], w[1], w[2] & (w[0] | w[1]), w[3] & 1u, moves_out + i * (size_t)max_moves, max_moves);
    counts_out[i] = (uint8_t)cnt;
  }
}

// mirrors oracle or_playouts_batch(): mode 0 random / 1 heuristic; order 0 canonical / 1 fast
void hb_playouts_batch(const uint32_t *packed, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base, int mode,
                       int order, int max_plies, int8_t *winners_out, uint32_t *plies_out, uint32_t *final_out,
                       uint64_t counters_out[4]) {
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  size_t total = n * (size_t)reps;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c0, c1, c2, c3)
  for (size_t wi = 0; wi < total; wi++) {
    const uint32_t *w = packed + 4 * (wi % n);
    Game g = load_game(w[0], w[1], w[2], w[3]);
    uint64_t pid = pid_base + wi;
    uint32_t ply = 0;
    int res = kRunning;
    for (;;) {
      if (max_plies >= 0 && (int)ply >= max_plies) {
        // stopped early: still report a finished game as finished (oracle does the same)
        Game probe = g;
        res = random_ply<kOrderFast>(probe, 0);
        if (res == kRunning) break;
        break;
      }
      if (mode == 1) {
        res = heuristic_ply(g, 0u, [&](int b) { return philox_block(key, pid, kDomainNoise | ((uint32_t)b << 8), ply); },
                            [](uint32_t r) { return gauss_sigma(r); }, [](uint32_t a, uint32_t b) { return (float)a / (float)b; });
      } else {
        Philox4 b = philox_block(key, pid, kDomainRandom, ply >> 2);
        uint32_t r = b.v[ply & 3];
        res = order == 1 ? random_ply<kOrderFast>(g, r) : random_ply<kOrderCanonical>(g, r);
      }
      if (res != kRunning) break;
      ply++;
    }
    if (winners_out) winners_out[wi] = (int8_t)res;
    if (plies_out) plies_out[wi] = ply;
    if (final_out) store_game(g, final_out + 4 * wi);
    if (res == -1) c0++; else if (res == 0) c1++; else if (res == 1) c2++;
    c3 += ply;
  }
  if (counters_out) { counters_out[0] = c0; counters_out[1] = c1; counters_out[2] = c2; counters_out[3] = c3; }
}

void hb_gen_leaves(size_t n, uint64_t key, uint64_t first_index, uint32_t *packed_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t j = 0; j < n; j++) {
    Game g = load_game(0x00000FFFu, 0xFFF00000u, 0u, 0u);
    uint64_t pid = first_index + j;
    uint32_t prefix = 1 + mulhi(philox_block(key, pid, kDomainLeaf, 0).v[0], 100u);
    for (uint32_t p = 0; p < prefix; p++) {
      uint32_t t = 1 + p;
      uint32_t r = philox_block(key, pid, kDomainLeaf, t >> 2).v[t & 3];
      if (random_ply<kOrderCanonical>(g, r) != kRunning) break;
    }
    store_game(g, packed_out + 4 * j);
  }
}

// experiment check: the two-level origin-major select returns the same (origin, slot) as the binary search
uint64_t hb_select_mismatches(uint64_t seed, uint64_t trials) {
  uint64_t bad = 0, x = seed * 0x9E3779B97F4A7C15ull + 1;
  auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (uint64_t t = 0; t < trials; t++) {
    uint32_t a[4];
    const uint64_t density = next() % 5;
    for (int d = 0; d < 4; d++) {
      a[d] = (uint32_t)next();
      for (uint64_t r = 0; r < density; r++) a[d] &= (uint32_t)next();
    }
    const int n = popc(a[0]) + popc(a[1]) + popc(a[2]) + popc(a[3]);
    for (int k = 0; k < n; k++) bad += select_origin_major(a, k) != select_origin_major_two_level(a, k);
  }
  return bad;
}

void hb_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  Philox4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; i++) out[i] = o.v[i];
}

}  // extern "C"
