// gpu_ai_b200/csrc/playout_core.cuh
//
// One ply of a playout on the register-resident bitboard.  This is the body of the canonical
// host loop of the reference,
//     while (!state.isGameOver()) { move = pick(state.getMoves()); state.move(move); }
//     result = state.getWinner();
// (HostPlayoutDriver::runPlayouts, src/playout.cpp:17-32; isGameOver/getWinner,
// src/state.cpp:16-23), NOT of the reference's single/coarse kernels, whose hop-by-hop
// capture handling deviates from the host rules.
//
// Everything a lane owns while it plays: 3 board words + turn + draw counter + ply counter.
#pragma once

#include "bitboard.cuh"
#include "philox.cuh"

namespace b2p {

constexpr uint32_t kDrawPlies = 50;  // NUM_DRAW_MOVES, src/state.hpp:14
constexpr int kRunning = 2;          // result code "not finished"; winners are -1 (none), 0, 1

struct Game {
  Pos pos;        // mover-normalised board
  uint32_t turn;  // 0 = PLAYER_1 to move, 1 = PLAYER_2
  uint32_t msc;   // movesSinceLastCapture
};

B2P_HD Game load_game(uint32_t p1, uint32_t p2, uint32_t kings, uint32_t meta) {
  Game g;
  g.turn = meta & 1u;
  g.msc = meta >> 8;
  kings &= p1 | p2;
  if (g.turn == 0) { g.pos.own = p1; g.pos.opp = p2; g.pos.kings = kings; }
  else { g.pos.own = brev(p2); g.pos.opp = brev(p1); g.pos.kings = brev(kings); }
  return g;
}

B2P_HD void store_game(const Game &g, uint32_t out[4]) {
  if (g.turn == 0) { out[0] = g.pos.own; out[1] = g.pos.opp; out[2] = g.pos.kings; }
  else { out[0] = brev(g.pos.opp); out[1] = brev(g.pos.own); out[2] = brev(g.pos.kings); }
  out[3] = g.turn | ((g.msc > 0xFFFFFFu ? 0xFFFFFFu : g.msc) << 8);
}

// One ply with a uniformly random legal move chosen by the 32-bit draw r:
// rank j = mulhi32(r, n) in the list order ORDER (bitboard.cuh).  Returns kRunning, or the
// winner when the game is over BEFORE a move is made: -1 if msc >= 50 (the draw test wins
// over "no moves", src/state.cpp:20-23), else the player who is not to move.
template <int ORDER>
B2P_HD int random_ply(Game &g, uint32_t r) {
  if (g.msc >= kDrawPlies) return -1;
  const Pos p = g.pos;
  const JumpMasks jm = jump_masks(p);
  uint32_t a[4];
  capture_origins(p, jm, a);
  uint32_t from, to, captured;
  if ((a[0] | a[1] | a[2] | a[3]) != 0) {
    if (!any_second_hop(p, jm, a)) {
      // every capture is a single hop: one list entry per (origin, direction)
      if (ORDER == kOrderCanonical) {
        // reference slot order at one origin: man UL,UR (left first); king UR,UL,DR,DL
        const uint32_t men = ~p.kings;
        const uint32_t s0 = (a[1] & men) | (a[0] & p.kings);
        const uint32_t s1 = (a[0] & men) | (a[1] & p.kings);
        a[0] = s0; a[1] = s1;
      }
      const int n0 = popc(a[0]), n1 = popc(a[1]), n2 = popc(a[2]);
      const int n = n0 + n1 + n2 + popc(a[3]);
      int k = (int)mulhi(r, (uint32_t)n);
      int sel;
      if (ORDER == kOrderCanonical) {
        if (g.turn) k = n - 1 - k;
        sel = select_origin_major(a, k);
      } else {
        sel = select_dir_major(a, n0, n1, n2, k);
      }
      const int o = sel & 31;
      int d = sel >> 5;
      if (ORDER == kOrderCanonical && !((p.kings >> o) & 1u)) d ^= 1;
      from = 1u << o;
      to = 1u << jump_target(o, d);
      captured = 1u << step_target(o, d);
    } else {
      // multi-hop sequences exist: count, then walk to the chosen one
      const int n = for_each_capture(p, jm, [](const CaptureMove &) { return false; });
      int k = (int)mulhi(r, (uint32_t)n);
      if (ORDER == kOrderCanonical && g.turn) k = n - 1 - k;
      from = to = captured = 0;
      for_each_capture(p, jm, [&](const CaptureMove &cm) {
        if (k-- != 0) return false;
        from = 1u << cm.from; to = 1u << cm.to; captured = cm.captured;
        return true;
      });
    }
    g.msc = 0;
  } else {
    step_origins(p, a);
    const int n0 = popc(a[0]), n1 = popc(a[1]), n2 = popc(a[2]);
    const int n = n0 + n1 + n2 + popc(a[3]);
    if (n == 0) return (int)(g.turn ^ 1u);
    int k = (int)mulhi(r, (uint32_t)n);
    int sel;
    if (ORDER == kOrderCanonical) {
      if (g.turn) k = n - 1 - k;
      sel = select_origin_major(a, k);
    } else {
      sel = select_dir_major(a, n0, n1, n2, k);
    }
    const int o = sel & 31, d = sel >> 5;
    from = 1u << o;
    to = 1u << step_target(o, d);
    captured = 0;
    g.msc++;
  }
  Pos q = p;
  apply_move(q, from, to, captured);
  g.pos = flip(q);
  g.turn ^= 1u;
  return kRunning;
}

// ---- heuristic playouts ---------------------------------------------------------------------
// reference: HostHeuristicPlayoutDriver::runPlayouts, src/heuristicPlayout.cpp:20-45 with
// scoreMove / getWeight, src/heuristic.cu:33-49.  For every legal move, in list order,
//   weight = float(my + 3*promoted) / float(opp - value of captured pieces) + noise
// and the FIRST maximum (strict '>') is played.  Material is recomputed from the board
// (popc) instead of being carried incrementally -- same value by construction.
// NoiseFn: float noise(int canonical_index).
template <class NoiseFn>
B2P_HD int heuristic_ply(Game &g, NoiseFn &&noise) {
  if (g.msc >= kDrawPlies) return -1;
  const Pos p = g.pos;
  const JumpMasks jm = jump_masks(p);
  uint32_t a[4];
  capture_origins(p, jm, a);
  const uint32_t my = material(p.own, p.kings), his = material(p.opp, p.kings);
  const uint32_t ownMen = p.own & ~p.kings;
  const bool rev = g.turn != 0;  // canonical index = n-1-normalised index for PLAYER_2
  uint32_t from = 0, to = 0, captured = 0;
  float best = -__builtin_inff();
  if ((a[0] | a[1] | a[2] | a[3]) != 0) {
    const int n = for_each_capture(p, jm, [](const CaptureMove &) { return false; });
    int i = 0;
    for_each_capture(p, jm, [&](const CaptureMove &cm) {
      const uint32_t promo = (((ownMen >> cm.from) & 1u) && cm.to >= 28) ? 3u : 0u;
      const uint32_t loss = (uint32_t)(popc(cm.captured) + 3 * popc(cm.captured & p.kings));
      const float w = (float)(my + promo) / (float)(his - loss) + noise(rev ? n - 1 - i : i);
      // first maximum in canonical order: ascending scan keeps '>', the reversed scan (PLAYER_2) takes '>='
      if (rev ? (w >= best) : (w > best)) { best = w; from = 1u << cm.from; to = 1u << cm.to; captured = cm.captured; }
      i++;
      return false;
    });
    g.msc = 0;
  } else {
    step_origins(p, a);
    const int n = popc(a[0]) + popc(a[1]) + popc(a[2]) + popc(a[3]);
    if (n == 0) return (int)(g.turn ^ 1u);
    const float plain = (float)my / (float)his, crowned = (float)(my + 3u) / (float)his;
    uint32_t origins = a[0] | a[1] | a[2] | a[3];
    int i = 0;
    while (origins) {
      const int o = lowbit(origins);
      origins &= origins - 1;
      const bool man = (ownMen >> o) & 1u;
      for (int d = 0; d < 4; d++) {
        if (!((a[d] >> o) & 1u)) continue;
        const int t = step_target(o, d);
        const float w = ((man && t >= 28) ? crowned : plain) + noise(rev ? n - 1 - i : i);
        if (rev ? (w >= best) : (w > best)) { best = w; from = 1u << o; to = 1u << t; }
        i++;
      }
    }
    g.msc++;
  }
  Pos q = p;
  apply_move(q, from, to, captured);
  g.pos = flip(q);
  g.turn ^= 1u;
  return kRunning;
}

}  // namespace b2p
