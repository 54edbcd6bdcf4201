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

// ---- multi-hop captures ----------------------------------------------------------------------
// Exact number of complete capture sequences of a MAN standing on (or landing on) square s, for
// all 32 squares at once.  Men only hop upward (UL = +7 through E1, UR = +9 through E0), the
// board is static during a sequence, so the sequences below s form a DAG of height <= 3 and
//   count(s) = 1                                   if no hop leaves s
//            = [E1(s)] count(s+7) + [E0(s)] count(s+9)   otherwise
// Three rounds of a bit-sliced adder (values <= 8: planes t[0..3]) replace the reference's
// recursion (genLocCaptureReg, src/state.cu:283-340) -- ~35 LOP3/SHF for the whole board.
struct ManCounts {
  uint32_t t[4];
};

B2P_HD ManCounts man_leaf_counts(uint32_t E0, uint32_t E1) {
  const uint32_t none = ~(E0 | E1);
  // height <= 1: 1 or 2
  const uint32_t b1 = E0 & E1, b0 = ~b1;
  // height <= 2
  uint32_t a0 = E1 & (b0 >> 7), a1 = E1 & (b1 >> 7);
  uint32_t d0 = E0 & (b0 >> 9), d1 = E0 & (b1 >> 9);
  uint32_t cy = a0 & d0;
  const uint32_t s0 = (a0 ^ d0) | none;
  const uint32_t s1 = a1 ^ d1 ^ cy;
  const uint32_t s2 = (a1 & d1) | (cy & (a1 ^ d1));
  // height <= 3
  a0 = E1 & (s0 >> 7); a1 = E1 & (s1 >> 7);
  const uint32_t a2 = E1 & (s2 >> 7);
  d0 = E0 & (s0 >> 9); d1 = E0 & (s1 >> 9);
  const uint32_t d2 = E0 & (s2 >> 9);
  ManCounts m;
  cy = a0 & d0;
  m.t[0] = (a0 ^ d0) | none;
  m.t[1] = a1 ^ d1 ^ cy;
  cy = (a1 & d1) | (cy & (a1 ^ d1));
  m.t[2] = a2 ^ d2 ^ cy;
  m.t[3] = (a2 & d2) | (cy & (a2 ^ d2));
  return m;
}

B2P_HD int man_count_at(const ManCounts &m, int s) {
  return (int)(((m.t[0] >> s) & 1u) | (((m.t[1] >> s) & 1u) << 1) | (((m.t[2] >> s) & 1u) << 2) | (((m.t[3] >> s) & 1u) << 3));
}

// The rare path: some capture can be extended by a second hop.  Picks sequence number
// mulhi32(r, n) (reversed for `reverse`) of the normalised canonical list.  Men are counted and
// selected through the bit-sliced DAG counts; only a KING that can hop twice falls back to the
// register DFS (visited-square rule, src/state.cu:134-139).
B2P_HD void pick_multi_hop_capture(const Pos &p, const JumpMasks &jm, const uint32_t cap[4], uint32_t r, bool reverse,
                                   uint32_t &from, uint32_t &to, uint32_t &captured) {
  const uint32_t K = p.kings;
  const uint32_t anyJ = jm.j[0] | jm.j[1] | jm.j[2] | jm.j[3];
  const uint32_t land_king = jumpUR(cap[0] & K) | jumpUL(cap[1] & K) | jumpDR(cap[2]) | jumpDL(cap[3]);
  if (land_king & anyJ) {
    // one DFS pass: remember the first kLeafBuf sequences, pick by index afterwards (a second
    // walk is needed only when the chosen index lies beyond the buffer -- practically never)
    constexpr int kLeafBuf = 12;
    uint32_t buf_captured[kLeafBuf];
    uint16_t buf_from_to[kLeafBuf];
    int seen = 0;
    const int n = for_each_capture(p, jm, [&](const CaptureMove &cm) {
      if (seen < kLeafBuf) {
        buf_captured[seen] = cm.captured;
        buf_from_to[seen] = (uint16_t)(cm.from | (cm.to << 5));
      }
      seen++;
      return false;
    });
    int k = (int)mulhi(r, (uint32_t)n);
    if (reverse) k = n - 1 - k;
    if (k < kLeafBuf) {
      from = 1u << (buf_from_to[k] & 31);
      to = 1u << (buf_from_to[k] >> 5);
      captured = buf_captured[k];
    } else {
      from = to = captured = 0;
      for_each_capture(p, jm, [&](const CaptureMove &cm) {
        if (k-- != 0) return false;
        from = 1u << cm.from; to = 1u << cm.to; captured = cm.captured;
        return true;
      });
    }
    return;
  }
  const uint32_t E0 = jm.j[0], E1 = jm.j[1];
  const ManCounts mc = man_leaf_counts(E0, E1);
  const uint32_t M = (cap[0] | cap[1]) & ~K;
  const int n = popc(mc.t[0] & M) + 2 * popc(mc.t[1] & M) + 4 * popc(mc.t[2] & M) + 8 * popc(mc.t[3] & M) +
                popc(cap[0] & K) + popc(cap[1] & K) + popc(cap[2]) + popc(cap[3]);
  int k = (int)mulhi(r, (uint32_t)n);
  if (reverse) k = n - 1 - k;
  uint32_t origins = cap[0] | cap[1] | cap[2] | cap[3];
  int o;
  bool king;
  for (;;) {
    o = lowbit(origins);
    king = (K >> o) & 1u;
    const int w = king ? (int)(((cap[0] >> o) & 1u) + ((cap[1] >> o) & 1u) + ((cap[2] >> o) & 1u) + ((cap[3] >> o) & 1u))
                       : man_count_at(mc, o);
    if (k < w) break;
    k -= w;
    origins &= origins - 1;
  }
  from = 1u << o;
  if (king) {
    uint32_t nib = ((cap[0] >> o) & 1u) | (((cap[1] >> o) & 1u) << 1) | (((cap[2] >> o) & 1u) << 2) | (((cap[3] >> o) & 1u) << 3);
    if (k >= 1) nib &= nib - 1;
    if (k >= 2) nib &= nib - 1;
    if (k >= 3) nib &= nib - 1;
    const int d = lowbit(nib);
    to = 1u << jump_target(o, d);
    captured = 1u << step_target(o, d);
  } else {
    int cur = o;
    captured = 0;
    while (((E0 | E1) >> cur) & 1u) {
      const int left = ((E1 >> cur) & 1u) ? man_count_at(mc, cur + 7) : 0;  // UL subtree first (left before right)
      if (k < left) {
        captured |= 1u << step_target(cur, 1);
        cur += 7;
      } else {
        k -= left;
        captured |= 1u << step_target(cur, 0);
        cur += 9;
      }
    }
    to = 1u << cur;
  }
}

// ---- one ply ----------------------------------------------------------------------------------
// Everything a ply needs to know about the position, computed once: 4 step pre-images of the
// empty set, 4 jump masks, capture origins per direction.
struct PlyMasks {
  uint32_t e[4];    // squares from which one step in direction d reaches an empty square
  JumpMasks jm;
  uint32_t cap[4];  // mover's squares with a first hop in direction d
  bool capture;     // captures are mandatory (src/state.cu:239-245)
};

B2P_HD PlyMasks ply_masks(const Pos &p) {
  PlyMasks m;
  const uint32_t empty = ~(p.own | p.opp);
  const uint32_t ownK = p.own & p.kings;
  m.e[0] = stepDL(empty); m.e[1] = stepDR(empty); m.e[2] = stepUL(empty); m.e[3] = stepUR(empty);
  m.jm.j[0] = stepDL(p.opp & m.e[0]);
  m.jm.j[1] = stepDR(p.opp & m.e[1]);
  m.jm.j[2] = stepUL(p.opp & m.e[2]);
  m.jm.j[3] = stepUR(p.opp & m.e[3]);
  m.cap[0] = p.own & m.jm.j[0];
  m.cap[1] = p.own & m.jm.j[1];
  m.cap[2] = ownK & m.jm.j[2];
  m.cap[3] = ownK & m.jm.j[3];
  m.capture = (m.cap[0] | m.cap[1] | m.cap[2] | m.cap[3]) != 0;
  return m;
}

// Direct moves and single-hop captures share one branch-free path (four origin masks -> popc ->
// rank -> k-th bit): the warp stays converged whether or not a lane must capture.
// Returns the number of legal moves (0: the mover is stuck; outputs untouched).
template <int ORDER>
B2P_HD int pick_single_hop(const Pos &p, const PlyMasks &m, uint32_t turn, uint32_t r, uint32_t &from, uint32_t &to,
                           uint32_t &captured) {
  const uint32_t ownK = p.own & p.kings;
  const bool capture = m.capture;
  uint32_t a[4];
  if (ORDER == kOrderCanonical) {
    // reference slot order at one origin: capturing man UL,UR (left first); everything else UR,UL,DR,DL
    const uint32_t men = ~p.kings;
    const uint32_t c0 = (m.cap[1] & men) | (m.cap[0] & p.kings), c1 = (m.cap[0] & men) | (m.cap[1] & p.kings);
    a[0] = capture ? c0 : (p.own & m.e[0]);
    a[1] = capture ? c1 : (p.own & m.e[1]);
  } else {
    a[0] = capture ? m.cap[0] : (p.own & m.e[0]);
    a[1] = capture ? m.cap[1] : (p.own & m.e[1]);
  }
  a[2] = capture ? m.cap[2] : (ownK & m.e[2]);
  a[3] = capture ? m.cap[3] : (ownK & m.e[3]);
  const int n0 = popc(a[0]), n1 = popc(a[1]), n2 = popc(a[2]);
  const int n = n0 + n1 + n2 + popc(a[3]);
  if (n == 0) return 0;
  int k = (int)mulhi(r, (uint32_t)n);
  // the move as single-bit MASKS (origin, square stepped to / jumped over, landing square): no bit index ->
  // arithmetic -> 1 << index round trip (+3 % playouts/s, profiles/r02j_ab.txt)
  int slot;
  if (ORDER == kOrderFast) {
    from = select_dir_major_mask(a, n0, n1, n2, k, slot);
  } else {
    if (turn) k = n - 1 - k;
    const int sel = select_origin_major(a, k);
    from = 1u << (sel & 31);
    slot = (sel >> 5) & 3;
    if (capture && !(p.kings & from)) slot ^= 1;  // capturing men try UL before UR (src/state.cu:326-337)
  }
  const uint32_t mid = step_mask(from, slot);
  to = capture ? jump_mask(from, slot) : mid;
  captured = capture ? mid : 0u;
  return n;
}

B2P_HD void finish_ply(Game &g, bool capture, uint32_t from, uint32_t to, uint32_t captured) {
  g.msc = capture ? 0u : g.msc + 1u;
  Pos q = g.pos;
  apply_move(q, from, to, captured);
  g.pos = flip(q);
  g.turn ^= 1u;
}

// Shape of a capture ply:
//   0  every capture is a single hop: one list entry per (origin, first direction);
//   1  some piece hops on, but never with a choice: men along forced chains (<= 3 hops), kings with at
//      most one forced second hop.  Still exactly one sequence per first hop, in first-hop order -- pick
//      the first hop on the common branch-free path, then follow the forced chain (a few instructions);
//   2  a landing square offers a choice, or a king could hop a third time: full enumeration.
// A man makes at most 3 hops, so its choices can only arise on the first two landing squares.  For a king
// the hop back over the piece it just jumped is never available on its FIRST landing square (its origin
// is still occupied on the unmodified board) and is explicitly excluded on the second.
B2P_HD int capture_shape(const Pos &p, const JumpMasks &jm, const uint32_t cap[4]) {
  const uint32_t K = p.kings, men = ~K;
  const uint32_t up = jm.j[0] | jm.j[1], down = jm.j[2] | jm.j[3];
  uint32_t full = 0, multi = 0;  // single exit, bitwise accumulation: keeps the warp on one path
  const uint32_t land_king = jumpUR(cap[0] & K) | jumpUL(cap[1] & K) | jumpDR(cap[2]) | jumpDL(cap[3]);
  if (land_king & (up | down)) {
    multi = 1;
    // two or more onward hops from a first landing square?
    full = land_king & ((jm.j[0] & jm.j[1]) | (jm.j[2] & jm.j[3]) | (up & down));
    // a third hop from a second landing square (arrival direction d, so opp(d) = 3 - d is excluded)?
    const uint32_t b0 = jumpUR(land_king & jm.j[0]), b1 = jumpUL(land_king & jm.j[1]);
    const uint32_t b2 = jumpDR(land_king & jm.j[2]), b3 = jumpDL(land_king & jm.j[3]);
    full |= (b0 & (up | jm.j[2])) | (b1 & (up | jm.j[3])) | (b2 & (down | jm.j[0])) | (b3 & (down | jm.j[1]));
  }
  const uint32_t l1 = jumpUR(cap[0] & men) | jumpUL(cap[1] & men);
  const uint32_t l2 = jumpUR(l1 & jm.j[0]) | jumpUL(l1 & jm.j[1]);
  multi |= l1 & up;
  full |= (jm.j[0] & jm.j[1]) & (l1 | l2);
  int shape = full ? 2 : (multi ? 1 : 0);
  B2P_PIN_INT(shape);  // opaque: the callers' common path must not be cloned per shape
  return shape;
}

// A capture that starts with a hop onto the single-bit mask `land` on a shape-1 ply: follows the forced
// continuation (men: up to two more hops, never a choice; kings: one forced second hop -- shape 1 guarantees there
// is no third), accumulating the jumped squares.  All single-bit masks, no bit indices.
B2P_HD void follow_forced_chain(const JumpMasks &jm, bool man, uint32_t &land, uint32_t &cap) {
  if (!man) {
    const uint32_t nib = ((jm.j[0] & land) ? 1u : 0u) | ((jm.j[1] & land) ? 2u : 0u) | ((jm.j[2] & land) ? 4u : 0u) |
                         ((jm.j[3] & land) ? 8u : 0u);
    if (nib) {
      const int d2 = lowbit(nib);
      cap |= step_mask(land, d2);
      land = jump_mask(land, d2);
    }
  } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int hop = 0; hop < 2; hop++) {  // a man makes at most 3 hops
      if (!((jm.j[0] | jm.j[1]) & land)) break;
      const int d2 = (jm.j[1] & land) ? 1 : 0;  // 1 = UL, 0 = UR (exactly one is available)
      cap |= step_mask(land, d2);
      land = jump_mask(land, d2);
    }
  }
}

// shapes 0 and 1: the move list has one entry per first hop (or step).  Returns the number of legal moves.
template <int ORDER>
B2P_HD int pick_first_hop_and_chain(const Pos &p, const PlyMasks &m, int shape, uint32_t turn, uint32_t r, uint32_t &from,
                                    uint32_t &to, uint32_t &captured) {
  const int n = pick_single_hop<ORDER>(p, m, turn, r, from, to, captured);
  if (n != 0 && shape == 1) follow_forced_chain(m.jm, !(p.kings & from), to, captured);
  return n;
}

// One ply with a uniformly random legal move chosen by the 32-bit draw r:
// rank j = mulhi32(r, n) in the list order ORDER (bitboard.cuh).  Returns kRunning, or the
// winner when the game is over BEFORE a move is made: -1 if msc >= 50 (the draw test wins
// over "no moves", src/state.cpp:20-23), else the player who is not to move.
template <int ORDER>
B2P_HD int random_ply(Game &g, uint32_t r) {
  if (g.msc >= kDrawPlies) return -1;
  const Pos p = g.pos;
  const PlyMasks m = ply_masks(p);
  uint32_t from, to, captured;
  // evaluated by every lane (all-zero masks when there is no capture): a lane-dependent branch around it
  // would split the warp before the common selection path and run that path twice
  const int shape = capture_shape(p, m.jm, m.cap);
  if (shape == 2) {
    pick_multi_hop_capture(p, m.jm, m.cap, r, ORDER == kOrderCanonical && g.turn != 0, from, to, captured);
  } else {
    if (pick_first_hop_and_chain<ORDER>(p, m, shape, g.turn, r, from, to, captured) == 0) return (int)(g.turn ^ 1u);
  }
  finish_ply(g, m.capture, from, to, captured);
  return kRunning;
}

// ---- heuristic playouts ---------------------------------------------------------------------
// reference: HostHeuristicPlayoutDriver::runPlayouts, src/heuristicPlayout.cpp:20-45 with
// scoreMove / getWeight, src/heuristic.cu:33-49.  For every legal move, in list order,
//   weight = float(my + 3*promoted) / float(opp - value of captured pieces) + noise
// and the FIRST maximum (strict '>') is played.  Material is recomputed from the board
// (popc) instead of being carried incrementally -- same value by construction.
//
// Candidate i (canonical list position) of the ply takes its noise from 16-bit half i&1 of word
// (i>>1)&3 of noise block i>>3 (philox.cuh): 8 candidates per Philox call.
// NoiseBlock: Philox4 block(int b);  Gauss: float gauss(uint32_t h16).
//
// Three kinds of ply, cheapest first:
//  SCAN  no capture (3 plies in 4).  All non-crowning candidates share ONE base weight, so the winner among
//        them is the candidate with the largest noise DRAW: gauss() is strictly increasing in the 16-bit draw
//        (steps of >= 4.2e-6 per unit, tests/test_host_logic.py) and float addition is monotone; a base weight is
//        at most 48 (eleven kings and a crowning man against one man), so every sum lies below 64 where one ulp
//        is 3.8e-6: distinct draws give distinct sums.  The scan is therefore an INTEGER
//        max over keys (draw << 8 | 255 - index): no table look-up, no float work per candidate; the table is
//        read once for the winner and once per crowning candidate (a handful per game).
//  LOOP  a capture ply with two or more sequences, all single hops or forced chains (shape 0 / 1): a plain loop
//        over the few candidates in list order, weight and noise in float exactly as the reference spells it.
//        (A ply with ONE list entry -- most capture plies -- weighs nothing: the move is forced.)
//  DFS   a capture ply with a choice after the first hop (shape 2, ~1 % of plies): register DFS.
struct HeurBest {
  float w;
  int idx;
};

B2P_HD bool heur_better(float w, int idx, const HeurBest &b) { return w > b.w || (w == b.w && idx < b.idx); }

B2P_HD uint32_t noise_half(const Philox4 &b, int q) {  // 16-bit draw of candidate q (0..7) of a noise block
  uint32_t r = b.v[0];
  r = (q >> 1) == 1 ? b.v[1] : r;
  r = (q >> 1) == 2 ? b.v[2] : r;
  r = (q >> 1) == 3 ? b.v[3] : r;
  return (q & 1) ? r >> 16 : r & 0xFFFFu;
}

// slots (directions, reference order) of the four candidate masks present at origin o, as a 4-bit value
B2P_HD uint32_t slots_at(const uint32_t a[4], int o) {
  return ((a[0] >> o) & 1u) | (((a[1] >> o) & 1u) << 1) | (((a[2] >> o) & 1u) << 2) | (((a[3] >> o) & 1u) << 3);
}

// Staged for SIMT: every lane of the calling group walks through the same stages, and the group is
// explicitly rejoined (B2P_REJOIN) before the converged ones.  Without the rejoin points ptxas keeps lanes
// that took different rare paths apart and runs the scan and the index -> move mapping once per sub-group.
// `lanes` = mask of the lanes that call this function together (device; ignored on the host).  No early
// return before the last rejoin point.
// Ratio: float ratio(uint32_t a, uint32_t b) = the correctly rounded IEEE quotient float(a) / float(b)
// (getWeight, src/heuristic.cu:44-49); the kernels serve it from a shared-memory table of all material pairs.
template <class NoiseBlock, class Gauss, class Ratio>
B2P_HD int heuristic_ply(Game &g, unsigned lanes, NoiseBlock &&noise_block, Gauss &&gauss, Ratio &&ratio) {
  (void)lanes;
  const Pos p = g.pos;
  const PlyMasks m = ply_masks(p);
  const bool drawn = g.msc >= kDrawPlies;
  const uint32_t my = material(p.own, p.kings), his = material(p.opp, p.kings);
  const uint32_t ownMen = p.own & ~p.kings;
  const uint32_t ownK = p.own & p.kings;
  const bool rev = g.turn != 0;  // canonical index = n-1-normalised index for PLAYER_2
  const int shape = capture_shape(p, m.jm, m.cap);
  // one list entry per (origin, slot): same four masks as the random path, canonical slot order
  const uint32_t men = ~p.kings;
  uint32_t a[4];
  a[0] = m.capture ? ((m.cap[1] & men) | (m.cap[0] & p.kings)) : (p.own & m.e[0]);
  a[1] = m.capture ? ((m.cap[0] & men) | (m.cap[1] & p.kings)) : (p.own & m.e[1]);
  a[2] = m.capture ? m.cap[2] : (ownK & m.e[2]);
  a[3] = m.capture ? m.cap[3] : (ownK & m.e[3]);
  const int n = popc(a[0]) + popc(a[1]) + popc(a[2]) + popc(a[3]);
  const bool dfs = !drawn && shape == 2;
  // with one list entry there is nothing to weigh: the move is forced (most capture plies)
  // ... and when the opponent has no piece left every weight is my / 0 = +inf: the first list entry wins
  const bool choice = !drawn && !dfs && n > 1 && his != 0u;
  const bool loop = choice && m.capture;
  const int n_scan = (choice && !loop) ? n : 0;  // candidates handled by the converged integer scan
  uint32_t from = 0, to = 0, captured = 0;
  HeurBest best;
  best.w = -__builtin_inff();
  best.idx = 0;  // no choice: list entry 0

  // noise block 0 serves candidates 0..7 -- all there are on most plies
  const Philox4 blk0 = noise_block(0);

  // ---- stage 1 (divergent, rare): capture plies with a choice -----------------------------------------------
  if (dfs) {
    // ~1 % of plies: one register-DFS pass buffers the sequences, then they are scored in list order
    int cached = 0;
    Philox4 nb = blk0;
    auto noise = [&](int idx) {
      const int b = idx >> 3;
      if (b != cached) { nb = noise_block(b); cached = b; }
      return gauss(noise_half(nb, idx & 7));
    };
    constexpr int kLeafBuf = 12;
    uint32_t buf_captured[kLeafBuf];
    uint16_t buf_from_to[kLeafBuf];
    int seen = 0;
    const int total = for_each_capture(p, m.jm, [&](const CaptureMove &cm) {
      if (seen < kLeafBuf) {
        buf_captured[seen] = cm.captured;
        buf_from_to[seen] = (uint16_t)(cm.from | (cm.to << 5));
      }
      seen++;
      return false;
    });
    auto score = [&](int i, int f, int t, uint32_t cap) {
      const int idx = rev ? total - 1 - i : i;
      const uint32_t promo = (((ownMen >> f) & 1u) && t >= 28) ? 3u : 0u;
      const uint32_t loss = (uint32_t)(popc(cap) + 3 * popc(cap & p.kings));
      const float w = ratio(my + promo, his - loss) + noise(idx);
      if (heur_better(w, idx, best)) { best.w = w; best.idx = idx; from = 1u << f; to = 1u << t; captured = cap; }
    };
    for (int i = 0; i < total && i < kLeafBuf; i++) score(i, buf_from_to[i] & 31, buf_from_to[i] >> 5, buf_captured[i]);
    if (total > kLeafBuf) {
      int i = 0;
      for_each_capture(p, m.jm, [&](const CaptureMove &cm) {
        if (i >= kLeafBuf) score(i, cm.from, cm.to, cm.captured);
        i++;
        return false;
      });
    }
  } else if (loop) {
    // a few % of plies: two or more captures (single hops or forced chains).  ONE loop body per candidate, walked in list order by peeling (origin, slot) pairs: every lane in
    // here runs the same instructions, only the trip count differs.  Scores only; the move is built in stage 3.
    uint32_t origins = a[0] | a[1] | a[2] | a[3];
    int o = lowbit(origins);
    uint32_t nib = slots_at(a, o);
    Philox4 nb = blk0;
    int cached = 0;
    for (int k = 0; k < n; k++) {
      const int slot = lowbit(nib);
      const uint32_t fbit = 1u << o;
      const bool man = (ownMen & fbit) != 0u;
      const int d = (m.capture && man) ? (slot ^ 1) : slot;
      const uint32_t mid = step_mask(fbit, d);
      uint32_t land = m.capture ? jump_mask(fbit, d) : mid;
      uint32_t cap = m.capture ? mid : 0u;
      if (shape == 1) follow_forced_chain(m.jm, man, land, cap);
      const int idx = rev ? n - 1 - k : k;
      const uint32_t promo = (man && (land & 0xF0000000u)) ? 3u : 0u;
      const uint32_t loss = (uint32_t)(popc(cap) + 3 * popc(cap & p.kings));
      if ((idx >> 3) != cached) { cached = idx >> 3; nb = noise_block(cached); }  // more than 8 candidates: rare
      const float w = ratio(my + promo, his - loss) + gauss(noise_half(nb, idx & 7));
      if (heur_better(w, idx, best)) { best.w = w; best.idx = idx; }
      // next (origin, slot) pair
      nib &= nib - 1;
      if (nib == 0u) {
        origins &= origins - 1;
        o = origins ? lowbit(origins) : 0;
        nib = slots_at(a, o);
      }
    }
  }

  // ---- stage 2 (converged): direct-move plies, integer scan of the noise draws ---------------------------
  B2P_REJOIN(lanes);
  {
    // crowning candidates (men stepping onto the last row) carry another base weight: a second integer maximum
    uint64_t crown_idx = 0;
    if (n_scan > 0) {
      for (uint32_t special = (a[0] | a[1]) & ownMen & 0x0F000000u; special; special &= special - 1) {
        const int o = lowbit(special);
        const uint32_t below = (1u << o) - 1u;
        const int i = popc(a[0] & below) + popc(a[1] & below) + popc(a[2] & below) + popc(a[3] & below);
        const int cnt = (int)((a[0] >> o) & 1u) + (int)((a[1] >> o) & 1u);  // a man has slots 0 and 1 only
        crown_idx |= (uint64_t)(cnt == 2 ? 3u : 1u) << i;                   // normalised positions i (, i + 1)
      }
      if (rev) {
        // canonical index = n - 1 - normalised index: reverse the low n bits
        const uint64_t r = ((uint64_t)brev((uint32_t)crown_idx) << 32) | (uint64_t)brev((uint32_t)(crown_idx >> 32));
        crown_idx = r >> (64 - n_scan);
      }
    }
    const uint64_t all = n_scan >= 64 ? ~0ull : ((1ull << n_scan) - 1ull);
    const uint64_t plain = all & ~crown_idx;
    uint32_t key0 = 0, key1 = 0;
    for (int b = 0; 8 * b < n_scan; b++) {
      const Philox4 blk = b == 0 ? blk0 : noise_block(b);
      const uint32_t plain8 = (uint32_t)(plain >> (8 * b)) & 0xFFu, crown8 = (uint32_t)(crown_idx >> (8 * b)) & 0xFFu;
      const uint32_t tail = 255u - 8u * (uint32_t)b;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int q = 0; q < 8; q++) {
        const uint32_t word = blk.v[q >> 1];
        const uint32_t h8 = (q & 1) ? ((word >> 8) & 0xFFFF00u) : ((word << 8) & 0xFFFF00u);
        const uint32_t key = h8 | (tail - (uint32_t)q);
        const uint32_t c0 = ((plain8 >> q) & 1u) ? key : 0u, c1 = ((crown8 >> q) & 1u) ? key : 0u;
        key0 = c0 > key0 ? c0 : key0;
        key1 = c1 > key1 ? c1 : key1;
      }
    }
    if (n_scan > 0) {
      // two float evaluations settle the ply: the best plain and the best crowning candidate
      const float s0 = key0 ? ratio(my, his) + gauss(key0 >> 8) : -__builtin_inff();
      const float s1 = key1 ? ratio(my + 3u, his) + gauss(key1 >> 8) : -__builtin_inff();
      const int i0 = 255 - (int)(key0 & 0xFFu), i1 = 255 - (int)(key1 & 0xFFu);
      const bool take1 = s1 > s0 || (s1 == s0 && i1 < i0);
      best.idx = take1 ? i1 : i0;
    }
  }

  // ---- stage 3 (converged): winning list position -> move --------------------------------------------------
  B2P_REJOIN(lanes);
  if (!dfs) {
    const int pick = rev ? n - 1 - best.idx : best.idx;
    const int sel = select_origin_major(a, n > 0 ? pick : 0);
    const uint32_t fbit = 1u << (sel & 31);
    const int slot = (sel >> 5) & 3;
    const bool man = (ownMen & fbit) != 0u;
    const int d = (m.capture && man) ? (slot ^ 1) : slot;
    const uint32_t mid = step_mask(fbit, d);
    uint32_t land = m.capture ? jump_mask(fbit, d) : mid;
    uint32_t cap = m.capture ? mid : 0u;
    if (shape == 1) follow_forced_chain(m.jm, man, land, cap);
    from = fbit;
    to = land;
    captured = cap;
  }

  // ---- stage 4: outcome ---------------------------------------------------------------------------------
  if (drawn) return -1;
  if (n == 0) return (int)(g.turn ^ 1u);
  finish_ply(g, m.capture, from, to, captured);
  return kRunning;
}

}  // namespace b2p
