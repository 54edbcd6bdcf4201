// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the UNMODIFIED reference rules.  This file is compiled (by
// oracle/Makefile, only where /root/reference exists) together with the reference's own
// src/state.cu, src/state.cpp, src/heuristic.cu, src/playout.cpp, src/heuristicPlayout.cpp,
// src/player.cpp and src/mcts.cpp -- taken from where they lie, never copied -- into
// oracle/_ref/libref_harness.so.  The harness owns only (a) the packed<->State conversion,
// (b) the Philox move chooser of chooser.h and (c) the loops around the reference calls;
// every rule (genMoves, move, isGameOver, getWinner, scoreMove, getWeight, the host playout
// drivers) is the reference's own object code.
//
// It pins oracle/checkers_oracle.c (tests/test_oracle.py) and produces the golden vectors
// under tests/golden/ (tools/make_golden.py).

#include "state.hpp"      // /root/reference/src (via -I)
#include "heuristic.hpp"  // /root/reference/src
#include "playout.hpp"    // /root/reference/src
#include "genMovesTest.hpp"
#include "mcts.hpp"        // /root/reference/src: GameTree

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

extern "C" {
#include "chooser.h"
}

// The reference's host objects need the five device-side symbols.  In the
// oracle build they are inert: the oracle is CPU-only by construction.
std::vector<PlayerId> DeviceSinglePlayoutDriver::runPlayouts(std::vector<State>) { throw std::runtime_error("oracle build: no device drivers"); }
std::vector<PlayerId> DeviceMultiplePlayoutDriver::runPlayouts(std::vector<State>) { throw std::runtime_error("oracle build: no device drivers"); }
std::vector<PlayerId> DeviceCoarsePlayoutDriver::runPlayouts(std::vector<State>) { throw std::runtime_error("oracle build: no device drivers"); }
std::vector<PlayerId> DeviceHeuristicPlayoutDriver::runPlayouts(std::vector<State>) { throw std::runtime_error("oracle build: no device drivers"); }
bool genMovesTest(State) { throw std::runtime_error("oracle build: no device drivers"); }

namespace {

inline Loc square_loc(int i) {
  int r = i >> 2;
  return Loc((uint8_t)r, (uint8_t)(2 * (i & 3) + (r % 2 == 0)));
}
inline uint32_t loc_square(Loc l) { return (uint32_t)l.row * 4u + l.col / 2u; }

State unpack(const uint32_t w[4]) {
  State s;
  std::memset(&s, 0, sizeof s);
  for (int i = 0; i < 32; i++) {
    Loc l = square_loc(i);
    BoardItem &b = s.board[l.row][l.col];
    uint32_t bit = 1u << i;
    if ((w[0] | w[1]) & bit) {
      b.occupied = true;
      b.owner = (w[0] & bit) ? PLAYER_1 : PLAYER_2;
      b.type = (w[2] & bit) ? CHECKER_KING : CHECKER;
    }
  }
  s.turn = (w[3] & 1u) ? PLAYER_2 : PLAYER_1;
  s.movesSinceLastCapture = w[3] >> 8;
  return s;
}

void pack(const State &s, uint32_t w[4]) {
  w[0] = w[1] = w[2] = 0;
  for (int i = 0; i < 32; i++) {
    Loc l = square_loc(i);
    const BoardItem &b = s.board[l.row][l.col];
    if (!b.occupied) continue;
    (b.owner == PLAYER_1 ? w[0] : w[1]) |= 1u << i;
    if (b.type == CHECKER_KING) w[2] |= 1u << i;
  }
  unsigned m = s.movesSinceLastCapture > 0xFFFFFFu ? 0xFFFFFFu : s.movesSinceLastCapture;
  w[3] = (s.turn == PLAYER_2 ? 1u : 0u) | (m << 8);
}

uint64_t encode(const Move &m) {
  uint64_t e = loc_square(m.from) | ((uint64_t)loc_square(m.to) << 5) | ((uint64_t)(m.jumps & 7u) << 10) |
               ((uint64_t)(m.promoted ? 1 : 0) << 13);
  for (int k = 0; k < m.jumps && k < 7; k++) e |= (uint64_t)loc_square(m.intermediate[k]) << (16 + 5 * k);
  return e;
}

// same definition as fast_order_pick() in checkers_oracle.c, on reference types
int fast_order_pick(const State &s, const Move *mv, int n, int j) {
  bool full = false;
  for (int i = 0; i < n && !full; i++) {
    if (mv[i].jumps >= 3 && s[mv[i].from].type == CHECKER_KING) full = true;
    for (int k = i + 1; k < n && !full; k++)
      if (mv[i].jumps >= 1 && mv[k].jumps >= 1 && mv[i].from == mv[k].from && mv[i].intermediate[0] == mv[k].intermediate[0]) full = true;
  }
  if (full) return s.turn == PLAYER_1 ? j : n - 1 - j;
  std::vector<int> key(n);
  for (int i = 0; i < n; i++) {
    const Loc first = mv[i].jumps ? mv[i].intermediate[0] : mv[i].to;  // first hop (or the step)
    int dr = first.row > mv[i].from.row ? 1 : -1, dc = first.col > mv[i].from.col ? 1 : -1;
    if (s.turn == PLAYER_2) { dr = -dr; dc = -dc; }
    int dir = dr > 0 ? (dc > 0 ? 0 : 1) : (dc > 0 ? 2 : 3);
    int origin = (int)loc_square(mv[i].from);
    if (s.turn == PLAYER_2) origin = 31 - origin;
    key[i] = dir * 32 + origin;
  }
  for (int i = 0; i < n; i++) {
    int below = 0;
    for (int k = 0; k < n; k++) below += key[k] < key[i];
    if (below == j) return i;
  }
  return -1;
}

int outcome(const State &s, int n) {  // State::isGameOver + getWinner semantics on an already generated list
  if (s.movesSinceLastCapture >= NUM_DRAW_MOVES) return -1;
  if (n == 0) return (int)s.getNextTurn();
  return 2;
}

int random_playout(State &s, uint64_t key, uint64_t pid, uint32_t domain, uint32_t first_draw, int order,
                   int max_plies, uint32_t *plies_out) {
  Move mv[MAX_MOVES];
  uint32_t ply = 0;
  int res;
  for (;;) {
    int n = s.genMoves(mv);
    res = outcome(s, n);
    if (res != 2) {
      // cross-check the shortcut against the reference's own predicates
      if (!s.isGameOver() || (int)s.getWinner() != res) throw std::logic_error("outcome() disagrees with reference");
      break;
    }
    if (max_plies >= 0 && (int)ply >= max_plies) break;
    uint32_t j = ch_mulhi32(ch_draw(key, pid, domain, first_draw + ply), (uint32_t)n);
    int pick = order == 1 ? fast_order_pick(s, mv, n, (int)j) : (int)j;
    s.move(mv[pick]);
    ply++;
  }
  if (plies_out) *plies_out = ply;
  return res;
}

int heuristic_playout(State &s, uint64_t key, uint64_t pid, int max_plies, uint32_t *plies_out) {
  Move mv[MAX_MOVES];
  unsigned stateScore[NUM_PLAYERS];
  scoreState(s, stateScore);
  uint32_t ply = 0;
  int res;
  for (;;) {
    int n = s.genMoves(mv);
    res = outcome(s, n);
    if (res != 2) break;
    if (max_plies >= 0 && (int)ply >= max_plies) break;
    std::vector<int> d(2 * n);
    int best = -1;
    float bestW = -INFINITY;
    for (int i = 0; i < n; i++) {
      int ms[NUM_PLAYERS];
      scoreMove(s, mv[i], ms);
      d[2 * i] = ms[0];
      d[2 * i + 1] = ms[1];
      float w = getWeight(s, stateScore, ms) + ch_gauss_sigma(ch_noise_draw(key, pid, ply, (uint32_t)i));
      if (w > bestW) { bestW = w; best = i; }
    }
    s.move(mv[best]);
    stateScore[0] += d[2 * best];
    stateScore[1] += d[2 * best + 1];
    ply++;
  }
  if (plies_out) *plies_out = ply;
  return res;
}

uint64_t perft(const State &s, int depth) {
  Move mv[MAX_MOVES];
  int n = s.genMoves(mv);
  if (depth == 1) return (uint64_t)n;
  uint64_t t = 0;
  for (int i = 0; i < n; i++) {
    State c = s;
    c.move(mv[i]);
    t += perft(c, depth - 1);
  }
  return t;
}

}  // namespace

extern "C" {

void ref_layout(int out[8]) {
  out[0] = (int)sizeof(State);
  out[1] = (int)sizeof(Move);
  out[2] = (int)sizeof(BoardItem);
  out[3] = (int)offsetof(State, turn);
  out[4] = (int)offsetof(State, movesSinceLastCapture);
  out[5] = (int)offsetof(BoardItem, type);
  out[6] = (int)offsetof(BoardItem, owner);
  out[7] = (int)sizeof(PlayerId);
}

void ref_start_state776(void *out) {
  State s = getStartingState();
  std::memcpy(out, &s, sizeof s);
}

void ref_pack776_batch(const void *states776, size_t n, uint32_t *packed_out) {
  const State *s = (const State *)states776;
  for (size_t i = 0; i < n; i++) pack(s[i], packed_out + 4 * i);
}

void ref_unpack776_batch(const uint32_t *packed, size_t n, void *states776_out) {
  State *s = (State *)states776_out;
  for (size_t i = 0; i < n; i++) s[i] = unpack(packed + 4 * i);
}

void ref_genmoves_batch(const uint32_t *packed, size_t n, int max_moves, uint64_t *moves_out, uint8_t *counts_out) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    State s = unpack(packed + 4 * i);
    std::vector<Move> mv = s.getMoves();  // State::getMoves, src/state.cpp:10-14
    counts_out[i] = (uint8_t)mv.size();
    for (size_t k = 0; k < mv.size() && (int)k < max_moves; k++) moves_out[i * (size_t)max_moves + k] = encode(mv[k]);
  }
}

void ref_playouts_batch(const uint32_t *packed, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base, int mode,
                        int order, int max_plies, int8_t *winners_out, uint32_t *plies_out, uint32_t *final_out,
                        uint64_t counters_out[4]) {
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  size_t total = n * (size_t)reps;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c0, c1, c2, c3)
  for (size_t w = 0; w < total; w++) {
    State s = unpack(packed + 4 * (w % n));
    uint32_t plies = 0;
    int res = mode == 1 ? heuristic_playout(s, key, pid_base + w, max_plies, &plies)
                        : random_playout(s, key, pid_base + w, CH_DOMAIN_RANDOM, 0, order, max_plies, &plies);
    if (winners_out) winners_out[w] = (int8_t)res;
    if (plies_out) plies_out[w] = plies;
    if (final_out) pack(s, final_out + 4 * w);
    if (res == -1) c0++; else if (res == 0) c1++; else if (res == 1) c2++;
    c3 += plies;
  }
  if (counters_out) { counters_out[0] = c0; counters_out[1] = c1; counters_out[2] = c2; counters_out[3] = c3; }
}

void ref_gen_leaves(size_t n, uint64_t key, uint64_t first_index, uint32_t *packed_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t j = 0; j < n; j++) {
    State s = getStartingState();
    uint64_t pid = first_index + j;
    int prefix = 1 + (int)ch_mulhi32(ch_draw(key, pid, CH_DOMAIN_LEAF, 0), 100u);
    random_playout(s, key, pid, CH_DOMAIN_LEAF, 1, 0, prefix, nullptr);
    pack(s, packed_out + 4 * j);
  }
}

uint64_t ref_perft_packed(const uint32_t packed[4], int depth) {
  State s = unpack(packed);
  return depth <= 0 ? 1 : perft(s, depth);
}

// The reference's own host drivers, RNG and all (glibc rand() / default_random_engine):
// the statistical oracle and the "reference" CPU baseline.
// mode 0: HostPlayoutDriver (src/playout.cpp:17-32); mode 1: HostHeuristicPlayoutDriver
// (src/heuristicPlayout.cpp:12-48).  winners_out: PlayerId values (-1/0/1).
int ref_host_driver_run(const uint32_t *packed, size_t n, int mode, int32_t *winners_out) {
  try {
    std::vector<State> states(n);
    for (size_t i = 0; i < n; i++) states[i] = unpack(packed + 4 * i);
    std::vector<PlayerId> res;
    if (mode == 1) {
      HostHeuristicPlayoutDriver d;
      res = d.runPlayouts(states);
    } else {
      HostPlayoutDriver d;
      res = d.runPlayouts(states);
    }
    for (size_t i = 0; i < n; i++) winners_out[i] = (int32_t)res[i];
    return 0;
  } catch (...) {
    return -1;
  }
}

// ---- the reference's GameTree (src/mcts.hpp, src/mcts.cpp), for pinning b2p_tree (tests/test_tree.py) ----
struct RefTree {
  std::shared_ptr<GameTree> tree;
};

void *ref_tree_create(const uint32_t packed_root[4]) {
  RefTree *t = new RefTree();
  t->tree = std::make_shared<GameTree>(unpack(packed_root));
  return t;
}

void ref_tree_destroy(void *h) { delete (RefTree *)h; }

size_t ref_tree_select(void *h, unsigned trials, uint32_t *packed_out) {
  std::vector<State> leaves = ((RefTree *)h)->tree->select(trials);
  for (size_t i = 0; i < leaves.size(); i++) pack(leaves[i], packed_out + 4 * i);
  return leaves.size();
}

void ref_tree_update(void *h, const int8_t *winners, size_t n) {
  std::vector<PlayerId> r(n);
  for (size_t i = 0; i < n; i++) r[i] = (PlayerId)winners[i];
  ((RefTree *)h)->tree->update(r);
}

uint64_t ref_tree_total(void *h) { return ((RefTree *)h)->tree->getTotalTrials(); }

uint64_t ref_tree_best_move(void *h, int player) { return encode(((RefTree *)h)->tree->getOptMove((PlayerId)player)); }

// GameTree::move with the move given as its compact record (matched against the root's own list)
int ref_tree_move(void *h, uint64_t move) {
  RefTree *t = (RefTree *)h;
  for (const Move &m : t->tree->state.getMoves())
    if (encode(m) == move) {
      t->tree = t->tree->move(m);
      return 0;
    }
  return -1;
}

void ref_tree_root_state(void *h, uint32_t packed_out[4]) { pack(((RefTree *)h)->tree->state, packed_out); }

}  // extern "C"
