// gpu_ai_b200/csrc/tree.cu -- packed-state MCTS tree: the caller side of the hot path.
//
// Mirrors GameTree of the reference (src/mcts.hpp:12-64, src/mcts.cpp:11-191) decision for decision --
// UCB1 weights, proportional trial allocation with the "extras by descending weight" rule, expansion
// when a leaf is given more than one trial, terminal nodes absorbing their trials, result slicing in
// update -- so that, fed the same playout results, it selects exactly the same leaves in exactly the
// same order (tests/test_tree.py checks this against the reference's own GameTree).  What changes is
// the data path around it:
//   * nodes hold 16-byte packed states in one arena (no shared_ptr graph, no 776-byte copies);
//   * select() writes the leaves straight into a caller buffer (pinned staging in b2p_tree_search):
//     the reference allocates and concatenates a vector<State> at every level (src/mcts.cpp:144-156),
//     which caps it at ~3e5 leaves/s;
//   * update() takes per-trial winners OR `reps` playouts per selected leaf, so one leaf selection
//     can be amortised over many GPU playouts (the kernel plays 2e9/s; the tree cannot select that fast).
// Move generation for node expansion uses the same bitboard code as the kernels (host instantiation of
// bitboard.cuh).  The tree lives on the host exactly as in the reference; the playouts never do.
#include "../../include/b2p.h"

#include <cmath>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "bitboard.cuh"

using namespace b2p;

namespace {

constexpr int kMaxMoves = 128;

struct Node {
  b2p_state16 state;
  int32_t parent = -1;
  uint32_t first_child = 0, n_children = 0;  // children are contiguous in the arena
  uint32_t first_move = 0, n_moves = 0;      // canonical move list (State::getMoves order)
  bool expanded = false;
  uint32_t assigned = 0;  // trials assigned in the last select (GameTree::assignedTrials)
  uint64_t total = 0;     // finished trials (GameTree::totalTrials)
  uint64_t wins[2] = {0, 0};
};

// State::move (src/state.cu:57-92) on the packed state, from a b2p_move_t record (absolute frame)
b2p_state16 apply_record(const b2p_state16 &s, b2p_move_t m) {
  const int from = (int)(m & 31), to = (int)((m >> 5) & 31), hops = (int)((m >> 10) & 7);
  const bool promoted = (m >> 13) & 1;
  uint32_t captured = 0;
  int prev = from;
  for (int k = 0; k < hops; k++) {
    const int land = (int)((m >> (16 + 5 * k)) & 31);
    const int delta = land - prev;
    const int d = delta == 9 ? 0 : delta == 7 ? 1 : delta == -7 ? 2 : 3;
    captured |= 1u << step_target(prev, d);
    prev = land;
  }
  const uint32_t turn = s.meta & 1u, msc = s.meta >> 8;
  const uint32_t fbit = 1u << from, tbit = 1u << to;
  b2p_state16 o = s;
  uint32_t &mine = turn == 0 ? o.p1 : o.p2;
  uint32_t &theirs = turn == 0 ? o.p2 : o.p1;
  const bool was_king = (s.kings & fbit) != 0;
  mine = (mine & ~fbit) | tbit;
  theirs &= ~captured;
  o.kings &= ~(captured | fbit);
  if (was_king || promoted) o.kings |= tbit;
  const uint32_t new_msc = hops ? 0u : (msc + 1u > 0xFFFFFFu ? 0xFFFFFFu : msc + 1u);
  o.meta = (turn ^ 1u) | (new_msc << 8);
  return o;
}

}  // namespace

struct b2p_tree {
  std::vector<Node> nodes;
  std::vector<b2p_move_t> moves;
  int32_t root = 0;
  std::string err;

  int32_t add_node(const b2p_state16 &s, int32_t parent) {
    Node n;
    n.state = s;
    n.state.kings &= s.p1 | s.p2;
    n.parent = parent;
    b2p_move_t buf[kMaxMoves];
    const int cnt = gen_moves_canonical(s.p1, s.p2, n.state.kings, s.meta & 1u, buf, kMaxMoves);
    n.first_move = (uint32_t)moves.size();
    n.n_moves = (uint32_t)(cnt < kMaxMoves ? cnt : kMaxMoves);
    moves.insert(moves.end(), buf, buf + n.n_moves);
    nodes.push_back(n);
    return (int32_t)nodes.size() - 1;
  }

  // State::isGameOver (src/state.cpp:16-18)
  bool game_over(const Node &n) const { return n.n_moves == 0 || (n.state.meta >> 8) >= 50u; }

  // GameTree::ucb1 (src/mcts.cpp:182-191), same expression types: the exploration term is long double
  double ucb1(const Node &n) const {
    if (n.total == 0) return INFINITY;
    const Node &p = nodes[n.parent];
    const unsigned turn = n.state.meta & 1u;
    return (double)n.wins[turn] / n.total + std::sqrt(2.0L * std::log((double)p.total) / n.total);
  }

  // GameTree::select (src/mcts.cpp:63-157).  Appends the selected leaf states to out.
  void select(int32_t id, uint32_t trials, b2p_state16 *out, uint32_t &pos) {
    if (game_over(nodes[id])) {
      nodes[id].assigned = trials;
      for (uint32_t i = 0; i < trials; i++) out[pos++] = nodes[id].state;
      return;
    }
    if (!nodes[id].expanded) {
      if (trials > 1) {
        const uint32_t first = (uint32_t)nodes.size(), cnt = nodes[id].n_moves, fm = nodes[id].first_move;
        const b2p_state16 st = nodes[id].state;
        for (uint32_t i = 0; i < cnt; i++) add_node(apply_record(st, moves[fm + i]), id);  // may reallocate `nodes`
        nodes[id].first_child = first;
        nodes[id].n_children = cnt;
        nodes[id].expanded = true;
      } else if (trials == 1) {
        nodes[id].assigned = 1;
        out[pos++] = nodes[id].state;
        return;
      } else {
        nodes[id].assigned = 0;
        return;
      }
    }
    const uint32_t nc = nodes[id].n_children, fc = nodes[id].first_child;
    uint32_t child_trials[kMaxMoves];
    double weights[kMaxMoves];
    bool extra[kMaxMoves];
    double total_w = 0;
    uint32_t untried = 0, assigned = 0;
    for (uint32_t i = 0; i < nc; i++) {
      weights[i] = ucb1(nodes[fc + i]);
      total_w += weights[i];
      if (nodes[fc + i].total == 0) untried++;
      child_trials[i] = 0;
      extra[i] = false;
    }
    for (uint32_t i = 0; i < nc; i++) {
      if (untried > 0) child_trials[i] = nodes[fc + i].total == 0 ? trials / untried : 0;
      else if (trials > 0) child_trials[i] = (uint32_t)(trials * (weights[i] / total_w));
      assigned += child_trials[i];
    }
    while (assigned < trials) {  // extras in descending weight order, first maximum wins
      double best = -INFINITY;
      int opt = -1;
      for (uint32_t i = 0; i < nc; i++)
        if (!extra[i] && weights[i] > best) { best = weights[i]; opt = (int)i; }
      if (opt < 0) break;  // the reference asserts here (src/mcts.cpp:135)
      extra[opt] = true;
      child_trials[opt]++;
      assigned++;
    }
    nodes[id].assigned = assigned;
    for (uint32_t i = 0; i < nc; i++) select((int32_t)(fc + i), child_trials[i], out, pos);
  }

  // GameTree::update (src/mcts.cpp:159-180): winners of this node's assigned trials start at `pos`.
  // With reps > 1 every assigned trial stands for `reps` playouts laid out [rep][trial] (stride = n).
  void update(int32_t id, const int8_t *winners, uint32_t n, uint32_t reps, uint32_t &pos) {
    Node &nd = nodes[id];
    const uint32_t mine = nd.assigned, begin = pos;
    nd.total += (uint64_t)mine * reps;
    if (nd.expanded) {
      const uint32_t nc = nd.n_children, fc = nd.first_child;
      for (uint32_t i = 0; i < nc; i++) update((int32_t)(fc + i), winners, n, reps, pos);
    } else {
      pos += mine;
    }
    Node &me = nodes[id];
    for (uint32_t r = 0; r < reps; r++)
      for (uint32_t t = begin; t < begin + mine; t++) {
        const int w = winners[(size_t)r * n + t];
        if (w == 0 || w == 1) me.wins[w]++;
      }
  }

  // update() for `reps` playouts per selected leaf given as win counts per leaf (b2p_run_counts)
  void update_counts(int32_t id, const uint32_t *wins, uint32_t reps, uint32_t &pos, uint64_t &w1, uint64_t &w2) {
    const uint32_t mine = nodes[id].assigned;
    nodes[id].total += (uint64_t)mine * reps;
    uint64_t a = 0, b = 0;
    if (nodes[id].expanded) {
      const uint32_t nc = nodes[id].n_children, fc = nodes[id].first_child;
      for (uint32_t i = 0; i < nc; i++) update_counts((int32_t)(fc + i), wins, reps, pos, a, b);
    } else {
      for (uint32_t t = pos; t < pos + mine; t++) { a += wins[2 * t]; b += wins[2 * t + 1]; }
      pos += mine;
    }
    nodes[id].wins[0] += a;
    nodes[id].wins[1] += b;
    w1 += a;
    w2 += b;
  }

  // GameTree::getScore (src/mcts.cpp:27-37)
  double score(const Node &n, int player) const {
    if (game_over(n)) return ((int)((n.state.meta & 1u) ^ 1u) == player) ? 1 : 0;
    return (double)n.wins[player] / n.total;
  }
};

extern "C" {

int b2p_tree_create(b2p_tree **out, const b2p_state16 *root) {
  if (!out || !root) return B2P_EINVAL;
  b2p_tree *t = new b2p_tree();
  t->nodes.reserve(1 << 16);
  t->moves.reserve(1 << 18);
  t->root = t->add_node(*root, -1);
  *out = t;
  return B2P_OK;
}

void b2p_tree_destroy(b2p_tree *t) { delete t; }

int b2p_tree_select(b2p_tree *t, uint32_t trials, b2p_state16 *leaves_out, uint32_t *n_out) {
  if (!t || (trials && !leaves_out)) return B2P_EINVAL;
  uint32_t pos = 0;
  t->select(t->root, trials, leaves_out, pos);
  if (n_out) *n_out = pos;
  return B2P_OK;
}

int b2p_tree_update(b2p_tree *t, const int8_t *winners, uint32_t n, uint32_t reps) {
  if (!t || (n && !winners) || reps == 0) return B2P_EINVAL;
  if (t->nodes[t->root].assigned != n) {
    t->err = "b2p_tree_update: result count does not match the last select";
    return B2P_EINVAL;  // the reference asserts (src/mcts.cpp:160)
  }
  uint32_t pos = 0;
  t->update(t->root, winners, n, reps, pos);
  return B2P_OK;
}

int b2p_tree_update_counts(b2p_tree *t, const uint32_t *wins, uint32_t n, uint32_t reps) {
  if (!t || (n && !wins) || reps == 0) return B2P_EINVAL;
  if (t->nodes[t->root].assigned != n) {
    t->err = "b2p_tree_update_counts: result count does not match the last select";
    return B2P_EINVAL;
  }
  uint32_t pos = 0;
  uint64_t a = 0, b = 0;
  t->update_counts(t->root, wins, reps, pos, a, b);
  return B2P_OK;
}

// GameTree::getOptMove (src/mcts.cpp:39-55): the child with the highest score for `player`, first maximum
int b2p_tree_best_move(const b2p_tree *t, int player, b2p_move_t *move_out) {
  if (!t || !move_out || player < 0 || player > 1) return B2P_EINVAL;
  const Node &r = t->nodes[t->root];
  if (!r.expanded) return B2P_EINVAL;
  double best = -INFINITY;
  int opt = -1;
  for (uint32_t i = 0; i < r.n_children; i++) {
    const double s = t->score(t->nodes[r.first_child + i], player);
    if (s > best) { best = s; opt = (int)i; }
  }
  if (opt < 0) return B2P_EINVAL;
  *move_out = t->moves[r.first_move + opt];
  return B2P_OK;
}

// GameTree::move (src/mcts.cpp:11-25): keep the chosen subtree (re-rooted in a fresh arena), or start a
// new tree from the successor state when the root was never expanded.
int b2p_tree_move(b2p_tree *t, b2p_move_t move) {
  if (!t) return B2P_EINVAL;
  const Node &r = t->nodes[t->root];
  int idx = -1;
  for (uint32_t i = 0; i < r.n_moves; i++)
    if (t->moves[r.first_move + i] == move) idx = (int)i;
  if (idx < 0) {
    t->err = "b2p_tree_move: not a legal move of the root";
    return B2P_EINVAL;
  }
  b2p_tree fresh;
  if (!r.expanded) {
    fresh.root = fresh.add_node(apply_record(r.state, move), -1);
  } else {
    // breadth-first copy of the subtree; children stay contiguous
    std::vector<int32_t> queue{(int32_t)(r.first_child + idx)};
    fresh.nodes.reserve(t->nodes.size() / 2 + 16);
    {
      Node n = t->nodes[queue[0]];
      n.parent = -1;
      fresh.nodes.push_back(n);
    }
    for (size_t head = 0; head < queue.size(); head++) {
      const Node old = t->nodes[queue[head]];
      Node &copy = fresh.nodes[head];
      copy.first_move = (uint32_t)fresh.moves.size();
      fresh.moves.insert(fresh.moves.end(), t->moves.begin() + old.first_move, t->moves.begin() + old.first_move + old.n_moves);
      if (old.expanded) {
        const uint32_t first = (uint32_t)fresh.nodes.size();
        fresh.nodes[head].first_child = first;
        for (uint32_t i = 0; i < old.n_children; i++) {
          Node c = t->nodes[old.first_child + i];
          c.parent = (int32_t)head;
          fresh.nodes.push_back(c);
          queue.push_back((int32_t)(old.first_child + i));
        }
      }
    }
    fresh.root = 0;
  }
  t->nodes.swap(fresh.nodes);
  t->moves.swap(fresh.moves);
  t->root = fresh.root;
  return B2P_OK;
}

int b2p_tree_info(const b2p_tree *t, b2p_tree_stats *out) {
  if (!t || !out) return B2P_EINVAL;
  const Node &r = t->nodes[t->root];
  out->nodes = t->nodes.size();
  out->total_trials = r.total;
  out->wins_p1 = r.wins[0];
  out->wins_p2 = r.wins[1];
  out->root_children = r.expanded ? r.n_children : 0;
  out->root_moves = r.n_moves;
  out->root_state = r.state;
  return B2P_OK;
}

int b2p_tree_root_moves(const b2p_tree *t, b2p_move_t *moves_out, uint64_t *trials_out, uint64_t *wins_p1_out,
                        uint64_t *wins_p2_out, uint32_t capacity) {
  if (!t) return B2P_EINVAL;
  const Node &r = t->nodes[t->root];
  for (uint32_t i = 0; i < r.n_moves && i < capacity; i++) {
    if (moves_out) moves_out[i] = t->moves[r.first_move + i];
    const bool have = r.expanded;
    if (trials_out) trials_out[i] = have ? t->nodes[r.first_child + i].total : 0;
    if (wins_p1_out) wins_p1_out[i] = have ? t->nodes[r.first_child + i].wins[0] : 0;
    if (wins_p2_out) wins_p2_out[i] = have ? t->nodes[r.first_child + i].wins[1] : 0;
  }
  return (int)r.n_moves;
}

const char *b2p_tree_last_error(const b2p_tree *t) { return t ? t->err.c_str() : ""; }

// The worker loop of MCTSPlayer (src/player.cpp:134-150) fused with the playout engine:
//   repeat { n = max(initial_batch, scale * leaf selections so far); leaves = select(n);
//            winners = playouts(leaves x reps); update(winners) }
// for `iterations` rounds or until `seconds` have passed (whichever comes first; 0 = no limit on that axis).
// Deliberate difference from the reference's batch policy: there the batch is `scale * totalTrials` and
// only falls back to the initial size when that product is 0 (src/player.cpp:137-139), so e.g.
// mcts_device_coarse drops from 4000 to ~4 leaves per launch after its first batch and the
// GPU idles; here the batch never shrinks below initial_batch.
int b2p_tree_search(b2p_ctx *ctx, b2p_tree *t, uint32_t iterations, double seconds, uint32_t initial_batch, float scale,
                    uint32_t reps, int mode, uint64_t key, uint64_t *playouts_out) {
  if (!ctx || !t || reps == 0 || initial_batch == 0) return B2P_EINVAL;
  std::vector<b2p_state16> leaves;
  std::vector<uint32_t> wins;
  uint64_t played = 0;
  struct timespec ts0;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  for (uint32_t it = 0; iterations == 0 || it < iterations; it++) {
    if (seconds > 0) {
      struct timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      if ((ts.tv_sec - ts0.tv_sec) + 1e-9 * (ts.tv_nsec - ts0.tv_nsec) >= seconds) break;
    } else if (iterations == 0) {
      break;
    }
    const uint64_t total = t->nodes[t->root].total;
    uint32_t n = (uint32_t)((double)(total / reps) * scale);
    if (n < initial_batch) n = initial_batch;
    leaves.resize(n);
    wins.resize((size_t)n * 2);
    uint32_t got = 0;
    t->select(t->root, n, leaves.data(), got);
    // 16 B per leaf up, 8 B per leaf down: the per-playout winners stay on the device
    int rc = b2p_run_counts(ctx, leaves.data(), got, reps, key + it, played, mode, B2P_SCHED_AUTO, B2P_ORDER_FAST, wins.data(), nullptr);
    if (rc != B2P_OK) {
      t->err = std::string("b2p_tree_search: ") + b2p_last_error(ctx);
      return rc;
    }
    uint32_t pos = 0;
    uint64_t a = 0, b = 0;
    t->update_counts(t->root, wins.data(), reps, pos, a, b);
    played += (uint64_t)got * reps;
  }
  if (playouts_out) *playouts_out = played;
  return B2P_OK;
}

}  // extern "C"
