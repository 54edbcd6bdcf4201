// gpu_ai_b200/csrc/tree.cu -- packed-state MCTS tree: the caller side of the hot path.
//
// Mirrors GameTree of the reference (src/mcts.hpp:12-64, src/mcts.cpp:11-191) decision for decision --
// UCB1 weights, proportional trial allocation with the "extras by descending weight" rule, expansion
// when a leaf is given more than one trial, terminal nodes absorbing their trials, result slicing in
// update -- so that, fed the same playout results, it selects exactly the same leaves in exactly the
// same order (tests/test_tree.py checks this against the reference's own GameTree).  What changes is
// the data path around it:
//   * nodes are 64-byte records (one cache line: packed 16-byte state + statistics) in a block arena that
//     never moves (no shared_ptr graph, no 776-byte copies, no per-node move list: a node's move list is
//     regenerated from its state in the rare places that need it);
//   * select() writes the leaves straight into a caller buffer: the reference allocates and concatenates a
//     vector<State> at every level (src/mcts.cpp:144-156), which caps it at ~3e5 leaves/s;
//   * subtrees that receive no trial are not walked (the reference, and round 1 of this file, recursed into
//     every node of the tree twice per batch just to write assignedTrials = 0; an epoch stamp replaces that);
//   * update() takes per-trial winners OR `reps` playouts per selected leaf as win counts, so one leaf
//     selection is amortised over many GPU playouts;
//   * b2p_tree_search_ex runs the MCTSPlayer::worker loop (src/player.cpp:134-150) as a PIPELINE: leaf
//     selection is spread over host worker threads (disjoint subtrees below the first tree levels), batches go
//     to the GPU asynchronously from pinned buffers (b2p_run_counts_async), and with depth >= 2 the selection
//     and statistics update of one batch overlap the playouts of the previous one (in-flight trials count as
//     visits without wins -- the usual virtual loss).  With depth == 1 it takes exactly the decisions of the
//     strictly serial select -> playouts -> update loop, whatever the number of threads or devices.
// Move generation for node expansion uses the same bitboard code as the kernels (host instantiation of
// bitboard.cuh).  The tree lives on the host exactly as in the reference; the playouts never do.
#include "../../include/b2p.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <sys/mman.h>

#include "bitboard.cuh"
#include "hostpool.h"

using namespace b2p;

namespace {

constexpr int kMaxMoves = 128;
constexpr uint32_t kBlockShift = 16, kBlockNodes = 1u << kBlockShift, kMaxBlocks = 1u << 15;  // 64 Ki nodes = 4 MiB per block
constexpr uint8_t kMovesUnknown = 255;
constexpr int kPipeSlots = 4;  // = the pipeline slots of b2p_run_counts_async

struct Node {
  b2p_state16 state;
  uint64_t total;     // trials counted so far (GameTree::totalTrials); in a pipelined search: finished + in flight
  uint64_t wins[2];
  uint32_t first_child;  // children are contiguous in the arena, child i <-> i-th move of State::getMoves()
  uint32_t assigned;     // trials assigned in the last b2p_tree_select (GameTree::assignedTrials), valid if epoch matches
  uint32_t epoch;
  uint8_t n_children, n_moves, expanded, pad8;
  uint32_t pad32[2];
};
static_assert(sizeof(Node) == 64, "one node per cache line");

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

// number of legal moves (State::genMoves count, src/state.cu:239-245) without building the list
int count_moves(const b2p_state16 &s) {
  Pos p;
  const uint32_t kings = s.kings & (s.p1 | s.p2);
  if ((s.meta & 1u) == 0) { p.own = s.p1; p.opp = s.p2; p.kings = kings; }
  else { p.own = brev(s.p2); p.opp = brev(s.p1); p.kings = brev(kings); }
  const JumpMasks jm = jump_masks(p);
  uint32_t cap[4];
  capture_origins(p, jm, cap);
  int n;
  if (cap[0] | cap[1] | cap[2] | cap[3]) {
    n = for_each_capture(p, jm, [](const CaptureMove &) { return false; });
  } else {
    uint32_t st[4];
    step_origins(p, st);
    n = popc(st[0]) + popc(st[1]) + popc(st[2]) + popc(st[3]);
  }
  return n < kMaxMoves ? n : kMaxMoves;
}

// Successor states of `s` in State::getMoves() order (what node expansion needs: src/mcts.cpp:75-80), written
// straight as packed states.  Positions without a capture -- three quarters of all -- skip the move records
// altogether: in the mover's frame a step list is "origins ascending, slots UR UL DR DL"; PLAYER_2's canonical
// list is that list reversed (bitboard.cuh).  Capture positions go through the move generator.
// Returns the number of legal moves (at most kMaxMoves are written).
int successors(const b2p_state16 &s, b2p_state16 *out) {
  const uint32_t turn = s.meta & 1u, msc = s.meta >> 8;
  const uint32_t kings = s.kings & (s.p1 | s.p2);
  Pos p;
  if (turn == 0) { p.own = s.p1; p.opp = s.p2; p.kings = kings; }
  else { p.own = brev(s.p2); p.opp = brev(s.p1); p.kings = brev(kings); }
  const JumpMasks jm = jump_masks(p);
  uint32_t cap[4];
  capture_origins(p, jm, cap);
  if (cap[0] | cap[1] | cap[2] | cap[3]) {
    b2p_move_t buf[kMaxMoves];
    int cnt = gen_moves_canonical(s.p1, s.p2, kings, turn, buf, kMaxMoves);
    if (cnt > kMaxMoves) cnt = kMaxMoves;
    for (int i = 0; i < cnt; i++) out[i] = apply_record(s, buf[i]);
    return cnt;
  }
  uint32_t a[4];
  step_origins(p, a);
  const int n = popc(a[0]) + popc(a[1]) + popc(a[2]) + popc(a[3]);
  const uint32_t new_msc = msc + 1u > 0xFFFFFFu ? 0xFFFFFFu : msc + 1u;
  const uint32_t meta = (turn ^ 1u) | (new_msc << 8);
  int k = turn == 0 ? 0 : n - 1;
  const int dk = turn == 0 ? 1 : -1;
  for (uint32_t origins = a[0] | a[1] | a[2] | a[3]; origins; origins &= origins - 1) {
    const int o = lowbit(origins);
    const uint32_t fbit = 1u << o;
    for (int d = 0; d < 4; d++) {
      if (!(a[d] & fbit)) continue;
      const uint32_t tbit = 1u << step_target(o, d);
      const uint32_t own = p.own ^ (fbit | tbit);
      uint32_t kg = p.kings & ~fbit;
      if ((p.kings & fbit) | (tbit & 0xF0000000u)) kg |= tbit;
      b2p_state16 &c = out[k];
      if (turn == 0) { c.p1 = own; c.p2 = p.opp; c.kings = kg; }
      else { c.p1 = brev(p.opp); c.p2 = brev(own); c.kings = brev(kg); }
      c.meta = meta;
      k += dk;
    }
  }
  return n;
}

// recycled arena blocks: a search allocates tens of MB of nodes per round, and fresh pages cost a fault each;
// blocks of destroyed / re-rooted trees are kept (up to kBlockCache of them) for the next tree
constexpr size_t kBlockCache = 1024;  // x 4 MiB
std::mutex g_block_mu;
std::vector<void *> g_block_pool;

void *take_block() {
  {
    std::lock_guard<std::mutex> l(g_block_mu);
    if (!g_block_pool.empty()) {
      void *p = g_block_pool.back();
      g_block_pool.pop_back();
      return p;
    }
  }
  const size_t bytes = (size_t)kBlockNodes * 64;
  void *p = std::aligned_alloc(2u << 20, bytes);
#if defined(MADV_HUGEPAGE)
  if (p) madvise(p, bytes, MADV_HUGEPAGE);
#endif
  return p;
}

void give_block(void *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> l(g_block_mu);
    if (g_block_pool.size() < kBlockCache) {
      g_block_pool.push_back(p);
      return;
    }
  }
  std::free(p);
}

// where a thread allocates: a private range inside one arena block
struct alignas(64) Cursor {  // one cache line each: neighbours in a vector are written by different threads
  uint32_t next = 0, end = 0;
  uint64_t made = 0;
};

// Block arena: node ids are stable, blocks never move, so worker threads can expand disjoint subtrees
// concurrently; the only shared step is taking a fresh block (one atomic add per 64 Ki nodes).
struct Arena {
  std::vector<Node *> blocks;
  std::atomic<uint32_t> used{0};
  Arena() : blocks(kMaxBlocks, nullptr) {}
  ~Arena() {
    const uint32_t u = std::min(used.load(), kMaxBlocks);
    for (uint32_t b = 0; b < u; b++) give_block(blocks[b]);
  }
  Arena(const Arena &) = delete;
  Arena &operator=(const Arena &) = delete;
  Node &at(uint32_t id) const { return blocks[id >> kBlockShift][id & (kBlockNodes - 1)]; }
  bool alloc(Cursor &c, uint32_t cnt, uint32_t *first) {
    if (c.next + cnt > c.end) {
      const uint32_t b = used.fetch_add(1);
      if (b >= kMaxBlocks) { used.store(kMaxBlocks); return false; }
      Node *p = (Node *)take_block();
      blocks[b] = p;
      if (!p) return false;
      c.next = b << kBlockShift;
      c.end = c.next + kBlockNodes;
    }
    *first = c.next;
    c.next += cnt;
    c.made += cnt;
    return true;
  }
};

void init_node(Node &n, const b2p_state16 &s) {
  n.state = s;
  n.state.kings &= s.p1 | s.p2;
  n.total = 0;
  n.wins[0] = n.wins[1] = 0;
  n.first_child = 0;
  n.assigned = 0;
  n.epoch = 0;
  n.n_children = 0;
  n.n_moves = kMovesUnknown;
  n.expanded = 0;
  n.pad8 = 0;
  n.pad32[0] = n.pad32[1] = 0;
}

// one node visited by a batch of the pipelined search: its trials are leaves [begin, end) of the batch
struct Visit {
  uint32_t node, begin, end;
};
struct Item {  // a subtree handed to a worker thread
  uint32_t node, trials, off;
  uint32_t worker, vbegin, vend;  // its visit records: [vbegin, vend) of that worker's list for the batch's slot
};
// per-thread scratch of the parallel walks, one cache line apart (the record lists are appended to on every node
// visit: vector headers side by side would ping-pong between the cores)
struct alignas(64) WorkerLists {
  std::vector<Visit> visits[kPipeSlots];
};

struct Batch {
  uint32_t n = 0;
  uint64_t pid_base = 0, key = 0;
  std::vector<Visit> top;                   // visits above the items (handled by the calling thread)
  std::vector<Item> items;                  // disjoint subtrees, ascending leaf offset, covering [0, n)
  std::vector<uint64_t> sum1, sum2;         // per item: wins of its leaf range
  int slot = 0;
  double t_launch = 0;
};

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

}  // namespace

struct b2p_tree {
  Arena *arena = new Arena();
  std::vector<Cursor> cursors = std::vector<Cursor>(1);  // [0] = calling thread, [1 + w] = search worker w
  std::vector<WorkerLists> lists;
  uint32_t root = 0;
  uint32_t epoch = 0;
  std::string err;
  bool out_of_memory = false;
  // pipelined search state; the page-locked leaf / count staging belongs to the context (b2p_slot_staging)
  b2p_state16 *h_leaves[kPipeSlots] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t *h_wins[kPipeSlots] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<uint32_t> pre1, pre2;  // prefix sums of the per-leaf win counts (update scratch)
  Batch batch[kPipeSlots];
  Pool pool;

  ~b2p_tree() { delete arena; }

  Node &at(uint32_t id) const { return arena->at(id); }
  uint64_t node_count() const {
    uint64_t n = 0;
    for (const Cursor &c : cursors) n += c.made;
    return n;
  }

  bool new_root(const b2p_state16 &s) {
    uint32_t id;
    if (!arena->alloc(cursors[0], 1, &id)) return false;
    init_node(at(id), s);
    root = id;
    return true;
  }

  static uint32_t moves_of(Node &n) {
    if (n.n_moves == kMovesUnknown) n.n_moves = (uint8_t)count_moves(n.state);
    return n.n_moves;
  }
  // State::isGameOver (src/state.cpp:16-18)
  static bool game_over(Node &n) { return (n.state.meta >> 8) >= 50u || moves_of(n) == 0; }

  // What a node that was handed `trials` > 0 trials is, in the order GameTree::select asks (src/mcts.cpp:63-91):
  // game over -> all trials are played from it; expanded -> hand the trials down; one trial -> it is the leaf;
  // else expand (children = successor states in State::getMoves() order, src/mcts.cpp:75-80) and hand down.
  // Returns true when the trials go to the children.  A node with one trial is a leaf whether or not the game
  // is over there, so its move list is never looked at; the expansion's move generation doubles as the
  // "no legal move" test.
  bool descend(Node &nd, uint32_t trials, Cursor &cur) {
    if (nd.expanded) return true;  // expanded nodes have moves and were below the draw limit when expanded
    if ((nd.state.meta >> 8) >= 50u || nd.n_moves == 0) return false;
    if (trials == 1) return false;
    b2p_state16 kids[kMaxMoves];
    int cnt = successors(nd.state, kids);
    if (cnt > kMaxMoves) cnt = kMaxMoves;
    nd.n_moves = (uint8_t)cnt;
    if (cnt == 0) return false;
    uint32_t first;
    if (!arena->alloc(cur, (uint32_t)cnt, &first)) {
      out_of_memory = true;
      return false;
    }
    for (int i = 0; i < cnt; i++) init_node(at(first + (uint32_t)i), kids[i]);
    nd.first_child = first;
    nd.n_children = (uint8_t)cnt;
    nd.expanded = 1;
    return true;
  }

  // GameTree::ucb1 (src/mcts.cpp:182-191), same expression types: the exploration term is long double.
  // log_parent = std::log((double)parent.total), the same value for every child of a node.
  static double ucb1(const Node &n, double log_parent) {
    if (n.total == 0) return INFINITY;
    const unsigned turn = n.state.meta & 1u;
    return (double)n.wins[turn] / n.total + std::sqrt(2.0L * log_parent / n.total);
  }

  // The allocation rule of GameTree::select (src/mcts.cpp:93-139) for an expanded node: trials of the untried
  // children first (equal shares), else shares proportional to UCB1, the remainder one by one in descending
  // weight order (first maximum wins).  Returns the number of trials handed out.
  //
  // EXACT = true evaluates UCB1 with the reference's expression types (x87 extended for the exploration term):
  // the serial interface and the depth-1 search use it and take the reference's decisions bit for bit.
  // EXACT = false (pipelined search, depth >= 2, where in-flight trials already make the statistics differ from
  // a serial run) evaluates the same rule in single precision -- a quarter of the cycles.
  template <bool EXACT>
  uint32_t distribute(const Node &nd, uint32_t trials, uint32_t *child_trials) const {
    using W = typename std::conditional<EXACT, double, float>::type;
    const uint32_t nc = nd.n_children, fc = nd.first_child;
    W weights[kMaxMoves];
    W total_w = 0, best_w = -INFINITY;
    uint32_t untried = 0, assigned = 0, first_best = 0;
    const double log_parent = EXACT ? std::log((double)nd.total) : (double)logf((float)nd.total);
    const float two_log = 2.0f * (float)log_parent;
    for (uint32_t i = 0; i < nc; i++) {
      const Node &c = at(fc + i);
      W w;
      if (EXACT) {
        w = (W)ucb1(c, log_parent);
      } else if (c.total == 0) {
        w = INFINITY;
      } else {
        const float t = (float)(int64_t)c.total;
        w = (W)((float)(int64_t)c.wins[c.state.meta & 1u] / t + sqrtf(two_log / t));
      }
      weights[i] = w;
      total_w += w;
      if (c.total == 0) untried++;
      if (w > best_w) { best_w = w; first_best = i; }
      child_trials[i] = 0;
    }
    if (trials == 1) {
      // one trial goes to the first maximum (an untried child weighs +inf): what the general rule below gives --
      // zero shares all round, then one extra -- without the share arithmetic
      child_trials[first_best] = 1;
      return 1;
    }
    for (uint32_t i = 0; i < nc; i++) {
      if (untried > 0) child_trials[i] = at(fc + i).total == 0 ? trials / untried : 0;
      else if (trials > 0) child_trials[i] = (uint32_t)(trials * (weights[i] / total_w));
      if (child_trials[i] > trials - assigned) child_trials[i] = trials - assigned;  // guards a rounded-up share
      assigned += child_trials[i];
    }
    if (assigned < trials) {
      bool extra[kMaxMoves];
      for (uint32_t i = 0; i < nc; i++) extra[i] = false;
      while (assigned < trials) {
        W best = -INFINITY;
        int opt = -1;
        for (uint32_t i = 0; i < nc; i++)
          if (!extra[i] && weights[i] > best) { best = weights[i]; opt = (int)i; }
        if (opt < 0) break;  // the reference asserts here (src/mcts.cpp:135)
        extra[opt] = true;
        child_trials[opt]++;
        assigned++;
      }
    }
    return assigned;
  }

  // B2P_POLICY_UCT: the node's trials go down ONE BY ONE, each to the child with the highest UCB1 value, every
  // assigned trial counting at once as `reps` visits without a win (virtual loss).  This is what a strictly
  // sequential UCT search does with the same budget; the reference's rule above shares a batch out in proportion
  // to the UCB1 values, which are all of the same size (0.4 ... 1.5), i.e. almost uniformly -- with batches of
  // thousands of leaves it explores the first tree levels breadth-first however good or bad a move looks
  // (measured: 300x the playouts of mcts_host buy a 57 % score, profiles/r02m_*).
  uint32_t distribute_uct(const Node &nd, uint32_t trials, uint32_t reps, uint32_t *child_trials) const {
    const uint32_t nc = nd.n_children, fc = nd.first_child;
    double wins[kMaxMoves], tot[kMaxMoves], w[kMaxMoves];
    const double L = 2.0 * std::log((double)(nd.total > 1 ? nd.total : 2) + (double)trials * (double)reps);
    const double step = (double)reps;
    // a move is as good as the share of playouts through it that the player MAKING it wins.  (GameTree::ucb1,
    // src/mcts.cpp:182-191, reads wins[state.turn] of the CHILD -- the opponent's wins; under its near-uniform
    // allocation that hardly shows, and the serial interface reproduces it for parity.  Under UCT it would steer
    // the search towards the mover's worst moves.)
    const unsigned mover = nd.state.meta & 1u;
    for (uint32_t i = 0; i < nc; i++) {
      const Node &c = at(fc + i);
      wins[i] = (double)c.wins[mover];
      tot[i] = (double)c.total;
      child_trials[i] = 0;
    }
    uint32_t left = trials;
    if (trials > 32 * nc) {
      // Many trials: the one-by-one rule is a water-filling -- every child ends up with just enough visits M_i for
      // its value W_i / M_i + sqrt(L / M_i) to sink to a common level lambda.  With y = 1 / sqrt(M) that is the
      // quadratic W y^2 + sqrt(L) y = lambda, so the visits needed at a given level have a closed form; a bisection
      // on lambda finds the highest level that fits the budget, the one-by-one loop below hands out the remainder.
      const double sL = std::sqrt(L);
      auto needed = [&](double lam, uint32_t *out) {
        uint64_t sum = 0;
        for (uint32_t i = 0; i < nc; i++) {
          const double y = wins[i] > 0 ? (std::sqrt(L + 4.0 * wins[i] * lam) - sL) / (2.0 * wins[i]) : lam / sL;
          const double m = 1.0 / (y * y);
          double t = m > tot[i] ? std::ceil((m - tot[i]) / step) : 0.0;
          if (t > (double)trials) t = (double)trials;
          out[i] = (uint32_t)t;
          sum += out[i];
        }
        return sum;
      };
      double lo = 0.0, hi = 1.0 + sL;  // at `hi` no child with a visit needs another one
      uint32_t tmp[kMaxMoves];
      for (int it = 0; it < 18; it++) {  // level to 2e-5: the loop below settles the rest
        const double mid = 0.5 * (lo + hi);
        if (needed(mid, tmp) > (uint64_t)trials) lo = mid;
        else hi = mid;
      }
      if (needed(hi, tmp) <= (uint64_t)trials)
        for (uint32_t i = 0; i < nc; i++) {
          child_trials[i] = tmp[i];
          tot[i] += step * (double)tmp[i];
          left -= tmp[i];
        }
    }
    for (uint32_t i = 0; i < nc; i++) w[i] = tot[i] == 0 ? INFINITY : wins[i] / tot[i] + std::sqrt(L / tot[i]);
    for (; left > 0; left--) {
      uint32_t best = 0;
      double bw = w[0];
      for (uint32_t i = 1; i < nc; i++)
        if (w[i] > bw) { bw = w[i]; best = i; }
      child_trials[best]++;
      tot[best] += step;
      w[best] = wins[best] / tot[best] + std::sqrt(L / tot[best]);
    }
    return trials;
  }

  uint32_t assigned_of(const Node &n) const { return n.epoch == epoch ? n.assigned : 0u; }

  // ---- the reference's serial interface ---------------------------------------------------------------------
  // GameTree::select (src/mcts.cpp:63-157).  Appends the selected leaf states to out.
  void select(uint32_t id, uint32_t trials, b2p_state16 *out, uint32_t &pos) {
    if (trials == 0) return;  // the reference walks the subtree writing assignedTrials = 0: the epoch stamp does that
    Node &nd = at(id);
    nd.epoch = epoch;
    if (!descend(nd, trials, cursors[0])) {
      nd.assigned = trials;
      for (uint32_t i = 0; i < trials; i++) out[pos++] = nd.state;
      return;
    }
    uint32_t child_trials[kMaxMoves];
    nd.assigned = distribute<true>(nd, trials, child_trials);
    const uint32_t nc = nd.n_children, fc = nd.first_child;
    for (uint32_t i = 0; i < nc; i++) select(fc + i, child_trials[i], out, pos);
  }

  // GameTree::update (src/mcts.cpp:159-180): winners of this node's assigned trials start at `pos`.
  // With reps > 1 every assigned trial stands for `reps` playouts laid out [rep][trial] (stride = n).
  void update(uint32_t id, const int8_t *winners, uint32_t n, uint32_t reps, uint32_t &pos) {
    Node &nd = at(id);
    const uint32_t mine = assigned_of(nd), begin = pos;
    if (mine == 0) return;
    nd.total += (uint64_t)mine * reps;
    if (nd.expanded) {
      const uint32_t nc = nd.n_children, fc = nd.first_child;
      for (uint32_t i = 0; i < nc; i++) update(fc + i, winners, n, reps, pos);
    } else {
      pos += mine;
    }
    for (uint32_t r = 0; r < reps; r++)
      for (uint32_t t = begin; t < begin + mine; t++) {
        const int w = winners[(size_t)r * n + t];
        if (w == 0 || w == 1) nd.wins[w]++;
      }
  }

  // update() for `reps` playouts per selected leaf given as win counts per leaf (b2p_run_counts)
  void update_counts(uint32_t id, const uint32_t *wins, uint32_t reps, uint32_t &pos, uint64_t &w1, uint64_t &w2) {
    Node &nd = at(id);
    const uint32_t mine = assigned_of(nd);
    if (mine == 0) return;
    nd.total += (uint64_t)mine * reps;
    uint64_t a = 0, b = 0;
    if (nd.expanded) {
      const uint32_t nc = nd.n_children, fc = nd.first_child;
      for (uint32_t i = 0; i < nc; i++) update_counts(fc + i, wins, reps, pos, a, b);
    } else {
      for (uint32_t t = pos; t < pos + mine; t++) { a += wins[2 * t]; b += wins[2 * t + 1]; }
      pos += mine;
    }
    nd.wins[0] += a;
    nd.wins[1] += b;
    w1 += a;
    w2 += b;
  }

  // GameTree::getScore (src/mcts.cpp:27-37)
  double score(Node &n, int player) const {
    if (game_over(n)) return ((int)((n.state.meta & 1u) ^ 1u) == player) ? 1 : 0;
    return (double)n.wins[player] / n.total;
  }

  int root_move_list(b2p_move_t *buf) const {
    const Node &r = at(root);
    const int cnt = gen_moves_canonical(r.state.p1, r.state.p2, r.state.kings, r.state.meta & 1u, buf, kMaxMoves);
    return cnt < kMaxMoves ? cnt : kMaxMoves;
  }

  // ---- pipelined search: selection ----------------------------------------------------------------------------
  // Same decisions as select() above.  Differences in mechanism only: every trial of the batch has a fixed leaf
  // index (a node's trials are leaves [off, off + trials)), visits are recorded per batch instead of in the
  // nodes, and the trials are added to `total` right away -- after the node's own children were weighted, so a
  // batch never sees its own trials -- so that a batch selected before the previous one has been updated
  // explores elsewhere.
  struct SelCtx {
    Cursor *cur;
    std::vector<Visit> *visits;
    b2p_state16 *leaves;
    uint32_t reps;
    int policy;  // 0: reference rule, reference arithmetic; 1: reference rule, single precision; 2: UCT
  };

  uint32_t hand_down(const Node &nd, uint32_t trials, uint32_t *child_trials, const SelCtx &c) const {
    if (c.policy == 2) return distribute_uct(nd, trials, c.reps, child_trials);
    return c.policy == 0 ? distribute<true>(nd, trials, child_trials) : distribute<false>(nd, trials, child_trials);
  }

  void emit_leaf(Node &nd, uint32_t trials, uint32_t off, const SelCtx &c) {
    for (uint32_t i = 0; i < trials; i++) c.leaves[off + i] = nd.state;
    nd.total += (uint64_t)trials * c.reps;
  }

  // One item = one subtree, walked depth-first: the children a node has just weighed are still in cache when the
  // walk descends into them.  (A breadth-first walk with the visit list as queue and prefetch a few entries
  // ahead was measured 25 % SLOWER: the walk is bound by the allocation arithmetic, not by memory latency.)
  void fast_select(uint32_t id, uint32_t trials, uint32_t off, const SelCtx &c) {
    Node &nd = at(id);
    c.visits->push_back({id, off, off + trials});
    if (!descend(nd, trials, *c.cur)) {
      emit_leaf(nd, trials, off, c);
      return;
    }
    uint32_t child_trials[kMaxMoves];
    const uint32_t given = hand_down(nd, trials, child_trials, c);
    if (given < trials) child_trials[0] += trials - given;  // never in practice (see distribute); keeps leaf indices dense
    nd.total += (uint64_t)trials * c.reps;
    const uint32_t nc = nd.n_children, fc = nd.first_child;
    for (uint32_t i = 0; i < nc; i++) {
      if (child_trials[i] == 0) continue;
      fast_select(fc + i, child_trials[i], off, c);
      off += child_trials[i];
    }
  }

  // the first levels, on the calling thread: cuts the batch into items of at most `grain` trials
  void top_select(uint32_t id, uint32_t trials, uint32_t off, int depth, uint32_t grain, Batch &b, const SelCtx &c) {
    Node &nd = at(id);
    if (trials <= grain || depth >= 48 || !descend(nd, trials, *c.cur)) {
      b.items.push_back({id, trials, off, 0u, 0u, 0u});
      return;
    }
    b.top.push_back({id, off, off + trials});
    uint32_t child_trials[kMaxMoves];
    const uint32_t given = hand_down(nd, trials, child_trials, c);
    if (given < trials) child_trials[0] += trials - given;
    nd.total += (uint64_t)trials * c.reps;
    const uint32_t nc = nd.n_children, fc = nd.first_child;
    for (uint32_t i = 0; i < nc; i++) {
      if (child_trials[i] == 0) continue;
      top_select(fc + i, child_trials[i], off, depth + 1, grain, b, c);
      off += child_trials[i];
    }
  }

  void select_batch(Batch &b, uint32_t n, uint32_t reps, b2p_state16 *leaves, size_t threads, int policy) {
    b.slot = (int)(&b - batch);
    b.n = n;
    b.top.clear();
    b.items.clear();
    if (n == 0) return;
    if (cursors.size() < threads + 1) cursors.resize(threads + 1);
    if (lists.size() < threads) lists.resize(threads);
    const uint32_t grain = std::max<uint32_t>(64u, n / (uint32_t)(threads * 16));
    SelCtx c0{&cursors[0], &b.top, leaves, reps, policy};
    top_select(root, n, 0, 0, threads <= 1 ? n : grain, b, c0);
    b.sum1.assign(b.items.size(), 0);
    b.sum2.assign(b.items.size(), 0);
    std::atomic<size_t> next{0}, worker_id{0};
    auto work = [&]() {
      const size_t me = worker_id.fetch_add(1);
      std::vector<Visit> &mine = lists[me].visits[b.slot];
      mine.clear();
      SelCtx c{&cursors[1 + me], &mine, leaves, reps, policy};
      for (;;) {
        const size_t k = next.fetch_add(1);
        if (k >= b.items.size()) return;
        Item &it = b.items[k];
        it.worker = (uint32_t)me;
        it.vbegin = (uint32_t)mine.size();
        fast_select(it.node, it.trials, it.off, c);
        it.vend = (uint32_t)mine.size();
      }
    };
    pool.run(std::min(threads, b.items.size()), work);
  }

  // ---- pipelined search: statistics update ---------------------------------------------------------------------
  // wins[2 * leaf + p] = playouts from that leaf won by PLAYER_(p+1).  Every visited node receives the wins of its
  // leaf range: per item, a prefix sum over the item's leaves and one subtraction per visit.
  void update_batch(Batch &b, const uint32_t *wins, size_t threads) {
    if (b.n == 0) return;
    if (pre1.size() < b.n) { pre1.resize(b.n); pre2.resize(b.n); }
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (;;) {
        const size_t k = next.fetch_add(1);
        if (k >= b.items.size()) return;
        const Item &it = b.items[k];
        uint32_t a = 0, c = 0;
        for (uint32_t i = it.off; i < it.off + it.trials; i++) {
          a += wins[2 * i];
          c += wins[2 * i + 1];
          pre1[i] = a;
          pre2[i] = c;
        }
        b.sum1[k] = a;
        b.sum2[k] = c;
        const Visit *vs = lists[it.worker].visits[b.slot].data() + it.vbegin;
        const size_t nv = it.vend - it.vbegin;
        for (size_t v = 0; v < nv; v++) {
          if (v + 24 < nv) __builtin_prefetch(&at(vs[v + 24].node), 1);
          const Visit &vi = vs[v];
          Node &nd = at(vi.node);
          nd.wins[0] += pre1[vi.end - 1] - (vi.begin > it.off ? pre1[vi.begin - 1] : 0u);
          nd.wins[1] += pre2[vi.end - 1] - (vi.begin > it.off ? pre2[vi.begin - 1] : 0u);
        }
      }
    };
    pool.run(std::min(threads, b.items.size()), work);
    // nodes above the items: their ranges are unions of whole items
    std::vector<uint64_t> base1(b.items.size() + 1, 0), base2(b.items.size() + 1, 0);
    for (size_t k = 0; k < b.items.size(); k++) {
      base1[k + 1] = base1[k] + b.sum1[k];
      base2[k + 1] = base2[k] + b.sum2[k];
    }
    auto item_at = [&](uint32_t off) {  // index of the item that starts at leaf `off` (items.size() for off == n)
      return (size_t)(std::lower_bound(b.items.begin(), b.items.end(), off, [](const Item &it, uint32_t o) { return it.off < o; }) - b.items.begin());
    };
    for (const Visit &v : b.top) {
      const size_t lo = item_at(v.begin), hi = item_at(v.end);
      Node &nd = at(v.node);
      nd.wins[0] += base1[hi] - base1[lo];
      nd.wins[1] += base2[hi] - base2[lo];
    }
    b.n = 0;  // folded in: the slot is free, and b2p_tree_move may re-root again
  }

  bool batches_pending() const {
    for (const Batch &b : batch)
      if (b.n != 0) return true;
    return false;
  }
};

extern "C" {

int b2p_tree_create(b2p_tree **out, const b2p_state16 *root) {
  if (!out || !root) return B2P_EINVAL;
  b2p_tree *t = new b2p_tree();
  if (!t->new_root(*root)) {
    delete t;
    return B2P_ENOMEM;
  }
  *out = t;
  return B2P_OK;
}

void b2p_tree_destroy(b2p_tree *t) { delete t; }

int b2p_tree_select(b2p_tree *t, uint32_t trials, b2p_state16 *leaves_out, uint32_t *n_out) {
  if (!t || (trials && !leaves_out)) return B2P_EINVAL;
  uint32_t pos = 0;
  t->epoch++;
  t->select(t->root, trials, leaves_out, pos);
  if (n_out) *n_out = pos;
  return B2P_OK;
}

int b2p_tree_update(b2p_tree *t, const int8_t *winners, uint32_t n, uint32_t reps) {
  if (!t || (n && !winners) || reps == 0) return B2P_EINVAL;
  if (t->assigned_of(t->at(t->root)) != n) {
    t->err = "b2p_tree_update: result count does not match the last select";
    return B2P_EINVAL;  // the reference asserts (src/mcts.cpp:160)
  }
  uint32_t pos = 0;
  t->update(t->root, winners, n, reps, pos);
  t->epoch++;  // a result set is consumed once
  return B2P_OK;
}

int b2p_tree_update_counts(b2p_tree *t, const uint32_t *wins, uint32_t n, uint32_t reps) {
  if (!t || (n && !wins) || reps == 0) return B2P_EINVAL;
  if (t->assigned_of(t->at(t->root)) != n) {
    t->err = "b2p_tree_update_counts: result count does not match the last select";
    return B2P_EINVAL;
  }
  uint32_t pos = 0;
  uint64_t a = 0, b = 0;
  t->update_counts(t->root, wins, reps, pos, a, b);
  t->epoch++;
  return B2P_OK;
}

// GameTree::getOptMove (src/mcts.cpp:39-55): the child with the highest score for `player`, first maximum
int b2p_tree_best_move(const b2p_tree *t, int player, b2p_move_t *move_out) {
  if (!t || !move_out || player < 0 || player > 1) return B2P_EINVAL;
  const Node &r = t->at(t->root);
  if (!r.expanded) return B2P_EINVAL;
  double best = -INFINITY;
  int opt = -1;
  for (uint32_t i = 0; i < r.n_children; i++) {
    const double s = t->score(t->at(r.first_child + i), player);
    if (s > best) { best = s; opt = (int)i; }
  }
  if (opt < 0) return B2P_EINVAL;
  b2p_move_t buf[kMaxMoves];
  t->root_move_list(buf);
  *move_out = buf[opt];
  return B2P_OK;
}

// The "robust child": the root move with the most trials (ties: the better score for `player`, then the first).
// The companion of B2P_POLICY_UCT, which visits moves very unevenly: the win rate of a move that was tried a few
// hundred times is noise next to one that was tried millions of times, and GameTree::getOptMove's "highest rate"
// picks exactly such outliers.  (With the reference's near-uniform allocation the two rules agree.)
int b2p_tree_robust_move(const b2p_tree *t, int player, b2p_move_t *move_out) {
  if (!t || !move_out || player < 0 || player > 1) return B2P_EINVAL;
  const Node &r = t->at(t->root);
  if (!r.expanded) return B2P_EINVAL;
  int opt = -1;
  uint64_t best_n = 0;
  double best_s = -INFINITY;
  for (uint32_t i = 0; i < r.n_children; i++) {
    Node &c = t->at(r.first_child + i);
    const double s = c.total ? t->score(c, player) : (b2p_tree::game_over(c) ? t->score(c, player) : 0.0);
    if (opt < 0 || c.total > best_n || (c.total == best_n && s > best_s)) { opt = (int)i; best_n = c.total; best_s = s; }
  }
  if (opt < 0) return B2P_EINVAL;
  b2p_move_t buf[kMaxMoves];
  t->root_move_list(buf);
  *move_out = buf[opt];
  return B2P_OK;
}

// GameTree::move (src/mcts.cpp:11-25): keep the chosen subtree (copied breadth-first into a fresh arena, the
// rest of the old tree is freed), or start a new tree from the successor state when the root was never expanded.
int b2p_tree_move(b2p_tree *t, b2p_move_t move) {
  if (!t) return B2P_EINVAL;
  if (t->batches_pending()) {
    // the visit records of a selected batch name nodes of the arena this call is about to free
    t->err = "b2p_tree_move: a batch selected with b2p_tree_select_batch has not been folded in (b2p_tree_update_batch) yet";
    return B2P_EINVAL;
  }
  const Node &r = t->at(t->root);
  b2p_move_t buf[kMaxMoves];
  const int n_moves = t->root_move_list(buf);
  int idx = -1;
  for (int i = 0; i < n_moves; i++)
    if (buf[i] == move) idx = i;
  if (idx < 0) {
    t->err = "b2p_tree_move: not a legal move of the root";
    return B2P_EINVAL;
  }
  Arena *fresh = new Arena();
  Cursor cur;
  uint32_t new_root = 0;
  bool ok = fresh->alloc(cur, 1, &new_root);
  if (ok && !r.expanded) {
    init_node(fresh->at(new_root), apply_record(r.state, move));
  } else if (ok) {
    // breadth-first copy; children stay contiguous
    std::vector<uint32_t> old_ids{r.first_child + (uint32_t)idx}, new_ids{new_root};
    fresh->at(new_root) = t->at(old_ids[0]);
    for (size_t head = 0; head < old_ids.size() && ok; head++) {
      const Node &old = t->at(old_ids[head]);
      if (!old.expanded) continue;
      uint32_t first;
      if (!(ok = fresh->alloc(cur, old.n_children, &first))) break;
      fresh->at(new_ids[head]).first_child = first;
      for (uint32_t i = 0; i < old.n_children; i++) {
        fresh->at(first + i) = t->at(old.first_child + i);
        old_ids.push_back(old.first_child + i);
        new_ids.push_back(first + i);
      }
    }
  }
  if (!ok) {
    delete fresh;
    t->err = "b2p_tree_move: out of memory";
    return B2P_ENOMEM;
  }
  delete t->arena;
  t->arena = fresh;
  t->cursors.assign(1, cur);
  t->root = new_root;
  t->epoch++;
  return B2P_OK;
}

int b2p_tree_info(const b2p_tree *t, b2p_tree_stats *out) {
  if (!t || !out) return B2P_EINVAL;
  Node &r = t->at(t->root);
  out->nodes = t->node_count();
  out->total_trials = r.total;
  out->wins_p1 = r.wins[0];
  out->wins_p2 = r.wins[1];
  out->root_children = r.expanded ? r.n_children : 0;
  out->root_moves = b2p_tree::moves_of(r);
  out->root_state = r.state;
  return B2P_OK;
}

int b2p_tree_root_moves(const b2p_tree *t, b2p_move_t *moves_out, uint64_t *trials_out, uint64_t *wins_p1_out,
                        uint64_t *wins_p2_out, uint32_t capacity) {
  if (!t) return B2P_EINVAL;
  const Node &r = t->at(t->root);
  b2p_move_t buf[kMaxMoves];
  const uint32_t n_moves = (uint32_t)t->root_move_list(buf);
  for (uint32_t i = 0; i < n_moves && i < capacity; i++) {
    if (moves_out) moves_out[i] = buf[i];
    const bool have = r.expanded;
    if (trials_out) trials_out[i] = have ? t->at(r.first_child + i).total : 0;
    if (wins_p1_out) wins_p1_out[i] = have ? t->at(r.first_child + i).wins[0] : 0;
    if (wins_p2_out) wins_p2_out[i] = have ? t->at(r.first_child + i).wins[1] : 0;
  }
  return (int)n_moves;
}

const char *b2p_tree_last_error(const b2p_tree *t) { return t ? t->err.c_str() : ""; }

// The two host halves of one pipelined round (what b2p_tree_search_ex does around b2p_run_counts_async), for a
// caller that runs the playouts itself.
int b2p_tree_select_batch(b2p_tree *t, int slot, uint32_t trials, uint32_t reps, int threads, int policy, b2p_state16 *leaves_out) {
  if (!t || slot < 0 || slot >= kPipeSlots || reps == 0 || (trials && !leaves_out)) return B2P_EINVAL;
  if (policy < 0 || policy > 2) return B2P_EINVAL;
  t->select_batch(t->batch[slot], trials, reps, leaves_out, (size_t)std::max(1, threads), policy);
  return B2P_OK;
}

int b2p_tree_update_batch(b2p_tree *t, int slot, const uint32_t *wins, int threads) {
  if (!t || slot < 0 || slot >= kPipeSlots) return B2P_EINVAL;
  if (t->batch[slot].n && !wins) return B2P_EINVAL;
  t->update_batch(t->batch[slot], wins, (size_t)std::max(1, threads));
  return B2P_OK;
}

// The worker loop of MCTSPlayer (src/player.cpp:134-150) fused with the playout engine and pipelined:
//   repeat { n = max(initial_batch, scale * leaf selections so far); leaves = select(n);
//            counts = playouts(leaves x reps) [asynchronous]; update(counts of the batch `depth` rounds back) }
// for `iterations` rounds or until `seconds` have passed (whichever comes first; 0 = no limit on that axis).
// Deliberate difference from the reference's batch policy: there the batch is `scale * totalTrials` and only
// falls back to the initial size when that product is 0 (src/player.cpp:137-139), so e.g. mcts_device_coarse
// drops from 4000 to ~4 leaves per launch after its first batch and the GPU idles; here the
// batch never shrinks below initial_batch.
int b2p_tree_search_ex(b2p_ctx *ctx, b2p_tree *t, const b2p_search_opts *o, b2p_search_stats *stats_out) {
  if (!ctx || !t || !o || o->reps == 0 || o->initial_batch == 0) return B2P_EINVAL;
  if (o->iterations == 0 && !(o->seconds > 0)) {
    if (stats_out) std::memset(stats_out, 0, sizeof *stats_out);
    return B2P_OK;
  }
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t threads = o->threads > 0 ? (size_t)o->threads : std::max<size_t>(1, std::min<size_t>(hw ? hw : 1, 64));
  const int depth = o->depth <= 0 ? 2 : std::min(o->depth, kPipeSlots);
  // one launch holds fewer than 2^31 playouts per device; a batch also never exceeds max_batch leaves
  uint64_t cap = o->max_batch ? o->max_batch : (1u << 20);
  cap = std::min<uint64_t>(cap, ((1ull << 31) - 1) / o->reps);
  if (cap == 0) return B2P_EINVAL;
  b2p_search_stats st;
  std::memset(&st, 0, sizeof st);
  const double t0 = now_s();
  uint64_t selected = 0;   // leaf selections so far (finished or in flight)
  int rc = B2P_OK;
  uint32_t launched = 0, retired = 0;
  // page-locked staging for every slot, sized once for the largest batch this search can reach
  {
    uint64_t reach = o->initial_batch;
    if (o->scale > 0) reach = cap;  // the batch grows with the tree
    reach = std::min<uint64_t>(std::max<uint64_t>(reach, 1u << 12), cap);
    for (int s = 0; s < depth; s++)
      if ((rc = b2p_slot_staging(ctx, s, reach, &t->h_leaves[s], &t->h_wins[s])) != B2P_OK) {
        t->err = std::string("b2p_tree_search: ") + b2p_last_error(ctx);
        return rc;
      }
  }

  auto retire = [&]() -> int {  // wait for the oldest batch in flight and fold its results into the tree
    const int slot = (int)(retired % (uint32_t)depth);
    Batch &b = t->batch[slot];
    float kms = 0.f;
    const double w0 = now_s();
    const int r = b2p_wait_slot(ctx, slot, nullptr, &kms);
    const double w1 = now_s();
    st.wait_s += w1 - w0;
    st.kernel_s += 1e-3 * kms;
    if (r != B2P_OK) {
      t->err = std::string("b2p_tree_search: ") + b2p_last_error(ctx);
      return r;
    }
    const uint64_t leaves_in_batch = b.n;
    t->update_batch(b, t->h_wins[slot], threads);
    st.update_s += now_s() - w1;
    st.playouts += leaves_in_batch * o->reps;
    retired++;
    return B2P_OK;
  };

  double iter_s = 0;  // running estimate of the wall time of one round
  for (uint32_t it = 0; o->iterations == 0 || it < o->iterations; it++) {
    const double t_it = now_s();
    if (o->seconds > 0 && t_it - t0 + iter_s >= o->seconds && it > 0) break;
    uint64_t want = (uint64_t)((double)selected * o->scale);
    if (want < o->initial_batch) want = o->initial_batch;
    const uint32_t n = (uint32_t)std::min<uint64_t>(want, cap);
    if (launched - retired == (uint32_t)depth && (rc = retire()) != B2P_OK) break;
    const int slot = (int)(launched % (uint32_t)depth);
    Batch &b = t->batch[slot];
    const double s0 = now_s();
    t->select_batch(b, n, o->reps, t->h_leaves[slot], threads, o->policy == B2P_POLICY_UCT ? 2 : (depth == 1 ? 0 : 1));
    st.select_s += now_s() - s0;
    b.pid_base = selected * o->reps;
    b.key = o->key + it;
    // 16 B per leaf up, 8 B per leaf down: the per-playout winners stay on the device
    rc = b2p_run_counts_async(ctx, slot, t->h_leaves[slot], n, o->reps, b.key, b.pid_base, o->mode, B2P_SCHED_AUTO, B2P_ORDER_FAST,
                              t->h_wins[slot]);
    if (rc != B2P_OK) {
      t->err = std::string("b2p_tree_search: ") + b2p_last_error(ctx);
      // the batch was counted into the nodes' totals at selection: fold it back in as all-draws so that the
      // tree stays consistent (total >= wins); the caller sees the error
      std::memset(t->h_wins[slot], 0, (size_t)n * 2 * sizeof(uint32_t));
      t->update_batch(b, t->h_wins[slot], threads);
      break;
    }
    launched++;
    selected += n;
    st.leaves += n;
    st.batches++;
    const double dt = now_s() - t_it;
    iter_s = it == 0 ? dt : 0.5 * (iter_s + dt);
  }
  while (retired < launched) {
    const int r = retire();
    if (r != B2P_OK) { rc = rc != B2P_OK ? rc : r; break; }
  }
  st.seconds = now_s() - t0;
  st.nodes = t->node_count();
  st.threads = (uint32_t)threads;
  st.depth = (uint32_t)depth;
  if (t->out_of_memory && rc == B2P_OK) t->err = "b2p_tree_search: node arena exhausted; unexpandable leaves were played as they are";
  if (stats_out) *stats_out = st;
  return rc;
}

int b2p_tree_search(b2p_ctx *ctx, b2p_tree *t, uint32_t iterations, double seconds, uint32_t initial_batch, float scale,
                    uint32_t reps, int mode, uint64_t key, uint64_t *playouts_out) {
  b2p_search_opts o;
  std::memset(&o, 0, sizeof o);
  o.iterations = iterations;
  o.seconds = seconds;
  o.initial_batch = initial_batch;
  o.scale = scale;
  o.reps = reps;
  o.mode = mode;
  o.key = key;
  b2p_search_stats st;
  const int rc = b2p_tree_search_ex(ctx, t, &o, &st);
  if (playouts_out) *playouts_out = st.playouts;
  return rc;
}

}  // extern "C"
