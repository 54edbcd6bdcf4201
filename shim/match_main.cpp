// shim/match_main.cpp -- BASELINE config 5: game-play tournaments against the reference's own `mcts_host` player.
//
//   match_b200 <mode> <games> <seconds A> <seconds B (whole seconds)> [batch] [reps] [scale] [policy: 1 UCT (default), 0 reference]
//
//   mode b200    side A = a Player that searches with b2p_tree_search_ex on the B200s (<seconds A> per move, fractional)
//   mode hybrid  side A = the reference's `mcts_hybrid` preset (MCTSPlayer(50, 0.02, T, HybridPlayoutDriver(1.2)),
//                src/player.cpp:173-175) running on the drop-in (shim/playout_shim.cpp + shim/hybrid_b200.cpp)
//   mode device  side A = the reference's `mcts_device_multiple` preset (MCTSPlayer(50, 0.02, T, DeviceMultiple...))
//   mode optimal side A = the reference's `mcts_optimal` preset (MCTSPlayer(50, 0.004, T, OptimalPlayoutDriver), its bandit
//                over {host, device_multiple, hybrid(host, device_coarse)} unchanged, src/playout.cpp:85-172)
//   side B is always the reference's `mcts_host` preset (MCTSPlayer(50, 0, T, HostPlayoutDriver), src/player.cpp:164-166).
// The presets' 7-second move time (an `unsigned` number of seconds slept in getMove, src/player.cpp:99) is replaced
// by the command-line budgets; everything else -- tree, pondering worker thread, playout drivers -- is the
// reference's own object code (built against the UNMODIFIED sources by shim/Makefile).  Colours alternate.
// One JSON line per game (winner, plies, playouts per move of BOTH sides: the reference players report their tree
// size through their verbose output, src/player.cpp:101-107) and a summary line.
#include "player.hpp"   // reference
#include "playout.hpp"  // reference
#include "state.hpp"    // reference

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../include/b2p.h"

namespace {

b2p_move_t encode(const Move &m) {
  auto sq = [](Loc l) { return (b2p_move_t)(l.row * 4 + l.col / 2); };
  b2p_move_t e = sq(m.from) | (sq(m.to) << 5) | ((b2p_move_t)(m.jumps & 7) << 10) | ((b2p_move_t)(m.promoted ? 1 : 0) << 13);
  for (int k = 0; k < m.jumps && k < 7; k++) e |= sq(m.intermediate[k]) << (16 + 5 * k);
  return e;
}

class B200TreePlayer : public Player {
 public:
  B200TreePlayer(double seconds, uint32_t batch, uint32_t reps, float scale, int policy)
      : seconds(seconds), batch(batch), reps(reps), scale(scale), policy(policy) {
    if (b2p_create(&ctx, nullptr, 0, 12345) != B2P_OK) throw std::runtime_error(b2p_last_error(nullptr));
    reset();
  }
  ~B200TreePlayer() {
    if (tree) b2p_tree_destroy(tree);
    b2p_destroy(ctx);
  }
  std::string getName() const { return "b200_tree"; }
  void start() { reset(); }
  Move getMove(const State &state, bool) {
    (void)state;
    b2p_search_opts o;
    std::memset(&o, 0, sizeof o);
    o.seconds = seconds;
    o.initial_batch = batch;
    o.scale = scale;
    o.max_batch = 1u << 18;
    o.reps = reps;
    o.mode = B2P_MODE_RANDOM;
    o.policy = policy;
    o.key = key++;
    b2p_search_stats st;
    if (b2p_tree_search_ex(ctx, tree, &o, &st) != B2P_OK) throw std::runtime_error(b2p_tree_last_error(tree));
    playouts += st.playouts;
    moves++;
    b2p_tree_stats ts;
    b2p_tree_info(tree, &ts);
    b2p_move_t best;
    const int me = (int)(ts.root_state.meta & 1u);
    // UCT visits moves unevenly: play the most-tried one; the reference allocation goes with the reference's rule
    if ((policy == B2P_POLICY_UCT ? b2p_tree_robust_move(tree, me, &best) : b2p_tree_best_move(tree, me, &best)) != B2P_OK)
      throw std::runtime_error("no best move");
    Move m;
    b2p_expand_move(best, &m);
    return m;
  }
  void move(const Move &m) {
    if (b2p_tree_move(tree, encode(m)) != B2P_OK) throw std::runtime_error(b2p_tree_last_error(tree));
  }
  uint64_t playouts = 0, moves = 0;

 private:
  void reset() {
    if (tree) b2p_tree_destroy(tree);
    State s = getStartingState();
    b2p_state16 packed;
    b2p_pack776(&s, 1, &packed);
    b2p_tree_create(&tree, &packed);
  }
  double seconds;
  uint32_t batch, reps;
  float scale;
  int policy;
  uint64_t key = 1;
  b2p_ctx *ctx = nullptr;
  b2p_tree *tree = nullptr;
};

// a reference MCTSPlayer asked for its move with verbose = true: "Tree size: N" is its trial count at that moment
Move verbose_move(Player &p, const State &s, uint64_t &tree_size_sum, uint64_t &count) {
  std::ostringstream cap;
  std::streambuf *old = std::cout.rdbuf(cap.rdbuf());
  Move m;
  try {
    m = p.getMove(s, true);
  } catch (...) {
    std::cout.rdbuf(old);
    throw;
  }
  std::cout.rdbuf(old);
  const std::string text = cap.str();
  const size_t at = text.find("Tree size: ");
  if (at != std::string::npos) {
    tree_size_sum += std::strtoull(text.c_str() + at + 11, nullptr, 10);
    count++;
  }
  return m;
}

}  // namespace

// a crash in a tournament that runs for an hour should say where
static void on_crash(int sig) {
  void *frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "match_b200: fatal signal, backtrace:\n";
  (void)!write(2, msg, sizeof msg - 1);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}

int main(int argc, char **argv) {
  signal(SIGSEGV, on_crash);
  signal(SIGABRT, on_crash);
  const std::string mode = argc > 1 ? argv[1] : "b200";
  const int games = argc > 2 ? std::atoi(argv[2]) : 2;
  const double seconds_a = argc > 3 ? std::atof(argv[3]) : 1.0;
  const unsigned seconds_b = argc > 4 ? (unsigned)std::atoi(argv[4]) : 1u;
  // search policy of the B200 player: fixed batches of 2048 leaves x 16 playouts scored best in self-play
  // (profiles/r02i_search_policy_selfplay.jsonl); scale > 0 lets the batch grow with the tree
  const uint32_t batch = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 2048, reps = argc > 6 ? (uint32_t)std::atoi(argv[6]) : 16;
  const float scale = argc > 7 ? (float)std::atof(argv[7]) : 0.0f;
  const int policy = argc > 8 ? std::atoi(argv[8]) : B2P_POLICY_UCT;  // 0 = the reference's allocation rule
  {
    // Create the process's CUDA context and load the kernels BEFORE any game clock starts.  The reference's players
    // move after a fixed sleep whatever their worker thread has achieved; with 1-second moves the very first device
    // batch (context creation + module load, about a second in a fresh process) may not be back yet, every root
    // child then has 0 trials, GameTree::getOptMove compares NaN scores, returns an indeterminate Move
    // (src/mcts.cpp:39-55, its assert is compiled out), GameTree::move finds no such child and returns a null tree
    // (src/mcts.cpp:11-25), and the next getMove dereferences it.  (Seen as a segfault in State::operator== at
    // the second move of `mcts_device_multiple` games; the reference's own 7-second moves hide it.)
    b2p_ctx *warm = nullptr;
    if (b2p_create(&warm, nullptr, 0, 1) != B2P_OK) throw std::runtime_error(b2p_last_error(nullptr));
    State s0 = getStartingState();
    b2p_state16 packed;
    b2p_pack776(&s0, 1, &packed);
    for (int mode_i : {B2P_MODE_RANDOM, B2P_MODE_HEURISTIC})
      for (int sched : {B2P_SCHED_THREAD, B2P_SCHED_WARP})  // lazy module loading: touch every kernel the drivers use
        b2p_run_packed(warm, &packed, 1, 8, 1, 0, mode_i, sched, B2P_ORDER_FAST, -1, nullptr, nullptr, nullptr, nullptr);
    b2p_destroy(warm);
  }
  int score[3] = {0, 0, 0};  // A wins, mcts_host wins, draws
  const char *name_a = mode == "b200" ? "b200_tree" : mode == "hybrid" ? "mcts_hybrid(drop-in)" : mode == "optimal" ? "mcts_optimal(drop-in)" : "mcts_device_multiple(drop-in)";
  for (int g = 0; g < games; g++) {
    B200TreePlayer *mine = nullptr;
    std::unique_ptr<Player> players[NUM_PLAYERS];
    const int seat_a = g % 2;  // colours alternate
    if (mode == "b200") {
      mine = new B200TreePlayer(seconds_a, batch, reps, scale, policy);
      players[seat_a] = std::unique_ptr<Player>(mine);
    } else if (mode == "hybrid") {
      players[seat_a] = std::make_unique<MCTSPlayer>(50, 0.02f, (unsigned)seconds_a, std::make_unique<HybridPlayoutDriver>(1.2f));
    } else if (mode == "optimal") {
      players[seat_a] = std::make_unique<MCTSPlayer>(50, 0.004f, (unsigned)seconds_a, std::make_unique<OptimalPlayoutDriver>());
    } else {
      players[seat_a] = std::make_unique<MCTSPlayer>(50, 0.02f, (unsigned)seconds_a, std::make_unique<DeviceMultiplePlayoutDriver>());
    }
    players[1 - seat_a] = std::make_unique<MCTSPlayer>(50, 0.0f, seconds_b, std::make_unique<HostPlayoutDriver>());
    State state = getStartingState();
    for (auto &p : players) p->start();
    int plies = 0;
    uint64_t tree_a = 0, moves_a = 0, tree_b = 0, moves_b = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (!state.isGameOver()) {
      const int seat = (int)state.turn;
      Move m;
      if (seat == seat_a && mine) m = players[seat]->getMove(state, false);
      else if (seat == seat_a) m = verbose_move(*players[seat], state, tree_a, moves_a);
      else m = verbose_move(*players[seat], state, tree_b, moves_b);
      for (auto &p : players) p->move(m);
      state.move(m);
      plies++;
    }
    const PlayerId winner = state.getWinner();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const char *who = winner == PLAYER_NONE ? "draw" : ((int)winner == seat_a ? name_a : "mcts_host");
    score[winner == PLAYER_NONE ? 2 : ((int)winner == seat_a ? 0 : 1)]++;
    const uint64_t a_per_move = mine ? (mine->moves ? mine->playouts / mine->moves : 0) : (moves_a ? tree_a / moves_a : 0);
    std::cout << "{\"game\": " << g << ", \"a\": \"" << name_a << "\", \"a_seat\": \"P" << seat_a + 1 << "\", \"winner\": \"" << who
              << "\", \"plies\": " << plies << ", \"seconds\": " << secs << ", \"a_playouts_per_move\": " << a_per_move
              << ", \"a_metric\": \"" << (mine ? "playouts run by the move's search" : "tree size (trials) when the move was made, pondering included")
              << "\", \"mcts_host_tree_size_per_move\": " << (moves_b ? tree_b / moves_b : 0) << "}" << std::endl;
    for (auto &p : players) p->stop();
  }
  // Wilson 95 % interval of A's score rate (draw = half a win)
  const double n = games, w = score[0] + 0.5 * score[2], z = 1.96;
  const double ph = n > 0 ? w / n : 0, den = 1 + z * z / n, centre = (ph + z * z / (2 * n)) / den,
               half = z * std::sqrt(ph * (1 - ph) / n + z * z / (4 * n * n)) / den;
  std::cout << "{\"summary\": {\"a\": \"" << name_a << "\", \"a_wins\": " << score[0] << ", \"mcts_host_wins\": " << score[1] << ", \"draws\": " << score[2]
            << ", \"games\": " << games << ", \"a_score_rate\": " << ph << ", \"wilson95\": [" << centre - half << ", " << centre + half << "]}"
            << ", \"a_policy\": \"" << (mode == "b200" ? (policy == B2P_POLICY_UCT ? "uct + most-tried move" : "reference allocation + best-rate move") : "reference")
            << "\", \"a_seconds_per_move\": " << seconds_a << ", \"mcts_host_seconds_per_move\": " << seconds_b
            << ", \"mcts_host\": \"reference preset MCTSPlayer(50, 0, T, HostPlayoutDriver), pondering on all host cores\"}" << std::endl;
  return 0;
}
