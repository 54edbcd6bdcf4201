// shim/match_main.cpp -- BASELINE config 5, literally: the reference's own `mcts_host` player (getPlayer("mcts_host"):
// MCTSPlayer(50, 0, 7 s, HostPlayoutDriver), src/player.cpp:164-166, pondering on the host cores in its worker
// thread) against a Player that searches with b2p_tree_search on the B200s.  Built against the UNMODIFIED reference
// sources (shim/Makefile, build container only); the game loop below is ours, written against the reference's
// Player interface (src/player.hpp:20-30).
//
//   match_b200 <games> <b200 seconds per move> [batch] [reps]
#include "player.hpp"   // reference
#include "state.hpp"    // reference

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>

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
  B200TreePlayer(double seconds, uint32_t batch, uint32_t reps) : seconds(seconds), batch(batch), reps(reps) {
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
    uint64_t played = 0;
    if (b2p_tree_search(ctx, tree, 0, seconds, batch, 0.02f, reps, B2P_MODE_RANDOM, key++, &played) != B2P_OK)
      throw std::runtime_error(b2p_tree_last_error(tree));
    playouts += played;
    moves++;
    b2p_tree_stats st;
    b2p_tree_info(tree, &st);
    b2p_move_t best;
    if (b2p_tree_best_move(tree, (int)(st.root_state.meta & 1u), &best) != B2P_OK) throw std::runtime_error("no best move");
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
  uint64_t key = 1;
  b2p_ctx *ctx = nullptr;
  b2p_tree *tree = nullptr;
};

}  // namespace

int main(int argc, char **argv) {
  const int games = argc > 1 ? std::atoi(argv[1]) : 2;
  const double seconds = argc > 2 ? std::atof(argv[2]) : 1.0;
  const uint32_t batch = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 8192, reps = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 32;
  int score[3] = {0, 0, 0};  // b200 wins, mcts_host wins, draws
  for (int g = 0; g < games; g++) {
    B200TreePlayer *mine = new B200TreePlayer(seconds, batch, reps);
    std::unique_ptr<Player> players[NUM_PLAYERS];
    const int my_seat = g % 2;  // colours alternate
    players[my_seat] = std::unique_ptr<Player>(mine);
    players[1 - my_seat] = getPlayer("mcts_host");
    State state = getStartingState();
    for (auto &p : players) p->start();
    int plies = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (!state.isGameOver()) {
      Move m = players[state.turn]->getMove(state, false);
      for (auto &p : players) p->move(m);
      state.move(m);
      plies++;
    }
    const PlayerId winner = state.getWinner();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const char *who = winner == PLAYER_NONE ? "draw" : ((int)winner == my_seat ? "b200_tree" : "mcts_host");
    score[winner == PLAYER_NONE ? 2 : ((int)winner == my_seat ? 0 : 1)]++;
    std::cout << "{\"game\": " << g << ", \"b200_seat\": \"P" << my_seat + 1 << "\", \"winner\": \"" << who << "\", \"plies\": " << plies
              << ", \"seconds\": " << secs << ", \"b200_playouts_per_move\": " << (mine->moves ? mine->playouts / mine->moves : 0) << "}"
              << std::endl;
    for (auto &p : players) p->stop();
  }
  std::cout << "{\"summary\": {\"b200_tree\": " << score[0] << ", \"mcts_host\": " << score[1] << ", \"draw\": " << score[2]
            << "}, \"b200_seconds_per_move\": " << seconds << ", \"mcts_host\": \"reference preset: 50 playouts per batch, 7 s per move, pondering\"}"
            << std::endl;
  return 0;
}
