// shim/playout_shim.cpp -- the link-time drop-in for the reference binary.
//
// The reference's host objects (mcts.cpp, player.cpp, playout.cpp, driver.cpp) need exactly five
// device-side symbols (`nm` of the reference objects):
//   DeviceSinglePlayoutDriver::runPlayouts    (reference definition: src/singlePlayout.cu:71-120)
//   DeviceMultiplePlayoutDriver::runPlayouts  (src/multiplePlayout.cu:53-98)
//   DeviceCoarsePlayoutDriver::runPlayouts    (src/coarsePlayout.cu:91-163)
//   DeviceHeuristicPlayoutDriver::runPlayouts (src/heuristicPlayout.cu:102-147)
//   bool genMovesTest(State)                  (src/genMovesTest.cu:26-100)
// This TU includes the reference's own headers (never copied: -I/root/reference/src at build time),
// defines those five symbols on top of the C ABI of include/b2p.h, and is linked INSTEAD OF the
// reference's singlePlayout.cu / multiplePlayout.cu / coarsePlayout.cu / heuristicPlayout.cu /
// genMovesTest.cu.  mcts.cpp's tree search and every run_ai player type then run unchanged on the
// B200 kernels.  Each class's runPlayouts is its key function, so this TU also emits the vtables.
#include "playout.hpp"       // reference: src/playout.hpp
#include "genMovesTest.hpp"  // reference: src/genMovesTest.hpp

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/b2p.h"

static_assert(sizeof(State) == 776, "b2p_run_states776 reads the reference State layout (SURVEY.md 8a)");
static_assert(sizeof(Move) == 38, "b2p_expand_move writes the reference Move layout");
static_assert(sizeof(PlayerId) == sizeof(int32_t), "PlayerId is a 4-byte enum");
static_assert(PLAYER_1 == B2P_PLAYER_1 && PLAYER_2 == B2P_PLAYER_2 && PLAYER_NONE == B2P_PLAYER_NONE, "winner encoding");

namespace {

// One context per calling thread: two MCTSPlayers run their workers concurrently
// (src/player.cpp:119-150) and HybridPlayoutDriver calls the device driver from an OpenMP thread
// (src/playout.cpp:51-56); a b2p context is single-caller, contexts are independent.
struct ThreadContext {
  b2p_ctx *ctx = nullptr;
  ThreadContext() {
    const char *env = std::getenv("B2P_DEVICES");  // e.g. B2P_DEVICES=2 -> devices 0,1; default: all visible
    int n = env ? std::atoi(env) : 0;
    if (b2p_create(&ctx, nullptr, n, 12345) != B2P_OK)  // SEED 12345: src/singlePlayout.cu:12
      throw std::runtime_error(std::string("b2p_create: ") + b2p_last_error(nullptr));
  }
  ~ThreadContext() { b2p_destroy(ctx); }
};

b2p_ctx *context() {
  thread_local ThreadContext tc;
  return tc.ctx;
}

std::vector<PlayerId> run(const std::vector<State> &states, int mode, int sched) {
  std::vector<PlayerId> results(states.size());
  if (states.empty()) return results;  // src/singlePlayout.cu:73-75
  b2p_ctx *ctx = context();
  if (b2p_run_states776(ctx, states.data(), states.size(), mode, sched, reinterpret_cast<int32_t *>(results.data())) != B2P_OK)
    throw std::runtime_error(std::string("b2p_run_states776: ") + b2p_last_error(ctx));
  return results;
}

}  // namespace

std::vector<PlayerId> DeviceSinglePlayoutDriver::runPlayouts(std::vector<State> states) { return run(states, B2P_MODE_RANDOM, B2P_SCHED_THREAD); }
std::vector<PlayerId> DeviceCoarsePlayoutDriver::runPlayouts(std::vector<State> states) { return run(states, B2P_MODE_RANDOM, B2P_SCHED_THREAD); }
std::vector<PlayerId> DeviceMultiplePlayoutDriver::runPlayouts(std::vector<State> states) { return run(states, B2P_MODE_RANDOM, B2P_SCHED_AUTO); }
std::vector<PlayerId> DeviceHeuristicPlayoutDriver::runPlayouts(std::vector<State> states) { return run(states, B2P_MODE_HEURISTIC, B2P_SCHED_AUTO); }

// Device move list vs host State::genMoves, element-wise Move::operator== (src/state.cu:440-454),
// same contract and same diagnostics as the reference's genMovesTest.
bool genMovesTest(State state) {
  b2p_ctx *ctx = context();
  b2p_state16 packed;
  b2p_move_t dev[MAX_MOVES];
  uint8_t devCount = 0;
  if (b2p_pack776(&state, 1, &packed) != B2P_OK || b2p_genmoves(ctx, &packed, 1, MAX_MOVES, dev, &devCount) != B2P_OK)
    throw std::runtime_error(std::string("b2p_genmoves: ") + b2p_last_error(ctx));

  Move cpuMoves[MAX_MOVES];
  uint8_t cpuCount = state.genMoves(cpuMoves);
  std::vector<Move> gpuMoves(devCount);
  for (uint8_t i = 0; i < devCount; i++) b2p_expand_move(dev[i], &gpuMoves[i]);

  bool match = cpuCount == devCount;
  for (uint8_t i = 0; match && i < cpuCount; i++) match = cpuMoves[i] == gpuMoves[i];
  if (!match) {
    std::cout << "Mismatch in CPU and GPU genMoves()" << std::endl << state << std::endl;
    std::cout << "CPU Moves: " << (int)cpuCount << std::endl;
    for (uint8_t i = 0; i < cpuCount; i++) std::cout << cpuMoves[i] << std::endl;
    std::cout << "GPU Moves: " << (int)devCount << std::endl;
    for (uint8_t i = 0; i < devCount; i++) std::cout << gpuMoves[i] << std::endl;
  }
  return match;
}
