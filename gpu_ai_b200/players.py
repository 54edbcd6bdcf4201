"""Python mirror of the reference's Player interface (src/player.hpp:20-30) for the B200 tree search.

    class Player { getName(); getMove(const State&, bool verbose); start(); stop(); move(const Move&); }

`B200MCTSPlayer` is the counterpart of MCTSPlayer (src/player.hpp:50-88, src/player.cpp:88-150): it keeps a search
tree (gpu_ai_b200.Tree == the reference's GameTree, decision for decision), re-roots it on every move of either
side (subtree reuse, GameTree::move) and, instead of pondering in a worker thread with 50-leaf batches, searches for
`seconds` inside getMove with large batches and `reps` GPU playouts per selected leaf (b2p_tree_search).
States and moves are the packed forms of include/b2p.h (b2p_state16 as 4 x uint32, b2p_move_t as int)."""
import numpy as np

from . import engine as _e

START_STATE = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)  # getStartingState, src/state.cpp:25-40


class Player:
    def getName(self):
        raise NotImplementedError

    def getMove(self, state, verbose=True):
        raise NotImplementedError

    def start(self):
        pass

    def stop(self):
        pass

    def move(self, move):
        pass


class B200MCTSPlayer(Player):
    def __init__(self, engine=None, seconds=1.0, initial_batch=2048, scale=0.0, reps=16, mode=_e.MODE_RANDOM, seed=1, policy=1):
        """policy 1 = B2P_POLICY_UCT with the most-tried root move (the strong setting); 0 = the reference's
        allocation rule and its highest-rate move (GameTree::select / getOptMove)."""
        self.engine = engine or _e.Engine()
        self.seconds, self.initial_batch, self.scale, self.reps, self.mode = seconds, initial_batch, scale, reps, mode
        self.policy = policy
        self.key = seed
        self.tree = None
        self.playouts = 0

    def getName(self):
        return "mcts_b200"

    def start(self):
        self.tree = _e.Tree(START_STATE)

    def stop(self):
        self.tree = None

    def getMove(self, state, verbose=True):
        state = np.ascontiguousarray(state, dtype=np.uint32).reshape(4)
        if self.tree is None or not np.array_equal(self.tree.info()["root_state"], state):
            self.tree = _e.Tree(state)  # src/player.cpp:95-97: a position the tree does not know starts a new tree
        self.key += 1
        st = self.tree.search_ex(self.engine, seconds=self.seconds, initial_batch=self.initial_batch, scale=self.scale,
                                 max_batch=max(self.initial_batch, 1 << 18), reps=self.reps, mode=self.mode, key=self.key,
                                 policy=self.policy)
        self.playouts += st["playouts"]
        if verbose:
            info = self.tree.info()
            print("Tree size: %d" % info["total_trials"])
        me = int(state[3] & 1)
        return self.tree.robust_move(me) if self.policy == 1 else self.tree.best_move(me)

    def move(self, move):
        if self.tree is not None:
            self.tree.move(move)
