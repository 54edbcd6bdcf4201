"""CPU: b2p_tree (gpu_ai_b200/csrc/tree.cu) against the reference's own GameTree (src/mcts.cpp), SURVEY.md 8f-1.

Both trees are driven with the same deterministic fake playout results (a hash of the leaf state), batch
after batch; every batch of selected leaves, the trial totals, the chosen move and the re-rooted subtree
must be identical.  The live comparison needs oracle/_ref (build container); the golden trace generated from
it (tests/golden/tree_trace.npz, tools/make_tree_golden.py) travels to the GPU box."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.pyoracle import START_PACKED, make_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_winners(leaves, salt):
    """deterministic 'playout result' per leaf: -1 / 0 / 1 from a hash of the packed state"""
    a = leaves.astype(np.uint64)
    h = (a[:, 0] * np.uint64(0x9E3779B1) ^ a[:, 1] * np.uint64(0x85EBCA77) ^ a[:, 2] * np.uint64(0xC2B2AE3D) ^ a[:, 3] + np.uint64(salt))
    h = (h ^ (h >> np.uint64(15))) * np.uint64(0x2545F491)
    return ((h >> np.uint64(7)) % np.uint64(3)).astype(np.int8) - 1


class RefTree:
    def __init__(self, ref, root):
        self.lib = ref.lib
        self.lib.ref_tree_create.restype = C.c_void_p
        self.lib.ref_tree_select.restype = C.c_size_t
        self.lib.ref_tree_total.restype = C.c_uint64
        self.lib.ref_tree_best_move.restype = C.c_uint64
        r = np.ascontiguousarray(root, dtype=np.uint32)
        self.h = C.c_void_p(self.lib.ref_tree_create(r.ctypes.data_as(C.c_void_p)))

    def select(self, trials):
        out = np.zeros((max(trials, 1), 4), dtype=np.uint32)
        n = self.lib.ref_tree_select(self.h, C.c_uint(trials), out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def update(self, winners):
        w = np.ascontiguousarray(winners, dtype=np.int8)
        self.lib.ref_tree_update(self.h, w.ctypes.data_as(C.c_void_p), C.c_size_t(w.size))

    def total(self):
        return int(self.lib.ref_tree_total(self.h))

    def best_move(self, player):
        return int(self.lib.ref_tree_best_move(self.h, player))

    def move(self, m):
        assert self.lib.ref_tree_move(self.h, C.c_uint64(m)) == 0

    def root_state(self):
        out = np.zeros(4, dtype=np.uint32)
        self.lib.ref_tree_root_state(self.h, out.ctypes.data_as(C.c_void_p))
        return out


BATCHES = [50, 1, 2, 0, 7, 50, 257, 50, 50, 1000, 3, 50, 4000, 80, 80, 80, 513, 50, 50, 2000]


def drive(tree, batches, salt, moves_after=(7, 14)):
    """MCTSPlayer-like session: select/update rounds, two moves played in between (subtree reuse)"""
    trace = []
    for i, n in enumerate(batches):
        leaves = np.array(tree.select(n), copy=True)
        trace.append(leaves)
        tree.update(fake_winners(leaves, salt + i) if len(leaves) else np.zeros(0, np.int8))
        if i in moves_after:
            root = tree.root_state() if hasattr(tree, "root_state") else tree.info()["root_state"]
            m = tree.best_move(int(root[3] & 1))
            trace.append(np.array([[m & 0xFFFFFFFF, m >> 32, 0, 0]], dtype=np.uint32))
            tree.move(m)
    return trace


@pytest.mark.parametrize("name,root", [("start", START_PACKED),
                                       ("endgame", make_state(p1_kings=[(2, 1)], p1_men=[(1, 4)], p2_men=[(5, 2), (6, 5)], p2_kings=[(7, 0)])),
                                       ("captures", make_state(p1_men=[(2, 1), (2, 3), (1, 6)], p2_men=[(3, 2), (5, 4), (5, 6), (6, 1)], turn=0))])
def test_tree_equals_reference_gametree_live(ref, name, root):
    import gpu_ai_b200 as b
    mine, theirs = b.Tree(root), RefTree(ref, root)
    a = drive(mine, BATCHES, salt=11)
    c = drive(theirs, BATCHES, salt=11)
    assert len(a) == len(c)
    for x, y in zip(a, c):
        assert x.shape == y.shape and np.array_equal(x, y)
    assert mine.info()["total_trials"] == theirs.total()


def test_tree_equals_golden_trace():
    import gpu_ai_b200 as b
    g = np.load(os.path.join(ROOT, "tests", "golden", "tree_trace.npz"))
    tree = b.Tree(g["root"])
    trace = drive(tree, [int(x) for x in g["batches"]], salt=int(g["salt"]))
    flat = np.concatenate([t for t in trace if len(t)])
    assert np.array_equal(flat, g["flat"])
    assert [len(t) for t in trace] == list(g["lengths"])
    assert tree.info()["total_trials"] == int(g["total"])


def test_tree_reps_and_errors():
    import gpu_ai_b200 as b
    t = b.Tree(START_PACKED)
    leaves = t.select(50)
    assert len(leaves) == 50
    with pytest.raises(b.B2PError):
        t.update(np.zeros(49, np.int8))                 # wrong result count (the reference asserts)
    t.update(np.tile(fake_winners(leaves, 1), 4), reps=4)   # 4 playouts per selected leaf
    assert t.info()["total_trials"] == 200
    mv, tr, w1, w2 = t.root_moves()
    assert len(mv) == 7 and tr.sum() == 200
    with pytest.raises(b.B2PError):
        t.move(0xFFFF)                                   # not a legal root move


def test_update_counts_equals_update_with_winners():
    import gpu_ai_b200 as b
    a, c = b.Tree(START_PACKED), b.Tree(START_PACKED)
    reps = 5
    for it, n in enumerate([50, 300, 7, 1000]):
        la, lc = a.select(n), c.select(n)
        assert np.array_equal(la, lc)
        w = np.stack([fake_winners(la, 100 * it + r) for r in range(reps)])       # [rep][leaf]
        a.update(w.reshape(-1), reps=reps)
        c.update_counts(np.stack([(w == 0).sum(axis=0), (w == 1).sum(axis=0)], axis=1), reps)
        assert a.info()["total_trials"] == c.info()["total_trials"] and a.info()["wins"] == c.info()["wins"]
        for x, y in zip(a.root_moves(), c.root_moves()):
            assert np.array_equal(x, y)


def test_tree_equals_reference_gametree_random_roots(ref, golden):
    """randomised sessions: reachable roots from the golden leaf set, random batch schedules, random re-rooting"""
    import gpu_ai_b200 as b
    rng = np.random.default_rng(17)
    roots = golden["leaves_states"][golden["leaves_counts"] > 0]
    for trial in range(12):
        root = roots[int(rng.integers(len(roots)))]
        batches = [int(x) for x in rng.choice([0, 1, 2, 3, 17, 50, 120, 700, 2500], size=10)]
        moves_after = tuple(sorted(int(x) for x in rng.choice(np.arange(2, 9), size=2, replace=False)))
        mine, theirs = b.Tree(root), RefTree(ref, root)
        try:
            a = drive(mine, batches, salt=100 + trial, moves_after=moves_after)
        except b.B2PError:
            a = None   # best_move on an unexpanded root: the reference asserts there (src/mcts.cpp:40)
        if a is None:
            continue
        c = drive(theirs, batches, salt=100 + trial, moves_after=moves_after)
        assert len(a) == len(c)
        for x, y in zip(a, c):
            assert x.shape == y.shape and np.array_equal(x, y)
        assert mine.info()["total_trials"] == theirs.total()


# ---- the pipelined search's host halves (b2p_tree_select_batch / b2p_tree_update_batch), no GPU needed -------------
def fake_counts(leaves, reps, salt):
    w = np.stack([fake_winners(leaves, salt + 1000 * r) for r in range(reps)])
    return np.stack([(w == 0).sum(axis=0), (w == 1).sum(axis=0)], axis=1).astype(np.uint32)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_batch_pipeline_depth1_equals_serial_interface(threads):
    """One batch in flight = the strictly serial loop: the parallel, record-based select/update makes exactly the
    decisions of b2p_tree_select / b2p_tree_update_counts (and therefore of the reference GameTree), for any
    number of host threads -- same leaves in the same order, same statistics."""
    import gpu_ai_b200 as b
    reps = 3
    for root in (START_PACKED, make_state(p1_kings=[(2, 1)], p1_men=[(1, 4)], p2_men=[(5, 2), (6, 5)], p2_kings=[(7, 0)])):
        a, c = b.Tree(root), b.Tree(root)
        for it, n in enumerate([50, 1, 2, 7, 300, 5000, 64, 20000, 3, 9000]):
            la = a.select(n)
            lc = c.select_batch(it % 4, n, reps=reps, threads=threads)
            assert np.array_equal(la, lc), "batch %d" % it
            wins = fake_counts(la, reps, 7 * it)
            a.update_counts(wins, reps)
            c.update_batch(it % 4, wins, threads=threads)
            ia, ic = a.info(), c.info()
            assert ia["total_trials"] == ic["total_trials"] and ia["wins"] == ic["wins"] and ia["nodes"] == ic["nodes"]
            for x, y in zip(a.root_moves(), c.root_moves()):
                assert np.array_equal(x, y)
            if it == 5:
                m = a.best_move(int(root[3] & 1))
                assert m == c.best_move(int(root[3] & 1))
                a.move(m)
                c.move(m)


def test_batch_pipeline_two_in_flight_is_consistent_and_deterministic():
    """Two batches selected before the first is updated (virtual loss): every playout is accounted for exactly
    once, wins never exceed trials anywhere the root can see, and the result does not depend on the thread count."""
    import gpu_ai_b200 as b
    reps = 4
    runs = []
    for threads in (1, 6):
        t = b.Tree(START_PACKED)
        sizes = [4000, 4000, 6000, 100, 12000, 12000, 50]
        pending = []
        total = 0
        for it, n in enumerate(sizes):
            leaves = t.select_batch(it % 2, n, reps=reps, threads=threads)
            assert len(leaves) == n
            total += n * reps
            assert t.info()["total_trials"] == total          # in-flight trials are visible as visits at once
            pending.append((it % 2, fake_counts(leaves, reps, it), leaves.copy()))
            if len(pending) == 2:
                slot, wins, _ = pending.pop(0)
                t.update_batch(slot, wins, threads=threads)
        for slot, wins, _ in pending:
            t.update_batch(slot, wins, threads=threads)
        info = t.info()
        mv, tr, w1, w2 = t.root_moves()
        assert tr.sum() == total and info["wins"][0] == w1.sum() and info["wins"][1] == w2.sum()
        assert ((w1 + w2) <= tr).all()
        runs.append((tr.copy(), w1.copy(), w2.copy(), info["nodes"]))
    for x, y in zip(runs[0], runs[1]):
        assert np.array_equal(x, y)


def test_uct_policy_is_consistent_deterministic_and_stronger(port):
    """B2P_POLICY_UCT (trials one by one to the best UCB1 child from the MOVER's point of view, virtual loss, most-tried
    root move) against the reference's allocation rule (shares in proportion to UCB1 values that read the opponent's
    wins, src/mcts.cpp:93-139,182-191; best-rate root move) at the SAME budget -- 1500 trials per move in batches of
    50, one CPU-oracle playout per leaf.  Bookkeeping first, then six games with alternating colours."""
    import gpu_ai_b200 as b
    from oracle.pyoracle import ORDER_FAST
    t = b.Tree(START_PACKED)
    for it, n in enumerate([50, 700, 3000, 50]):
        leaves = t.select_batch(it % 2, n, reps=2, threads=3, policy=2)
        assert len(leaves) == n
        t.update_batch(it % 2, fake_counts(leaves, 2, it), threads=3)
    mv, tr, w1, w2 = t.root_moves()
    assert tr.sum() == 2 * 3800 and ((w1 + w2) <= tr).all() and t.robust_move(0) in set(int(x) for x in mv)
    t2 = b.Tree(START_PACKED)
    for it, n in enumerate([50, 700, 3000, 50]):
        leaves = t2.select_batch(it % 2, n, reps=2, threads=1, policy=2)
        t2.update_batch(it % 2, fake_counts(leaves, 2, it), threads=1)
    assert np.array_equal(t2.root_moves()[1], tr)          # same tree for any thread count

    def think(tree, policy, key):
        for it in range(30):
            leaves = tree.select_batch(0, 50, reps=1, threads=1, policy=policy)
            w, _, _, _ = port.playouts(leaves, key=key + it, order=ORDER_FAST)
            tree.update_batch(0, np.stack([w == 0, w == 1], axis=1).astype(np.uint32), threads=1)

    points = 0.0
    for g in range(6):
        uct_seat = g % 2
        trees = [b.Tree(START_PACKED), b.Tree(START_PACKED)]
        plies = 0
        while True:
            info = trees[0].info()
            st = info["root_state"]
            turn, msc = int(st[3] & 1), int(st[3] >> 8)
            if info["root_moves"] == 0 or msc >= 50:
                points += 0.5 if msc >= 50 else (1.0 if (turn ^ 1) == uct_seat else 0.0)
                break
            uct = turn == uct_seat
            think(trees[turn], 2 if uct else 0, key=10000 * g + 100 * plies)
            m = trees[turn].robust_move(turn) if uct else trees[turn].best_move(turn)
            for tr_ in trees:
                tr_.move(m)
            plies += 1
    assert points >= 4.5, points


def test_rerooting_is_refused_while_a_batch_is_outstanding():
    """the visit records of a selected batch name nodes of the arena that b2p_tree_move frees"""
    import gpu_ai_b200 as b
    t = b.Tree(START_PACKED)
    first = t.select_batch(0, 500, reps=2, threads=2)
    t.update_batch(0, fake_counts(first, 2, 1), threads=2)
    pend = t.select_batch(1, 64, reps=2, threads=1)
    mv, tr, _, _ = t.root_moves()
    with pytest.raises(b.B2PError, match="not been folded in"):
        t.move(int(mv[0]))
    t.update_batch(1, fake_counts(pend, 2, 99), threads=1)
    mv, tr, _, _ = t.root_moves()
    t.move(int(mv[0]))
    assert t.info()["total_trials"] == int(tr[0])
