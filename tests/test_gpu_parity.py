"""GPU (B200): the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Three parity gates of BASELINE.json's north_star:
  1. move generation bit-exact against the reference's State::getMoves (>= 10^6 positions);
  2. playouts driven by an injected deterministic move-choice sequence give bit-exact winners
     (and ply counts, and final states);
  3. random and heuristic win rates match the reference's host drivers within a binomial CI.
Plus size-independent properties at the full benchmark size (2^20 leaves).
"""
import numpy as np
import pytest

from conftest import fast_synthetic, unflatten
from oracle.pyoracle import MODE_HEURISTIC, MODE_RANDOM, ORDER_CANONICAL, ORDER_FAST, START_PACKED

pytestmark = pytest.mark.gpu


# ---- gate 1: move generation ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["leaves", "synth", "kat"])
def test_genmoves_equals_golden(engine, golden, name):
    st, cnt = golden[name + "_states"], golden[name + "_counts"]
    mv, c = engine.genmoves(st, 64)
    assert np.array_equal(c, cnt)
    assert np.array_equal(mv, unflatten(golden[name + "_moves_flat"], cnt))


def test_genmoves_million_positions_bit_exact(engine, port):
    leaves = engine.gen_leaves(1 << 20, key=31337)
    assert np.array_equal(leaves[:50000], port.gen_leaves(50000, key=31337))
    st = np.concatenate([leaves, fast_synthetic(1 << 19, 3)])
    a, ca = engine.genmoves(st, 48)
    b, cb = port.genmoves(st, 48)
    assert cb.max() <= 48
    assert np.array_equal(ca, cb)
    assert np.array_equal(a, b)


def test_genmoves_truncation_and_edge_sizes(engine, golden):
    st, cnt = golden["leaves_states"][:100], golden["leaves_counts"][:100]
    mv, c = engine.genmoves(st, 2)           # max_moves smaller than the list: counts stay true
    assert np.array_equal(c, cnt)
    full = unflatten(golden["leaves_moves_flat"], golden["leaves_counts"])[:100]
    assert np.array_equal(mv, full[:, :2])
    mv, c = engine.genmoves(np.zeros((0, 4), np.uint32), 8)   # empty input
    assert mv.shape == (0, 8) and c.shape == (0,)
    mv, c = engine.genmoves(st[:1], 64)      # single state
    assert c[0] == cnt[0]


# ---- gate 2: deterministic replay -----------------------------------------------------------------
@pytest.mark.parametrize("name", ["leaves", "synth"])
@pytest.mark.parametrize("tag,mode,order", [("rc", MODE_RANDOM, ORDER_CANONICAL), ("rf", MODE_RANDOM, ORDER_FAST),
                                            ("h", MODE_HEURISTIC, ORDER_CANONICAL)])
def test_playouts_equal_golden(engine, golden, name, tag, mode, order):
    st = golden[name + "_states"]
    w, p, f, c = engine.run_packed(st, reps=2, key=12345, pid_base=1000, mode=mode, order=order,
                                   want_plies=True, want_final=True)
    assert np.array_equal(w, golden["%s_%s_winners" % (name, tag)])
    assert np.array_equal(p, golden["%s_%s_plies" % (name, tag)])
    assert np.array_equal(f, golden["%s_%s_final" % (name, tag)])
    assert np.array_equal(c, golden["%s_%s_counters" % (name, tag)])


@pytest.mark.parametrize("mode,order", [(MODE_RANDOM, ORDER_CANONICAL), (MODE_RANDOM, ORDER_FAST), (MODE_HEURISTIC, ORDER_CANONICAL)])
def test_playouts_equal_oracle_large(engine, port, mode, order):
    st = np.concatenate([engine.gen_leaves(150000, key=77), fast_synthetic(50000, 19)])
    w, p, f, c = engine.run_packed(st, key=99, pid_base=12, mode=mode, order=order, want_plies=True, want_final=True)
    ow, op, of, oc = port.playouts(st, key=99, pid_base=12, mode=mode, order=order, want_final=True)
    assert np.array_equal(w, ow) and np.array_equal(p, op) and np.array_equal(f, of) and np.array_equal(c, oc)


def test_fast_kernel_without_optional_outputs_matches(engine, port):
    """the lean kernel variant (winners + counters only) is the one the benchmark times"""
    st = engine.gen_leaves(100000, key=5)
    w, _, _, c = engine.run_packed(st, reps=3, key=31, mode=MODE_RANDOM, order=ORDER_FAST)
    ow, _, _, oc = port.playouts(st, reps=3, key=31, mode=MODE_RANDOM, order=ORDER_FAST)
    assert np.array_equal(w, ow) and np.array_equal(c, oc)


@pytest.mark.parametrize("name", ["leaves", "synth"])
def test_truncated_playouts_equal_golden(engine, golden, name):
    w, p, f, c = engine.run_packed(golden[name + "_states"], key=12345, max_plies=5, want_plies=True, want_final=True)
    assert np.array_equal(w, golden[name + "_cut5_winners"])
    assert np.array_equal(f, golden[name + "_cut5_final"])


def test_leafgen_equals_golden(engine, golden):
    assert np.array_equal(engine.gen_leaves(4096, key=2016), golden["leaves_states"])
    # leaf j does not depend on the batch it was generated in
    assert np.array_equal(engine.gen_leaves(1000, key=2016, first_index=3000), golden["leaves_states"][3000:4000])


def test_terminal_and_degenerate_inputs(engine, golden):
    names = list(golden["kat_names"])
    w, _, _, _ = engine.run_packed(golden["kat_states"], key=12345)
    assert np.array_equal(w, golden["kat_rc_winners"])
    assert w[names.index("no_pieces_to_move")] == 1 and w[names.index("draw_counter")] == -1
    w, _, _, c = engine.run_packed(np.zeros((0, 4), np.uint32))          # empty batch
    assert w.shape == (0,) and c.sum() == 0
    w, _, _, _ = engine.run_packed(START_PACKED.reshape(1, 4), reps=1)   # batch of one
    assert w[0] in (-1, 0, 1)


def test_results_do_not_depend_on_batch_split(engine):
    """RNG is keyed by the global playout id: playing a batch in two halves gives the same winners."""
    st = engine.gen_leaves(20001, key=8)
    w, _, _, _ = engine.run_packed(st, key=5, pid_base=0, order=ORDER_FAST)
    a, _, _, _ = engine.run_packed(st[:7777], key=5, pid_base=0, order=ORDER_FAST)
    b, _, _, _ = engine.run_packed(st[7777:], key=5, pid_base=7777, order=ORDER_FAST)
    assert np.array_equal(w, np.concatenate([a, b]))


# ---- the reference-facing 776-byte entry point -----------------------------------------------------------
def test_run_states776_drop_in(engine, port, golden):
    import gpu_ai_b200 as b
    st = golden["leaves_states"]
    s776 = port.unpack776(st)
    for name in ("device_single", "device_multiple", "device_coarse", "device_heuristic"):
        drv = b.getPlayoutDriver(name)
        assert drv.getName() == name
        res = drv.runPlayouts(s776)
        assert res.dtype == np.int32 and res.shape == (len(st),)
        assert set(np.unique(res)) <= {-1, 0, 1}
        # states that are already over must come back with their winner (src/mcts.cpp:65-68)
        term = golden["leaves_counts"] == 0
        assert np.array_equal(res[term], golden["leaves_rc_winners"][:len(st)][term])
        assert drv.runPlayouts(np.zeros((0, 776), np.uint8)).shape == (0,)


# ---- gate 3: win-rate statistics vs the reference's own host drivers ------------------------------------
def _z(count_a, n_a, count_b, n_b):
    pa, pb = count_a / n_a, count_b / n_b
    p = (count_a + count_b) / (n_a + n_b)
    se = np.sqrt(p * (1 - p) * (1 / n_a + 1 / n_b))
    return (pa - pb) / se


@pytest.mark.parametrize("tag,mode", [("host", MODE_RANDOM), ("host_heuristic", MODE_HEURISTIC)])
def test_winrates_match_reference_host_driver(engine, tag, mode):
    """Two-proportion z-test on draws / P1 wins / P2 wins: GPU (Philox; heuristic: table Gaussian) against the
    reference's HostPlayoutDriver / HostHeuristicPlayoutDriver (glibc rand / std::normal_distribution) on the same
    2^20 D_ref leaves -- >= 10^6 playouts on BOTH sides, overall and per piece-count stratum of the
    leaf (0-5, 6-9, 10-14, 15-24 pieces).  Tolerance: |z| < 4 on each outcome: +-0.25 % absolute on the whole
    set (two-sided p ~ 6e-5 per test, 15 tests per mode)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tallies.npz"))
    n = int(g["n"])
    ref_tally = g["ref_%s_tally" % tag]
    st = engine.gen_leaves(n, key=int(g["leaf_key"]))
    assert np.uint64(st.astype(np.uint64).sum()) == g["leaves_checksum"]
    edges = g["strata_edges"]
    pieces = np.unpackbits((st[:, 0] | st[:, 1]).astype(np.uint32).view(np.uint8)).reshape(n, 32).sum(axis=1)
    strata = np.digitize(pieces, edges[1:-1])
    reps = 4
    w, _, _, c = engine.run_packed(st, reps=reps, key=424242, mode=mode, order=ORDER_CANONICAL if mode else ORDER_FAST)
    w = w.reshape(reps, n)
    assert int(c[0] + c[1] + c[2]) == n * reps
    for s in list(range(len(edges) - 1)) + [-1]:
        sel = slice(None) if s == -1 else (strata == s)
        n_ref = int(ref_tally[s].sum())
        n_gpu = int(n_ref * reps)
        assert n_ref == (n if s == -1 else int(sel.sum()))
        for k, v in enumerate((-1, 0, 1)):
            cnt = int((w[:, sel] == v).sum())
            z = _z(cnt, n_gpu, int(ref_tally[s][k]), n_ref)
            assert abs(z) < 4.0, "stratum %d outcome %d: gpu %.4f vs reference %.4f (z = %.2f, n_ref = %d)" % (
                s, v, cnt / n_gpu, ref_tally[s][k] / n_ref, z, n_ref)


# ---- size-independent properties at the full benchmark size ------------------------------------------------
def test_full_size_properties(engine):
    n = 1 << 20
    st = engine.gen_leaves(n, key=2016)
    w, p, f, c = engine.run_packed(st, key=1, order=ORDER_FAST, want_plies=True, want_final=True)
    # counters are a checksum of the per-playout outputs
    assert int(c[0]) == int((w == -1).sum()) and int(c[1]) == int((w == 0).sum()) and int(c[2]) == int((w == 1).sum())
    assert int(c[3]) == int(p.sum(dtype=np.uint64))
    # every playout ends, and ends in a terminal state: a second pass from the final states plays 0 plies
    assert set(np.unique(w)) <= {-1, 0, 1}
    w2, p2, f2, _ = engine.run_packed(f, key=2, order=ORDER_FAST, want_plies=True, want_final=True)
    assert p2.max() == 0 and np.array_equal(w2, w) and np.array_equal(f2, f)
    # a draw is declared exactly when the counter reached 50
    assert np.array_equal((f[:, 3] >> 8) >= 50, w == -1)
    # material never appears from nowhere
    pop = lambda a: np.unpackbits(a.view(np.uint8)).reshape(len(a), -1).sum(axis=1)  # noqa: E731
    assert (pop(f[:, 0].copy()) <= pop(st[:, 0].copy())).all() and (pop(f[:, 1].copy()) <= pop(st[:, 1].copy())).all()
    # determinism: same key -> same winners; different key -> different stream
    w3, _, _, _ = engine.run_packed(st, key=1, order=ORDER_FAST)
    assert np.array_equal(w3, w)
    w4, _, _, _ = engine.run_packed(st, key=2, order=ORDER_FAST)
    assert 0.5 < (w4 == w).mean() < 0.95


# ---- warp-per-playout scheduling (B2P_SCHED_WARP) gives the same answers ------------------------------------
@pytest.mark.parametrize("name", ["leaves", "synth"])
@pytest.mark.parametrize("tag,mode,order", [("rc", MODE_RANDOM, ORDER_CANONICAL), ("rf", MODE_RANDOM, ORDER_FAST),
                                            ("h", MODE_HEURISTIC, ORDER_CANONICAL)])
def test_warp_scheduler_equals_golden(engine, golden, name, tag, mode, order):
    import gpu_ai_b200 as b
    st = golden[name + "_states"]
    w, p, f, c = engine.run_packed(st, reps=2, key=12345, pid_base=1000, mode=mode, order=order, sched=b.SCHED_WARP,
                                   want_plies=True, want_final=True)
    assert np.array_equal(w, golden["%s_%s_winners" % (name, tag)])
    assert np.array_equal(p, golden["%s_%s_plies" % (name, tag)])
    assert np.array_equal(f, golden["%s_%s_final" % (name, tag)])
    assert np.array_equal(c, golden["%s_%s_counters" % (name, tag)])


@pytest.mark.parametrize("mode,order", [(MODE_RANDOM, ORDER_FAST), (MODE_HEURISTIC, ORDER_CANONICAL)])
def test_warp_scheduler_equals_thread_scheduler(engine, mode, order):
    import gpu_ai_b200 as b
    st = np.concatenate([engine.gen_leaves(60000, key=21), fast_synthetic(20000, 23)])
    a = engine.run_packed(st, key=3, mode=mode, order=order, sched=b.SCHED_THREAD, want_plies=True, want_final=True)
    c = engine.run_packed(st, key=3, mode=mode, order=order, sched=b.SCHED_WARP, want_plies=True, want_final=True)
    for x, y in zip(a, c):
        assert np.array_equal(x, y)
    # AUTO routes small batches to the warp kernel and large ones to the lane kernel: same answers either way
    d = engine.run_packed(st[:500], key=3, mode=mode, order=order, sched=b.SCHED_AUTO, want_plies=True, want_final=True)
    assert np.array_equal(d[0], a[0][:500]) and np.array_equal(d[1], a[1][:500])


def test_warp_scheduler_truncated(engine, golden):
    import gpu_ai_b200 as b
    w, p, f, c = engine.run_packed(golden["synth_states"], key=12345, max_plies=5, sched=b.SCHED_WARP, want_plies=True, want_final=True)
    assert np.array_equal(w, golden["synth_cut5_winners"])
    assert np.array_equal(f, golden["synth_cut5_final"])


def test_run_states776_bit_exact_through_the_chunk_pipeline(port):
    """The reference-facing entry point (pack -> chunked H2D/kernel/D2H pipeline) is itself replayable:
    call c of a context uses key = seed + c * 0x9E3779B97F4A7C15 and global leaf indices as playout ids."""
    import gpu_ai_b200 as b
    eng = b.Engine(devices=1, seed=4711)
    st = eng.gen_leaves(200003, key=2016)          # > 65536: exercises the multi-chunk path, ragged last chunk
    s776 = port.unpack776(st)
    for call in range(2):
        key = (4711 + call * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        res = eng.run_states776(s776, mode=b.MODE_RANDOM)
        ow, _, _, _ = port.playouts(st, key=key, order=ORDER_FAST)
        assert np.array_equal(res, ow.astype(np.int32))
    res = eng.run_states776(s776[:3000], mode=b.MODE_HEURISTIC, sched=b.SCHED_AUTO)   # small batch -> warp kernel
    key = (4711 + 2 * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    ow, _, _, _ = port.playouts(st[:3000], key=key, mode=MODE_HEURISTIC)
    assert np.array_equal(res, ow.astype(np.int32))


# ---- boundary behaviour: threading and errors -------------------------------------------------------------------
def test_two_contexts_from_two_threads(port):
    """Two MCTSPlayer workers call runPlayouts concurrently (src/player.cpp:119-150): contexts are independent."""
    import threading
    import gpu_ai_b200 as b
    st = b.Engine(devices=1).gen_leaves(150000, key=3)
    expect = {}
    for key in (11, 22):
        expect[key] = port.playouts(st[:20000], key=key, order=ORDER_FAST)[0]
    out = {}

    def work(key):
        eng = b.Engine(devices=1, seed=key)
        for _ in range(3):
            w, _, _, _ = eng.run_packed(st, key=key, order=ORDER_FAST)
        out[key] = w

    th = [threading.Thread(target=work, args=(k,)) for k in (11, 22)]
    [t.start() for t in th]
    [t.join() for t in th]
    for key in (11, 22):
        assert np.array_equal(out[key][:20000], expect[key])


def test_errors_are_codes_not_crashes(engine):
    import ctypes as C
    import gpu_ai_b200 as b
    st = engine.gen_leaves(16, key=1)
    with pytest.raises(b.B2PError, match="unknown mode"):
        engine.run_packed(st, mode=7)
    with pytest.raises(b.B2PError):
        engine.run_packed(st, reps=2 ** 31 - 1)                     # n * reps >= 2^31
    with pytest.raises(b.B2PError):
        engine.run_states776(np.zeros(776 * 4, np.uint8), mode=9)
    lib = b.load_library()
    assert lib.b2p_run_states776(None, None, 5, 0, 0, None) == -1     # B2P_EINVAL on a NULL context
    assert lib.b2p_run_states776(engine.ctx, None, 5, 0, 0, None) == -1
    assert b"NULL" in lib.b2p_last_error(engine.ctx)
    with pytest.raises(b.B2PError, match="out of range"):
        b.Engine(devices=[99])
    # the context is still usable after errors
    w, _, _, _ = engine.run_packed(st)
    assert w.shape == (16,)


# ---- next row: tree search fused with the playout engine ---------------------------------------------
def test_tree_search_on_gpu(engine, port):
    import gpu_ai_b200 as b
    t = b.Tree(START_PACKED)
    played = t.search(engine, iterations=20, initial_batch=200, scale=0.02, reps=8, key=5)
    info = t.info()
    assert played == info["total_trials"] and played >= 20 * 200 * 8  # batch never shrinks below initial_batch
    assert info["wins"][0] + info["wins"][1] <= played
    mv, tr, w1, w2 = t.root_moves()
    assert len(mv) == 7 and tr.sum() == played               # every playout is accounted for at the root's children
    best = t.best_move(0)
    legal, cnt = port.genmoves(START_PACKED.reshape(1, 4), 64)
    assert best in set(int(x) for x in legal[0, :cnt[0]])
    # the search is deterministic for a given key
    t2 = b.Tree(START_PACKED)
    t2.search(engine, iterations=20, initial_batch=200, scale=0.02, reps=8, key=5)
    assert np.array_equal(t2.root_moves()[1], tr) and np.array_equal(t2.root_moves()[2], w1)
    # subtree reuse keeps the statistics of the chosen child
    t.move(best)
    assert t.info()["total_trials"] == int(tr[list(mv).index(best)])


def test_run_counts_equals_winner_tally(engine):
    import gpu_ai_b200 as b
    st = engine.gen_leaves(30000, key=4)
    for sched, n in ((b.SCHED_THREAD, 30000), (b.SCHED_WARP, 3000)):
        w, _, _, c = engine.run_packed(st[:n], reps=6, key=8, pid_base=77, order=ORDER_FAST, sched=sched)
        wins, c2 = engine.run_counts(st[:n], reps=6, key=8, pid_base=77, order=ORDER_FAST, sched=sched)
        w = w.reshape(6, n)
        assert np.array_equal(wins[:, 0], (w == 0).sum(axis=0)) and np.array_equal(wins[:, 1], (w == 1).sum(axis=0))
        assert np.array_equal(c, c2)


# ---- in-process multi-device sharding (what the shim uses: b2p_create with several devices) ------------------------
def _multi_device_ids():
    """All visible GPUs when there are several; on a 1-GPU box the SAME device three times: the context then
    owns three independent Device records (streams, staging, buffers, queue-head rings), so the sharding,
    gather-at-offset and host-combine logic runs exactly as it does over distinct GPUs."""
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n >= 2 else [0, 0, 0]


def test_in_process_multi_device_equals_single_device(engine):
    import gpu_ai_b200 as b
    ids = _multi_device_ids()
    multi = b.Engine(devices=ids, seed=12345)
    assert multi.device_count == len(ids)
    st = np.concatenate([engine.gen_leaves(70001, key=12), fast_synthetic(10000, 29)])
    assert np.array_equal(multi.gen_leaves(70001, key=12), st[:70001])
    for mode, order in ((MODE_RANDOM, ORDER_FAST), (MODE_HEURISTIC, ORDER_CANONICAL)):
        a = engine.run_packed(st, reps=3, key=6, pid_base=40, mode=mode, order=order, want_plies=True, want_final=True)
        c = multi.run_packed(st, reps=3, key=6, pid_base=40, mode=mode, order=order, want_plies=True, want_final=True)
        for x, y in zip(a, c):
            assert np.array_equal(x, y)
        wa, ca = engine.run_counts(st, reps=5, key=7, mode=mode, order=order)
        wc, cc = multi.run_counts(st, reps=5, key=7, mode=mode, order=order)
        assert np.array_equal(wa, wc) and np.array_equal(ca, cc)
    ma, na = engine.genmoves(st, 40)
    mc, nc = multi.genmoves(st, 40)
    assert np.array_equal(ma, mc) and np.array_equal(na, nc)
    s776 = b.engine.unpack776(st)
    e1, e2 = b.Engine(devices=1, seed=99), b.Engine(devices=ids, seed=99)
    assert np.array_equal(e1.run_states776(s776), e2.run_states776(s776))
    # below the minimum shard (8192 playouts per device) a batch stays on one device: same answers
    assert np.array_equal(e1.run_states776(s776[:5000]), e2.run_states776(s776[:5000]))
    t1, t2 = b.Tree(START_PACKED), b.Tree(START_PACKED)
    t1.search(engine, iterations=6, initial_batch=30000, reps=4, key=2)
    t2.search(multi, iterations=6, initial_batch=30000, reps=4, key=2)
    assert np.array_equal(t1.root_moves()[1], t2.root_moves()[1]) and np.array_equal(t1.root_moves()[2], t2.root_moves()[2])


@pytest.mark.parametrize("n", [1, 50, 511, 512, 513, 2049, 8191, 8193, 10239, 10240, 20000, 40961, 65536, 65537, 131073])
def test_run_states776_every_batch_size_class(port, n):
    """Pack tasks (512 states) and launch segments (8192..32768 leaves, short tails merged) at their edges: the
    reference-facing call returns the replayable winners for every size class of tests.sh."""
    import gpu_ai_b200 as b
    eng = b.Engine(devices=_multi_device_ids(), seed=7)
    st = eng.gen_leaves(n, key=2016, first_index=11)
    res = eng.run_states776(port.unpack776(st), mode=b.MODE_RANDOM, sched=b.SCHED_THREAD)
    ow, _, _, _ = port.playouts(st, key=7, order=ORDER_FAST)
    assert np.array_equal(res, ow.astype(np.int32))


def test_many_launches_in_flight_on_caller_streams(engine, port):
    """More launches than the 256-slot queue-head ring, spread over several caller streams without a
    synchronisation in between: every launch still sees a zeroed head of its own (event-guarded ring)."""
    import torch
    dev = torch.device("cuda", 0)
    n = 3000
    st = engine.gen_leaves(n, key=9)
    d_states = torch.from_numpy(st.view(np.int32)).to(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    outs = [torch.empty(n, dtype=torch.int8, device=dev) for _ in range(600)]
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        engine.run_packed_device(d_states.data_ptr(), n, key=1000 + (i % 3), order=ORDER_FAST, d_winners=o.data_ptr(),
                                 stream=streams[i % 4].cuda_stream)
    torch.cuda.synchronize()
    expect = [port.playouts(st, key=1000 + k, order=ORDER_FAST)[0] for k in range(3)]
    for i, o in enumerate(outs):
        assert np.array_equal(o.cpu().numpy(), expect[i % 3]), "launch %d" % i


def test_async_count_slots_equal_the_blocking_call(engine):
    """b2p_run_counts_async / b2p_wait_slot (what the pipelined search is built on): three batches in flight on three
    slots from page-locked buffers give the counts of the blocking call; a busy slot refuses a second batch."""
    import gpu_ai_b200 as b
    from gpu_ai_b200.engine import PinnedArray
    multi = b.Engine(devices=_multi_device_ids(), seed=1)
    sizes = (30000, 777, 52001)
    bufs = []
    for slot, n in enumerate(sizes):
        st = engine.gen_leaves(n, key=40 + slot)
        ps, pw = PinnedArray((n, 4), np.uint32), PinnedArray((n, 2), np.uint32)
        ps.array[:] = st
        bufs.append((st, ps, pw))
        multi.run_counts_async(slot, ps.array, pw.array, reps=5, key=9 + slot, pid_base=100 * slot)
    with pytest.raises(b.B2PError, match="in flight"):
        multi.run_counts_async(1, bufs[1][1].array, bufs[1][2].array, reps=5)
    for slot in (2, 0, 1):
        st, ps, pw = bufs[slot]
        counters, kernel_ms = multi.wait_slot(slot)
        wins, c = engine.run_counts(st, reps=5, key=9 + slot, pid_base=100 * slot)
        assert np.array_equal(pw.array, wins) and np.array_equal(counters, c) and kernel_ms > 0
    counters, kernel_ms = multi.wait_slot(3)        # idle slot: nothing to wait for
    assert counters.sum() == 0 and kernel_ms == 0


def test_uct_search_on_gpu(engine):
    """B2P_POLICY_UCT through the pipelined search: deterministic, every playout accounted for, and it concentrates
    its trials (the reference rule spreads them almost evenly over the root moves)."""
    import gpu_ai_b200 as b
    runs = []
    for policy in (1, 1, 0):
        t = b.Tree(START_PACKED)
        st = t.search_ex(engine, iterations=40, initial_batch=2048, max_batch=2048, reps=8, key=5, depth=2, policy=policy)
        mv, tr, w1, w2 = t.root_moves()
        assert st["playouts"] == 40 * 2048 * 8 == tr.sum() == t.info()["total_trials"] and ((w1 + w2) <= tr).all()
        runs.append(tr.copy())
        assert t.robust_move(0) in set(int(x) for x in mv)
    assert np.array_equal(runs[0], runs[1])
    assert runs[0].max() / runs[0].sum() > runs[2].max() / runs[2].sum()


def test_pinned_host_buffers(engine):
    import gpu_ai_b200 as b
    from gpu_ai_b200.engine import PinnedArray
    st = engine.gen_leaves(50000, key=3)
    pin_s, pin_w = PinnedArray((50000, 4), np.uint32), PinnedArray((50000,), np.int8)
    pin_s.array[:] = st
    w, _, _, c = engine.run_packed(st, key=4, order=ORDER_FAST)
    w2, _, _, c2 = engine.run_packed(pin_s.array, key=4, order=ORDER_FAST, winners_out=pin_w.array)
    assert w2 is pin_w.array and np.array_equal(w, w2) and np.array_equal(c, c2)
    lib = b.load_library()
    assert lib.b2p_alloc_host(None, 16) == -1 and lib.b2p_free_host(None) == 0


# ---- the real drop-in binary: the reference's run_ai linked against shim/playout_shim.cpp + libb2p.so ---------------
def _run_ai(*args, timeout=600):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "shim", "_ref", "run_ai_b200")
    if not os.path.exists(exe):
        pytest.skip("shim/_ref/run_ai_b200 not built (needs /root/reference in the build container: make -C shim)")
    r = subprocess.run([exe] + list(args), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout,
                       env=dict(os.environ, B2P_ROUTING_REPORT="1"))
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout


def _tallies(text):
    import re
    d = [int(x) for x in re.findall(r"Games drawn: (\d+)", text)]
    a = [int(x) for x in re.findall(r"Player 1 wins: (\d+)", text)]
    c = [int(x) for x in re.findall(r"Player 2 wins: (\d+)", text)]
    assert len(d) == 2 and len(a) == 2 and len(c) == 2, text[-1500:]
    return [(d[i], a[i], c[i]) for i in range(2)]


def test_drop_in_binary_hybrid_routes_tiny_batches_to_the_host():
    """below the device's launch-latency floor (a handful of playouts) the re-tuned hybrid driver plays on the host"""
    import json
    out = _run_ai("-m", "playout_test", "-n", "1", "-1", "hybrid", "-2", "hybrid")
    rep = [json.loads(line) for line in out.splitlines() if line.startswith('{"b2p_routing"')]
    assert rep and rep[0]["calls_host"] >= 1 and rep[0]["calls_host"] + rep[0]["calls_device"] == 2, rep


def test_drop_in_binary_gen_moves_test():
    """`run_ai -m gen_moves_test` (src/driver.cpp:106-117): the reference's own host State::genMoves against
    the B200 move generator through the shim's genMovesTest, 20000 random states of ITS genRandomStates."""
    out = _run_ai("-m", "gen_moves_test", "-n", "20000")
    assert "Passed" in out and "Mismatch" not in out and "Failed" not in out, out[-1500:]


@pytest.mark.parametrize("dev,host", [("device_single", "host"), ("device_multiple", "host"), ("device_coarse", "host"),
                                      ("device_heuristic", "host_heuristic"), ("hybrid", "host"), ("optimal", "host"),
                                      ("optimal_heuristic", "host_heuristic")])
def test_drop_in_binary_playout_test(dev, host):
    """`run_ai -m playout_test -n 100000 -1 <device driver> -2 <host driver>` (src/driver.cpp:119-170): both
    drivers play the same 100000 leaves inside the reference binary; outcome tallies within |z| < 4."""
    n = 100000
    out = _run_ai("-m", "playout_test", "-n", str(n), "-1", dev, "-2", host)
    (d0, a0, c0), (d1, a1, c1) = _tallies(out)
    if dev in ("hybrid", "optimal_heuristic"):
        # the B200 re-tuning of HybridPlayoutDriver / OptimalHeuristicPlayoutDriver (shim/hybrid_b200.cpp): a batch
        # of this size goes whole to the device
        import json
        rep = [json.loads(line) for line in out.splitlines() if line.startswith('{"b2p_routing"')]
        assert rep and rep[0]["playouts_device"] == n and rep[0]["playouts_host"] == 0, rep
    assert d0 + a0 + c0 == n and d1 + a1 + c1 == n
    for x, y in ((d0, d1), (a0, a1), (c0, c1)):
        assert abs(_z(x, n, y, n)) < 4.0, "%s %s vs %s %s" % (dev, (d0, a0, c0), host, (d1, a1, c1))


def test_b200_mcts_player_plays_legal_moves(engine, port):
    """Player-interface mirror: a short game of the B200 searcher against uniformly random replies."""
    import gpu_ai_b200 as b
    rng = np.random.default_rng(3)
    me = b.B200MCTSPlayer(engine=engine, seconds=0.05, initial_batch=2048, reps=8)
    me.start()
    state_tree = b.Tree(START_PACKED)   # the arbiter's copy of the game
    for ply in range(24):
        info = state_tree.info()
        state = info["root_state"]
        if info["root_moves"] == 0 or (state[3] >> 8) >= 50:
            break
        legal, cnt = port.genmoves(state.reshape(1, 4), 64)
        legal = [int(x) for x in legal[0, :cnt[0]]]
        mv = me.getMove(state, verbose=False) if ply % 2 == 0 else legal[int(rng.integers(len(legal)))]
        assert mv in legal
        me.move(mv)
        state_tree.move(mv)
    assert me.playouts > 0 and me.getName() == "mcts_b200"
