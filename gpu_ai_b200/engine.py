"""ctypes binding of include/b2p.h (libb2p.so).  Thin: argument marshalling and error mapping only."""
import ctypes as C
import os

import numpy as np

MODE_RANDOM, MODE_HEURISTIC = 0, 1
SCHED_THREAD, SCHED_WARP, SCHED_AUTO = 0, 1, 2
ORDER_CANONICAL, ORDER_FAST = 0, 1
UNFINISHED = 2

STATE776_BYTES = 776  # sizeof(State) of the reference (src/state.hpp:119-122)
MOVE_BYTES = 38       # sizeof(Move)  of the reference (src/state.hpp:253-296)

_HERE = os.path.dirname(os.path.abspath(__file__))


class B2PError(RuntimeError):
    pass


def lib_path():
    # B2P_LIB_PATH: load an experiment build (make -C gpu_ai_b200/csrc variant ...) instead of the shipped library
    return os.environ.get("B2P_LIB_PATH") or os.path.join(_HERE, "libb2p.so")


class DevInfo(C.Structure):
    _fields_ = [("device_id", C.c_int), ("sm_count", C.c_int), ("clock_khz", C.c_int), ("cc_major", C.c_int),
                ("cc_minor", C.c_int), ("total_mem", C.c_size_t), ("name", C.c_char * 128)]


# every symbol include/b2p.h declares: (name, restype, argtypes)
_VP, _U64, _U32, _SZ, _INT = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int
ABI = [
    ("b2p_create", _INT, [C.POINTER(_VP), C.POINTER(C.c_int), _INT, _U64]),
    ("b2p_destroy", None, [_VP]),
    ("b2p_last_error", C.c_char_p, [_VP]),
    ("b2p_device_count", _INT, [_VP]),
    ("b2p_device_info", _INT, [_VP, _INT, C.POINTER(DevInfo)]),
    ("b2p_version", C.c_char_p, []),
    ("b2p_run_states776", _INT, [_VP, _VP, _SZ, _INT, _INT, _VP]),
    ("b2p_run_packed", _INT, [_VP, _VP, _SZ, _U32, _U64, _U64, _INT, _INT, _INT, _INT, _VP, _VP, _VP, _VP]),
    ("b2p_run_counts", _INT, [_VP, _VP, _SZ, _U32, _U64, _U64, _INT, _INT, _INT, _VP, _VP]),
    ("b2p_run_counts_async", _INT, [_VP, _INT, _VP, _SZ, _U32, _U64, _U64, _INT, _INT, _INT, _VP]),
    ("b2p_wait_slot", _INT, [_VP, _INT, _VP, C.POINTER(C.c_float)]),
    ("b2p_slot_staging", _INT, [_VP, _INT, _SZ, C.POINTER(_VP), C.POINTER(_VP)]),
    ("b2p_genmoves", _INT, [_VP, _VP, _SZ, _INT, _VP, _VP]),
    ("b2p_run_packed_device", _INT, [_VP, _INT, _VP, _SZ, _U32, _U64, _U64, _INT, _INT, _INT, _INT, _VP, _VP, _VP, _VP, _VP]),
    ("b2p_genmoves_device", _INT, [_VP, _INT, _VP, _SZ, _INT, _VP, _VP, _VP]),
    ("b2p_gen_leaves_device", _INT, [_VP, _INT, _SZ, _U64, _U64, _VP, _VP]),
    ("b2p_gen_leaves", _INT, [_VP, _SZ, _U64, _U64, _VP]),
    ("b2p_sync", _INT, [_VP]),
    ("b2p_alloc_host", _INT, [C.POINTER(_VP), _SZ]),
    ("b2p_free_host", _INT, [_VP]),
    ("b2p_pack776", _INT, [_VP, _SZ, _VP]),
    ("b2p_pack776_impl", C.c_char_p, []),
    ("b2p_unpack776", _INT, [_VP, _SZ, _VP]),
    ("b2p_expand_move", _INT, [_U64, _VP]),
    ("b2p_microbench", _INT, [_VP, _INT, _INT, _INT, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("b2p_launch_count", _U64, [_VP]),
    ("b2p_tree_create", _INT, [C.POINTER(_VP), _VP]),
    ("b2p_tree_destroy", None, [_VP]),
    ("b2p_tree_select", _INT, [_VP, _U32, _VP, C.POINTER(_U32)]),
    ("b2p_tree_update", _INT, [_VP, _VP, _U32, _U32]),
    ("b2p_tree_update_counts", _INT, [_VP, _VP, _U32, _U32]),
    ("b2p_tree_best_move", _INT, [_VP, _INT, C.POINTER(_U64)]),
    ("b2p_tree_robust_move", _INT, [_VP, _INT, C.POINTER(_U64)]),
    ("b2p_tree_move", _INT, [_VP, _U64]),
    ("b2p_tree_info", _INT, [_VP, _VP]),
    ("b2p_tree_root_moves", _INT, [_VP, _VP, _VP, _VP, _VP, _U32]),
    ("b2p_tree_last_error", C.c_char_p, [_VP]),
    ("b2p_tree_search", _INT, [_VP, _VP, _U32, C.c_double, _U32, C.c_float, _U32, _INT, _U64, C.POINTER(_U64)]),
    ("b2p_tree_search_ex", _INT, [_VP, _VP, _VP, _VP]),
    ("b2p_tree_select_batch", _INT, [_VP, _INT, _U32, _U32, _INT, _INT, _VP]),
    ("b2p_tree_update_batch", _INT, [_VP, _INT, _VP, _INT]),
]

_lib = None


def load_library():
    """Load libb2p.so and bind every ABI symbol.  Fails loudly if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise B2PError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C gpu_ai_b200/csrc`. There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, res, args in ABI:
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _as_packed(states):
    a = np.ascontiguousarray(states, dtype=np.uint32)
    if a.ndim == 1:
        a = a.reshape(-1, 4)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("packed states must have shape (n, 4) uint32")
    return a


class Engine:
    """One b2p context (= the set of GPUs one playout driver uses)."""

    def __init__(self, devices=None, seed=12345):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if devices is None:
            ids, n = None, 0
        elif isinstance(devices, int):
            ids, n = None, devices
        else:
            devices = list(devices)
            ids, n = (C.c_int * len(devices))(*devices), len(devices)
        rc = self.lib.b2p_create(C.byref(self.ctx), ids, n, seed)
        if rc != 0:
            msg = self.lib.b2p_last_error(None)
            self.ctx = None
            raise B2PError("b2p_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.b2p_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise B2PError("b2p error %d: %s" % (rc, self.lib.b2p_last_error(self.ctx).decode()))

    # ---- info ----------------------------------------------------------------------------------
    @property
    def device_count(self):
        return self.lib.b2p_device_count(self.ctx)

    def device_info(self, i=0):
        d = DevInfo()
        self._check(self.lib.b2p_device_info(self.ctx, i, C.byref(d)))
        return {"device_id": d.device_id, "sm_count": d.sm_count, "clock_khz": d.clock_khz,
                "cc": (d.cc_major, d.cc_minor), "total_mem": d.total_mem, "name": d.name.decode()}

    @property
    def launch_count(self):
        return int(self.lib.b2p_launch_count(self.ctx))

    def sync(self):
        self._check(self.lib.b2p_sync(self.ctx))

    # ---- host-buffer calls ------------------------------------------------------------------------
    def run_states776(self, states776, mode=MODE_RANDOM, sched=SCHED_THREAD, out=None):
        """states776: bytes-like / uint8 array of n reference `State` objects.  Returns int32 PlayerId[n]."""
        buf = np.ascontiguousarray(np.frombuffer(states776, dtype=np.uint8) if not isinstance(states776, np.ndarray) else states776, dtype=np.uint8).reshape(-1)
        if buf.size % STATE776_BYTES:
            raise ValueError("buffer is not a whole number of 776-byte States")
        n = buf.size // STATE776_BYTES
        if out is None:
            out = np.empty(n, dtype=np.int32)
        self._check(self.lib.b2p_run_states776(self.ctx, _ptr(buf), n, mode, sched, _ptr(out)))
        return out

    def run_packed(self, states, reps=1, key=12345, pid_base=0, mode=MODE_RANDOM, sched=SCHED_THREAD,
                   order=ORDER_CANONICAL, max_plies=-1, want_winners=True, want_plies=False, want_final=False,
                   winners_out=None):
        """winners_out: optional preallocated int8[n*reps] (e.g. a PinnedArray's .array) to receive the winners."""
        a = _as_packed(states)
        n = a.shape[0]
        total = n * reps
        winners = (winners_out if winners_out is not None else np.empty(total, dtype=np.int8)) if want_winners else None
        plies = np.empty(total, dtype=np.uint32) if want_plies else None
        final = np.empty((total, 4), dtype=np.uint32) if want_final else None
        counters = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.b2p_run_packed(self.ctx, _ptr(a), n, reps, key, pid_base, mode, sched, order, max_plies,
                                            _ptr(winners), _ptr(plies), _ptr(final), _ptr(counters)))
        return winners, plies, final, counters

    def run_counts(self, states, reps, key=12345, pid_base=0, mode=MODE_RANDOM, sched=SCHED_THREAD, order=ORDER_FAST,
                   wins_out=None):
        """`reps` playouts per leaf; returns (wins[n, 2] uint32, counters[4])."""
        a = _as_packed(states)
        n = a.shape[0]
        wins = wins_out if wins_out is not None else np.zeros((n, 2), dtype=np.uint32)
        counters = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.b2p_run_counts(self.ctx, _ptr(a), n, reps, key, pid_base, mode, sched, order, _ptr(wins), _ptr(counters)))
        return wins, counters

    def run_counts_async(self, slot, states, wins_out, reps, key=12345, pid_base=0, mode=MODE_RANDOM, sched=SCHED_THREAD,
                         order=ORDER_FAST):
        """b2p_run_counts_async: `states` (n, 4) uint32 and `wins_out` (n, 2) uint32 must stay alive (ideally
        PinnedArray views) until wait_slot(slot) returns."""
        n = states.shape[0]
        self._check(self.lib.b2p_run_counts_async(self.ctx, slot, _ptr(states), n, reps, key, pid_base, mode, sched, order,
                                                  _ptr(wins_out)))

    def wait_slot(self, slot):
        """returns (counters[4], kernel_ms)"""
        counters = np.zeros(4, dtype=np.uint64)
        ms = C.c_float()
        self._check(self.lib.b2p_wait_slot(self.ctx, slot, _ptr(counters), C.byref(ms)))
        return counters, float(ms.value)

    def genmoves(self, states, max_moves=64):
        a = _as_packed(states)
        n = a.shape[0]
        moves = np.zeros((n, max_moves), dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.b2p_genmoves(self.ctx, _ptr(a), n, max_moves, _ptr(moves), _ptr(counts)))
        return moves, counts

    def gen_leaves(self, n, key=2016, first_index=0):
        out = np.empty((n, 4), dtype=np.uint32)
        self._check(self.lib.b2p_gen_leaves(self.ctx, n, key, first_index, _ptr(out)))
        return out

    # ---- device-resident calls (raw device pointers, e.g. torch tensor.data_ptr()) ---------------------
    def run_packed_device(self, d_states, n, reps=1, key=12345, pid_base=0, mode=MODE_RANDOM, sched=SCHED_THREAD,
                          order=ORDER_FAST, max_plies=-1, d_winners=None, d_plies=None, d_final=None, d_counters=None,
                          stream=None, dev_index=0):
        self._check(self.lib.b2p_run_packed_device(self.ctx, dev_index, d_states, n, reps, key, pid_base, mode, sched,
                                                   order, max_plies, d_winners, d_plies, d_final, d_counters, stream))

    def genmoves_device(self, d_states, n, max_moves, d_moves, d_counts, stream=None, dev_index=0):
        self._check(self.lib.b2p_genmoves_device(self.ctx, dev_index, d_states, n, max_moves, d_moves, d_counts, stream))

    def gen_leaves_device(self, n, d_out, key=2016, first_index=0, stream=None, dev_index=0):
        self._check(self.lib.b2p_gen_leaves_device(self.ctx, dev_index, n, key, first_index, d_out, stream))

    def microbench(self, which, iters=2000, dev_index=0):
        ops, ms = C.c_double(), C.c_double()
        self._check(self.lib.b2p_microbench(self.ctx, dev_index, which, iters, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value


class SearchOpts(C.Structure):
    _fields_ = [("iterations", _U32), ("seconds", C.c_double), ("initial_batch", _U32), ("scale", C.c_float),
                ("max_batch", _U32), ("reps", _U32), ("mode", _INT), ("key", _U64), ("threads", _INT), ("depth", _INT),
                ("policy", _INT)]


class SearchStats(C.Structure):
    _fields_ = [("playouts", _U64), ("leaves", _U64), ("batches", _U64), ("nodes", _U64), ("seconds", C.c_double),
                ("select_s", C.c_double), ("update_s", C.c_double), ("wait_s", C.c_double), ("kernel_s", C.c_double),
                ("threads", _U32), ("depth", _U32)]


class TreeStats(C.Structure):
    _fields_ = [("nodes", _U64), ("total_trials", _U64), ("wins_p1", _U64), ("wins_p2", _U64), ("root_children", _U32),
                ("root_moves", _U32), ("root_state", _U32 * 4)]


class Tree:
    """b2p_tree: the packed-state counterpart of the reference's GameTree (src/mcts.hpp:12-64)."""

    def __init__(self, root_state):
        self.lib = load_library()
        self.h = C.c_void_p()
        root = np.ascontiguousarray(root_state, dtype=np.uint32).reshape(4)
        if self.lib.b2p_tree_create(C.byref(self.h), _ptr(root)) != 0:
            raise B2PError("b2p_tree_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2p_tree_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise B2PError("b2p_tree error %d: %s" % (rc, self.lib.b2p_tree_last_error(self.h).decode()))

    def select(self, trials):
        out = np.empty((max(trials, 1), 4), dtype=np.uint32)
        n = _U32()
        self._check(self.lib.b2p_tree_select(self.h, trials, _ptr(out), C.byref(n)))
        return out[: n.value]

    def update(self, winners, reps=1):
        w = np.ascontiguousarray(winners, dtype=np.int8)
        self._check(self.lib.b2p_tree_update(self.h, _ptr(w), w.size // reps, reps))

    def update_counts(self, wins, reps):
        w = np.ascontiguousarray(wins, dtype=np.uint32)
        self._check(self.lib.b2p_tree_update_counts(self.h, _ptr(w), w.size // 2, reps))

    def best_move(self, player):
        m = _U64()
        self._check(self.lib.b2p_tree_best_move(self.h, player, C.byref(m)))
        return int(m.value)

    def robust_move(self, player):
        m = _U64()
        self._check(self.lib.b2p_tree_robust_move(self.h, player, C.byref(m)))
        return int(m.value)

    def move(self, move):
        self._check(self.lib.b2p_tree_move(self.h, int(move)))

    def info(self):
        st = TreeStats()
        self._check(self.lib.b2p_tree_info(self.h, C.byref(st)))
        return {"nodes": st.nodes, "total_trials": st.total_trials, "wins": (st.wins_p1, st.wins_p2),
                "root_children": st.root_children, "root_moves": st.root_moves,
                "root_state": np.array(list(st.root_state), dtype=np.uint32)}

    def root_moves(self):
        cap = 128
        mv, tr = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        w1, w2 = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        n = self.lib.b2p_tree_root_moves(self.h, _ptr(mv), _ptr(tr), _ptr(w1), _ptr(w2), cap)
        return mv[:n], tr[:n], w1[:n], w2[:n]

    def search(self, engine, iterations=0, seconds=0.0, initial_batch=50, scale=0.02, reps=1, mode=MODE_RANDOM, key=1):
        played = _U64()
        rc = self.lib.b2p_tree_search(engine.ctx, self.h, iterations, seconds, initial_batch, scale, reps, mode, key, C.byref(played))
        self._check(rc)
        return int(played.value)

    def select_batch(self, slot, trials, reps=1, threads=1, exact=True, policy=None):
        """policy: 0 reference rule / reference arithmetic (exact=True), 1 reference rule / float, 2 UCT"""
        out = np.empty((max(trials, 1), 4), dtype=np.uint32)
        pol = policy if policy is not None else (0 if exact else 1)
        self._check(self.lib.b2p_tree_select_batch(self.h, slot, trials, reps, threads, pol, _ptr(out)))
        return out[:trials]

    def update_batch(self, slot, wins, threads=1):
        w = np.ascontiguousarray(wins, dtype=np.uint32)
        self._check(self.lib.b2p_tree_update_batch(self.h, slot, _ptr(w), threads))

    def search_ex(self, engine, iterations=0, seconds=0.0, initial_batch=50, scale=0.02, max_batch=0, reps=1,
                  mode=MODE_RANDOM, key=1, threads=0, depth=0, policy=0):
        """b2p_tree_search_ex: returns the b2p_search_stats as a dict.  policy: 0 reference allocation, 1 UCT."""
        o = SearchOpts(iterations, seconds, initial_batch, scale, max_batch, reps, mode, key, threads, depth, policy)
        st = SearchStats()
        self._check(self.lib.b2p_tree_search_ex(engine.ctx, self.h, C.byref(o), C.byref(st)))
        return {k: getattr(st, k) for k, _ in SearchStats._fields_}


class PinnedArray:
    """numpy view of a page-locked buffer from b2p_alloc_host (freed when the object dies)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        self.ptr = C.c_void_p()
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dt.itemsize
        rc = self.lib.b2p_alloc_host(C.byref(self.ptr), nbytes)
        if rc != 0:
            raise B2PError("b2p_alloc_host failed (%d): %s" % (rc, (self.lib.b2p_last_error(None) or b"").decode()))
        buf = (C.c_uint8 * nbytes).from_address(self.ptr.value) if nbytes else (C.c_uint8 * 0)()
        self.array = np.frombuffer(buf, dtype=dt).reshape(shape)

    def __del__(self):
        if getattr(self, "ptr", None) and self.ptr.value:
            self.array = None
            self.lib.b2p_free_host(self.ptr)
            self.ptr = C.c_void_p()


# ---- converters (no context, no device) -------------------------------------------------------------
def pack776(states776):
    lib = load_library()
    buf = np.ascontiguousarray(states776, dtype=np.uint8).reshape(-1)
    n = buf.size // STATE776_BYTES
    out = np.empty((n, 4), dtype=np.uint32)
    rc = lib.b2p_pack776(_ptr(buf), n, _ptr(out))
    if rc:
        raise B2PError("b2p_pack776 failed: %d" % rc)
    return out


def unpack776(states):
    lib = load_library()
    a = _as_packed(states)
    out = np.empty((a.shape[0], STATE776_BYTES), dtype=np.uint8)
    rc = lib.b2p_unpack776(_ptr(a), a.shape[0], _ptr(out))
    if rc:
        raise B2PError("b2p_unpack776 failed: %d" % rc)
    return out


def expand_move(move):
    lib = load_library()
    out = np.zeros(MOVE_BYTES, dtype=np.uint8)
    rc = lib.b2p_expand_move(int(move), _ptr(out))
    if rc:
        raise B2PError("b2p_expand_move failed: %d" % rc)
    return out
