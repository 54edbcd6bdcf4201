"""ctypes front-end for the CPU checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Loads oracle/liboracle.so (plain-C restatement, `make -C oracle`) and, when it has been built
in the container that holds /root/reference, oracle/_ref/libref_harness.so (the unmodified
reference rules).  Both expose the same batch API, so `Checker("port")` and `Checker("reference")`
are interchangeable in tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under gpu_ai_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_harness.so")

ORDER_CANONICAL = 0
ORDER_FAST = 1
MODE_RANDOM = 0
MODE_HEURISTIC = 1


def build_port():
    """Compile the C restatement (gcc only; works on the GPU box)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def have_reference():
    return os.path.exists(REF_SO)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Checker:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "checkers_oracle.c")):
                build_port()
            self.lib = C.CDLL(PORT_SO)
            self.pfx = "or_"
        elif kind == "reference":
            self.lib = C.CDLL(REF_SO)
            self.pfx = "ref_"
        else:
            raise ValueError(kind)

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    # ---- move lists -----------------------------------------------------
    def genmoves(self, packed, max_moves=64):
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1, 4)
        n = packed.shape[0]
        moves = np.zeros((n, max_moves), dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint8)
        f = self._f("genmoves_batch")
        f.restype = None
        f(_p(packed, C.c_uint32), C.c_size_t(n), C.c_int(max_moves), _p(moves, C.c_uint64), _p(counts, C.c_uint8))
        return moves, counts

    # ---- playouts -------------------------------------------------------
    def playouts(self, packed, reps=1, key=12345, pid_base=0, mode=MODE_RANDOM, order=ORDER_CANONICAL,
                 max_plies=-1, want_final=False):
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1, 4)
        n = packed.shape[0]
        total = n * reps
        winners = np.zeros(total, dtype=np.int8)
        plies = np.zeros(total, dtype=np.uint32)
        final = np.zeros((total, 4), dtype=np.uint32) if want_final else None
        counters = np.zeros(4, dtype=np.uint64)
        f = self._f("playouts_batch")
        f.restype = None
        f(_p(packed, C.c_uint32), C.c_size_t(n), C.c_uint32(reps), C.c_uint64(key), C.c_uint64(pid_base), C.c_int(mode),
          C.c_int(order), C.c_int(max_plies), _p(winners, C.c_int8), _p(plies, C.c_uint32),
          _p(final, C.c_uint32) if want_final else None, _p(counters, C.c_uint64))
        return winners, plies, final, counters

    def gen_leaves(self, n, key=2016, first_index=0):
        out = np.zeros((n, 4), dtype=np.uint32)
        f = self._f("gen_leaves")
        f.restype = None
        f(C.c_size_t(n), C.c_uint64(key), C.c_uint64(first_index), _p(out, C.c_uint32))
        return out

    def perft(self, packed, depth):
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(4)
        f = self._f("perft_packed")
        f.restype = C.c_uint64
        return int(f(_p(packed, C.c_uint32), C.c_int(depth)))

    def pack776(self, states776):
        buf = np.ascontiguousarray(states776, dtype=np.uint8).reshape(-1, 776)
        out = np.zeros((buf.shape[0], 4), dtype=np.uint32)
        f = self._f("pack776_batch")
        f.restype = None
        f(_p(buf, C.c_uint8), C.c_size_t(buf.shape[0]), _p(out, C.c_uint32))
        return out

    def unpack776(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros((packed.shape[0], 776), dtype=np.uint8)
        f = self._f("unpack776_batch")
        f.restype = None
        f(_p(packed, C.c_uint32), C.c_size_t(packed.shape[0]), _p(out, C.c_uint8))
        return out

    # ---- reference-only -------------------------------------------------
    def host_driver(self, packed, mode=MODE_RANDOM):
        """The reference's own HostPlayoutDriver / HostHeuristicPlayoutDriver (its RNG and all)."""
        assert self.kind == "reference"
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros(packed.shape[0], dtype=np.int32)
        f = self.lib.ref_host_driver_run
        f.restype = C.c_int
        rc = f(_p(packed, C.c_uint32), C.c_size_t(packed.shape[0]), C.c_int(mode), _p(out, C.c_int32))
        if rc != 0:
            raise RuntimeError("reference host driver failed")
        return out

    def layout(self):
        assert self.kind == "reference"
        out = (C.c_int * 8)()
        self.lib.ref_layout(out)
        return list(out)

    # ---- port-only helpers ---------------------------------------------
    def philox(self, ctr, key):
        assert self.kind == "port"
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.or_philox(c, k, o)
        return list(o)

    def draw(self, key, pid, domain, t):
        assert self.kind == "port"
        self.lib.or_draw.restype = C.c_uint32
        return int(self.lib.or_draw(C.c_uint64(key), C.c_uint64(pid), C.c_uint32(domain), C.c_uint32(t)))

    def gauss(self, r):
        assert self.kind == "port"
        self.lib.or_gauss.restype = C.c_float
        return float(self.lib.or_gauss(C.c_uint32(r)))


# ---- pure-python helpers on the packed format (tests only) -----------------

START_PACKED = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)


def sq(row, col):
    """(row, col) -> square index i = row*4 + col//2 (dark squares only)."""
    assert (row + col) % 2 == 1, "not a dark square"
    return row * 4 + col // 2


def make_state(p1_men=(), p1_kings=(), p2_men=(), p2_kings=(), turn=0, msc=0):
    p1 = p2 = k = 0
    for (r, c) in p1_men:
        p1 |= 1 << sq(r, c)
    for (r, c) in p1_kings:
        p1 |= 1 << sq(r, c)
        k |= 1 << sq(r, c)
    for (r, c) in p2_men:
        p2 |= 1 << sq(r, c)
    for (r, c) in p2_kings:
        p2 |= 1 << sq(r, c)
        k |= 1 << sq(r, c)
    return np.array([p1, p2, k, (turn & 1) | (msc << 8)], dtype=np.uint32)


def decode_move(e):
    e = int(e)
    hops = (e >> 10) & 7
    return {
        "from": e & 31, "to": (e >> 5) & 31, "hops": hops, "promoted": (e >> 13) & 1,
        "via": [(e >> (16 + 5 * k)) & 31 for k in range(hops)],
    }


def rc(i):
    r = i >> 2
    return (r, 2 * (i & 3) + (1 - (r & 1)))
