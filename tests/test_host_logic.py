"""CPU: the product's bit logic (bitboard.cuh / philox.cuh / playout_core.cuh compiled with g++ by
tests/host_build) against the oracle.  Same code the kernels inline; no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from conftest import fast_synthetic, unflatten
from oracle.pyoracle import MODE_HEURISTIC, MODE_RANDOM, ORDER_CANONICAL, ORDER_FAST


def test_hostbuild_philox(hostbuild):
    o = (C.c_uint32 * 4)()
    hostbuild.lib.hb_philox((C.c_uint32 * 4)(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
                            (C.c_uint32 * 2)(0xa4093822, 0x299f31d0), o)
    assert list(o) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("name", ["leaves", "synth", "kat"])
def test_bitboard_movelists_equal_golden(hostbuild, golden, name):
    st, cnt = golden[name + "_states"], golden[name + "_counts"]
    mv, c = hostbuild.genmoves(st, 64)
    assert np.array_equal(c, cnt)
    assert np.array_equal(mv, unflatten(golden[name + "_moves_flat"], cnt))


def test_bitboard_movelists_million_positions(hostbuild, port):
    """bit-exact genMoves on >= 10^6 positions: reachable leaves + king-rich synthetic boards"""
    st = np.concatenate([port.gen_leaves(600000, key=31337), fast_synthetic(600000, 3)])
    a, ca = hostbuild.genmoves(st, 64)
    b, cb = port.genmoves(st, 64)
    assert ca.max() <= 64
    assert np.array_equal(ca, cb)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["leaves", "synth"])
@pytest.mark.parametrize("tag,mode,order", [("rc", MODE_RANDOM, ORDER_CANONICAL), ("rf", MODE_RANDOM, ORDER_FAST),
                                            ("h", MODE_HEURISTIC, ORDER_CANONICAL)])
def test_ply_step_equals_golden(hostbuild, golden, name, tag, mode, order):
    st = golden[name + "_states"]
    w, p, f, c = hostbuild.playouts(st, reps=2, key=12345, pid_base=1000, mode=mode, order=order, want_final=True)
    assert np.array_equal(w, golden["%s_%s_winners" % (name, tag)])
    assert np.array_equal(p, golden["%s_%s_plies" % (name, tag)])
    assert np.array_equal(f, golden["%s_%s_final" % (name, tag)])
    assert np.array_equal(c, golden["%s_%s_counters" % (name, tag)])


@pytest.mark.parametrize("mode,order", [(MODE_RANDOM, ORDER_CANONICAL), (MODE_RANDOM, ORDER_FAST), (MODE_HEURISTIC, ORDER_CANONICAL)])
def test_ply_step_equals_port_large(hostbuild, port, mode, order):
    st = np.concatenate([port.gen_leaves(40000, key=77), fast_synthetic(40000, 19)])
    a = hostbuild.playouts(st, key=99, pid_base=12, mode=mode, order=order, want_final=True)
    b = port.playouts(st, key=99, pid_base=12, mode=mode, order=order, want_final=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_leafgen_equals_golden(hostbuild, golden):
    assert np.array_equal(hostbuild.gen_leaves(4096, key=2016), golden["leaves_states"])


def test_playouts_from_the_initial_position(hostbuild, port):
    """D_start (BASELINE configs[0]): many playouts of the one starting state -- long games (mean ~68 plies),
    every rule exercised from a reachable root; bit-exact against the port in both orders."""
    from oracle.pyoracle import START_PACKED
    st = np.tile(START_PACKED, (1, 1))
    for order in (ORDER_CANONICAL, ORDER_FAST):
        a = hostbuild.playouts(st, reps=20000, key=5, order=order, want_final=True)
        b = port.playouts(st, reps=20000, key=5, order=order, want_final=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert 60 < a[1].mean() < 76   # 68.4 plies per playout from the start


def test_bitboard_movelists_hypothesis(hostbuild, port):
    """property test: arbitrary piece placements (any overlap-free p1/p2/kings words, any turn/msc)"""
    from hypothesis import given, settings, strategies as st32
    word = st32.integers(min_value=0, max_value=2**32 - 1)

    @settings(max_examples=300, deadline=None)
    @given(word, word, word, st32.integers(0, 1), st32.integers(0, 60))
    def check(a, b, k, turn, msc):
        state = np.array([[a & ~b, b & ~a, k, turn | (msc << 8)]], dtype=np.uint32)
        m1, c1 = hostbuild.genmoves(state, 100)
        m2, c2 = port.genmoves(state, 100)
        assert c1[0] == c2[0] and np.array_equal(m1, m2)

    check()


def test_two_level_select_experiment_matches(hostbuild):
    """B2P_SELECT2 candidate (bitboard.cuh): same (origin, slot) as the shipped origin-major search on random masks"""
    f = hostbuild.lib.hb_select_mismatches
    f.restype = C.c_uint64
    assert f(C.c_uint64(7), C.c_uint64(200000)) == 0


def test_integer_noise_scan_premises(port):
    """heuristic_ply's SCAN path replaces 'first maximum of base + gauss(draw)' by 'maximum draw, lowest index'.
    That is exact iff (1) gauss() is strictly increasing in the 16-bit draw and (2) its smallest step is larger
    than one ulp of any sum the scan sees (base weights are at most 48 = eleven kings and a crowning man against
    one man, noise below 0.39 => sums below 64, where one ulp is 2^-18)."""
    g = np.array([port.gauss(h) for h in range(0, 65536)], dtype=np.float32)
    step = np.diff(g.astype(np.float64))
    assert (step > 0).all()
    ulp_max = float(np.spacing(np.float32(63.9)))
    assert ulp_max == 2.0 ** -18 and step.min() > 4.2e-6 and step.min() > 1.1 * ulp_max
    assert abs(g).max() < 0.39
