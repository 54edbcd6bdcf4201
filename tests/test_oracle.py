"""CPU: pin the C restatement (oracle/checkers_oracle.c) against the reference.

Three anchors: (1) golden vectors produced by the unmodified reference (tools/make_golden.py),
(2) the live reference build oracle/_ref when present, (3) published known answers: the perft
table of SURVEY.md 8c and the Random123 Philox4x32-10 known-answer vectors."""
import numpy as np
import pytest

from conftest import fast_synthetic, unflatten
from oracle.pyoracle import (MODE_HEURISTIC, MODE_RANDOM, ORDER_CANONICAL, ORDER_FAST, START_PACKED, decode_move, rc)

PERFT = [7, 49, 302, 1469, 7361, 36768, 179740, 845931, 3963680, 18391564, 85242128, 388623644]


def test_philox_known_answers(port):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert port.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert port.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_perft_matches_reference_table(port, golden):
    assert list(golden["perft_start"]) == PERFT[:10]
    for d in range(1, 10):
        assert port.perft(START_PACKED, d) == PERFT[d - 1]


def test_reference_layout(golden):
    # sizeof(State), sizeof(Move), sizeof(BoardItem), offsetof(turn), offsetof(msc), type, owner, sizeof(PlayerId)
    assert list(golden["layout"]) == [776, 38, 12, 768, 772, 4, 8, 4]


@pytest.mark.parametrize("name", ["leaves", "synth", "kat"])
def test_port_movelists_equal_golden(port, golden, name):
    st, cnt = golden[name + "_states"], golden[name + "_counts"]
    mv, c = port.genmoves(st, 64)
    assert np.array_equal(c, cnt)
    assert np.array_equal(mv, unflatten(golden[name + "_moves_flat"], cnt))


def test_known_answer_positions(port, golden):
    names = list(golden["kat_names"])
    mv = unflatten(golden["kat_moves_flat"], golden["kat_counts"])

    def moves(n):
        i = names.index(n)
        return [decode_move(x) for x in mv[i, :golden["kat_counts"][i]]]

    start = moves("start")
    assert [(rc(m["from"]), rc(m["to"])) for m in start] == [((2, 1), (3, 2)), ((2, 1), (3, 0)), ((2, 3), (3, 4)),
                                                             ((2, 3), (3, 2)), ((2, 5), (3, 6)), ((2, 5), (3, 4)),
                                                             ((2, 7), (3, 6))]
    cyc = moves("king_cycle")  # king may not land on its origin: two 3-hop sequences
    assert [m["hops"] for m in cyc] == [3, 3]
    assert [rc(s) for s in cyc[0]["via"]] == [(4, 3), (2, 5), (0, 3)]
    assert [rc(s) for s in cyc[1]["via"]] == [(0, 3), (2, 5), (4, 3)]
    fan = moves("man_fan")  # left before right
    assert [[rc(s) for s in m["via"]] for m in fan] == [[(3, 0), (5, 2)], [(3, 4), (5, 2)], [(3, 4), (5, 6)]]
    promo = moves("promotion_ends_capture")
    assert len(promo) == 1 and promo[0]["hops"] == 1 and promo[0]["promoted"] == 1 and rc(promo[0]["to"]) == (7, 4)
    direct = moves("direct_order")
    assert [(rc(m["from"]), rc(m["to"]), m["promoted"]) for m in direct] == [((0, 1), (1, 2), 0), ((0, 1), (1, 0), 0),
                                                                             ((6, 1), (7, 2), 1), ((6, 1), (7, 0), 1)]
    w = dict(zip(names, golden["kat_rc_winners"]))
    assert w["no_pieces_to_move"] == 1   # side to move has no pieces: the other player wins
    assert w["draw_counter"] == -1       # msc >= 50 beats everything


@pytest.mark.parametrize("name", ["leaves", "synth"])
@pytest.mark.parametrize("tag,mode,order", [("rc", MODE_RANDOM, ORDER_CANONICAL), ("rf", MODE_RANDOM, ORDER_FAST),
                                            ("h", MODE_HEURISTIC, ORDER_CANONICAL)])
def test_port_playouts_equal_golden(port, golden, name, tag, mode, order):
    st = golden[name + "_states"]
    w, p, f, c = port.playouts(st, reps=2, key=12345, pid_base=1000, mode=mode, order=order, want_final=True)
    assert np.array_equal(w, golden["%s_%s_winners" % (name, tag)])
    assert np.array_equal(p, golden["%s_%s_plies" % (name, tag)])
    assert np.array_equal(f, golden["%s_%s_final" % (name, tag)])
    assert np.array_equal(c, golden["%s_%s_counters" % (name, tag)])


@pytest.mark.parametrize("name", ["leaves", "synth"])
def test_port_truncated_playouts_equal_golden(port, golden, name):
    w, p, f, c = port.playouts(golden[name + "_states"], key=12345, max_plies=5, want_final=True)
    assert np.array_equal(w, golden[name + "_cut5_winners"])
    assert np.array_equal(f, golden[name + "_cut5_final"])


def test_port_leaves_equal_golden(port, golden):
    assert np.array_equal(port.gen_leaves(4096, key=2016), golden["leaves_states"])


def test_port_pack776_roundtrip(port, golden):
    s776 = golden["states776_sample"]
    assert np.array_equal(port.pack776(s776), golden["leaves_states"][:64])
    assert np.array_equal(port.unpack776(golden["leaves_states"][:64]), s776)


# ---- live reference (only in the build container) ---------------------------------------------------
def test_port_equals_live_reference_movelists(port, ref):
    st = np.concatenate([ref.gen_leaves(200000, key=99), fast_synthetic(200000, 5)])
    a, ca = port.genmoves(st, 64)
    b, cb = ref.genmoves(st, 64)
    assert np.array_equal(ca, cb) and np.array_equal(a, b)


@pytest.mark.parametrize("mode,order", [(MODE_RANDOM, ORDER_CANONICAL), (MODE_RANDOM, ORDER_FAST), (MODE_HEURISTIC, ORDER_CANONICAL)])
def test_port_equals_live_reference_playouts(port, ref, mode, order):
    st = np.concatenate([ref.gen_leaves(15000, key=7), fast_synthetic(15000, 11)])
    a = port.playouts(st, key=4242, pid_base=5, mode=mode, order=order, want_final=True)
    b = ref.playouts(st, key=4242, pid_base=5, mode=mode, order=order, want_final=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_live_reference_perft_10(ref):
    assert ref.perft(START_PACKED, 10) == PERFT[9]
