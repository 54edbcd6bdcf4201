#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_harness.so).

Run in the build container (the only place /root/reference exists):
    make -C oracle ref && python tools/make_golden.py
The reference has no golden vectors of its own, so these fixtures are
outputs of the reference rules executed here; the GPU box checks the C restatement and the CUDA
path against them without needing /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import (Checker, MODE_HEURISTIC, MODE_RANDOM, ORDER_CANONICAL, ORDER_FAST, START_PACKED,  # noqa: E402
                             make_state)

OUT = os.path.join(ROOT, "tests", "golden")
N_LEAVES = 4096
N_SYNTH = 4096
MAX_MOVES = 64
KEY_LEAF = 2016
KEY_PLAY = 12345


def synthetic_positions(n, seed):
    """King-rich random placements (not necessarily reachable): stresses king multi-jumps, the
    origin/visited rule, promotion rows and the draw counter far more than random-prefix leaves."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4), dtype=np.uint32)
    for i in range(n):
        total = int(rng.integers(2, 21))
        squares = rng.choice(32, size=total, replace=False)
        n1 = int(rng.integers(1, total))
        p1 = p2 = k = 0
        for j, s in enumerate(squares):
            if j < n1:
                p1 |= 1 << int(s)
            else:
                p2 |= 1 << int(s)
            if rng.random() < 0.5:
                k |= 1 << int(s)
        msc = int(rng.choice([0, 1, 10, 48, 49, 50, 51]))
        out[i] = (p1, p2, k, int(rng.integers(0, 2)) | (msc << 8))
    return out


def flatten(moves, counts):
    flat = np.concatenate([moves[i, :counts[i]] for i in range(len(counts))]) if len(counts) else np.zeros(0, np.uint64)
    return flat.astype(np.uint64)


def kat_positions():
    """The known-answer positions of SURVEY.md 8c, as packed states."""
    return {
        "start": START_PACKED.copy(),
        "king_cycle": make_state(p1_kings=[(2, 1)], p2_men=[(3, 2), (3, 4), (1, 4), (1, 2)]),
        "man_fan": make_state(p1_men=[(1, 2)], p2_men=[(2, 1), (2, 3), (4, 1), (4, 3), (4, 5)]),
        "promotion_ends_capture": make_state(p1_men=[(5, 2)], p2_men=[(6, 3), (6, 5)]),
        "direct_order": make_state(p1_kings=[(0, 1)], p1_men=[(6, 1)], p2_men=[(7, 6)]),
        "no_pieces_to_move": make_state(p1_men=[], p2_men=[(5, 0)], turn=0),
        "draw_counter": make_state(p1_men=[(0, 1)], p2_men=[(7, 0)], turn=0, msc=50),
        "p2_mirror_fan": make_state(p2_men=[(6, 5)], p1_men=[(5, 6), (5, 4), (3, 6), (3, 4), (3, 2)], turn=1),
    }


def main():
    ref = Checker("reference")
    os.makedirs(OUT, exist_ok=True)

    layout = np.array(ref.layout(), dtype=np.int32)
    perft = np.array([ref.perft(START_PACKED, d) for d in range(1, 11)], dtype=np.uint64)

    sets = {"leaves": ref.gen_leaves(N_LEAVES, key=KEY_LEAF), "synth": synthetic_positions(N_SYNTH, 7)}
    blob = {"layout": layout, "perft_start": perft}
    for name, st in sets.items():
        mv, cnt = ref.genmoves(st, MAX_MOVES)
        assert cnt.max() <= MAX_MOVES
        blob[name + "_states"] = st
        blob[name + "_counts"] = cnt
        blob[name + "_moves_flat"] = flatten(mv, cnt)
        for tag, mode, order in (("rc", MODE_RANDOM, ORDER_CANONICAL), ("rf", MODE_RANDOM, ORDER_FAST),
                                 ("h", MODE_HEURISTIC, ORDER_CANONICAL)):
            w, p, f, c = ref.playouts(st, reps=2, key=KEY_PLAY, pid_base=1000, mode=mode, order=order, want_final=True)
            blob["%s_%s_winners" % (name, tag)] = w
            blob["%s_%s_plies" % (name, tag)] = p.astype(np.uint16)
            blob["%s_%s_final" % (name, tag)] = f
            blob["%s_%s_counters" % (name, tag)] = c
        # truncated playouts (max_plies) exercise the "advance" path used for leaf generation
        w, p, f, c = ref.playouts(st, reps=1, key=KEY_PLAY, pid_base=0, mode=MODE_RANDOM, order=ORDER_CANONICAL,
                                  max_plies=5, want_final=True)
        blob[name + "_cut5_winners"] = w
        blob[name + "_cut5_final"] = f

    kats = kat_positions()
    names = sorted(kats)
    st = np.stack([kats[k] for k in names])
    mv, cnt = ref.genmoves(st, MAX_MOVES)
    blob["kat_names"] = np.array(names)
    blob["kat_states"] = st
    blob["kat_counts"] = cnt
    blob["kat_moves_flat"] = flatten(mv, cnt)
    w, _, _, _ = ref.playouts(st, key=KEY_PLAY)
    blob["kat_rc_winners"] = w

    # the reference's own 776-byte layout for a few states (pins b2p_pack776 / b2p_unpack776)
    blob["states776_sample"] = ref.unpack776(sets["leaves"][:64])

    # reference host drivers with their own RNG (glibc rand / default_random_engine): win-rate
    # tallies for the statistical parity test (draws, P1, P2) on the first 4096 D_ref leaves x 16
    big = ref.gen_leaves(65536, key=KEY_LEAF)
    for tag, mode in (("host", MODE_RANDOM), ("host_heuristic", MODE_HEURISTIC)):
        res = ref.host_driver(big, mode)
        blob["ref_%s_tally_65536" % tag] = np.array([(res == -1).sum(), (res == 0).sum(), (res == 1).sum()], dtype=np.int64)

    path = os.path.join(OUT, "reference_vectors.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("perft_start", "ref_host_tally_65536", "ref_host_heuristic_tally_65536", "kat_counts"):
        print(k, blob[k])


if __name__ == "__main__":
    main()
