#!/usr/bin/env python3
"""Reference-side tallies for the statistical parity gate (check 3): the reference's OWN host drivers
(HostPlayoutDriver: glibc rand(); HostHeuristicPlayoutDriver: std::normal_distribution noise) play 2^20 D_ref leaves
once each; outcome tallies overall and per piece-count stratum of the leaf go to tests/golden/reference_tallies.npz.
Needs oracle/_ref (build container).  ~1 minute on 8 cores."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Checker, MODE_HEURISTIC, MODE_RANDOM  # noqa: E402

N = 1 << 20
KEY_LEAF = 2016
STRATA = np.array([0, 6, 10, 15, 25])  # total pieces on the leaf: [0,6) [6,10) [10,15) [15,25)


def stratum_of(leaves):
    pieces = np.unpackbits((leaves[:, 0] | leaves[:, 1]).astype(np.uint32).view(np.uint8)).reshape(len(leaves), 32).sum(axis=1)
    return np.digitize(pieces, STRATA[1:-1])


def tallies(res, strata):
    out = np.zeros((len(STRATA), 3), dtype=np.int64)   # row 0..3 = strata, last row = all
    for k, v in enumerate((-1, 0, 1)):
        out[-1, k] = (res == v).sum()
        for s in range(len(STRATA) - 1):
            out[s, k] = ((res == v) & (strata == s)).sum()
    return out


def main():
    ref = Checker("reference")
    leaves = ref.gen_leaves(N, key=KEY_LEAF)
    strata = stratum_of(leaves)
    blob = {"n": np.int64(N), "strata_edges": STRATA, "leaf_key": np.int64(KEY_LEAF),
            "leaves_checksum": np.uint64(leaves.astype(np.uint64).sum())}
    for tag, mode in (("host", MODE_RANDOM), ("host_heuristic", MODE_HEURISTIC)):
        res = ref.host_driver(leaves, mode)
        blob["ref_%s_tally" % tag] = tallies(res, strata)
        print(tag, blob["ref_%s_tally" % tag].tolist())
    path = os.path.join(ROOT, "tests", "golden", "reference_tallies.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
