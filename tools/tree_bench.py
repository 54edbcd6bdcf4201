#!/usr/bin/env python3
"""Host throughput of the tree operations: b2p_tree vs the reference's GameTree (when oracle/_ref is present).
select + update with free (fake) playout results, so only the caller-side cost is measured (SURVEY.md 6:
the reference sustains ~3e5 leaves/s)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_ai_b200 as b  # noqa: E402
from oracle import pyoracle  # noqa: E402
from test_tree import RefTree, fake_winners  # noqa: E402


def run(tree, batches):
    t_sel = t_upd = 0.0
    n_leaves = 0
    for i, n in enumerate(batches):
        t0 = time.perf_counter()
        leaves = tree.select(n)
        t1 = time.perf_counter()
        w = fake_winners(leaves, i)
        t2 = time.perf_counter()
        tree.update(w)
        t3 = time.perf_counter()
        t_sel += t1 - t0
        t_upd += t3 - t2
        n_leaves += len(leaves)
    return {"leaves": n_leaves, "select_leaves_per_s": n_leaves / t_sel, "update_leaves_per_s": n_leaves / t_upd,
            "select_plus_update_leaves_per_s": n_leaves / (t_sel + t_upd)}


batches = [50] + [4000] * 60
out = {"b2p_tree": run(b.Tree(pyoracle.START_PACKED), batches)}
if pyoracle.have_reference():
    out["reference_GameTree"] = run(RefTree(pyoracle.Checker("reference"), pyoracle.START_PACKED), batches)
    out["cores"] = os.cpu_count()
print(json.dumps(out, indent=1))
