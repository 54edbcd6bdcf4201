#!/usr/bin/env python3
"""tests/golden/tree_trace.npz: leaf-selection trace of the reference's GameTree (src/mcts.cpp) under
deterministic fake playout results, generated through oracle/_ref (build container only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Checker, START_PACKED  # noqa: E402
from test_tree import BATCHES, RefTree, drive  # noqa: E402

ref = Checker("reference")
t = RefTree(ref, START_PACKED)
trace = drive(t, BATCHES, salt=5)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tree_trace.npz"), root=START_PACKED, batches=np.array(BATCHES), salt=5,
                    flat=np.concatenate([x for x in trace if len(x)]), lengths=np.array([len(x) for x in trace]), total=t.total())
print("trace entries", len(trace), "leaves", sum(len(x) for x in trace), "total trials", t.total())
