"""CPU: the C-ABI library builds, loads and exports every symbol include/b2p.h declares; the
host-only converters work; and there is no CPU execution path (context creation fails loudly
without a GPU)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "gpu_ai_b200", "libb2p.so")):
        g.build()
    from gpu_ai_b200 import engine
    return engine.load_library()


def test_every_declared_symbol_is_exported(lib):
    from gpu_ai_b200 import engine
    hdr = open(os.path.join(ROOT, "include", "b2p.h")).read()
    declared = set(re.findall(r"\b(b2p_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    bound = {name for name, _, _ in engine.ABI}
    assert declared == bound, "header and binding disagree: %s" % sorted(declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)


def test_pack776_roundtrip_matches_reference_layout(lib, golden):
    from gpu_ai_b200 import engine
    s776 = golden["states776_sample"]
    packed = engine.pack776(s776)
    assert np.array_equal(packed, golden["leaves_states"][:64])
    assert np.array_equal(engine.unpack776(packed), s776)


def test_pack776_ignores_stale_fields_of_vacated_squares(lib, golden):
    # State::move clears only `occupied` (src/state.cu:78-84): type/owner of empty squares are garbage
    from gpu_ai_b200 import engine
    s = golden["states776_sample"].copy()
    for sq in range(64):
        off = 12 * sq
        empty = s[:, off] == 0
        s[empty, off + 4] = 1
        s[empty, off + 8] = 1
    assert np.array_equal(engine.pack776(s), golden["leaves_states"][:64])


def test_both_packers_give_the_same_bits(lib, port):
    """The reference-facing call packs 776-byte States with AVX-512BW + BMI2 where the CPU has them and with a
    portable scalar loop elsewhere (gpu_ai_b200/csrc/pack776.cpp): both against the oracle's converter on 100 000
    reachable positions with garbage in every field the reference leaves stale (type / owner of vacated squares,
    including PLAYER_NONE = -1 owners) and in the padding bytes of `bool occupied`."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import gpu_ai_b200 as b
from gpu_ai_b200 import engine
from oracle.pyoracle import Checker
port = Checker("port")
leaves = np.concatenate([port.gen_leaves(60000, key=5), np.load(%r)["synth_states"]])
s = port.unpack776(leaves).reshape(-1, 64 * 12 + 8)
board = s[:, :768].reshape(-1, 64, 12)
rng = np.random.default_rng(1)
occ = board[:, :, 0] != 0
for byte in (4, 8, 9, 10, 11):
    junk = rng.integers(0, 256, size=occ.shape).astype(np.uint8)
    board[:, :, byte] = np.where(occ, board[:, :, byte], junk if byte != 4 else junk & 1)
for byte in (1, 2, 3):
    board[:, :, byte] = rng.integers(0, 256, size=occ.shape).astype(np.uint8)
out = engine.pack776(s)
assert np.array_equal(out, leaves), "packer mismatch"
# arbitrary bytes (not valid States): both implementations are defined on the low byte of every field, so the
# digest of the packed output must not depend on which one ran
fuzz = np.random.default_rng(7).integers(0, 256, size=(20000, 776), dtype=np.uint8)
fuzz[:, ::3] &= np.random.default_rng(8).integers(0, 2, size=fuzz[:, ::3].shape, dtype=np.uint8) * 255
digest = int(engine.pack776(fuzz).astype(np.uint64).sum())
print(digest)
print(b.load_library().b2p_pack776_impl().decode())
""" % (ROOT, os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    seen, digests = set(), set()
    for force in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B2P_PACK_SCALAR=force), stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = r.stdout.strip().splitlines()
        seen.add(lines[-1])
        digests.add(lines[-2])
    assert "scalar" in seen and seen <= {"scalar", "avx512bw+bmi2"}
    assert len(digests) == 1


def test_expand_move_matches_reference_move_layout(lib, golden):
    # Move: from@0 to@2 removed@4.. intermediate@20.. jumps@36 promoted@37 (38 bytes)
    from gpu_ai_b200 import engine
    from oracle.pyoracle import decode_move, rc
    names = list(golden["kat_names"])
    i = names.index("king_cycle")
    pos = int(golden["kat_counts"][:i].sum())
    m = golden["kat_moves_flat"][pos]
    raw = engine.expand_move(m)
    d = decode_move(m)
    assert (raw[0], raw[1]) == rc(d["from"]) and (raw[2], raw[3]) == rc(d["to"])
    assert raw[36] == 3 and raw[37] == 0
    assert [(raw[20 + 2 * k], raw[21 + 2 * k]) for k in range(3)] == [(4, 3), (2, 5), (0, 3)]
    assert [(raw[4 + 2 * k], raw[5 + 2 * k]) for k in range(3)] == [(3, 2), (3, 4), (1, 4)]


def test_no_cpu_fallback(lib):
    """Without a GPU the engine must refuse to exist; with one this test is vacuous."""
    import torch
    from gpu_ai_b200 import B2PError, Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B2PError):
        Engine()


def test_driver_factory_names():
    import gpu_ai_b200
    with pytest.raises(RuntimeError, match="Unknown playout type"):
        gpu_ai_b200.getPlayoutDriver("nonsense")


def test_product_does_not_reference_oracle():
    """The shipped package must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "gpu_ai_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "checkers_oracle" not in txt, f
                assert not re.search(r'#include\s+"[^"]*oracle/', txt), f


def test_missing_extension_fails_loudly():
    """No silent fallback: if libb2p.so is not there, importing the engine and creating anything raises."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['B2P_LIB_PATH'] = '/nonexistent/libb2p.so'\n"
            "import gpu_ai_b200 as b\n"
            "try:\n    b.load_library()\nexcept b.B2PError as e:\n    print('RAISED', 'no CPU fallback' in str(e).lower() or 'missing' in str(e))\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "RAISED True" in r.stdout, r.stdout + r.stderr
