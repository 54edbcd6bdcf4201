"""gpu_ai_b200 -- B200-native batched checkers playouts (the data-parallel hot path of krame505/gpu_ai).

The product is the C-ABI shared library ``gpu_ai_b200/libb2p.so`` (include/b2p.h): hand-written
sm_100a kernels plus the C++ host code that packs, shards and launches.  This package is only
its Python binding (ctypes) -- used by tests/, bench.py and anything that wants to drive the
engine from Python.  There is no CPU fallback: importing works without a GPU (so the ABI can be
inspected), but creating an Engine without a usable B200 raises.
"""
from .engine import (  # noqa: F401
    B2PError,
    Engine,
    MODE_HEURISTIC,
    MODE_RANDOM,
    ORDER_CANONICAL,
    ORDER_FAST,
    SCHED_AUTO,
    SCHED_THREAD,
    SCHED_WARP,
    Tree,
    lib_path,
    load_library,
)
from .players import B200MCTSPlayer, Player  # noqa: F401
from .drivers import (  # noqa: F401
    DeviceCoarsePlayoutDriver,
    DeviceHeuristicPlayoutDriver,
    DeviceMultiplePlayoutDriver,
    DeviceSinglePlayoutDriver,
    PlayoutDriver,
    getPlayoutDriver,
)

__all__ = [
    "B2PError", "Engine", "Tree", "load_library", "lib_path",
    "MODE_RANDOM", "MODE_HEURISTIC", "SCHED_THREAD", "SCHED_WARP", "SCHED_AUTO", "ORDER_CANONICAL", "ORDER_FAST",
    "PlayoutDriver", "DeviceSinglePlayoutDriver", "DeviceMultiplePlayoutDriver", "DeviceCoarsePlayoutDriver",
    "DeviceHeuristicPlayoutDriver", "getPlayoutDriver", "Player", "B200MCTSPlayer",
]
