"""Python mirror of the reference's playout-driver interface, for the device drivers only.

reference: class PlayoutDriver { runPlayouts(std::vector<State>) -> std::vector<PlayerId>; getName(); }
(src/playout.hpp:27-33) and the four device drivers (src/playout.hpp:56-92), constructed by name
through getPlayoutDriver (src/playout.cpp:189-223).  Same names, same argument meaning
(`states` = n reference `State` objects, here an (n, 776) uint8 array or any buffer of n*776
bytes), same result meaning (`result[i]` = PlayerId winner of a playout from `states[i]`:
0, 1, or -1 for a draw), same error behaviour for unknown names (runtime_error ->
RuntimeError("Unknown playout type")).  The C++ drop-in for the reference binary itself is
shim/playout_shim.cpp; this mirror exists so that the parity tests read like the reference's.

All four names run the host-rule semantics (the reference's single/coarse kernels
deviate from its host rules; we do not reproduce the drift).  They differ in scheduling only.
"""
import numpy as np

from . import engine as _e

_shared_engine = None


def _engine():
    global _shared_engine
    if _shared_engine is None:
        _shared_engine = _e.Engine()
    return _shared_engine


class PlayoutDriver:
    """Abstract base, as in src/playout.hpp:27-33."""

    def runPlayouts(self, states):
        raise NotImplementedError

    def getName(self):
        raise NotImplementedError


class _DeviceDriver(PlayoutDriver):
    _name = None
    _mode = _e.MODE_RANDOM
    _sched = _e.SCHED_THREAD

    def __init__(self, engine=None):
        self.engine = engine or _engine()

    def runPlayouts(self, states):
        buf = np.ascontiguousarray(states, dtype=np.uint8).reshape(-1)
        if buf.size == 0:
            return np.empty(0, dtype=np.int32)  # empty in, empty out (src/singlePlayout.cu:73-75)
        return self.engine.run_states776(buf, mode=self._mode, sched=self._sched)

    def getName(self):
        return self._name


class DeviceSinglePlayoutDriver(_DeviceDriver):
    """replaces src/singlePlayout.cu: one lane per playout"""
    _name = "device_single"


class DeviceCoarsePlayoutDriver(_DeviceDriver):
    """replaces src/coarsePlayout.cu: persistent lanes + work queue (what SCHED_THREAD already is)"""
    _name = "device_coarse"


class DeviceMultiplePlayoutDriver(_DeviceDriver):
    """replaces src/multiplePlayout.cu: one warp per playout"""
    _name = "device_multiple"
    _sched = _e.SCHED_AUTO


class DeviceHeuristicPlayoutDriver(_DeviceDriver):
    """replaces src/heuristicPlayout.cu: heuristic-guided playouts"""
    _name = "device_heuristic"
    _mode = _e.MODE_HEURISTIC


def getPlayoutDriver(name):
    """src/playout.cpp:189-223, device names only (the host/hybrid/optimal drivers are reference code
    that stays as it is and is out of scope here)."""
    table = {
        "device_single": DeviceSinglePlayoutDriver,
        "device_multiple": DeviceMultiplePlayoutDriver,
        "device_coarse": DeviceCoarsePlayoutDriver,
        "device_heuristic": DeviceHeuristicPlayoutDriver,
    }
    if name not in table:
        raise RuntimeError("Unknown playout type")
    return table[name]()
