#!/usr/bin/env python3
"""bench.py -- checkers playouts/s on N B200s (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, through the C ABI)
    python bench.py --impl reference [...]                     the reference's own CPU path, same metric

Workload (config.workload): "D_ref" = the reference's genRandomStates leaf distribution
(src/driver.cpp:76-104; made reproducible with Philox), 2^20 leaves PER GPU (weak
scaling; rank r owns leaves [r*2^20, (r+1)*2^20)), every leaf played `reps` times per step by the
random-playout kernel (device_single-equivalent, BASELINE configs[1]).  A step = one kernel launch
over the resident batch + one 32-byte NCCL all-reduce of the win counters (N > 1).

value  : playouts/s, whole job, leaves resident in HBM, CUDA-event time, max over ranks.
e2e    : playouts/s through the reference-facing call b2p_run_states776 (host buffer of 776-byte
         reference `State`s in, int32 PlayerId out; pack + H2D + kernel + D2H inside the timed region).
roofline: INT32 issue rate: achieved = plies/s x 180 thread-ops (frozen per-ply
         model) against the ALU-pipe rate measured live by b2p_microbench (LOP3).
cpu_baseline: the reference's HostPlayoutDriver (oracle/_ref, kind "reference") or the C
         restatement (kind "port") on this box's host cores, bounded sample; plus the 1-thread number, the
         thread-local-Philox port ("fair" CPU number) and the literal `run_ai -m playout_test ... host host`.
Extra keys of the default command (the other BASELINE configs, each measured in the same run):
  heuristic   : configs[2] -- heuristic kernel on the same leaves, roofline on W_ply,heur = 380, CPU = HostHeuristicPlayoutDriver
  d_start / d_live : configs[0]'s position (copies of the initial position) and D_ref without terminal leaves
  mcts_search : configs[3] -- b2p_tree_search_ex from the initial position on the N GPUs of the job (in-process sharding)
  shard_invariance / in_process_multi_device : results do not depend on how leaves are split over ranks / devices
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_PLY = 180.0          # frozen algorithmic INT32 thread-ops per ply (SURVEY.md 8d / BASELINE.md 4)
W_PLY_HEUR = 380.0     # the same for a heuristic ply
BYTES_PER_PLAYOUT = 17  # 16 B leaf in + 1 B winner out
LEAVES_PER_GPU = 1 << 20
LEAF_KEY = 2016
PLAY_KEY = 12345


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2p", choices=["b2p", "reference"])
    ap.add_argument("--reps", type=int, default=64, help="playouts per leaf per step")
    ap.add_argument("--leaves", type=int, default=LEAVES_PER_GPU, help="leaves per GPU")
    ap.add_argument("--mode", default="random", choices=["random", "heuristic"])
    ap.add_argument("--order", default="fast", choices=["fast", "canonical"])
    ap.add_argument("--leaf-set", default="ref", choices=["ref", "live", "start"],
                    help="D_ref = genRandomStates recipe (default, 36 %% terminal); D_live = D_ref without terminal leaves; "
                         "D_start = copies of the initial position (SURVEY.md 8d)")
    ap.add_argument("--ref-seconds", type=float, default=0.0,
                    help="--impl reference: CPU seconds per step (0 = automatic: bounded so that the whole run ends within minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the heuristic / d_start / d_live / mcts_search sections")
    ap.add_argument("--search-seconds", type=float, default=1.0, help="wall-clock budget of each mcts_search configuration")
    return ap.parse_args()


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.  NVML in a thread (10 ms period, no
    process start-up latency, so even a 100 ms region gets samples); falls back to the nvidia-smi loop of
    the profiling recipe when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.samples = []
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML device through its UUID/PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.gpu)
            handle = None
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(bus.uuid)).encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = (pynvml, handle)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        k = 0
        pw = 0.0
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = reasons(h)
                if k % 4 == 0:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((sm, mx, pw, rs))
                k += 1
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nvml[0]
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, b in bits.items() if any(s[3] & b for s in self.samples))
            sm = [s[0] for s in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(s[1] for s in self.samples) if sm else None,
                    "power_w_max": max(s[2] for s in self.samples) if sm else None, "samples": len(sm), "reasons": reasons,
                    "source": "nvml, 10 ms period, during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 100"}


# ---- CPU baseline ------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_omp_threads(n):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not silently inherit that.  Sets the
    OpenMP thread count of the already loaded libgomp and returns what the runtime then reports."""
    import ctypes
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return None


def cpu_playouts_per_s(leaves, mode, budget_s=12.0, threads=None, force_port=False):
    """Time the reference's own host playout driver (or the C port) on a bounded sample of `leaves`."""
    from oracle import pyoracle
    threads = threads or host_threads()
    if pyoracle.have_reference() and not force_port:
        chk, kind = pyoracle.Checker("reference"), "reference"
        run = lambda st: chk.host_driver(st, 1 if mode == "heuristic" else 0)  # noqa: E731
        what = "reference Host%sPlayoutDriver::runPlayouts (OpenMP, %s)" % (
            ("Heuristic", "std::normal_distribution noise") if mode == "heuristic" else ("", "glibc rand() behind its lock"))
    else:
        chk, kind = pyoracle.Checker("port"), "port"
        run = lambda st: chk.playouts(st, key=PLAY_KEY, mode=1 if mode == "heuristic" else 0)  # noqa: E731
        what = "oracle/checkers_oracle.c (OpenMP, thread-local Philox chooser)"
    actual = set_omp_threads(threads) or threads
    n = min(len(leaves), 2048)
    t0 = time.perf_counter()
    run(leaves[:n])
    dt = time.perf_counter() - t0
    rate = n / max(dt, 1e-6)
    n = int(min(len(leaves), max(2048, rate * budget_s)))
    t0 = time.perf_counter()
    run(leaves[:n])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "playouts/s", "cores": actual, "kind": kind,
            "sample": "%d leaves, 1 playout each, %s, %d OpenMP threads, %.1f s" % (n, what, actual, dt)}, n, dt


def run_ai_literal(threads, n=20000):
    """BASELINE configs[0] as the reference spells it: `run_ai -m playout_test -n N -1 host -2 host` (the reference
    binary built from /root/reference by oracle/Makefile; its own genRandomStates, its own timer)."""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "run_ai_ref")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    try:
        r = subprocess.run([exe, "-m", "playout_test", "-n", str(n), "-1", "host", "-2", "host"], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=120)
    except (OSError, subprocess.TimeoutExpired):
        return None
    secs = [float(x) for x in re.findall(r"Elapsed time: ([0-9.eE+-]+) seconds", r.stdout)]
    if not secs:
        return None
    return {"threads": threads, "n": n, "playouts_per_s": n / min(secs), "elapsed_s": secs}


def cpu_baselines(sets, budget_s):
    """reference host driver at nproc and at 1 thread, the thread-local-Philox port (the
    "fair" CPU number: no rand() lock), on D_ref; the reference driver on D_live and D_start; host_heuristic; the
    literal run_ai command.  `sets` = {"ref": packed leaves, "live": ..., "start": ...}."""
    nt = host_threads()
    base, _, _ = cpu_playouts_per_s(sets["ref"], "random", budget_s=budget_s, threads=nt)
    extra = {}
    one, _, _ = cpu_playouts_per_s(sets["ref"], "random", budget_s=budget_s / 4, threads=1)
    extra["reference_1_thread"] = {"value": one["value"], "cores": one["cores"], "sample": one["sample"]}
    fair, _, _ = cpu_playouts_per_s(sets["ref"], "random", budget_s=budget_s / 4, threads=nt, force_port=True)
    extra["port_thread_local_philox"] = {"value": fair["value"], "cores": fair["cores"], "kind": "port", "sample": fair["sample"]}
    for name in ("live", "start"):
        if name in sets:
            r, _, _ = cpu_playouts_per_s(sets[name], "random", budget_s=budget_s / 4, threads=nt)
            extra["reference_D_" + name] = {"value": r["value"], "cores": r["cores"], "sample": r["sample"]}
    h, _, _ = cpu_playouts_per_s(sets["ref"], "heuristic", budget_s=budget_s / 3, threads=nt)
    extra["reference_host_heuristic"] = {"value": h["value"], "cores": h["cores"], "sample": h["sample"]}
    lit = [x for x in (run_ai_literal(1), run_ai_literal(nt)) if x]
    if lit:
        extra["run_ai_playout_test_host_host"] = lit
    set_omp_threads(nt)
    base["more"] = extra
    return base


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores (all of them:
    the OpenMP thread count is set explicitly and reported -- torchrun would otherwise pin it to 1)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    threads = set_omp_threads(host_threads()) or host_threads()
    chk = pyoracle.Checker("port")
    leaves = chk.gen_leaves(1 << 17, key=LEAF_KEY)   # bounded sample of the same D_ref stream (leaves 0 .. 131071)
    per_step_budget = args.ref_seconds if args.ref_seconds > 0 else max(1.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    base, n, _ = cpu_playouts_per_s(leaves, args.mode, budget_s=per_step_budget, threads=threads)
    if pyoracle.have_reference():
        c = pyoracle.Checker("reference")
        run = lambda: c.host_driver(leaves[:n], 1 if args.mode == "heuristic" else 0)  # noqa: E731
    else:
        run = lambda: chk.playouts(leaves[:n], key=PLAY_KEY, mode=1 if args.mode == "heuristic" else 0)  # noqa: E731
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    base.update(value=value, cores=threads,
                sample="first %d leaves of the D_ref stream per step x %d steps, %d OpenMP threads, %.1f s" % (n, args.steps, threads, dt))
    print(json.dumps({
        "impl": "reference", "metric": "checkers_playouts_per_sec", "value": value, "unit": "playouts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "D_ref leaves (seed 2016), %s playouts to the end; the CPU arm plays a bounded SAMPLE of the same leaf "
                               "stream per step (a rate on the same distribution, not the same number of leaves)" % args.mode,
                   "leaves_per_step": n, "playouts_per_step": n, "omp_threads": threads, "host_cpus": host_threads()},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "playouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- our arm --------------------------------------------------------------------------------------------
def time_kernel(torch, eng, d_states, n, reps, mode, order, steps, warmup, flush, d_winners, d_counters, stream, key0, pid_base=0):
    """`steps` timed launches over the resident batch (CUDA events on the launching stream, L2 flushed between
    launches, outside the per-launch events).  Returns mean ms per launch and the last launch's counters."""
    for i in range(warmup):
        d_counters.zero_()
        eng.run_packed_device(d_states.data_ptr(), n, reps=reps, key=key0 + i, pid_base=pid_base, mode=mode, order=order,
                              d_winners=d_winners.data_ptr(), d_counters=d_counters.data_ptr(), stream=stream)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()
        d_counters.zero_()
        ev[i][0].record()
        eng.run_packed_device(d_states.data_ptr(), n, reps=reps, key=key0 + warmup + i, pid_base=pid_base, mode=mode, order=order,
                              d_winners=d_winners.data_ptr(), d_counters=d_counters.data_ptr(), stream=stream)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    c = d_counters.cpu().numpy().astype(np.int64)
    return ms, c


def side_workload(torch, eng, d_states, n, reps, mode, order, w_ply, peak_alu, flush, d_winners, d_counters, stream, steps, key0):
    ms, c = time_kernel(torch, eng, d_states, n, reps, mode, order, steps, 2, flush, d_winners, d_counters, stream, key0)
    played = int(c[:3].sum())
    ppp = float(c[3]) / max(1, played)
    rate = n * reps / (ms * 1e-3)
    achieved = ppp * rate * w_ply
    return {"value": rate, "unit": "playouts/s (this rank's GPU, leaves resident, CUDA events)", "kernel_ms": ms, "leaves": n, "reps": reps,
            "launches_timed": steps, "plies_per_playout": ppp,
            "win_counts": {"draws": int(c[0]), "p1": int(c[1]), "p2": int(c[2])},
            "roofline": {"bound": "int32_issue", "achieved": achieved / 1e12, "peak": peak_alu / 1e12, "unit": "T thread-op/s",
                         "frac": achieved / peak_alu, "model": "W_ply = %d INT32 thread-ops/ply x %.2f plies/playout" % (w_ply, ppp)}}


def winners_checksum(torch, d_winners, n, first_index):
    """order-sensitive checksum of int8 winners of leaves [first_index, first_index + n)"""
    idx = torch.arange(first_index, first_index + n, device=d_winners.device, dtype=torch.int64)
    h = (idx * 2654435761 + 12345) % 1000003
    return int(((d_winners[:n].to(torch.int64) + 2) * h).sum().item())


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import gpu_ai_b200 as b
    from gpu_ai_b200 import sharding
    from gpu_ai_b200.engine import PinnedArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side barriers: an NCCL barrier would spin on the idle GPUs

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)

    eng = b.Engine(devices=[local], seed=PLAY_KEY)
    info = eng.device_info(0)
    n, reps = args.leaves, args.reps
    mode = b.MODE_HEURISTIC if args.mode == "heuristic" else b.MODE_RANDOM
    order = b.ORDER_CANONICAL if (args.order == "canonical" or mode == b.MODE_HEURISTIC) else b.ORDER_FAST
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic input, generated on the device by the leaf kernel (bit-exact vs oracle: tests) ----
    def make_leaf_set(which, count, first):
        t = torch.empty((count, 4), dtype=torch.int32, device=dev)
        if which == "start":
            t.copy_(torch.tensor([0x00000FFF, 0xFFF00000 - (1 << 32), 0, 0], dtype=torch.int32, device=dev).expand(count, 4))
        elif which == "live":
            # D_live: draw D_ref leaves until `count` non-terminal ones are found (terminal = no legal move or msc >= 50)
            got, nxt, chunks = 0, 2 * first, []
            while got < count:
                cand = eng.gen_leaves(count, key=LEAF_KEY, first_index=nxt)
                _, cnt = eng.genmoves(cand, 1)
                live = cand[(cnt > 0) & ((cand[:, 3] >> 8) < 50)]
                chunks.append(live)
                got += len(live)
                nxt += count
            t.copy_(torch.from_numpy(np.concatenate(chunks)[:count].view(np.int32)).to(dev))
        else:
            eng.gen_leaves_device(count, t.data_ptr(), key=LEAF_KEY, first_index=first, stream=stream)
        return t

    leaf_lo, _ = sharding.weak_shard(n, rank)
    d_states = make_leaf_set(args.leaf_set, n, leaf_lo)
    d_winners = torch.empty(n * reps, dtype=torch.int8, device=dev)
    d_counters = torch.zeros(4, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    torch.cuda.synchronize()

    def step(i):
        d_counters.zero_()
        eng.run_packed_device(d_states.data_ptr(), n, reps=reps, key=PLAY_KEY + i, pid_base=rank * n * reps, mode=mode,
                              order=order, d_winners=d_winners.data_ptr(), d_counters=d_counters.data_ptr(), stream=stream)
        sharding.allreduce_counters(d_counters)   # the single small NCCL all-reduce per iteration (32 bytes)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()

    # ---- integer-pipe peak, measured on this GPU right now ------------------------------------------
    peak_alu, _ = eng.microbench(0, iters=4000)       # LOP3 thread-ops/s
    peak_mix, _ = eng.microbench(5, iters=4000)       # LOP3 + IMAD dual-pipe

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        flush.zero_()                                  # L2 flush between timed iterations
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = t0.elapsed_time(t1)
    kernel_ms = [a.elapsed_time(bb) for a, bb in ev]
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    counters = d_counters.cpu().numpy().astype(np.int64)   # last step, summed over ranks if world > 1
    plies_per_playout = float(counters[3]) / float(max(1, counters[:3].sum()))

    elapsed_ms = sharding.max_over_ranks(elapsed_ms, dev)
    playouts_per_step = n * reps * world
    value = playouts_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank's GPU) ---------------------------------------------
    w_ply = W_PLY_HEUR if mode == b.MODE_HEURISTIC else W_PLY
    k_ms = float(np.mean(kernel_ms))
    plies_per_launch = plies_per_playout * n * reps
    achieved = plies_per_launch * w_ply / (k_ms * 1e-3)
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this very
    # command (profiles/traffic.json, written by tools/extract_traffic.py); null when the configuration differs
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if (tj["leaves"], tj["reps"], tj["mode"], tj["order"]) == (n, reps, args.mode, args.order):
            traffic = tj["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "bound": "int32_issue", "achieved": achieved / 1e12, "peak": peak_alu / 1e12, "unit": "T thread-op/s",
        "frac": achieved / peak_alu, "traffic": traffic,
        "peak_source": "b2p_microbench LOP3 (ALU pipe) measured live on this GPU; LOP3+IMAD dual-pipe %.2f T/s" % (peak_mix / 1e12),
        "model": "W_ply = %d INT32 thread-ops/ply (SURVEY.md 8d) x %.2f plies/playout counted by the kernel" % (w_ply, plies_per_playout),
        "plies_per_s": plies_per_launch / (k_ms * 1e-3), "kernel_ms": k_ms,
        "hbm": {"achieved_GBs": BYTES_PER_PLAYOUT * n * reps / (k_ms * 1e-3) / 1e9, "note": "not the bound (< 1% of HBM peak)"},
    }

    out = {
        "metric": "checkers_playouts_per_sec", "value": value, "unit": "playouts/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "D_%s: %s, %s playouts to the end, device_single-equivalent (thread-per-playout persistent lanes)"
                               % (args.leaf_set, {"ref": "2^20-class random reachable leaves per GPU (reference genRandomStates recipe, seed 2016)",
                                                  "live": "D_ref leaves with the terminal ones rejected", "start": "copies of the initial position"}[args.leaf_set],
                                  args.mode),
                   "leaves_per_gpu": n, "leaves_per_step": n * world, "reps_per_step": reps, "playouts_per_step": playouts_per_step,
                   "move_order": args.order,
                   "plies_per_playout": plies_per_playout, "parallelism": "leaf-sharded x%d, one 32-byte all-reduce per step" % world,
                   "l2": "256 MiB memset between timed steps (inside the timed region)", "gpu": info["name"], "sms": info["sm_count"]},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "win_counts_last_step": {"draws": int(counters[0]), "p1": int(counters[1]), "p2": int(counters[2]), "plies": int(counters[3])},
    }

    # ---- shard invariance: this rank's winners for its GLOBAL leaf range == the same leaves inside one big launch ------
    # every rank plays leaves [rank*n, (rank+1)*n) once with global playout ids; rank 0 then plays ALL world*n leaves of
    # the stream in a single launch on its own GPU (the N = 1 way of doing the job) and compares range by range
    m = min(n, 1 << 18)
    inv_states = make_leaf_set("ref", m, rank * m)
    inv_w = torch.empty(m, dtype=torch.int8, device=dev)
    eng.run_packed_device(inv_states.data_ptr(), m, reps=1, key=4242, pid_base=rank * m, mode=b.MODE_RANDOM, order=b.ORDER_FAST,
                          d_winners=inv_w.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    mine = winners_checksum(torch, inv_w, m, rank * m)
    sums = torch.tensor([mine], dtype=torch.int64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)
        per_rank = [int(g.item()) for g in gathered]
    else:
        per_rank = [mine]
    if rank == 0:
        whole = make_leaf_set("ref", m * world, 0)
        whole_w = torch.empty(m * world, dtype=torch.int8, device=dev)
        eng.run_packed_device(whole.data_ptr(), m * world, reps=1, key=4242, pid_base=0, mode=b.MODE_RANDOM, order=b.ORDER_FAST,
                              d_winners=whole_w.data_ptr(), stream=stream)
        # ... and once more as two unequal launches (a different split of the same ids)
        cut = (m * world) // 3
        split_w = torch.empty(m * world, dtype=torch.int8, device=dev)
        eng.run_packed_device(whole.data_ptr(), cut, reps=1, key=4242, pid_base=0, mode=b.MODE_RANDOM, order=b.ORDER_FAST,
                              d_winners=split_w.data_ptr(), stream=stream)
        eng.run_packed_device(whole[cut:].data_ptr(), m * world - cut, reps=1, key=4242, pid_base=cut, mode=b.MODE_RANDOM,
                              order=b.ORDER_FAST, d_winners=split_w[cut:].data_ptr(), stream=stream)
        torch.cuda.synchronize()
        single = [winners_checksum(torch, whole_w[r * m:(r + 1) * m], m, r * m) for r in range(world)]
        out["shard_invariance"] = bool(single == per_rank and torch.equal(whole_w, split_w))
        out["shard_invariance_detail"] = {"leaves_per_rank": m, "ranks": world, "checksums_by_rank": per_rank,
                                          "checksums_single_launch": single, "two_unequal_launches_equal": bool(torch.equal(whole_w, split_w))}
        del whole, whole_w, split_w
    del inv_states, inv_w

    # ---- single-pass latency figure: 1M playouts, reps = 1 (BASELINE configs[1] read literally) --------------
    if rank == 0:
        torch.cuda.synchronize()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms1 = []
        for i in range(5):
            d_counters.zero_()
            a.record()
            eng.run_packed_device(d_states.data_ptr(), n, reps=1, key=99 + i, mode=mode, order=order,
                                  d_winners=d_winners.data_ptr(), d_counters=d_counters.data_ptr(), stream=stream)
            bb.record()
            torch.cuda.synchronize()
            ms1.append(a.elapsed_time(bb))
        out["single_pass"] = {"playouts": n, "ms": float(np.median(ms1)), "playouts_per_s": n / (float(np.median(ms1)) * 1e-3)}

    # ---- the same workload with the strictly canonical move order (rank j -> j-th move of State::getMoves()) ------
    if rank == 0 and mode == b.MODE_RANDOM and order == b.ORDER_FAST:
        ms, _ = time_kernel(torch, eng, d_states, n, reps, mode, b.ORDER_CANONICAL, 3, 2, flush, d_winners, d_counters, stream, 7)
        out["canonical_order"] = {"playouts_per_s_per_gpu": n * reps / (ms * 1e-3),
                                  "note": "B2P_ORDER_CANONICAL on this rank's GPU; the headline uses B2P_ORDER_FAST (same uniform law, "
                                          "both bit-exact against the oracle)"}

    # ---- the other BASELINE configurations, on rank 0's GPU while the other ranks wait on the host ----------------
    sets_for_cpu = {}
    if rank == 0 and not args.no_extras and mode == b.MODE_RANDOM and args.leaf_set == "ref":
        k = max(3, min(args.steps, 6))
        hreps = max(1, reps // 4)
        out["heuristic"] = side_workload(torch, eng, d_states, n, hreps, b.MODE_HEURISTIC, b.ORDER_CANONICAL, W_PLY_HEUR, peak_alu,
                                         flush, d_winners, d_counters, stream, k, 501)
        out["heuristic"]["config"] = "BASELINE configs[2]: device_heuristic-equivalent, the same 2^20 D_ref leaves, %d playouts per leaf per launch" % hreps
        d_start = make_leaf_set("start", n, 0)
        out["d_start"] = side_workload(torch, eng, d_start, n, hreps, b.MODE_RANDOM, b.ORDER_FAST, W_PLY, peak_alu, flush, d_winners,
                                       d_counters, stream, k, 601)
        out["d_start"]["config"] = "BASELINE configs[0]'s position: random playouts from 2^20 copies of the initial position"
        sets_for_cpu["start"] = d_start[:4096].cpu().numpy().view(np.uint32)
        del d_start
        d_live = make_leaf_set("live", n, 0)
        out["d_live"] = side_workload(torch, eng, d_live, n, hreps, b.MODE_RANDOM, b.ORDER_FAST, W_PLY, peak_alu, flush, d_winners,
                                      d_counters, stream, k, 701)
        out["d_live"]["config"] = "D_ref with terminal leaves rejected (what an MCTS tree actually sends)"
        sets_for_cpu["live"] = d_live[:1 << 15].cpu().numpy().view(np.uint32)
        del d_live
    host_barrier()

    # ---- e2e: the reference-facing call, host buffers in and out ------------------------------------------------
    if not args.no_e2e:
        from gpu_ai_b200 import engine as eng_mod
        packed = d_states.cpu().numpy().view(np.uint32)
        s776 = eng_mod.unpack776(packed)            # n reference `State` objects (776 B each) in host memory
        res = np.empty(n, dtype=np.int32)
        mode_i = mode
        for _ in range(2):
            eng.run_states776(s776, mode=mode_i, out=res)
        host_barrier()
        k = max(3, min(args.steps, 10))
        t_start = time.perf_counter()
        for _ in range(k):
            eng.run_states776(s776, mode=mode_i, out=res)
        dt = time.perf_counter() - t_start
        dt = sharding.max_over_ranks(dt, dev)
        out["e2e"] = {"value": n * world * k / dt, "unit": "playouts/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": n,
                      "ms_per_call": 1e3 * dt / k, "host_input_bytes_per_step": 776 * n,
                      "api": "b2p_run_states776(host State[776 B] x n) -> int32 PlayerId[n]; pack + H2D + kernel + D2H timed"}
        del s776
        # the same call on 16-byte packed leaves in page-locked caller memory (b2p_alloc_host): what a caller that keeps
        # leaves packed pays -- 16 B up, 1 B down per playout
        pin_s, pin_w = PinnedArray((n, 4), np.uint32), PinnedArray((n,), np.int8)
        pin_s.array[:] = packed
        for _ in range(2):
            eng.run_packed(pin_s.array, reps=1, key=1, mode=mode_i, order=order, winners_out=pin_w.array)
        host_barrier()
        t_start = time.perf_counter()
        for i in range(k):
            eng.run_packed(pin_s.array, reps=1, key=2 + i, mode=mode_i, order=order, winners_out=pin_w.array)
        dtp = sharding.max_over_ranks(time.perf_counter() - t_start, dev)
        out["e2e_packed"] = {"value": n * world * k / dtp, "unit": "playouts/s", "ms_per_call": 1e3 * dtp / k,
                             "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": n,
                             "api": "b2p_run_packed(b2p_state16 x n in b2p_alloc_host memory) -> int8 winners + counters"}
        del pin_s, pin_w
    host_barrier()

    # ---- configs[3]: full MCTS move search from the initial position, leaf batches sharded over the job's GPUs -----
    # rank 0 drives ONE tree over an in-process context that owns all `world` GPUs (what the drop-in shim does with
    # B2P_DEVICES); the other ranks idle on a host barrier meanwhile
    if rank == 0 and not args.no_extras:
        start = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)
        ndev = min(world, torch.cuda.device_count())
        seng = eng if ndev == 1 else b.Engine(devices=list(range(ndev)), seed=PLAY_KEY)
        rows = []
        # allocation policy 0 = the reference's rule (GameTree::select), 1 = B2P_POLICY_UCT (what the B200 player uses)
        for batch, mx, sreps, depth, policy in ((65536, 1 << 20, 32, 2, 0), (65536, 1 << 20, 8, 2, 0), (65536, 1 << 20, 256, 2, 0),
                                                (65536, 1 << 20, 32, 1, 0), (4096, 4096, 32, 2, 0), (65536, 1 << 20, 32, 2, 1),
                                                (2048, 2048, 16, 2, 1)):
            t = b.Tree(start)
            t.search_ex(seng, iterations=3, initial_batch=batch, max_batch=mx, reps=sreps, key=1, depth=depth, policy=policy)   # warm-up
            t = b.Tree(start)
            st = t.search_ex(seng, seconds=args.search_seconds, initial_batch=batch, scale=0.02 if mx > batch else 0.0, max_batch=mx,
                             reps=sreps, key=3, depth=depth, policy=policy)
            rows.append({"initial_batch": batch, "max_batch": mx, "reps_per_leaf": sreps, "depth": st["depth"], "host_threads": st["threads"],
                         "policy": "uct" if policy else "reference",
                         "seconds": st["seconds"], "playouts_per_s": st["playouts"] / st["seconds"],
                         "leaf_selections_per_s": st["leaves"] / st["seconds"], "batches": st["batches"], "tree_nodes": st["nodes"],
                         "gpu_busy": st["kernel_s"] / st["seconds"], "select_s": st["select_s"], "update_s": st["update_s"],
                         "wait_s": st["wait_s"]})
            del t
        best = max(r["playouts_per_s"] for r in rows if r["reps_per_leaf"] <= 32)
        out["mcts_search"] = {"api": "b2p_tree_search_ex from the initial position, random playouts, one host tree, leaf batches sharded "
                                     "in-process over %d GPU(s)" % ndev,
                              "devices": ndev, "playouts_per_s_at_le_32_reps": best, "configs": rows}
        if ndev > 1:
            # the in-process multi-device path gives the single-device answers (playout ids are global)
            st_h = eng.gen_leaves(70001, key=12)
            a1 = eng.run_packed(st_h, reps=3, key=6, pid_base=40, order=b.ORDER_FAST, want_plies=True)
            a2 = seng.run_packed(st_h, reps=3, key=6, pid_base=40, order=b.ORDER_FAST, want_plies=True)
            w1, c1 = eng.run_counts(st_h, reps=5, key=7)
            w2, c2 = seng.run_counts(st_h, reps=5, key=7)
            e1, e2 = b.Engine(devices=[local], seed=99), b.Engine(devices=list(range(ndev)), seed=99)
            s7 = b.engine.unpack776(st_h)
            r1, r2 = e1.run_states776(s7), e2.run_states776(s7)
            out["in_process_multi_device"] = {
                "devices": ndev,
                "equal": bool(np.array_equal(a1[0], a2[0]) and np.array_equal(a1[1], a2[1]) and np.array_equal(a1[3], a2[3])
                              and np.array_equal(w1, w2) and np.array_equal(c1, c2) and np.array_equal(r1, r2)),
                "checked": "b2p_run_packed (winners, plies, counters), b2p_run_counts, b2p_run_states776 on 70001 leaves: "
                           "1 device vs %d devices in one context" % ndev}
    host_barrier()

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) --------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sets_for_cpu["ref"] = d_states.cpu().numpy().view(np.uint32)
        if args.no_extras or mode != b.MODE_RANDOM or args.leaf_set != "ref":
            base, _, _ = cpu_playouts_per_s(sets_for_cpu["ref"], args.mode)
        else:
            base = cpu_baselines(sets_for_cpu, budget_s=10.0)
            if "heuristic" in out:
                out["heuristic"]["cpu_baseline"] = dict(base["more"]["reference_host_heuristic"], kind=base["kind"], unit="playouts/s")
            if "d_start" in out and "reference_D_start" in base["more"]:
                out["d_start"]["cpu_baseline"] = dict(base["more"]["reference_D_start"], kind=base["kind"], unit="playouts/s")
        out["cpu_baseline"] = base

    if rank == 0:
        print(json.dumps(out))
    host_barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
