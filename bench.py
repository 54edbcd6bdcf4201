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
         restatement (kind "port") on this box's host cores, bounded sample.
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
BYTES_PER_PLAYOUT = 17  # 16 B leaf in + 1 B winner out
LEAVES_PER_GPU = 1 << 20
LEAF_KEY = 2016
PLAY_KEY = 12345


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2p", choices=["b2p", "reference"])
    ap.add_argument("--reps", type=int, default=32, help="playouts per leaf per step")
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
def cpu_playouts_per_s(leaves, mode, budget_s=12.0):
    """Time the reference's own host playout driver (or the C port) on a bounded sample of `leaves`."""
    from oracle import pyoracle
    threads = os.cpu_count() or 1
    if pyoracle.have_reference():
        chk, kind = pyoracle.Checker("reference"), "reference"
        run = lambda st: chk.host_driver(st, 1 if mode == "heuristic" else 0)  # noqa: E731
        what = "reference Host%sPlayoutDriver::runPlayouts (OpenMP, glibc rand)" % ("Heuristic" if mode == "heuristic" else "")
    else:
        chk, kind = pyoracle.Checker("port"), "port"
        run = lambda st: chk.playouts(st, key=PLAY_KEY, mode=1 if mode == "heuristic" else 0)  # noqa: E731
        what = "oracle/checkers_oracle.c (OpenMP, Philox chooser)"
    n = 4096
    t0 = time.perf_counter()
    run(leaves[:n])
    dt = time.perf_counter() - t0
    rate = n / max(dt, 1e-6)
    n = int(min(len(leaves), max(4096, rate * budget_s)))
    t0 = time.perf_counter()
    run(leaves[:n])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "playouts/s", "cores": threads, "kind": kind,
            "sample": "%d D_ref leaves, 1 playout each, %s, %.1f s" % (n, what, dt)}, n, dt


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    chk = pyoracle.Checker("port")
    leaves = chk.gen_leaves(1 << 17, key=LEAF_KEY)   # bounded sample of the same D_ref workload
    per_step_budget = args.ref_seconds if args.ref_seconds > 0 else max(1.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    base, n, _ = cpu_playouts_per_s(leaves, args.mode, budget_s=per_step_budget)
    from oracle import pyoracle as po
    if po.have_reference():
        c = po.Checker("reference")
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
    base.update(value=value, sample="%d D_ref leaves per step x %d steps, %.1f s" % (n, args.steps, dt))
    print(json.dumps({
        "impl": "reference", "metric": "checkers_playouts_per_sec", "value": value, "unit": "playouts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "D_ref leaves (seed 2016), %s playouts, bounded sample of %d leaves per step on host cores" % (args.mode, n)},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "playouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- our arm --------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import gpu_ai_b200 as b
    from gpu_ai_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    eng = b.Engine(devices=[local], seed=PLAY_KEY)
    info = eng.device_info(0)
    n, reps = args.leaves, args.reps
    mode = b.MODE_HEURISTIC if args.mode == "heuristic" else b.MODE_RANDOM
    order = b.ORDER_CANONICAL if (args.order == "canonical" or mode == b.MODE_HEURISTIC) else b.ORDER_FAST
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic input, generated on the device by the leaf kernel (bit-exact vs oracle: tests) ----
    d_states = torch.empty((n, 4), dtype=torch.int32, device=dev)
    leaf_lo, _ = sharding.weak_shard(n, rank)
    if args.leaf_set == "start":
        d_states.copy_(torch.tensor([0x00000FFF, 0xFFF00000 - (1 << 32), 0, 0], dtype=torch.int32, device=dev).expand(n, 4))
    elif args.leaf_set == "live":
        # D_live: draw D_ref leaves until n non-terminal ones are found (terminal = no legal move or msc >= 50)
        got, first, chunks = 0, 2 * leaf_lo, []
        while got < n:
            cand = eng.gen_leaves(n, key=LEAF_KEY, first_index=first)
            _, cnt = eng.genmoves(cand, 1)
            live = cand[(cnt > 0) & ((cand[:, 3] >> 8) < 50)]
            chunks.append(live)
            got += len(live)
            first += n
        d_states.copy_(torch.from_numpy(np.concatenate(chunks)[:n].view(np.int32)).to(dev))
    else:
        eng.gen_leaves_device(n, d_states.data_ptr(), key=LEAF_KEY, first_index=leaf_lo, stream=stream)
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
    kernel_ms = []
    plies_total = 0
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
    k_ms = float(np.mean(kernel_ms))
    plies_per_launch = plies_per_playout * n * reps
    achieved = plies_per_launch * W_PLY / (k_ms * 1e-3)
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
        "model": "W_ply = 180 INT32 thread-ops/ply (SURVEY.md 8d) x %.2f plies/playout counted by the kernel" % plies_per_playout,
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
                   "leaves_per_gpu": n, "reps_per_step": reps, "playouts_per_step": playouts_per_step, "move_order": args.order,
                   "plies_per_playout": plies_per_playout, "parallelism": "leaf-sharded x%d, one 32-byte all-reduce per step" % world,
                   "l2": "256 MiB memset between timed steps (inside the timed region)", "gpu": info["name"], "sms": info["sm_count"]},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "win_counts_last_step": {"draws": int(counters[0]), "p1": int(counters[1]), "p2": int(counters[2]), "plies": int(counters[3])},
    }

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
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(2):
            eng.run_packed_device(d_states.data_ptr(), n, reps=reps, key=7 + i, mode=mode, order=b.ORDER_CANONICAL,
                                  d_winners=d_winners.data_ptr(), d_counters=d_counters.data_ptr(), stream=stream)
        a.record()
        for i in range(3):
            eng.run_packed_device(d_states.data_ptr(), n, reps=reps, key=17 + i, mode=mode, order=b.ORDER_CANONICAL,
                                  d_winners=d_winners.data_ptr(), d_counters=d_counters.data_ptr(), stream=stream)
        bb.record()
        torch.cuda.synchronize()
        out["canonical_order"] = {"playouts_per_s_per_gpu": 3 * n * reps / (a.elapsed_time(bb) * 1e-3),
                                  "note": "B2P_ORDER_CANONICAL on this rank's GPU; the headline uses B2P_ORDER_FAST (same uniform law, "
                                          "both bit-exact against the oracle)"}

    # ---- e2e: the reference-facing call, host buffers in and out ------------------------------------------------
    if not args.no_e2e:
        from gpu_ai_b200 import engine as eng_mod
        packed = d_states.cpu().numpy().view(np.uint32)
        s776 = eng_mod.unpack776(packed)            # n reference `State` objects (776 B each) in host memory
        res = np.empty(n, dtype=np.int32)
        mode_i = mode
        for _ in range(2):
            eng.run_states776(s776, mode=mode_i, out=res)
        if world > 1:
            dist.barrier()
        k = max(3, min(args.steps, 10))
        t_start = time.perf_counter()
        for _ in range(k):
            eng.run_states776(s776, mode=mode_i, out=res)
        dt = time.perf_counter() - t_start
        dt = sharding.max_over_ranks(dt, dev)
        out["e2e"] = {"value": n * world * k / dt, "unit": "playouts/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": n,
                      "ms_per_call": 1e3 * dt / k, "host_input_bytes_per_step": 776 * n,
                      "api": "b2p_run_states776(host State[776 B] x n) -> int32 PlayerId[n]; pack + H2D + kernel + D2H timed"}
        # the same call on 16-byte packed leaves (what a caller that keeps leaves packed would pay)
        for _ in range(2):
            eng.run_packed(packed, reps=1, key=1, mode=mode_i, order=order)
        t_start = time.perf_counter()
        for i in range(k):
            eng.run_packed(packed, reps=1, key=2 + i, mode=mode_i, order=order)
        dtp = sharding.max_over_ranks(time.perf_counter() - t_start, dev)
        out["e2e_packed"] = {"value": n * world * k / dtp, "unit": "playouts/s", "ms_per_call": 1e3 * dtp / k,
                             "api": "b2p_run_packed(host b2p_state16 x n) -> int8 winners + counters"}
        del s776

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) --------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        packed = d_states[: 1 << 17].cpu().numpy().view(np.uint32)
        base, _, _ = cpu_playouts_per_s(packed, args.mode)
        out["cpu_baseline"] = base

    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
