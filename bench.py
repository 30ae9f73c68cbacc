#!/usr/bin/env python
"""bench.py — particle-steps/s of the kD-tree N-body step (build + walk + kick/drift) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--number 1000000] [--impl kdnb|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[2] — circular-orbit ring, N=1,000,000 (+ the central body), dt=1e-3, THETA=0.3,
MAX_PARTS=8; a "step" is one simple_sim step (Parallel/RustVersion/src/array_kd_tree.rs:632-663) through the C ABI.
Prints ONE JSON line (rank 0).  The oracle is used only for the cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

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

DT = 1e-3
SEED = 12345
METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"


def workload_name(n: int) -> str:
    return f"circular-orbit ring, N={n:,} (+1 central body), dt=1e-3, THETA=0.3, MAX_PARTS=8, full step = build+walk+kick/drift (BASELINE configs[2] shape)"


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle sampling through NVML during the timed region."""

    def __init__(self, index: int, period: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self._stop_ev = index, period, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self) -> dict:
        self._stop_ev.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU legs

def cpu_port(n: int, steps: int, warmup: int):
    """The oracle's OpenMP restatement of the Rust rayon path (build_tree_par4 + calc_accel + kick/drift) on all host cores."""
    from oracle.okd import ORDER_FAITHFUL, Oracle, build
    build()
    orc = Oracle()
    threads = orc.max_threads()
    parts = orc.circular_orbits(n, seed=SEED)
    if warmup:
        orc.simple_sim(parts, DT, warmup, order=ORDER_FAITHFUL, seed=1, threads=threads)
    t0 = time.perf_counter()
    orc.simple_sim(parts, DT, steps, order=ORDER_FAITHFUL, seed=2, threads=threads)
    t = time.perf_counter() - t0
    return (n + 1) * steps / t, t, threads


def cpu_ref_binary(n: int, steps: int):
    """The reference's own compiled sibling (Parallel/CppVersion, built by oracle/Makefile into oracle/_ref) — whole-process
    wall clock including its IC generation, exactly how the reference's times.txt were taken (benchmark.sh)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "kdtree-sim-cpp")
    if not os.path.exists(exe):
        return None
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    r = subprocess.run([exe, str(steps), str(n), str(threads)], capture_output=True, text=True)
    t = time.perf_counter() - t0
    if r.returncode != 0:
        return None
    return n * steps / t, t, threads


def cpu_sample_steps(n: int) -> int:
    # about 10-30 s of CPU work: ~4-5 us per particle-step per core-equivalent on 8 cores (SURVEY.md §6 probe)
    cores = os.cpu_count() or 1
    est_step = n * 40e-6 / max(1, cores)
    return int(max(1, min(10, round(15.0 / max(est_step, 1e-3)))))


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.number
    steps, warmup = args.steps, min(args.warmup, 1)
    v_port, t_port, threads = cpu_port(n, steps, warmup)
    ref = cpu_ref_binary(n, steps)
    kind, value, t = "port", v_port, t_port
    if ref and ref[0] > v_port:
        kind, value, t = "reference", ref[0], ref[1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "host": f"{os.cpu_count()} logical cores"},
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{steps} full steps at N={n:,} on all host threads",
            "port_value": v_port, "reference_cpp_value": ref[0] if ref else None,
            "note": "port = oracle OpenMP restatement of Parallel/RustVersion (rayon path); reference = the reference's own "
                    "Parallel/CppVersion built from /root/reference (MAX_PARTS=7, -Ofast, wall clock incl. IC generation); "
                    "the Rust crate itself cannot be built in this image (no rustc/cargo)",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm

def run_kdnb(args) -> None:
    import ctypes as C

    import torch

    import multilanguagekdtree_b200 as kd
    from multilanguagekdtree_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libkdnb has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, K, W = args.number, args.steps, max(args.warmup, 3)
    L = _lib.load()
    nb = (n + 1) * 64
    host_ptr = L.kdnb_host_alloc(nb)                      # pinned host buffer for the e2e leg
    assert host_ptr, "kdnb_host_alloc failed"
    host = np.ctypeslib.as_array(C.cast(host_ptr, C.POINTER(C.c_uint8)), shape=(nb,)).view(kd.PARTICLE)
    ics = kd.circular_orbits(n, seed=SEED)
    host[:] = ics

    sim = kd.KDTreeSim(device=local)                      # timed arm: plain context (the step is replayed as a CUDA graph)
    if world > 1:
        ids = [kd.KDTreeSim.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.comm_init(ids[0], rank, world)
    sim.upload(host)
    sim.simple_sim(DT, W)                                  # warm-up (untimed)
    sim.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sim.launch_count
    barrier()
    sim.stopwatch_begin()
    sim.simple_sim(DT, K)                                  # EXACTLY K steps, inputs resident in HBM
    ms = sim.stopwatch_end()
    barrier()
    launches = sim.launch_count - launches0
    ms = max_over_ranks(ms)
    # per-stage CUDA-event times: a second, profiled context on the same state (plain launches, events between stages)
    simp = kd.KDTreeSim(device=local, flags=kd.FLAG_PROFILE)
    if world > 1:
        ids2 = [kd.KDTreeSim.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids2, src=0)
        simp.comm_init(ids2[0], rank, world)
    simp.upload(host)
    simp.simple_sim(DT, W)
    simp.stage_reset()
    simp.simple_sim(DT, K)
    stage, nsteps = simp.stage_ms()
    simp.close()
    value = (n + 1) * K / (ms * 1e-3)

    # ---- e2e: the reference-facing call with HOST buffers, copies inside the timed region, every step.
    # N > 1: every rank owns the slice host_shard_range(n, rank, world) of the host array (sharded host state, as a
    # multi-process caller would hold it): its PCIe traffic is 1/N of the state, the rest travels over NVLink.
    host[:] = ics
    if world > 1:
        first, cnt = kd.host_shard_range(n + 1, rank, world)
        shard = host[first:first + cnt]

        def e2e_call(k):
            sim.simple_sim_bodies_sharded(shard, n + 1, DT, k)
    else:
        def e2e_call(k):
            sim.simple_sim_bodies(host, DT, k)
    e2e_call(1)                                             # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_call(1)                                         # upload 64 B/particle, one step, download 64 B/particle
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    e2e_call(K)                                             # the reference's own call shape: simple_sim(bodies, dt, K)
    e2e_amort_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (walk): counted flops / measured DFMA peak
        with kd.KDTreeSim(device=local, flags=kd.FLAG_WALK_COUNTS) as cs:
            cs.upload(ics)
            cs.build_tree()
            cs.calc_accel()
            cnt = cs.walk_counts().sum(axis=0).astype(np.float64)
            fp64_peak = cs.fp64_peak_tflops()
        V, A, LV, P = cnt
        flops_step = 10 * V + 6 * A + 3 * (V - A) + 18 * P       # SURVEY.md §8(d) flop model, counted for this run
        walk_ms = stage["walk"] / max(1, nsteps)
        build_ms = stage["build"] / max(1, nsteps)
        kick_ms = stage["kick"] / max(1, nsteps)
        exch_ms = stage["exchange"] / max(1, nsteps)
        achieved = flops_step / world / (walk_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        levels = int(np.ceil(np.log2((n + 1) / 8)))
        build_bytes = (n + 1) * (32.0 * levels) + 80.0 * 2 ** (levels + 1)   # SURVEY.md §8(d) lower-bound model
        kick_bytes = (n + 1) * 120.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(n), "parallelism": f"replicated tree, walk sharded over {world} GPU(s) by tree-ordered ranges, accelerations exchanged by peer stores over NVLink from inside the walk kernel (ncclAllGather fallback)",
                "l2": "no explicit flush: at N=1M every step streams ~0.5 GB (radix-sort ping-pong, level partitions, bottom build) through the 126 MB L2 before the walk; at N>=10M the inputs themselves exceed L2",
                "timer": "CUDA events on the library's stream around K steps (step replayed as a CUDA graph), max over ranks; stage_ms from a second, profiled context (plain launches)",
            },
            "stage_ms_per_step": {"build": build_ms, "walk": walk_ms, "kick": kick_ms, "exchange": exch_ms},
            "roofline": {
                "kernel": "walk2_kernel", "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                # DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture
                # profiles/r01_walk_kernel_ncu_full_raw.csv (53.0 MB read + 7.1 MB written); measured at N=1M on one GPU only
                "traffic": 60.2e6 if (n == 1_000_000 and world == 1) else None,
                "peak_source": "DFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_per_particle_step": flops_step / (n + 1),
                "counts_per_particle": {"node_tests": V / (n + 1), "accepts": A / (n + 1), "leaf_visits": LV / (n + 1), "pairs": P / (n + 1)},
                "share_of_step": walk_ms / (ms / K),
                # ncu of the committed capture (profiles/r01_walk_kernel_ncu_full_raw.csv, N=1M, one GPU): the flop model
                # counts sqrt and divide as one flop each, the pipe-utilisation figure is the gauge of the FP64 pipe
                "fp64_pipe_active_pct_ncu": 65.8 if (n == 1_000_000 and world == 1) else None,
            },
            "roofline_hbm": {
                "peak": hbm_peak, "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "build": {"bytes_model": build_bytes, "achieved": build_bytes / (build_ms * 1e-3) / 1e9, "frac": build_bytes / (build_ms * 1e-3) / 1e9 / hbm_peak},
                "kick": {"bytes_model": kick_bytes, "achieved": kick_bytes / (kick_ms * 1e-3) / 1e9, "frac": kick_bytes / (kick_ms * 1e-3) / 1e9 / hbm_peak},
            },
            "e2e": {"value": (n + 1) * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb,
                    "note": "K calls of kdnb_simple_sim_bodies[_sharded](host, dt, 1): pinned-host upload + one step + download per call; bytes are the sum over ranks (each rank moves its 1/N host shard over PCIe, NVLink all-gather for the rest)"},
            "e2e_amortized": {"value": (n + 1) * K / e2e_amort_s, "unit": UNIT, "note": "one call simple_sim(bodies, dt, K) with host buffers"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            s = cpu_sample_steps(n)
            v, t, threads = cpu_port(n, s, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{s} full step(s) at N={n:,} ({t:.1f} s) with the oracle's OpenMP restatement of the rayon path"}
    sim.close()
    L.kdnb_host_free(host_ptr)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--number", "-n", type=int, default=1_000_000)
    ap.add_argument("--impl", default="kdnb", choices=["kdnb", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_kdnb(args)


if __name__ == "__main__":
    main()
