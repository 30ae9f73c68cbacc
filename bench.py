#!/usr/bin/env python
"""bench.py — particle-steps/s of the kD-tree N-body step (build + walk + kick/drift) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--number 1000000] [--impl kdnb|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[2] — circular-orbit ring, N=1,000,000 (+ the central body), dt=1e-3, THETA=0.3,
MAX_PARTS=8; a "step" is one simple_sim step (Parallel/RustVersion/src/array_kd_tree.rs:632-663) through the C ABI.
Prints ONE JSON line (rank 0); the line also carries an "n10m" sub-record (BASELINE configs[3], N=10,000,000, measured
in the same invocation) and, at N > 1, a "parity" object (every rank's final state hashed and compared with a 1-GPU run).
The oracle is used only for the cpu_baseline / --impl reference legs.
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

def host_threads() -> int:
    """All host threads this process may use — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port(n: int, steps: int, warmup: int):
    """The oracle's OpenMP restatement of the Rust rayon path (build_tree_par4 + calc_accel + kick/drift) on all host cores."""
    from oracle.okd import ORDER_FAITHFUL, Oracle, build
    build()
    orc = Oracle()
    threads = host_threads()
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
    threads = host_threads()
    env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    t0 = time.perf_counter()
    r = subprocess.run([exe, str(steps), str(n), str(threads)], capture_output=True, text=True, env=env)
    t = time.perf_counter() - t0
    if r.returncode != 0:
        return None
    return n * steps / t, t, threads


def cpu_sample_steps(n: int) -> int:
    # about 10-30 s of CPU work: ~4-5 us per particle-step per core-equivalent on 8 cores (SURVEY.md §6 probe)
    cores = host_threads()
    est_step = n * 40e-6 / max(1, cores)
    return int(max(1, min(10, round(15.0 / max(est_step, 1e-3)))))


def run_reference(args) -> None:
    """CPU arm: ALWAYS the same program (the oracle's OpenMP port of the rayon path, kind "port") on all host threads,
    with the caller's steps and warm-up; the reference's own C++ sibling is timed beside it as an extra field."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.number
    steps, warmup = args.steps, args.warmup
    v_port, t_port, threads = cpu_port(n, steps, warmup)
    ref = cpu_ref_binary(n, min(steps, 3)) if not args.no_ref_binary else None
    line = {
        "impl": "reference", "metric": METRIC, "value": v_port, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * t_port / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "host": f"{host_threads()} host threads"},
        "cpu_baseline": {
            "value": v_port, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} full steps (+{warmup} warm-up) at N={n:,} on all host threads",
            "reference_cpp_value": ref[0] if ref else None,
            "note": "port = oracle OpenMP restatement of Parallel/RustVersion (rayon path), thread count set explicitly (torchrun "
                    "exports OMP_NUM_THREADS=1); reference_cpp_value = the reference's own Parallel/CppVersion built from "
                    "/root/reference (MAX_PARTS=7, -Ofast, wall clock incl. IC generation, <= 3 steps), reported beside it; the Rust "
                    "crate itself cannot be built in this image (no rustc/cargo)",
        },
        "e2e": {"value": v_port, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm

def nvlink_counters(index: int):
    """Cumulative NVLink data bytes (tx, rx) of one GPU from NVML, or None where the counters are not exposed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ids = [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX]
        vals = pynvml.nvmlDeviceGetFieldValues(h, ids)
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)   # KiB counters
        return tuple(out)
    except Exception:
        return None


def ncu_summary(n: int, world: int):
    """Counters of the committed ncu capture of the walk kernel for this (N, GPUs), if one exists (profiles/ncu_walk.json
    names the capture and the commit it was taken at).  Never a literal in this file."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_walk.json")))
        return table.get(f"n={n},gpus={world}")
    except Exception:
        return None


def measure(kd, L, torch, dist, local, rank, world, n, K, W, with_counts=True):
    """One workload size: device-timed value, stage times (same execution mode: graph replay), e2e with host buffers."""
    import ctypes as C

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

    def new_sim(flags=0):
        sim = kd.KDTreeSim(device=local, flags=flags)
        if world > 1:
            ids = [kd.KDTreeSim.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            sim.comm_init(ids[0], rank, world)
        return sim

    nb = (n + 1) * 64
    host_ptr = L.kdnb_host_alloc(nb)                      # pinned host buffer for the e2e leg
    assert host_ptr, "kdnb_host_alloc failed"
    host = np.ctypeslib.as_array(C.cast(host_ptr, C.POINTER(C.c_uint8)), shape=(nb,)).view(kd.PARTICLE)
    ics = kd.circular_orbits(n, seed=SEED)
    host[:] = ics

    sim = new_sim()                                        # timed arm: plain context (the step is replayed as a CUDA graph)
    sim.upload(host)
    sim.simple_sim(DT, W)                                  # warm-up (untimed)
    sim.synchronize()
    launches0 = sim.launch_count
    nv0 = nvlink_counters(local) if world > 1 else None
    barrier()
    sim.stopwatch_begin()
    sim.simple_sim(DT, K)                                  # EXACTLY K steps, inputs resident in HBM
    ms = sim.stopwatch_end()
    barrier()
    nv1 = nvlink_counters(local) if world > 1 else None
    launches = sim.launch_count - launches0
    ms = max_over_ranks(ms)
    # per-stage CUDA-event times: a second context on the same state in the SAME execution mode (the step graph carries
    # event-record nodes at the stage boundaries; every step is waited for so that the events can be read)
    simp = new_sim(kd.FLAG_PROFILE)
    simp.upload(host)
    simp.simple_sim(DT, W)
    simp.stage_reset()
    simp.stopwatch_begin()
    simp.simple_sim(DT, K)
    prof_ms = simp.stopwatch_end()
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
    sim.close()
    L.kdnb_host_free(host_ptr)

    rec = {
        "value": value, "ms_per_step": ms / K, "steps": K, "warmup": W, "gpu_launches": int(launches),
        "stage_ms_per_step": {k: stage[k] / max(1, nsteps) for k in ("build", "walk", "kick", "exchange")},
        # the context the stage times come from: its own K steps, device-timed (five event-record nodes per step and a
        # host wait after every step, so slightly above ms_per_step; the stage times add up to at most this)
        "stage_context_ms_per_step": prof_ms / K,
        "e2e": {"value": (n + 1) * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb,
                "note": "K calls of kdnb_simple_sim_bodies[_sharded](host, dt, 1): pinned-host upload + one step + download per call; bytes are the sum over ranks (each rank moves its 1/N host shard over PCIe, NVLink all-gather for the rest)"},
        "e2e_amortized": {"value": (n + 1) * K / e2e_amort_s, "unit": UNIT, "note": "one call simple_sim(bodies, dt, K) with host buffers"},
    }
    if world > 1:
        algo = 24.0 * (n + 1) * (world - 1) / world       # bytes one rank receives per step: everybody else's accelerations
        rec["nvlink"] = {
            "algorithmic_rx_bytes_per_step_per_gpu": algo,
            "measured_rx_bytes_per_step": (nv1[1] - nv0[1]) / K if nv0 and nv1 else None,
            "measured_tx_bytes_per_step": (nv1[0] - nv0[0]) / K if nv0 and nv1 else None,
            "source": "NVML NVLINK_THROUGHPUT_DATA_{TX,RX} of this rank's GPU around the timed region (rank 0)" if nv0 and nv1 else "NVML counters not exposed",
        }
    if rank == 0 and with_counts:
        # ---- roofline of the dominant kernel (walk): counted flops / measured DFMA peak
        with kd.KDTreeSim(device=local, flags=kd.FLAG_WALK_COUNTS) as cs:
            cs.upload(ics)
            cs.build_tree()
            cs.calc_accel()
            cnt = cs.walk_counts().sum(axis=0).astype(np.float64)
            fp64_peak = cs.fp64_peak_tflops()
        V, A, LV, P = cnt
        flops_step = 10 * V + 6 * A + 3 * (V - A) + 18 * P       # SURVEY.md §8(d) flop model, counted for this run
        walk_ms = rec["stage_ms_per_step"]["walk"]
        achieved = flops_step / world / (walk_ms * 1e-3) / 1e12
        ncu = ncu_summary(n, world) or {}
        rec["roofline"] = {
            "kernel": "walk2_kernel", "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak if fp64_peak else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed capture named in
            # profiles/ncu_walk.json for this (N, GPUs) — null when no capture of this configuration is committed
            "traffic": ncu.get("dram_bytes"),
            "peak_source": "DFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
            "flops_per_particle_step": flops_step / (n + 1),
            "counts_per_particle": {"node_tests": V / (n + 1), "accepts": A / (n + 1), "leaf_visits": LV / (n + 1), "pairs": P / (n + 1)},
            "share_of_step": walk_ms / (ms / K),
            "ncu": ncu or None,
        }
    return rec


def parity_digest(kd, torch, dist, local, rank, world, n=200_000, steps=3):
    """Multi-GPU parity, outside every timed region: every rank runs `steps` steps of the sharded path on a fresh context
    and hashes its final state + accelerations; rank 0 repeats the run on ONE GPU.  All digests must be equal."""
    import hashlib
    ics = kd.circular_orbits(n, seed=4242)
    sim = kd.KDTreeSim(device=local)
    ids = [kd.KDTreeSim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.comm_init(ids[0], rank, world)
    sim.upload(ics)
    sim.build_tree()
    sim.calc_accel()
    acc = sim.accel()
    sim.simple_sim(DT, steps)
    out = sim.download()
    sim.close()
    dig = hashlib.sha256(out.tobytes() + acc.tobytes()).hexdigest()
    digs = [None] * world
    dist.all_gather_object(digs, dig)
    res = None
    if rank == 0:
        with kd.KDTreeSim(device=local) as one:
            one.upload(ics)
            one.build_tree()
            one.calc_accel()
            acc1 = one.accel()
            one.simple_sim(DT, steps)
            out1 = one.download()
        dig1 = hashlib.sha256(out1.tobytes() + acc1.tobytes()).hexdigest()
        res = {"ranks_bit_identical": all(d == digs[0] for d in digs), "equals_single_gpu": digs[0] == dig1,
               "n": n + 1, "steps": steps, "digest": digs[0][:16],
               "what": "sha256 of final particle state + first-step accelerations on every rank vs a 1-GPU run on rank 0"}
    dist.barrier()
    return res


def run_kdnb(args) -> None:
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

    n, K, W = args.number, args.steps, max(args.warmup, 3)
    L = _lib.load()
    sampler = ClockSampler(local)
    sampler.start()
    top = measure(kd, L, torch, dist, local, rank, world, n, K, W)
    clocks = sampler.stop()
    n10 = None
    if not args.no_10m and n != 10_000_000:
        # BASELINE configs[3]: the north-star size, same invocation, fewer steps (a step is ~26 ms on one GPU)
        s10 = ClockSampler(local)
        s10.start()
        n10 = measure(kd, L, torch, dist, local, rank, world, 10_000_000, max(3, min(K, 5)), 3)
        n10["clocks"] = s10.stop()
        n10["workload"] = workload_name(10_000_000)
    parity = parity_digest(kd, torch, dist, local, rank, world) if world > 1 else None

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

        def hbm(nn, rec):
            levels = int(np.ceil(np.log2((nn + 1) / 8)))
            build_bytes = (nn + 1) * (32.0 * levels) + 80.0 * 2 ** (levels + 1)   # SURVEY.md §8(d) lower-bound model
            kick_bytes = (nn + 1) * 120.0
            b, k = rec["stage_ms_per_step"]["build"], rec["stage_ms_per_step"]["kick"]
            return {
                "peak": hbm_peak, "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "build": {"bytes_model": build_bytes, "achieved": build_bytes / (b * 1e-3) / 1e9, "frac": build_bytes / (b * 1e-3) / 1e9 / hbm_peak},
                "kick": {"bytes_model": kick_bytes, "achieved": kick_bytes / (k * 1e-3) / 1e9, "frac": kick_bytes / (k * 1e-3) / 1e9 / hbm_peak},
            }

        line = {
            "metric": METRIC, "value": top["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(n), "parallelism": f"replicated tree, walk sharded over {world} GPU(s) by tree-ordered ranges, accelerations exchanged by peer stores over NVLink from inside the walk kernel (ncclAllGather fallback)",
                "l2": "no explicit flush: at N=1M every step streams ~0.5 GB (radix-sort ping-pong, level partitions, bottom build) through the 126 MB L2 before the walk; at N>=10M the inputs themselves exceed L2",
                "timer": "CUDA events on the library's stream around K steps (step replayed as a CUDA graph), max over ranks; stage_ms from a second context in the same mode (graph replay with event-record nodes at the stage boundaries)",
            },
            "stage_ms_per_step": top["stage_ms_per_step"],
            "stage_context_ms_per_step": top["stage_context_ms_per_step"],
            "roofline": top.get("roofline"),
            "roofline_hbm": hbm(n, top),
            "e2e": top["e2e"], "e2e_amortized": top["e2e_amortized"],
            "gpu_launches": top["gpu_launches"], "clocks": clocks,
        }
        if "nvlink" in top:
            line["nvlink"] = top["nvlink"]
        if parity is not None:
            line["parity"] = parity
        if n10 is not None:
            n10["roofline_hbm"] = hbm(10_000_000, n10)
            n10["unit"] = UNIT
            line["n10m"] = n10
        if not args.no_cpu and world == 1:
            s = cpu_sample_steps(n)
            v, t, threads = cpu_port(n, s, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{s} full step(s) at N={n:,} ({t:.1f} s) with the oracle's OpenMP restatement of the rayon path"}
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
    ap.add_argument("--no-10m", action="store_true", help="skip the N=10M sub-record")
    ap.add_argument("--no-ref-binary", action="store_true", help="reference arm: do not time the reference's C++ sibling beside the port")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_kdnb(args)


if __name__ == "__main__":
    main()
