#!/usr/bin/env python
"""bench.py -- headline benchmark of the unified_rectrxm! hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): Float64 left / lower / no-trans TRSM, A 16384 x 16384, 16384 right-hand sides
per GPU, through unified_rectrxm (nla_rectrxm in the C ABI).  One "step" = one full solve of one batch of RHS.
Metric: TFLOP/s = n^2 * m_total / time.  N > 1: RHS columns are sharded by rank (16384 per GPU, weak
scaling), rank 0 owns A and broadcasts it over NCCL inside every timed step; no other exchange.

Timing: per step a CUDA-event pair on the launching stream around [broadcast +] solve; B is restored from a pristine
copy between steps OUTSIDE the event pair (the operation is in place); the K step times are summed, max over ranks.
Inputs (A 2 GiB + B 2 GiB) are far larger than the 126 MB L2, so no explicit flush is needed.

--impl reference: the reference is Julia and cannot run in this image (no Julia); this arm times the
oracle's C/OpenMP restatement of the reference algorithm (oracle/nla_oracle.c) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ORDER = 16384          # order of A
M_PER_GPU = 16384        # right-hand sides per GPU
FP64_PEAK_TFLOPS = 37.0  # measured DMMA.8x8x4 issue-rate peak on this pool (profiles/r01_probe_dmma_peak.txt); nominal 37.2
GEMM_TRAFFIC_BYTES = 14.73e9  # dram read+write of the top-level update (K=8192, 8192 tiles) from ncu --set full, profiles/r01_ncu_prof_gemm_top_summary.csv
METRIC = "fp64_trsm_left_lower_n16384_tflops"
UNIT = "TFLOP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_ORDER, help="override the order of A (debug only; changes the workload name)")
    ap.add_argument("--m", type=int, default=M_PER_GPU, help="override RHS per GPU (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="N > 1: one blocking broadcast of A before the solve instead of the panel pipeline")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("NLA_STREAMS", "0")), help="0 = library default")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def cpu_port_time(n: int, m: int, reps: int = 1):
    """Times the oracle's C/OpenMP restatement of the reference algorithm (FP64 L/L/N TRSM) on the host cores."""
    import numpy as np
    from oracle import c_port
    from oracle import reference_port as rp

    A, B0 = rp.make_inputs(n, m, "L", "L", np.float64, seed=99, recipe="scaled")
    best = 1e30
    for _ in range(reps):
        B = B0.copy(order="F")
        t0 = time.perf_counter()
        c_port.unified_rectrxm("L", "L", "N", 1.0, "S", A, B)
        best = min(best, time.perf_counter() - t0)
    return best, c_port.num_threads()


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (restated, Julia being unavailable) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun silently exports OMP_NUM_THREADS=1 to every rank; this arm is the reference's CPU path "with all the host threads it can
        # use", and only rank 0 runs it: give it the cores of the box back (must happen before the OpenMP runtime of the oracle loads)
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    n = m = 2048  # bounded sample: 1/512 of the flops of the headline workload per step
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_time(n, m)
    t = 0.0
    cores = 1
    for _ in range(args.steps):
        dt, cores = cpu_port_time(n, m)
        t += dt
    val = args.steps * float(n) * n * m / t * 1e-12
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Float64 left/lower/no-trans TRSM via unified_rectrxm!, A 16384x16384, 16384 RHS per GPU (BASELINE configs[1])",
                       "sample": f"n={n}, m={m} slice of that workload per step (same algorithm, thresholds and split rule)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle/nla_oracle.c (C/OpenMP restatement of src/rectrxm.jl+trsm.jl+matmul.jl; Julia not installed), n=m={n}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nla = ge.load_package()
    h = nla.Handle(local)
    h.set_option("streams", args.streams)

    n, m = args.n, args.m
    dt = torch.float64
    # synthetic inputs, "scaled" recipe of SURVEY.md 8(d): strict triangle U(-1,1)/sqrt(n), diagonal U(1,2), B = U(0,1)+1
    g = torch.Generator(device=dev).manual_seed(1234 + 1)
    A = torch.empty((n, n), dtype=dt, device=dev).t()  # column-major
    if rank == 0:
        A.copy_((2 * torch.rand(n, n, dtype=dt, device=dev, generator=g) - 1) / n ** 0.5)
        A.copy_(torch.tril(A, -1) + torch.diag(1 + torch.rand(n, dtype=dt, device=dev, generator=g)))
    else:
        A.zero_()
    gb = torch.Generator(device=dev).manual_seed(777 + rank)
    B0 = torch.empty((m, n), dtype=dt, device=dev).t()
    B0.copy_(torch.rand(n, m, dtype=dt, device=dev, generator=gb) + 1)
    X = torch.empty((m, n), dtype=dt, device=dev).t()
    A_store = A.t()  # contiguous (n x n) storage for the broadcast

    from importlib import import_module
    sharded = import_module(nla.__name__ + ".sharded")

    def step():
        # the only exchange: A from its owner to every GPU (NCCL over NVLink/NVSwitch).  Default: column panels of A in the order the
        # schedule consumes them, on a side stream, the solve gated per panel (nla_rectrxm_gated) -> the broadcast hides behind the solve
        if world > 1 and not args.no_pipeline:
            sharded.unified_rectrxm_pipelined("L", "L", "N", 1.0, "S", A, X, src=0, panels=8, handle=h)
            return
        if world > 1:
            dist.broadcast(A_store, src=0)
        nla.unified_rectrxm("L", "L", "N", 1.0, "S", A, X, handle=h)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        X.copy_(B0)
        step()
    sync_all()

    # ---- timed region: K steps, CUDA events per step on the launching stream ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    h.launch_count(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        X.copy_(B0)  # restore the in-place operand, outside the event pair
        ev[k][0].record()
        step()
        ev[k][1].record()
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = h.launch_count()
    # per-launch device times for the roofline: two extra steps right after the timed region with the RHS slabs serialised on
    # one stream (with concurrent slabs the per-launch event intervals would overlap and over-count)
    h.set_option("streams", 1)
    h.set_option("profile", 1)
    prof_steps = 2
    for _ in range(prof_steps):
        X.copy_(B0)
        nla.unified_rectrxm("L", "L", "N", 1.0, "S", A, X, handle=h)
    torch.cuda.synchronize()
    prof = h.profile_read()
    h.set_option("profile", 0)
    h.set_option("streams", args.streams)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = total_ms.item()
    flops_per_step_all = float(n) * n * m * world
    value = args.steps * flops_per_step_all / (total_ms * 1e-3) * 1e-12

    # ---- accuracy gate on the last timed result (backward error, independent cuBLAS FP64 product) ----
    R = torch.tril(A) @ X - B0
    berr = (torch.linalg.norm(R) / (torch.linalg.norm(torch.tril(A)) * torch.linalg.norm(X) + torch.linalg.norm(B0))).item()
    del R

    # ---- roofline of the dominant kernel (GEMM update): algorithmic flops / launch-time, from the per-launch events ----
    gemm = [(f, ms) for k, f, ms in prof if k == 1]
    leafs = [(f, ms) for k, f, ms in prof if k == 0]
    g_fl, g_ms = sum(f for f, _ in gemm), sum(ms for _, ms in gemm)
    l_fl, l_ms = sum(f for f, _ in leafs), sum(ms for _, ms in leafs)
    top = max(gemm, key=lambda r: r[0]) if gemm else (0.0, 1.0)
    achieved = g_fl / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_f64_tma_kernel (FP64 DMMA update, recursion levels above the fused-slab cutoff)", "achieved": achieved, "peak": FP64_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": GEMM_TRAFFIC_BYTES,
                "peak_source": "measured DMMA.8x8x4 issue-rate probe on this pool (profiles/r01_probe_dmma_peak.txt); MEASURED_PEAKS.json has no FP64 figure; "
                               "nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2; cuBLAS DGEMM measured 35.4",
                "launches_per_step": len(gemm) // prof_steps, "avg_launch_ms": g_ms / max(1, len(gemm)),
                "flops_per_step": g_fl / prof_steps,
                "how": "algorithmic flops of all GEMM-update launches / their CUDA-event durations, 2 profiled steps run right after the timed region on one stream",
                "top_level_launch": {"flops": top[0], "ms": top[1], "tflops": top[0] / (top[1] * 1e-3) * 1e-12},
                "gemm_share_of_step": g_ms / max(1e-9, g_ms + l_ms), "leaf_share_of_step": l_ms / max(1e-9, g_ms + l_ms),
                "fused_slab_launches_per_step": len(leafs) // prof_steps,
                "leaf_tflops": l_fl / (l_ms * 1e-3) * 1e-12 if l_ms > 0 else None,
                "whole_step_frac_of_peak": value / world / FP64_PEAK_TFLOPS}

    # ---- e2e: the same call with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        es = 8
        hostA = torch.empty((n, n), dtype=dt, pin_memory=True) if rank == 0 or world == 1 else None
        hostB = torch.empty((m, n), dtype=dt, pin_memory=True)
        if hostA is not None:
            hostA.copy_(A_store)
        hostB.copy_(B0.t())
        hostX = torch.empty((m, n), dtype=dt, pin_memory=True)
        lib = nla.load_library()
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            hostX.copy_(hostB)  # restore (host side, outside the timed call)
            t0 = time.perf_counter()
            if world == 1:
                rc = lib.nla_rectrxm_host(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, hostA.data_ptr(), n, hostX.data_ptr(), n)
                assert rc == 0, rc
            else:
                # host A -> owner GPU -> all GPUs panel by panel; every rank streams its own B through the library's host pipeline
                sharded.unified_rectrxm_pipelined_host("L", "L", "N", 1.0, "S", A, hostA.t() if hostA is not None else None, hostX.t(),
                                                       src=0, panels=8, handle=h)
                torch.cuda.synchronize()
            return time.perf_counter() - t0

        e2e_step()
        sync_all()
        tt = 0.0
        for _ in range(e2e_steps):
            sync_all()
            tt += e2e_step()
        tmax = torch.tensor([tt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tt = tmax.item()
        # bytes actually copied: only the referenced triangle of A crosses PCIe -- as 1024 x 1024 tiles on or below the diagonal
        # (nla_rectrxm_host) or as the trapezoid of each of the 8 column panels (multi-GPU path) -- plus every rank's B in and out
        if world == 1:
            nt = -(-n // 1024)
            a_bytes = sum(min(1024, n - i * 1024) * min(1024, n - j * 1024) for i in range(nt) for j in range(i + 1)) * es
        else:
            pc, npan = sharded.panel_geometry(n, 8)
            a_bytes = sum((n - p * pc) * (min(n, (p + 1) * pc) - p * pc) for p in range(npan)) * es
        e2e = {"value": e2e_steps * flops_per_step_all / tt * 1e-12, "unit": UNIT,
               "h2d_bytes_per_step": a_bytes + n * m * world * es, "d2h_bytes_per_step": n * m * world * es, "steps": e2e_steps,
               "ms_per_step": tt / e2e_steps * 1e3, "api": "nla_rectrxm_host (pinned host A and B in, B out)" if world == 1 else
               "sharded.unified_rectrxm_pipelined_host: H2D(A panels) + ncclBroadcast per panel + nla_rectrxm_hostb_gated (B streamed in chunks, launches gated on the panels of A)"}
        # check the e2e result too
        Xh = hostX.to(dev).t()
        R = torch.tril(A) @ Xh - B0
        e2e["backward_error"] = (torch.linalg.norm(R) / (torch.linalg.norm(torch.tril(A)) * torch.linalg.norm(Xh) + torch.linalg.norm(B0))).item()
        del R, Xh

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cn = 8192
        sec, cores = cpu_port_time(cn, cn)
        cpu = {"value": float(cn) ** 3 / sec * 1e-12, "unit": UNIT, "cores": cores, "kind": "port", "seconds": sec,
               "sample": f"oracle/nla_oracle.c (C/OpenMP restatement of the reference algorithm; Julia unavailable), FP64 L/L/N TRSM n=m={cn} "
                         f"(1/8 of the headline flops), {cores} OpenMP threads"}
        try:
            from scipy.linalg import blas
            from oracle import reference_port as rp

            Ah, Bh = rp.make_inputs(cn, cn, "L", "L", np.float64, seed=5, recipe="scaled")
            t0 = time.perf_counter()
            blas.dtrsm(1.0, Ah, Bh, side=0, lower=1, trans_a=0, diag=0)
            cpu["openblas_dtrsm_tflops"] = float(cn) ** 3 / (time.perf_counter() - t0) * 1e-12
            cpu["openblas_note"] = "scipy.linalg.blas.dtrsm (OpenBLAS, the reference tests' own oracle) on the same sample, all cores"
        except Exception as e:  # noqa: BLE001
            cpu["openblas_dtrsm_tflops"] = f"unavailable: {e}"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"Float64 left/lower/no-trans TRSM via unified_rectrxm!, A {n}x{n}, {m} RHS per GPU (BASELINE configs[1])",
                           "n": n, "rhs_per_gpu": m, "rhs_total": m * world, "alpha": 1.0, "inputs": "scaled recipe (SURVEY 8(d)), seed 1235/777+rank",
                           "l2": "inputs (A 2 GiB + B 2 GiB per GPU) larger than L2; B restored from a pristine copy between steps outside the timed events",
                           "parallelism": (f"rhs-sharded x{world}, A broadcast by NCCL inside every step "
                                           + ("(one blocking broadcast)" if args.no_pipeline else "(8 column panels pipelined with the solve)")) if world > 1 else "single GPU",
                           "streams": args.streams or "auto", "leaf": h.get_option("leaf"), "macro": h.get_option("macro")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "backward_error": berr, "tolerance": 1e-13, "wall_ms_per_step_incl_restore": t_wall / args.steps * 1e3,
                "pct_of_fp64_peak": 100.0 * value / world / FP64_PEAK_TFLOPS}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    assert berr < 1e-13, f"backward error {berr} above tolerance"


if __name__ == "__main__":
    main()
