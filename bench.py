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

After the headline the same process measures, under the key "extra", the other legs BASELINE.json's metric names
("fp64/fp32/fp16 at 1/2/4/8"): C3 (Float32 left/upper/transposed TRMM n = m = 16384), Float16 and Float32 solves at
n = 16384, C4 (Float16 right/lower TRSM n = 32768; 131072 RHS rows sharded over the GPUs, one GPU: a 16384-row slice) and
C5 (Float64 TRMM n = 32768, 8192 RHS per GPU) -- each with its backward error against the north_star tolerance.

--impl reference: the reference is Julia and cannot run in this image (no Julia); this arm times the
oracle's C/OpenMP restatement of the reference algorithm (oracle/nla_oracle.c) on the host cores, on a bounded sample of
the SAME workload: the same 16384 x 16384 matrix (same recursion depth, same GEMM K extents), 256 of the right-hand sides.
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
FP64_PEAK_TFLOPS = 37.0  # DMMA.8x8x4 issue-rate peak measured on this pool (profiles/r01_probe_dmma_peak.txt); nominal 148 x 64 FMA x 1.965 GHz = 37.2
METRIC = "fp64_trsm_left_lower_n16384_tflops"
UNIT = "TFLOP/s"
REF_SAMPLE_M = 256       # reference arm: right-hand sides per step (bounded sample of the headline workload, same A)
TOL = {"float64": 1e-13, "float32": 1e-5, "float16": 1e-2}   # north_star backward-error tolerances


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
    ap.add_argument("--no-extra", action="store_true", help="skip the fp32 / fp16 / C4 / C5 legs")
    ap.add_argument("--no-pipeline", action="store_true", help="N > 1: one blocking broadcast of A before the solve instead of the panel pipeline")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("NLA_STREAMS", "0")), help="0 = library default")
    return ap.parse_args()


def workload_config(n: int, m: int, world: int, pipelined: bool = True):
    """The workload description shared verbatim by both arms (`--impl ours` and `--impl reference`)."""
    return {"workload": f"Float64 left/lower/no-trans TRSM via unified_rectrxm!, A {n}x{n}, {m} RHS per GPU (BASELINE configs[1])",
            "n": n, "rhs_per_gpu": m, "rhs_total": m * world, "alpha": 1.0, "inputs": "scaled recipe (SURVEY 8(d)): strict triangle U(-1,1)/sqrt(n), diagonal U(1,2), B = U(0,1)+1",
            "l2": "inputs (A 2 GiB + B 2 GiB per GPU) larger than L2; B restored from a pristine copy between steps outside the timed events",
            "parallelism": (f"rhs-sharded x{world}, A broadcast by NCCL inside every step "
                            + ("(8 column panels pipelined with the solve)" if pipelined else "(one blocking broadcast)")) if world > 1 else "single GPU"}


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


_cpu_inputs = {}


def cpu_port_time(n: int, m: int, reps: int = 1):
    """Times the oracle's C/OpenMP restatement of the reference algorithm (FP64 L/L/N TRSM) on the host cores."""
    import numpy as np
    from oracle import c_port
    from oracle import reference_port as rp

    if (n, m) not in _cpu_inputs:
        _cpu_inputs.clear()
        _cpu_inputs[(n, m)] = rp.make_inputs(n, m, "L", "L", np.float64, seed=99, recipe="scaled")
    A, B0 = _cpu_inputs[(n, m)]
    best = 1e30
    for _ in range(reps):
        B = B0.copy(order="F")
        t0 = time.perf_counter()
        c_port.unified_rectrxm("L", "L", "N", 1.0, "S", A, B)
        best = min(best, time.perf_counter() - t0)
    return best, c_port.num_threads()


def openblas_time(n: int, m: int):
    """OpenBLAS dtrsm (the routine the reference's tests use as their oracle, test/unified_rectrxm.jl:36-40) on the same sample."""
    from scipy.linalg import blas

    A, B0 = _cpu_inputs[(n, m)]
    t0 = time.perf_counter()
    blas.dtrsm(1.0, A, B0, side=0, lower=1, trans_a=0, diag=0)
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (restated, Julia being unavailable) with all host threads, on a bounded sample of the
    headline workload: the SAME n = 16384 matrix (same recursion depth and GEMM K extents), 256 right-hand sides per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun silently exports OMP_NUM_THREADS=1 to every rank; this arm is the reference's CPU path "with all the host threads it can
        # use", and only rank 0 runs it: give it the cores of the box back (must happen before the OpenMP runtime of the oracle loads)
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    n, m = args.n, REF_SAMPLE_M
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_time(n, m)
    t = 0.0
    cores = 1
    for _ in range(args.steps):
        dt, cores = cpu_port_time(n, m)
        t += dt
    val = args.steps * float(n) * n * m / t * 1e-12
    try:
        ob = float(n) * n * m / openblas_time(n, m) * 1e-12
    except Exception as e:  # noqa: BLE001
        ob = f"unavailable: {e}"
    sample = (f"oracle/nla_oracle.c (C/OpenMP restatement of src/rectrxm.jl + trsm.jl + matmul.jl; Julia is not installed), the headline matrix "
              f"(n = {n}: same recursion depth, thresholds, split rule and GEMM K extents) with {m} of the {args.m} right-hand sides per step "
              f"(the reference's kernels process RHS columns independently, so TFLOP/s does not depend on how many there are); {cores} OpenMP threads")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.n, args.m, args.gpus, not args.no_pipeline),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "openblas_dtrsm_tflops": ob,
                             "openblas_note": "scipy.linalg.blas.dtrsm (OpenBLAS, the reference tests' own oracle) on the same sample, all cores"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# device-side helpers shared by the headline and the extra legs
# ---------------------------------------------------------------------------------------------------------------------------
def make_device_inputs(torch, n, m, side, uplo, dt, dev, seed_a, seed_b, fill_a=True):
    """Scaled recipe generated on the device in column panels (no n x n temporaries at n = 32768).  A column-major n x n, B0 column-major."""
    A = torch.empty((n, n), dtype=dt, device=dev).t()
    if fill_a:
        g = torch.Generator(device=dev).manual_seed(seed_a)
        blk = 4096
        for c0 in range(0, n, blk):
            c1 = min(n, c0 + blk)
            P = ((2 * torch.rand(n, c1 - c0, dtype=torch.float32 if dt != torch.float64 else torch.float64, device=dev, generator=g) - 1) / n ** 0.5)
            P = torch.tril(P, -c0 - 1) if uplo == "L" else torch.triu(P, -c0 + 1)
            A[:, c0:c1].copy_(P.to(dt))
            del P
        A.diagonal().copy_((1 + torch.rand(n, dtype=torch.float64, device=dev, generator=g)).to(dt))
    else:
        A.zero_()
    gb = torch.Generator(device=dev).manual_seed(seed_b)
    shape = (n, m) if side == "L" else (m, n)
    B0 = torch.empty((shape[1], shape[0]), dtype=dt, device=dev).t()
    blk = 4096
    for c0 in range(0, shape[1], blk):
        c1 = min(shape[1], c0 + blk)
        B0[:, c0:c1].copy_((torch.rand(shape[0], c1 - c0, dtype=torch.float32 if dt != torch.float64 else torch.float64, device=dev, generator=gb) + 1).to(dt))
    return A, B0


def backward_error(torch, side, uplo, trans, alpha, func, A, B0, X, blk=4096):
    """Normwise backward error in FP64 with an independent cuBLAS DGEMM product, panel by panel."""
    n = A.shape[0]
    Xd, Bd = X.double(), B0.double()
    V = Xd if func == "S" else Bd
    R = torch.zeros_like(Bd)                     # op(A) * V   (side R: V * op(A))
    nA2 = 0.0
    for c0 in range(0, n, blk):
        c1 = min(n, c0 + blk)
        P = A[:, c0:c1].double()
        P = torch.tril(P, -c0) if uplo == "L" else torch.triu(P, -c0)
        nA2 += float((P * P).sum())
        if side == "L":
            if trans == "N":
                R += P @ V[c0:c1, :]
            else:
                R[c0:c1, :] += P.t() @ V
        else:
            if trans == "N":
                R[:, c0:c1] += V @ P
            else:
                R += V[:, c0:c1] @ P.t()
        del P
    nA = nA2 ** 0.5
    if func == "S":
        return (torch.linalg.norm(R - alpha * Bd) / (nA * torch.linalg.norm(Xd) + abs(alpha) * torch.linalg.norm(Bd))).item()
    return (torch.linalg.norm(Xd - alpha * R) / (abs(alpha) * nA * torch.linalg.norm(Bd))).item()


def timed_steps(torch, dist, world, step, restore, steps, warmup, dev, min_warm_s=0.0):
    """W warm-up steps (and at least `min_warm_s` seconds of them: a leg of a few milliseconds per step would otherwise be timed
    while the GPU is still ramping its clocks up from the idle CPU leg before it), then K steps each bracketed by a CUDA-event pair
    on the launching stream (restore outside the pair); returns total device milliseconds, max over ranks."""
    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    t0 = time.perf_counter()
    done = 0
    while True:
        restore()
        step()
        done += 1
        if done >= warmup:
            torch.cuda.synchronize()
            flag = torch.tensor([1.0 if time.perf_counter() - t0 < min_warm_s else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)   # all ranks take the same number of warm-up steps (collectives inside step)
            if flag.item() == 0.0:
                break
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync_all()
    for k in range(steps):
        restore()
        ev[k][0].record()
        step()
        ev[k][1].record()
    sync_all()
    total = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    return total.item()


def run_lu_leg(torch, nla, dev, n, steps, warmup):
    """SURVEY.md 8(f2), not part of BASELINE.json's metric: the reference's recursive LU (getrf2!, src/lu.jl:185-299) through nla_getrf2 on
    a random Float64 matrix resident in HBM; 2/3 n^3 flops; residual ||P A - L U||_F / ||A||_F checked once with an FP64 product."""
    import numpy as np

    g = torch.Generator(device=dev).manual_seed(77)
    A0 = (torch.rand(n, n, dtype=torch.float64, device=dev, generator=g) - 0.5).t()
    A = torch.empty_like(A0, memory_format=torch.preserve_format)
    ipiv = torch.empty(n, dtype=torch.int64, device=dev)
    info = torch.zeros((), dtype=torch.int32, device=dev)
    times = []
    for it in range(warmup + steps):
        A.copy_(A0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nla.getrf2(A, ipiv, info)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    perm = np.arange(n)
    for i, p in enumerate((ipiv - 1).cpu().numpy()):
        if p != i:
            perm[i], perm[p] = perm[p], perm[i]
    R = A0[torch.from_numpy(perm).to(dev)]
    L = torch.tril(A, -1)
    L.diagonal().fill_(1.0)
    R -= L @ torch.triu(A)
    res = float(torch.linalg.norm(R) / torch.linalg.norm(A0))
    del R, L
    flops = 2.0 / 3.0 * float(n) ** 3
    out = {"config": f"LU: Float64 getrf2! (recursive, partial pivoting) n = {n}, whole factorisation on the device (nla_getrf2)",
           "note": "SURVEY.md 8(f2) row, outside BASELINE.json's metric", "dtype": "f64", "value": flops / ms / 1e9, "unit": "TFLOP/s",
           "ms_per_step": ms, "steps": steps, "info": int(info.item()), "residual": res, "tolerance": 1e-12}
    assert res < 1e-12 and out["info"] == 0, out
    del A, A0
    torch.cuda.empty_cache()
    return out


def run_extra_leg(torch, dist, nla, sharded, h, world, rank, dev, name, dts, side, uplo, trans, func, n, m_local, steps, warmup, peaks, note):
    """One extra leg: synthetic inputs on the device, A owned by rank 0 and broadcast (pipelined) inside every step when world > 1."""
    dt = getattr(torch, dts)
    A, B0 = make_device_inputs(torch, n, m_local, side, uplo, dt, dev, seed_a=4321 + len(name), seed_b=999 + rank, fill_a=(rank == 0))
    X = torch.empty_like(B0)
    launches0 = h.launch_count(reset=True)

    def step():
        if world > 1:
            sharded.unified_rectrxm_pipelined(side, uplo, trans, 1.0, func, A, X, src=0, panels=8, handle=h)
        else:
            nla.unified_rectrxm(side, uplo, trans, 1.0, func, A, X, handle=h)

    # two regimes, both with W warm-up + K timed steps: right after a short warm-up ("burst": comparable with MEASURED_PEAKS' best-of-10
    # cuBLAS figure) and after half a second of back-to-back steps ("sustained": the power-limited clock the GPU settles at under
    # tensor-core load, comparable with MEASURED_PEAKS' sustained figure).  `value` is the burst one.
    ms = timed_steps(torch, dist, world, step, lambda: X.copy_(B0), steps, warmup, dev)
    ms_sus = timed_steps(torch, dist, world, step, lambda: X.copy_(B0), steps, warmup, dev, min_warm_s=0.5)
    flops = float(n) * n * m_local * world
    val = steps * flops / (ms * 1e-3) * 1e-12
    err = torch.tensor([backward_error(torch, side, uplo, trans, 1.0, func, A, B0, X)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    err = err.item()
    per_gpu = val / world
    out = {"config": name, "dtype": dts, "call": f"{side}/{uplo}/{trans}/{func}", "n": n, "rhs_per_gpu": m_local, "rhs_total": m_local * world,
           "value": val, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "backward_error": err, "tolerance": TOL[dts],
           "within_tolerance": bool(err < TOL[dts]), "gpu_launches": h.launch_count() if launches0 is not None else None, "note": note,
           "sustained": {"value": steps * flops / (ms_sus * 1e-3) * 1e-12, "ms_per_step": ms_sus / steps,
                         "how": "same K timed steps after >= 0.5 s of back-to-back warm-up steps (power-limited clocks)"}}
    if dts == "float64":
        out["frac_of_fp64_dmma_peak"] = per_gpu / peaks["fp64"]
    elif dts == "float32":
        out["frac_of_nominal_tf32_third"] = per_gpu / (1125.0 / 3)
        out["frac_of_measured_bf16_third"] = per_gpu / (peaks["bf16"] / 2 / 3) if peaks.get("bf16") else None
        out["fp32_roofline_note"] = "3xTF32 (three tcgen05 kind::tf32 passes per product: the 1e-5 tolerance rules out single-pass TF32); denominators: nominal dense TF32 1125 TFLOP/s / 3, and measured bf16 / 2 / 3"
    else:
        out["frac_of_measured_bf16_burst"] = per_gpu / peaks["bf16"] if peaks.get("bf16") else None
        out["sustained"]["frac_of_measured_bf16_sustained"] = out["sustained"]["value"] / world / peaks["bf16_sustained"] if peaks.get("bf16_sustained") else None
        out["frac_of_nominal_fp16"] = per_gpu / 2250.0
    del A, B0, X
    torch.cuda.empty_cache()
    return out


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
    if os.environ.get("NLA_GATED_MACRO"):
        h.set_option("gated_macro", int(os.environ["NLA_GATED_MACRO"]))

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

    # ---- FP64 tensor-core peak measured in THIS run on THIS GPU (DMMA.8x8x4 issue-rate microbenchmark inside the library) ----
    live_peak = None
    try:
        live_peak = h.probe_fp64_peak()
    except Exception:  # noqa: BLE001
        live_peak = None
    peak = max(FP64_PEAK_TFLOPS, live_peak or 0.0)

    # ---- roofline of the dominant kernel: algorithmic flops / launch time, from the per-launch events ----
    # With enough right-hand sides the whole solve is ONE launch of the row-split fused slab kernel (slab2_f64_kernel: left-looking
    # DMMA main loop + in-register triangular phase); otherwise the GEMM updates of the recursion dominate.
    gemm = [(f, ms) for k, f, ms in prof if k == 1]
    leafs = [(f, ms) for k, f, ms in prof if k == 0]
    g_fl, g_ms = sum(f for f, _ in gemm), sum(ms for _, ms in gemm)
    l_fl, l_ms = sum(f for f, _ in leafs), sum(ms for _, ms in leafs)
    slab_dominant = l_ms >= g_ms
    d_fl, d_ms, d_n = (l_fl, l_ms, len(leafs)) if slab_dominant else (g_fl, g_ms, len(gemm))
    top = max(gemm, key=lambda r: r[0]) if gemm else None
    achieved = d_fl / (d_ms * 1e-3) * 1e-12 if d_ms > 0 else 0.0
    tkey = "slab2_f64_kernel_whole_solve" if slab_dominant else "gemm_f64_tma_kernel_top_level"
    traffic, traffic_src = None, "no ncu capture committed"
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed `ncu --set full` capture (per launch)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic, traffic_src = tj[tkey]["dram_bytes"], tj[tkey]["source"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "tensor",
                "kernel": ("slab2_f64_kernel (row-split fused slab: the whole left-lower solve in one launch, TMA-fed DMMA main loop + in-register substitution)"
                           if slab_dominant else "gemm_f64_tma_kernel (FP64 DMMA update, recursion levels above the fused-slab cutoff)"),
                "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_flops_per_launch": d_fl / max(1, d_n),
                "algorithmic_bytes_note": ("per launch: the lower triangle of A once (1.07 GB) + B read and written once (2 x 2.15 GB) = 5.4 GB; the left-looking kernel "
                                           "re-reads the solved rows of its own vectors for every block row (L2 / DRAM traffic in `traffic`)"
                                           if slab_dominant else "per launch of the top-level update (M = K = 8192, N = 16384): 0.54 GB of A + 1.07 GB of X + 2 x 1.07 GB of the updated block"),
                "peak_source": "FP64 DMMA.8x8x4 issue-rate peak: max(probe constant 37.0 of profiles/r01_probe_dmma_peak.txt, the same microbenchmark run live in this "
                               "process through nla_probe_fp64_peak); MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2; cuBLAS DGEMM measured 35.4",
                "peak_measured_this_run": live_peak,
                "launches_per_step": d_n // prof_steps, "avg_launch_ms": d_ms / max(1, d_n),
                "flops_per_step": d_fl / prof_steps,
                "how": "algorithmic flops of the dominant kernel's launches / their CUDA-event durations, 2 profiled steps run right after the timed region on one stream",
                "top_level_gemm_launch": ({"flops": top[0], "ms": top[1], "tflops": top[0] / (top[1] * 1e-3) * 1e-12} if top else None),
                "gemm_share_of_step": g_ms / max(1e-9, g_ms + l_ms), "leaf_share_of_step": l_ms / max(1e-9, g_ms + l_ms),
                "gemm_tflops": g_fl / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else None,
                "fused_slab_launches_per_step": len(leafs) // prof_steps,
                "leaf_tflops": l_fl / (l_ms * 1e-3) * 1e-12 if l_ms > 0 else None,
                "whole_step_frac_of_peak": value / world / peak}

    # ---- e2e: the same call with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        es = 8
        e2e_steps = max(2, min(args.steps, 5))
        lib = nla.load_library()
        if world == 1:
            hostA = torch.empty((n, n), dtype=dt, pin_memory=True)
            hostA.copy_(A_store)
            hostB = torch.empty((m, n), dtype=dt, pin_memory=True)
            hostB.copy_(B0.t())
            hostX = torch.empty((m, n), dtype=dt, pin_memory=True)

            def e2e_step():
                hostX.copy_(hostB)  # restore (host side, outside the timed call)
                t0 = time.perf_counter()
                rc = lib.nla_rectrxm_host(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, hostA.data_ptr(), n, hostX.data_ptr(), n)
                assert rc == 0, rc
                return time.perf_counter() - t0

            api = "nla_rectrxm_host (pinned host A and B in, B out)"
            extra_e2e = {}
        else:
            # A lives in host memory every rank can read (POSIX shared memory, page-locked by every rank): rank r uploads the column panels
            # p = r (mod N) through ITS OWN PCIe link and is the NCCL root for them, so the upload of A is spread over all links instead of
            # funnelled through rank 0 next to its own B; every rank streams its own B through the library's host pipeline
            hp = sharded.HostSharedMatrix(n, dt, rank, local, world, name="nla_bench_A")
            if rank == 0:
                hp.tensor.copy_(A_store)
            dist.barrier()
            sharded.bind_numa_local(local)
            hostB = torch.empty((m, n), dtype=dt, pin_memory=True)
            hostB.copy_(B0.t())
            hostX = torch.empty((m, n), dtype=dt, pin_memory=True)
            A.fill_(float("nan"))   # the device copy is rebuilt from the host inside every e2e step

            def e2e_step():
                hostX.copy_(hostB)
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                sharded.unified_rectrxm_pipelined_host("L", "L", "N", 1.0, "S", A, hp.tensor.t(), hostX.t(), src=None, panels=8, handle=h)
                torch.cuda.synchronize()
                return time.perf_counter() - t0

            api = ("sharded.unified_rectrxm_pipelined_host: A in shared pinned host memory, rank r uploads panels p = r mod N over its own PCIe link and "
                   "is their NCCL root (ncclBroadcast per panel, consumption order) + nla_rectrxm_hostb_gated (B streamed in chunks, launches gated on the panels of A)")
            extra_e2e = {"numa": sharded.numa_report(local)}

        # the box's own ceiling for this traffic pattern: every rank at once copies its 2 GiB of B host -> device and 2 GiB device -> host
        # (pinned buffers, two streams, nothing else running); e2e can never beat max(compute, bytes / this)
        link = None
        try:
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            best = 1e30
            for _ in range(2):
                sync_all()
                t0 = time.perf_counter()
                with torch.cuda.stream(s_up):
                    X.t().copy_(hostB, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    hostX.copy_(B0.t(), non_blocking=True)
                torch.cuda.synchronize()
                dtw = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(dtw, op=dist.ReduceOp.MAX)
                best = min(best, dtw.item())
            nb = n * m * es
            link = {"concurrent_h2d_plus_d2h_seconds": best, "per_rank_gbs_each_direction": nb / best * 1e-9,
                    "aggregate_gbs_both_directions": 2 * nb * world / best * 1e-9,
                    "what": f"all {world} ranks at once: {nb / 2**30:.0f} GiB pinned host -> device and {nb / 2**30:.0f} GiB device -> pinned host per rank on two streams"}
        except Exception as e:  # noqa: BLE001
            link = {"error": repr(e)}

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
        sec = tt / e2e_steps
        e2e = {"value": e2e_steps * flops_per_step_all / tt * 1e-12, "unit": UNIT,
               "h2d_bytes_per_step": a_bytes + n * m * world * es, "d2h_bytes_per_step": n * m * world * es, "steps": e2e_steps,
               "ms_per_step": sec * 1e3, "api": api,
               "h2d_gbs_per_rank": (a_bytes / world + n * m * es) / sec * 1e-9, "d2h_gbs_per_rank": n * m * es / sec * 1e-9,
               "host_buffers": "pinned (cudaHostAlloc / cudaHostRegister)", "host_link_ceiling": link}
        if link and "concurrent_h2d_plus_d2h_seconds" in link:
            # lower bound of any end-to-end step on this box: the B traffic alone at the measured ceiling, or the device-resident step
            e2e["floor_ms_per_step"] = max(link["concurrent_h2d_plus_d2h_seconds"] * 1e3, total_ms / args.steps)
            e2e["limiter"] = ("host link: the step moves %.1f GB through one host at %.0f GB/s aggregate (measured ceiling of this box for this pattern)"
                              % ((e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) * 1e-9, link["aggregate_gbs_both_directions"])
                              if link["concurrent_h2d_plus_d2h_seconds"] * 1e3 > total_ms / args.steps else "compute: transfers hide behind the solve")
        e2e.update(extra_e2e)
        # check the e2e result too
        Xh = hostX.to(dev).t()
        if world > 1:
            torch.cuda.synchronize()
        R = torch.tril(torch.nan_to_num(A)) @ Xh - B0
        e2e["backward_error"] = (torch.linalg.norm(R) / (torch.linalg.norm(torch.tril(torch.nan_to_num(A))) * torch.linalg.norm(Xh) + torch.linalg.norm(B0))).item()
        del R, Xh
        # pageable host buffers (what a plain Julia Array is): same call, one step, reported beside the pinned number
        if world == 1:
            try:
                pA = np.asfortranarray(A.cpu().numpy())
                pX = np.asfortranarray(B0.cpu().numpy())
                t0 = time.perf_counter()
                rc = lib.nla_rectrxm_host(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, pA.ctypes.data, n, pX.ctypes.data, n)
                tp = time.perf_counter() - t0
                assert rc == 0, rc
                e2e["pageable"] = {"value": flops_per_step_all / tp * 1e-12, "ms_per_step": tp * 1e3,
                                   "note": "same nla_rectrxm_host call on pageable (malloc) host arrays, one step; the driver stages pageable copies itself"}
                del pA, pX
            except Exception as e:  # noqa: BLE001
                e2e["pageable"] = {"error": str(e)}
        del hostB, hostX

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cn, cm = n, 2048
        sec, cores = cpu_port_time(cn, cm)
        cpu = {"value": float(cn) * cn * cm / sec * 1e-12, "unit": UNIT, "cores": cores, "kind": "port", "seconds": sec,
               "sample": f"oracle/nla_oracle.c (C/OpenMP restatement of the reference algorithm; Julia unavailable), FP64 L/L/N TRSM with the headline order n = {cn} "
                         f"and {cm} of the {m} right-hand sides (1/8 of the headline flops, same recursion depth and GEMM K extents), {cores} OpenMP threads"}
        try:
            cpu["openblas_dtrsm_tflops"] = float(cn) * cn * cm / openblas_time(cn, cm) * 1e-12
            cpu["openblas_note"] = "scipy.linalg.blas.dtrsm (OpenBLAS, the reference tests' own oracle) on the same sample, all cores"
        except Exception as e:  # noqa: BLE001
            cpu["openblas_dtrsm_tflops"] = f"unavailable: {e}"
        _cpu_inputs.clear()

    # ---- the other legs of the metric (fp32 / fp16, C3 / C4 / C5) ----
    del A, B0, X, A_store
    torch.cuda.empty_cache()
    extra = None
    if not args.no_extra and n == N_ORDER:
        peaks = {"fp64": peak}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peaks["bf16"], peaks["bf16_sustained"] = mp.get("bf16_tflops"), mp.get("bf16_tflops_sustained")
        except Exception:  # noqa: BLE001
            peaks["bf16"], peaks["bf16_sustained"] = 1590.0, None   # B200_PROFILING.md fallback
        xs, xw = max(3, min(args.steps, 5)), 3
        legs = []
        if world == 1:
            legs.append(("C3: Float32 left/upper/transposed TRMM n = m = 16384", "float32", "L", "U", "T", "M", 16384, 16384, "BASELINE configs[2]"))
            legs.append(("Float32 left/lower TRSM n = m = 16384", "float32", "L", "L", "N", "S", 16384, 16384, "the metric's fp32 leg on the headline shape"))
            legs.append(("Float16 left/lower TRSM n = m = 16384 (FP32 accumulate)", "float16", "L", "L", "N", "S", 16384, 16384, "the metric's fp16 leg on the headline shape"))
            legs.append(("C4 slice: Float16 right/lower TRSM n = 32768, 16384 RHS rows", "float16", "R", "L", "N", "S", 32768, 16384,
                         "BASELINE configs[3] is defined on 2/4/8 GPUs (131072 RHS rows in total); this is one GPU's share at 8 GPUs"))
        else:
            legs.append((f"C4: Float16 right/lower TRSM n = 32768, 131072 RHS rows sharded over {world} GPUs", "float16", "R", "L", "N", "S", 32768, 131072 // world,
                         "BASELINE configs[3]; strong scaling in the number of GPUs; A broadcast (pipelined) inside every step"))
        legs.append((f"C5: Float64 left/lower TRMM n = 32768, 8192 RHS per GPU x {world}", "float64", "L", "L", "N", "M", 32768, 8192,
                     "BASELINE configs[4] (side/uplo/trans unspecified there: left/lower/no-trans); weak scaling; A broadcast (pipelined) inside every step"))
        extra = []
        for (name, dts, sd, up, tr, fn, nn, mm, note) in legs:
            try:
                extra.append(run_extra_leg(torch, dist, nla, sharded, h, world, rank, dev, name, dts, sd, up, tr, fn, nn, mm, xs, xw, peaks, note))
            except Exception as e:  # noqa: BLE001
                extra.append({"config": name, "error": repr(e)})
                torch.cuda.empty_cache()

        if world == 1:
            try:
                extra.append(run_lu_leg(torch, nla, dev, 16384, xs, xw))
            except Exception as e:  # noqa: BLE001
                extra.append({"config": "LU: Float64 getrf2! n = 16384", "error": repr(e)})
                torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n, m, world, not args.no_pipeline),
                "options": {"streams": args.streams or "auto", "leaf": h.get_option("leaf"), "macro": h.get_option("macro")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "backward_error": berr, "tolerance": 1e-13, "wall_ms_per_step_incl_restore": t_wall / args.steps * 1e3,
                "pct_of_fp64_peak": 100.0 * value / world / peak, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    assert berr < 1e-13, f"backward error {berr} above tolerance"


if __name__ == "__main__":
    main()
