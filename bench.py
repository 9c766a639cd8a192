#!/usr/bin/env python
"""bench.py -- BP4 merged-CG throughput of the B200-native hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--degree P] [--s S] [--solver merged|plain]
  python bench.py --impl reference ...      # the reference algorithm's CPU restatement on host cores

A "step" is one call of the reference's `run_cg_solver` plugin (benchmark.h:193: x0 = 0,
ReductionControl(100, 1e-15, 1e-8)) on the synthetic BP4 problem; the metric is the
reference's own `dofs/s/it` column (benchmark.h:222): n_dofs * n_iterations / solver_time.

  value     device-timed (CUDA events on the solver's stream), right-hand side resident in HBM
  e2e       same call with HOST buffers: b uploaded and x downloaded inside the timed region
  roofline  merged cell kernel: algorithmic bytes (58.67 B/DoF + 300 B/cell) over the
            CUDA-event duration of its launches in the timed region, vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (C/OpenMP restatement of the reference algorithm, kind "port":
            the real reference needs deal.II + p4est + MPI, absent here) on a bounded sample
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BP4 CG GDoF/s (DoFs x iterations / s)"
UNIT = "GDoF/s"


def n_dofs_of(p, s):
    nr, rem = divmod(s, 3)
    n = [((2 if d < rem else 1) << nr) * p + 1 for d in range(3)]
    return 3 * n[0] * n[1] * n[2], (1 << s)


def algorithmic_bytes_per_iteration(p, s, merged=True):
    """merged = 7.333 doubles/DoF, plain = 17.333 doubles/DoF, + 300 B/cell metadata"""
    nd, nc = n_dofs_of(p, s)
    return (58.0 + 2.0 / 3.0 if merged else 138.0 + 2.0 / 3.0) * nd + 300.0 * nc


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def cpu_oracle_run(p, s, merged, steps, warmup, max_its=100):
    """time the CPU oracle's CG (C/OpenMP restatement of the reference algorithm) on host cores"""
    from oracle import bp4_oracle as O
    from oracle import c_oracle
    c_oracle.build()
    rd = O.build_problem(p, s)[0]
    co = c_oracle.COracle(rd)
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    for _ in range(warmup):
        co.cg(rd.rhs, prec, merged, max_steps=min(max_its, 5))
    t0 = time.perf_counter()
    its = 0
    for _ in range(steps):
        _, it, _ = co.cg(rd.rhs, prec, merged, max_steps=max_its)
        its += it
    dt = time.perf_counter() - t0
    nd, _ = n_dofs_of(p, s)
    return nd * its / dt * 1e-9, dt / steps, c_oracle.lib().oracle_num_threads(), its // steps


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the real reference cannot be
    built here) on this box's host cores, same metric/config, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = args.degree
    s = min(args.s, args.cpu_s)
    merged = args.solver == "merged"
    val, sec, cores, its = cpu_oracle_run(p, s, merged, max(1, min(args.steps, 2)), min(args.warmup, 1))
    nd, nc = n_dofs_of(p, s)
    sample = f"degree {p}, s={s} ({nc} cells, {nd} DoFs) instead of s={args.s}; {its} CG iterations per step"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
           "steps": max(1, min(args.steps, 2)), "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": workload_config(args, 1),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(args, n_gpus, s_run=None):
    s_run = args.s if s_run is None else s_run
    nd, nc = n_dofs_of(args.degree, s_run)
    plug = "benchmark_precond_merged" if args.solver == "merged" else "benchmark_precond"
    return {"workload": f"BP4 Q{args.degree} (q={args.degree + 2}) vector Laplace, curved mesh s={s_run}: "
                        f"{nc} cells, {nd} DoFs on {n_gpus} GPU(s), {plug} (ReductionControl(100,1e-15,1e-8), "
                        f"Jacobi, RHS i%8), Renumber(0,1,2)",
            "degree": args.degree, "s": s_run, "n_dofs": nd, "n_dofs_per_gpu": nd // n_gpus, "solver": args.solver,
            "parallelism": "single GPU" if n_gpus == 1 else
            f"domain decomposition over {n_gpus} GPUs (contiguous chunks of the cell order), NCCL ghost "
            f"exchange + 7-double all-reduce per iteration; weak scaling s = {args.s} + log2(N)",
            "l2": "vectors (%.0f MB per GPU each) exceed the 126 MB L2; no flush needed" % (nd * 8 / 1e6 / n_gpus)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--s", type=int, default=18)
    ap.add_argument("--cpu-s", type=int, default=17, help="mesh size of the bounded CPU sample")
    ap.add_argument("--solver", default="merged", choices=["merged", "plain"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from mf_data_locality_b200 import capi, host
    import ctypes as C

    # weak scaling: the mesh grows with the GPU count (s + log2 N: twice the cells per doubling,
    # SURVEY 8d config 5), partitioned along the renumbered cell order, one partition per GPU
    s_run = args.s + max(0, world.bit_length() - 1)
    nccl_id = None
    if world > 1:
        idt = torch.tensor(list(capi.unique_id() if rank == 0 else bytes(128)), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, src=0)
        nccl_id = bytes(idt.cpu().tolist())
    prob = host.Problem(args.degree, s_run, plugin=args.solver, device=local_rank, n_ranks=world, rank=rank,
                        nccl_id=nccl_id)
    ctx = C.c_void_p(prob.ctx_handle())
    L = capi.lib()
    stream_ptr = C.c_void_p()
    L.bp4_ctx_stream(ctx, C.byref(stream_ptr))
    stream = torch.cuda.ExternalStream(stream_ptr.value or 0, device=local_rank)
    n_dofs = prob.n_dofs
    kernel_id = capi.K_MERGED if args.solver == "merged" else capi.K_VMULT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------
    for _ in range(args.warmup):
        prob.run_cg_solver(None, want_x=False)
    L.bp4_profile_reset(ctx)
    L.bp4_profile_enable(ctx, 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    iters = 0
    for _ in range(args.steps):
        _, it = prob.run_cg_solver(None, want_x=False)
        iters += it
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    kms, kcnt = C.c_double(), C.c_uint64()
    L.bp4_profile_get(ctx, kernel_id, C.byref(kms), C.byref(kcnt))
    kernel_name = "cell_kernel_merged (fused pre + cells + post)" if kernel_id == capi.K_MERGED else "cell_kernel_plain"
    if kcnt.value == 0:   # merged solver running its three-kernel variant: the cell kernel dominates
        kernel_id, kernel_name = capi.K_VMULT, "cell_kernel_plain (merged CG = pre + cell + post kernels)"
        L.bp4_profile_get(ctx, kernel_id, C.byref(kms), C.byref(kcnt))
    parts = {}
    for nm, kid in (("cells", capi.K_VMULT), ("merged", capi.K_MERGED), ("pre", capi.K_PRE), ("post", capi.K_POST), ("blas1", capi.K_BLAS1)):
        a_, b_ = C.c_double(), C.c_uint64()
        L.bp4_profile_get(ctx, kid, C.byref(a_), C.byref(b_))
        parts[nm] = {"ms_total": a_.value, "launches": int(b_.value)}
    launches = C.c_uint64()
    L.bp4_launch_count(ctx, C.byref(launches))
    L.bp4_profile_enable(ctx, 0)
    value = n_dofs * iters / (total_ms * 1e-3) * 1e-9      # n_dofs is the global count

    # ---- end to end: host buffers through the plugin call --------------------------------
    b_host = torch.from_numpy(prob.rhs()).pin_memory()
    x_host = torch.empty(prob.n_owned, dtype=torch.float64).pin_memory()
    b_np, x_np = b_host.numpy(), x_host.numpy()
    it_c = C.c_uint()
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        prob._chk(prob.l.bp4h_run_cg_solver(prob.h, b_np.ctypes.data_as(C.c_void_p),
                                           x_np.ctypes.data_as(C.c_void_p), C.byref(it_c)))
        e2e_iters += it_c.value
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = n_dofs * e2e_iters / float(dt.item()) * 1e-9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    merged = args.solver == "merged"
    # per launch = per GPU: this rank's share of the DoFs and cells
    share = prob.n_owned / n_dofs
    fused = kernel_id == capi.K_MERGED
    alg_bytes = (algorithmic_bytes_per_iteration(args.degree, s_run, True) if fused else
                 16.0 * n_dofs + 300.0 * (1 << s_run)) * share
    # whole CG iteration against the same roof: SURVEY 8(d) bytes per iteration over the
    # device time of one iteration (all kernels, exchanges and the host round trip for the sums)
    it_bytes = algorithmic_bytes_per_iteration(args.degree, s_run, merged) * share
    it_ms = total_ms / max(iters, 1)
    avg_ms = kms.value / max(kcnt.value, 1)
    achieved = alg_bytes / (avg_ms * 1e-3) * 1e-9 if avg_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.solver}_q{args.degree}_s{args.s}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": kernel_name, "kernels_in_timed_region": parts,
                "kernel_ms_avg": avg_ms, "kernel_launches": int(kcnt.value),
                "kernel_share_of_step": kms.value / total_ms if total_ms else None,
                "algorithmic_bytes_per_launch": alg_bytes,
                # second roof: the cell kernel is FP64-pipe bound.  336 FP64
                # warp-lane instructions per DoF were counted by ncu at Q4 (533.9 M FP64 warp instructions per apply,
                # profiles/r01_final3_cell_kernel_q4_s18.txt + its source page); the peak is
                # 148 SMs x 64 lanes x sm_max_mhz
                "fp64_pipe": ({"instr_per_dof": 336, "peak_lane_instr_per_s": 148 * 64 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6,
                               "frac": 336.0 * prob.n_owned / (avg_ms * 1e-3) /
                                       (148 * 64 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6)}
                              if args.degree == 4 and avg_ms > 0 and not fused else None),
                "iteration": {"algorithmic_bytes": it_bytes, "ms": it_ms,
                              "achieved": it_bytes / (it_ms * 1e-3) * 1e-9,
                              "frac": it_bytes / (it_ms * 1e-3) * 1e-9 / peak}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            s_cpu = min(args.s, args.cpu_s)
            val, sec, cores, its = cpu_oracle_run(args.degree, s_cpu, merged, 1, 1)
            nd_c, nc_c = n_dofs_of(args.degree, s_cpu)
            cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"CPU oracle (C/OpenMP restatement of the reference algorithm), degree {args.degree}, "
                             f"s={s_cpu} ({nc_c} cells, {nd_c} DoFs), 1 solve of {its} iterations, {sec:.1f} s"}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world, s_run), "iterations_per_step": iters // args.steps,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(prob.n_owned * 8),
                   "d2h_bytes_per_step": int(prob.n_owned * 8)},
           "gpu_launches": int(launches.value), "clocks": clocks, "roofline": roofline,
           "cpu_baseline": cpu, "setup_seconds": prob.setup_seconds}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
