#!/usr/bin/env python
"""bench.py - "assembled DoFs/s + CG solve s (SWIPDG p1, 16M cells)" on N B200s of one node.

A step is one pass of the hot path over the workload: assemble every affine part of the system matrix and of the rhs
(K2 + K3), then freeze and CG-solve to ||r||/||b|| <= 1e-10 (K4 - K6), on the 4096^2 structured Q1 grid with ESV2007 data
(BASELINE.json configs[4], the configuration the metric is quoted on; --grid scales it down for a quick look).

    value          assembled DoFs/s over all ranks: K * N_dofs / sum of the K assembly times (device time, max over ranks)
    cg_solve_s     mean CG time-to-solution of the K steps (device time), with cg_iterations and cg_s_per_iteration;
                   default solver "cg.mg" (CG with the two-level multigrid preconditioner); cg_diagonal holds the plain
                   Jacobi-CG solve of the same system, run once after the timed steps
    ms_per_step    whole step (assembly + solve), wall clock between barriers / K
    e2e            the same through the public API from HOST buffers: grid arrays -> hdd_mesh_create (H2D) -> init -> solve
                   -> solution back on the host (D2H), every step
    roofline       the dominant kernel (CG SpMV with fused p.Ap): algorithmic bytes / CUDA-event time / measured HBM peak
    cpu_baseline   the CPU oracle (a port: the reference cannot be built here) on a bounded sample, rank 0, N = 1 only

`--impl reference` times the CPU restatement of the reference path (oracle/, serial walk like the reference) instead.
N > 1: one rank per GPU (torchrun); the 8 x 8 subdomains of BlockSWIPDG are dealt to the ranks in contiguous slabs,
NCCL carries the coupling-face halo of the CG direction and the dot-product all-reduces (strong scaling of one grid).
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

METRIC = "assembled DoFs/s + CG solve s (SWIPDG p1, 16M cells)"
PRECISION = 1e-10


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clocks and throttle reasons during the timed region: through NVML in-process (no fork, no driver
    re-initialisation next to a launch-bound step), falling back to spawning nvidia-smi"""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt, self._was_started = index, [], threading.Event(), False
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK -> physical device: respect CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self._handle)
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        return [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        return [x.strip() for x in out.strip().split(",")]

    def start(self):
        self._was_started = True
        super().start()

    def run(self):
        while not self._halt.is_set():
            try:
                f = self._sample_nvml() if self._nvml else self._sample_smi()
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.05 if self._nvml else 0.2)

    def stop(self):
        self._halt.set()
        if self._was_started:
            self.join(timeout=3)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for k, n in enumerate(self.NAMES) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self._nvml else "nvidia-smi"}


def _cpu_sample(n_cpu, cg_iters, threads):
    """one pass of the oracle over the sample with `threads` host threads -> (pattern s, assembly s, s per CG iteration, iterations, DoFs)"""
    from oracle import oracle as o
    m = o.mesh_cube(n_cpu, n_cpu, -1.0, 1.0, -1.0, 1.0)
    o.set_threads(threads)
    try:
        t0 = time.perf_counter()
        rp, col = o.pattern(m)
        t_pat = time.perf_counter() - t0
        t0 = time.perf_counter()
        A = o.assemble_lhs(m, o.const(1.0), None, rp, col)
        b = o.assemble_rhs(m, o.esv2007_force())
        t_asm = time.perf_counter() - t0
        t0 = time.perf_counter()
        x, it, rr = o.cg(rp, col, A, b, precond=1, rtol=1e-30, maxit=cg_iters)
        t_cg = time.perf_counter() - t0
    finally:
        o.set_threads(1)
    return t_pat, t_asm, t_cg / max(it, 1), it, m.n_dofs


def _cpu_record(n_cpu, sample, cores):
    t_pat, t_asm, t_it, it, n_dofs = sample
    return {"value": n_dofs / t_asm, "unit": "DoFs/s", "cores": cores, "kind": "port",
            "sample": "%dx%d Q1 cells (ESV2007 data): assembly walk + %d Jacobi-CG iterations, %d host thread%s"
                      % (n_cpu, n_cpu, it, cores, "" if cores == 1 else "s"),
            "assemble_s": t_asm, "pattern_s": t_pat, "cg_s_per_iteration": t_it,
            "note": "CPU restatement of the reference (oracle/); the reference itself needs un-vendored DUNE modules"}


def cpu_baseline(n_cpu, cg_iters, all_cores=True):
    """oracle (port of the reference's serial walk) on a bounded sample: n_cpu^2 Q1 cells, cg_iters CG iterations.
    `value` is the faithful one-core number (the reference's walk is serial, discretizations/swipdg.hh:485); all_cores
    repeats the sample with the oracle's optional host threading (SURVEY 8d asks for both)."""
    out = _cpu_record(n_cpu, _cpu_sample(n_cpu, cg_iters, 1), 1)
    if all_cores:
        cores = os.cpu_count() or 1
        rec = _cpu_record(n_cpu, _cpu_sample(n_cpu, cg_iters, cores), cores)
        out["all_cores"] = {"cores": cores, "value": rec["value"], "assemble_s": rec["assemble_s"],
                            "cg_s_per_iteration": rec["cg_s_per_iteration"],
                            "note": "same sample with the oracle's own threading (atomic scatter); the reference has none"}
    return out


def run_reference(args):
    """The CPU arm: the oracle port with all the host threads it can use (its own threading; the reference's walk is
    serial, the one-core figure is reported next to it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = args.cpu_n
    cores = os.cpu_count() or 1
    steps = []
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rec = _cpu_record(n_cpu, _cpu_sample(n_cpu, args.cpu_cg_iters, cores), cores)
        steps.append((time.perf_counter() - t0, rec))
    timed = steps[args.warmup:]
    cb = timed[-1][1]
    value = float(np.mean([s[1]["value"] for s in timed]))
    cb["value"] = value
    one = _cpu_record(n_cpu, _cpu_sample(n_cpu, args.cpu_cg_iters, 1), 1)  # outside the timed steps
    cb["one_core"] = {"cores": 1, "value": one["value"], "assemble_s": one["assemble_s"],
                      "cg_s_per_iteration": one["cg_s_per_iteration"],
                      "note": "the same sample as the reference runs it: serial walk (discretizations/swipdg.hh:485)"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "DoFs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([s[0] for s in timed])),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: SWIPDG p1 (Q1), structured [-1,1]^2 grid, ESV2007 data; CPU sample %dx%d cells"
                       % (n_cpu, n_cpu), "cells": n_cpu * n_cpu},
            "cg_s_per_iteration": float(np.mean([s[1]["cg_s_per_iteration"] for s in timed])),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", dest="n", type=int, default=4096, help="cells per side of the structured grid")
    ap.add_argument("--cpu-n", type=int, default=768)
    ap.add_argument("--cpu-cg-iters", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--solver", default="cg.mg", help="cg.mg | cg.blockdiagonal | cg.diagonal | cg.identity")
    ap.add_argument("--no-jacobi", action="store_true", help="skip the extra Jacobi-CG solve reported next to --solver")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import dune_hdd_b200 as hdd
    from dune_hdd_b200 import capi
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = hdd.parallel.init_comm(rank, world, local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    n = args.n
    parts = (8, 8) if n % 8 == 0 else (1, 1)
    t0 = time.perf_counter()
    grid = hdd.grids.cube(n, partitions=parts, pinned=True)  # page-locked host arrays (hdd_host_alloc)
    t_grid = time.perf_counter() - t0
    problem = hdd.problems.ESV2007()
    roff = hdd.parallel.rank_cell_offsets(grid, world)
    cell_range = (int(roff[rank]), int(roff[rank + 1]))
    options = {"type": args.solver, "precision": PRECISION, "max_iter": 200000}

    def make():
        d = hdd.BlockSWIPDG(grid, problem, device=local_rank, cell_range=cell_range, comm=comm)
        return d

    d = make()
    d.init()
    n_dofs = d.num_dofs()
    launches0 = capi.kernel_launches()

    def step():
        ta = d.assemble()
        _, info = d.uncached_solve(options, return_info=True, copy_to_host=False)
        return ta, info

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches1 = capi.kernel_launches()
    t0 = time.perf_counter()
    results = [step() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    launches = capi.kernel_launches() - launches1
    clocks = sampler.stop() if rank == 0 else None

    t_asm = sum(r[0] for r in results)
    t_cg = sum(r[1]["seconds"] for r in results)
    iters = results[-1][1]["iterations"]
    stats = torch.tensor([t_asm, t_cg, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(stats, op=torch.distributed.ReduceOp.MAX)
    t_asm, t_cg, wall = [float(v) for v in stats.cpu()]

    # the plain Jacobi-preconditioned CG on the same system, once, outside the timed steps: its SpMV / update / direction
    # kernels are the ones the roofline numbers below are taken from
    jacobi = None
    if not args.no_jacobi or args.solver != "cg.diagonal":
        jopt = dict(options, type="cg.diagonal")
        if args.no_jacobi:
            jopt["max_iter"] = 50
        try:
            _, ji = d.uncached_solve(jopt, return_info=True, copy_to_host=False)
        except hdd.discretizations.linear_solver_failed:
            ji = None
        if ji is not None:
            js = torch.tensor([ji["seconds"]], dtype=torch.float64, device="cuda")
            if world > 1:
                torch.distributed.all_reduce(js, op=torch.distributed.ReduceOp.MAX)
            jacobi = {"solve_s": float(js.cpu()[0]), "iterations": ji["iterations"],
                      "s_per_iteration": float(js.cpu()[0]) / max(ji["iterations"], 1),
                      "halo": ("peer-memory SpMV (CUDA IPC over NVLink)" if ji.get("peer_memory") else "NCCL send/recv")
                      if world > 1 else "none"}

    # roofline of the dominant kernel: CG SpMV, timed alone with CUDA events on the library's stream
    L = capi.lib()
    roof = {}
    peak, peak_kind = measured_peak_gbs()
    # DRAM traffic per launch from the committed `ncu --set full` capture of this workload (profiles/), if it matches
    traffic = {}
    for fname, kmap in (("r01c_traffic_n4096.json", {"cg_update": "k_cg_update", "cg_direction": "k_cg_direction"}),
                        ("r01f_traffic_n4096.json", {"spmv": "k_cg_spmv_tma", "assembly": "k_assemble_lhs_cube<0, 4, 1>"})):
        try:
            with open(os.path.join(ROOT, "profiles", fname)) as f:
                tj = json.load(f)
            if tj.get("cells") == n * n and world == 1:
                traffic.update({k: tj["kernels"][v]["dram_bytes"] for k, v in kmap.items() if v in tj["kernels"]})
        except Exception:
            pass
    for which, name in ((0, "spmv"), (1, "cg_update"), (2, "cg_direction"), (3, "assembly")):
        sec, byt = C.c_double(), C.c_double()
        capi.check(L.hdd_profile_kernel(d._h, which, 20 if which != 3 else 5, C.byref(sec)))
        capi.check(L.hdd_kernel_bytes(d._h, which, C.byref(byt)))
        roof[name] = {"bound": "hbm", "achieved": byt.value / sec.value / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": byt.value / sec.value / 1e9 / peak, "traffic": traffic.get(name), "ms": sec.value * 1e3,
                      "algorithmic_bytes": byt.value, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)"}

    # end to end through the public API from host buffers (H2D of the grid + problem, D2H of the solution), every step
    e2e = None
    if not args.no_e2e:
        del d
        torch.cuda.empty_cache()
        e_asm, e_cg = [], []
        own_dofs = (cell_range[1] - cell_range[0]) * 4
        x_host = capi.pinned_empty((own_dofs,), np.float64)  # page-locked result buffer, reused by every step
        for k in range(1 + args.steps):  # one warm-up
            barrier()
            t0 = time.perf_counter()
            d2 = make()
            d2.init()
            d2._sync = capi.check(L.hdd_sync(d2._h))
            t1 = time.perf_counter()
            u, info = d2.uncached_solve(options, return_info=True, copy_to_host=True, out=x_host)
            t2 = time.perf_counter()
            del d2
            if k > 0:
                e_asm.append(t1 - t0)
                e_cg.append(t2 - t1)
        ev = torch.tensor([sum(e_asm), sum(e_cg)], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(ev, op=torch.distributed.ReduceOp.MAX)
        ea, ec = [float(v) for v in ev.cpu()]
        own_cells = cell_range[1] - cell_range[0]
        h2d = grid.xy.nbytes + grid.cell_verts.nbytes + grid.cell_neigh.nbytes + grid.cell_subdomain.nbytes
        e2e = {"value": args.steps * n_dofs / ea, "unit": "DoFs/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(own_cells * 4 * 8), "cg_solve_s": ec / args.steps,
               "setup_s": ea / args.steps,
               "note": "value = DoFs / (hdd_mesh_create from page-locked host arrays + hdd_swipdg_create + init), incl. "
                       "host-side localisation of the grid; cg_solve_s includes the D2H copy of the solution"}

    if rank == 0:
        line = {"metric": METRIC, "value": args.steps * n_dofs / t_asm, "unit": "DoFs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config5: SWIPDG p1 (Q1) on the %dx%d structured grid [-1,1]^2, ESV2007 data, "
                                       "8x8 BlockSWIPDG partition" % (n, n), "cells": n * n, "dofs": n_dofs,
                           "cg": "%s to ||r||/||b|| <= 1e-10" % args.solver, "l2_flush": "inputs >> L2 (matrix "
                           "%.1f GB per part)" % (8.0 * 16 * (n * n + 2 * 2 * n * (n - 1)) / 1e9),
                           "parallelism": "subdomain slabs x%d" % world,
                           "halo": ("peer-memory SpMV (CUDA IPC over NVLink)" if results[-1][1].get("peer_memory") else
                                    "NCCL send/recv of the CG direction" + (", all-reduce of the restricted residual (cg.mg)"
                                                                           if args.solver == "cg.mg" else ""))
                           if world > 1 else "none"},
                "assemble_ms": 1e3 * t_asm / args.steps, "cg_solve_s": t_cg / args.steps, "cg_iterations": iters,
                "cg_s_per_iteration": t_cg / args.steps / max(iters, 1), "cg_diagonal": jacobi,
                "roofline": roof["spmv"], "roofline_assembly": roof["assembly"], "roofline_cg_update": roof["cg_update"],
                "roofline_cg_direction": roof["cg_direction"],
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "grid_generation_s": t_grid}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_n, args.cpu_cg_iters)
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
