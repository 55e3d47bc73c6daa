#!/usr/bin/env python
"""bench.py -- MG-PCG DoFs/s of the 64k-atom (atom_n20_64000) solve + Gaussian-charge RHS assembly.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one pass of the hot path on the LAST refinement cycle of the reference's 64k-atom cluster run
(`Cluster runs output and postprocessing/SSOR_64k_atoms.o876224`: 5 cycles, final hierarchy
1771561 / 170516 / 14336 level DoFs): charge densities + load vector (RHS) and the GMG-preconditioned CG
solve from the transferred initial guess to 1e-8 |b|.  The earlier cycles run once, untimed, as set-up.
`value` = DoFs / s with every input resident in HBM; `e2e` = the same through the host-buffer entry points
(LaplaceProblem::compute_charge_densities / assemble rhs / solve(): atoms, cells, CSR matrices and vectors
cross PCIe inside the timed region).  `--impl reference` times the oracle's plain-C restatement of the same
path on the host cores (the reference itself needs deal.II + Trilinos + p4est + MPI and cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "geometric-multigrid-preconditioners-for-long-range-coulomb-interaction_b200"
METRIC = "mg_pcg_dofs_per_s"
UNIT = "DoF/s"


def pkg():
    import importlib
    return importlib.import_module(PKG)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--atoms-n", type=int, default=20, help="NaCl lattice of n^3 unit cells = 8 n^3 atoms (20 -> 64000)")
    ap.add_argument("--cycles", type=int, default=5)
    ap.add_argument("--smoother", default="MulticolourSSOR", choices=["MulticolourSSOR", "SSOR", "Jacobi", "Chebyshev"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--assembly", default="device", choices=["device", "host"],
                    help="e2e leg: system / level-0 matrices assembled on the device at the hand-over (default) or handed over "
                         "assembled (always measured as well)")
    ap.add_argument("--coarse-levels", type=int, default=0,
                    help="experimental (SURVEY 8f N4): multigrid levels below the base lattice; 0 = the reference's hierarchy "
                         "(the headline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    return ap.parse_args()


def workload_config(args, extra=None):
    n = args.atoms_n
    cfg = {
        "workload": f"atom_n{n}_{8 * n ** 3} (synthetic NaCl lattice = reference atom/atom_n{n}_{8 * n ** 3}.data), "
                    f"cycle {args.cycles - 1} of {args.cycles}: Gaussian-charge RHS + GMG-PCG solve",
        "prm": "Mesh size 0.25, Vacuum repetitions 10, smoothing length 0.5, cutoff 3.5, RHS quadrature 2^3, "
               "Homogeneous BC, Kelly marking (the cluster-log build)",
        "smoother": f"{args.smoother}(0.5) x 2", "coarse": "CG on level 0 to 1e-10 (abs), <= 1000 its",
        "tolerance": "1e-8 * |b|_2",
        "l2_hygiene": "inputs larger than L2: every outer iteration streams the 600 MB system matrix (SELL) and the step starts with "
                      "the RHS over 1.85 M cells / 118 M cell-atom pairs, so no coarse solve starts with a warm L2; inside a coarse "
                      "solve the row-pattern CG keeps its 57 MB working set L2-resident by design",
    }
    if getattr(args, "coarse_levels", 0):
        cfg["coarse"] = (f"EXPERIMENTAL, not the reference's algorithm: {args.coarse_levels} multigrid levels below the base "
                         f"lattice, CG on the coarsest to 1e-10")
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic():
    """dram__bytes_read+write per inner CG iteration from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "cg_persistent_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def write_atoms(args):
    lat = pkg().lattice
    pos, q = lat.nacl_lattice(args.atoms_n)
    d = tempfile.mkdtemp(prefix="gmg_bench_")
    path = os.path.join(d, f"atom_n{args.atoms_n}_{len(q)}.data")
    lat.write_lammps(path, pos, q)
    return path, pos, q


# ===================================================================================== B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # control plane only (barriers, handle exchange, max over ranks of the timings): the data path is peer memory
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = pkg()
    P.load_library()
    atom_file, pos, q = write_atoms(args)
    t0 = time.time()
    prm = P.lattice.cluster_prm(atom_file, args.atoms_n, cycles=args.cycles, smoother=args.smoother, device=local,
                                coarse_levels=args.coarse_levels)
    connect = None
    if world > 1:
        def connect(gmg):
            def all_gather_bytes(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
            P.capi.connect_ranks(gmg, rank, world, all_gather_bytes)
            dist.barrier()
    B = P.hostapi.BenchProblem(prm, connect=connect)
    setup_s = time.time() - t0
    g = B.gmg
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    traffic = g.matrix_traffic(P.capi.GMG_LEVEL, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = g.launch_count()
        g.coarse_profile(True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            out = [fn() for _ in range(steps)]
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, g.launch_count() - launches0, g.coarse_profile(True)

    for _ in range(max(args.warmup, 3)):
        B.step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ms, outs, launches, prof = timed(B.step_device, args.steps)
    clocks = sampler.stop()
    its = [o[0] for o in outs]
    # N > 1: ONE problem row-partitioned over the GPUs (strong scaling): every rank times the same solve
    dofs_total = B.n_dofs * args.steps
    value = dofs_total / (ms * 1e-3)

    # end to end through the host-buffer entry points (the LaplaceProblem methods themselves)
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 5))
    full_e2e = world == 1  # N > 1: the hierarchy hand-over (partitioning) is set-up; rhs + solve go through host buffers
    B.step_host(full_e2e)
    g.transfer_bytes(True)
    ms_e, outs_e, _, _ = timed(lambda: B.step_host(full_e2e), e2e_steps)
    h2d, d2h = g.transfer_bytes(True)
    e2e_value = B.n_dofs * e2e_steps / (ms_e * 1e-3)
    ms_e2, _, _, _ = timed(lambda: B.step_host(False), e2e_steps)
    # the same step with the system / level-0 matrices assembled on the device at the hand-over (gmg_assemble_matrix:
    # cell -> dof maps H2D instead of 1.2 GB of assembled CSR; the CSR built there is bit-identical, tests/test_gpu_assembly.py)
    e2e_dev = None
    if full_e2e and args.assembly == "device":
        try:
            B.set_device_assembly(True)
            for _ in range(3):  # (the arenas of the assembly reach their final size at the second hand-over)
                B.step_host(True)
            g.transfer_bytes(True)
            ms_d, outs_d, _, _ = timed(lambda: B.step_host(True), e2e_steps)
            h2d_d, d2h_d = g.transfer_bytes(True)
            if [o[0] for o in outs_d] != [o[0] for o in outs_e]:
                raise RuntimeError(f"device-assembled step took {outs_d} iterations, host-assembled {outs_e}")
            e2e_dev = {"ms": ms_d, "h2d": h2d_d, "d2h": d2h_d}
        except Exception as exc:  # the host-assembled hand-over above stays the reported number
            print(f"bench.py: device assembly leg failed: {exc}", file=sys.stderr)
        finally:
            B.set_device_assembly(False)

    peak, peak_src = measured_peak()
    # algorithmic bytes per inner iteration: CSR-equivalent 12 B per stored entry + 4 (n + 1) + 88 n;
    # next to it the bytes of the format actually held on the device (of this rank's row block)
    cg_bytes, stored_bytes = traffic["csr_cg_iter_bytes"], traffic["cg_iter_bytes"]
    if world > 1:
        t = torch.tensor([cg_bytes, stored_bytes], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        cg_bytes, stored_bytes = float(t[0].item()), float(t[1].item())  # all ranks together, per inner iteration
        peak_scale = world
    else:
        peak_scale = 1
    per_s = prof["iterations"] / (prof["ms"] * 1e-3) / 1e9 if prof["ms"] > 0 else 0.0
    achieved = cg_bytes * per_s
    tr = profile_traffic()
    fmt = traffic.get("format", 0)
    kernel = {4: "gmg::cg_persistent_win<2> (coarse-level CG on the row-pattern matrix, TMA-filled shared-memory windows, row "
                 "codes in global memory; one cooperative launch per V-cycle)",
              3: "gmg::cg_persistent_win<2> (coarse-level CG on the row-pattern matrix, TMA-filled shared-memory windows; one "
                 "cooperative launch per V-cycle)",
              2: "gmg::cg_persistent<512, PatView> (coarse-level CG on the row-pattern matrix, L1 gathers)",
              1: "gmg::cg_persistent<512, CsellView> (coarse-level CG, one cooperative launch per V-cycle)",
              0: "gmg::cg_persistent<512, SellView> (coarse-level CG, one cooperative launch per V-cycle)"}[
        g.coarse_kernel(P.capi.GMG_LEVEL, 0)]
    if world > 1:
        kernel = "gmg::cg_persistent_dist<512> (distributed coarse-level CG, halo + all-reduce over peer memory inside the kernel)"
    roofline = {
        "bound": "hbm", "kernel": kernel,
        "achieved": achieved, "peak": peak * peak_scale, "unit": "GB/s", "frac": achieved / (peak * peak_scale),
        "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
        "algorithmic_bytes_per_inner_iteration": cg_bytes, "stored_nnz_level0": traffic["nnz"],
        "stored_format": {"format": {0: "SELL-32 (12 B/entry)", 1: "CSELL (4 B/entry)", 2: "row-pattern dictionary (4 B/row)"}[fmt],
                          "bytes_per_inner_iteration": stored_bytes, "achieved_gbs": stored_bytes * per_s,
                          "frac_of_hbm_peak": stored_bytes * per_s / (peak * peak_scale)},
        "note": ("achieved = CSR-equivalent algorithmic bytes / time (SURVEY 8d).  With the row-pattern format the matrix is not "
                 "streamed at all and the CG's four vectors (%d MB) %s the 126 MB L2: frac > 1 means the kernel has left the HBM "
                 "roofline; it is bound by shared-memory/L1 wavefronts, L2 bandwidth (vector updates at ~10 TB/s) and three grid "
                 "barriers per iteration (profiles/)" % (4 * 8 * B.level_n[0] // 10 ** 6, "stay in" if 4 * 8 * B.level_n[0] < 100e6
                                                          else "no longer fit")) if fmt == 2 and world == 1 else None,
        "inner_iterations_per_launch": prof["iterations"] / max(prof["launches"], 1),
        "launches_in_timed_region": prof["launches"], "avg_launch_ms": prof["ms"] / max(prof["launches"], 1),
        "share_of_step": prof["ms"] / ms if ms > 0 else None,
        "traffic": (tr["dram_bytes_per_inner_iteration"] * prof["iterations"] / max(prof["launches"], 1)) if tr else None,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "b200",
        "config": workload_config(args, {
            "n_dofs": B.n_dofs, "n_dofs_level": B.level_n, "n_active_cells": B.n_cells, "n_atoms": B.n_atoms,
            "cell_atom_pairs": B.n_pairs, "outer_iterations": its[-1], "parallelism": "1 GPU" if world == 1 else
            f"{world} GPUs: level 0 + system matrix row-partitioned in z-slabs, patch levels replicated, halos / all-reduces "
            f"over NVLink peer memory inside the kernels (no NCCL on the data path)",
            "v_cycle_ms": None, "setup_seconds_untimed": setup_s}),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps,
                "ms_per_step_hierarchy_already_on_device": ms_e2 / e2e_steps,
                "what": ("compute_charge_densities + rhs assembly + solve() with host buffers: atoms/cells/CSR matrices/vectors H2D, "
                         "densities/rhs/solution D2H") if full_e2e else
                        "compute_charge_densities + rhs assembly + gmg_pcg_solve with host buffers (hierarchy partitioned at set-up)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if e2e_dev:
        host_leg = dict(line["e2e"])
        line["e2e"].update({
            "value": B.n_dofs * e2e_steps / (e2e_dev["ms"] * 1e-3), "ms_per_step": e2e_dev["ms"] / e2e_steps,
            "h2d_bytes_per_step": e2e_dev["h2d"] // e2e_steps, "d2h_bytes_per_step": e2e_dev["d2h"] // e2e_steps,
            "what": "compute_charge_densities + rhs assembly + solve() with host buffers and `Matrix assembly = Device`: "
                    "atoms/cells/cell-dof maps/constraints/patch-level CSR/vectors H2D, system + level-0 matrices assembled on "
                    "the device (bit-identical CSR), densities/rhs/solution D2H",
            "host_assembled_matrices": {k: host_leg[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step",
                                                                 "d2h_bytes_per_step")}})
    # V-cycle time: first (largest) and mean over the solve, from the coarse profile + one direct measurement
    src, dst = g.vec_alloc(B.n_dofs), g.vec_alloc(B.n_dofs)
    g.vec_upload(src, B.get("rhs"))
    for _ in range(3):
        g.vcycle_dev(src, dst)
    msv, _, _, pv = timed(lambda: g.vcycle_dev(src, dst), 10)
    line["config"]["v_cycle_ms"] = msv / 10
    line["config"]["v_cycle_inner_iterations"] = pv["iterations"] / max(pv["launches"], 1)
    # the one kernel of the step that still streams a matrix from HBM: the system-matrix SpMV of the outer PCG
    # (SELL, 12 B per entry; the refined mesh makes its rows too irregular for the row-pattern format)
    if world == 1:
        trs = g.matrix_traffic(P.capi.GMG_SYSTEM, 0)
        ms_s, _, _, _ = timed(lambda: g.spmv_dev(P.capi.GMG_SYSTEM, 0, src, dst), 20)
        gbs = trs["spmv_bytes"] * 20 / (ms_s * 1e-3) / 1e9
        line["roofline_level_spmv"] = {
            "bound": "hbm", "kernel": "gmg::sell_spmv<0, 0> (system matrix, %d rows, %d stored entries, SELL-32)" % (
                B.n_dofs, int(trs["nnz"])),
            "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_launch": trs["spmv_bytes"],
            "avg_launch_ms": ms_s / 20, "format": trs.get("format", 0),
            "note": "20 back-to-back launches; the matrix (%d MB) exceeds the 126 MB L2, x and y stay in it" % (
                trs["spmv_bytes"] // 10 ** 6)}
    g.vec_free(src)
    g.vec_free(dst)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_from(B, pos, q, args)
    B.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        _emit(line)


def cpu_baseline_from(B, pos, q, args):
    """The oracle's C port timed on the host cores on the same last-cycle problem (one full step)."""
    from oracle import cport, cpu_arm
    lo, H, reps = B.mesh()
    lists = (B.get("list_ptr"), B.get("list_atoms"))
    step = cpu_arm.CpuStep(B.get, B.n_levels, lo, H, pos, q, lists, B.nq, smoother="ssor")
    r = step.run(B.get("x0"))
    return {"value": B.n_dofs / r["seconds"], "unit": UNIT, "cores": cport.max_threads(), "kind": "port",
            "sample": f"1 full step (densities + load vector {r['rhs_seconds']:.2f} s, MG-PCG {r['solve_seconds']:.2f} s, "
                      f"{r['its']} outer its, processor-block SSOR with {step.n_blocks} blocks) of the same last-cycle problem",
            "outer_iterations": r["its"]}


# ===================================================================================== reference (CPU) arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport, cpu_arm
    P = pkg()
    pos, q = P.lattice.nacl_lattice(args.atoms_n)
    t0 = time.time()
    step, x0 = cpu_arm.adaptive_run_on_cpu(P.hostapi, pos, q, args.atoms_n, args.cycles, smoother="ssor",
                                           log=lambda s: print(s, file=sys.stderr, flush=True))
    setup_s = time.time() - t0
    # bounded: as many of the requested steps as fit the budget (at least one warm-up and one timed step)
    t = time.perf_counter()
    r = step.run(x0)
    first = time.perf_counter() - t
    k = max(1, min(args.steps, int(args.cpu_budget_s / max(first, 1e-3)) - 1))
    w = 1 if first * (k + args.warmup) > args.cpu_budget_s else max(0, min(args.warmup, 3) - 1)
    for _ in range(w):
        step.run(x0)
    t = time.perf_counter()
    for _ in range(k):
        r = step.run(x0)
    sec = time.perf_counter() - t
    value = step.n_dofs * k / sec
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w + 1,
        "ms_per_step": 1e3 * sec / k, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(args, {"n_dofs": step.n_dofs, "outer_iterations": r["its"], "smoother":
                                         f"processor-block SSOR(0.5) x 2, {step.n_blocks} blocks (= threads, as MPI ranks)",
                                         "setup_seconds_untimed": setup_s, "requested_steps": args.steps}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cport.max_threads(), "kind": "port",
                         "sample": f"{k} full steps (densities + load vector + MG-PCG) of the last-cycle problem"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def _emit(line):
    """The ONE JSON line on the real stdout (everything else the process prints goes to stderr, see below)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    a = parse_args()
    # libraries print to stdout too (e.g. "NCCL version ..." from the native side): keep stdout for the JSON line alone
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
