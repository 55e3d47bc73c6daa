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

Every b200 line carries a `parity` block: outer / coarse iteration counts and the solution norms of the timed
solve against the cluster log's printed values (tests/golden/reference_goldens.json), and at N = 1 the relative
L2 distance of the GPU load vector and solution from the CPU port's.
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
    ap.add_argument("--smoother", default="MulticolourSSOR", choices=["MulticolourSSOR", "SSOR", "Jacobi", "Chebyshev"],
                    help="smoother of the headline number.  MulticolourSSOR is the product's stand-in for the reference's "
                         "sequential SSOR (BASELINE.json north_star); the lexicographic SSOR step is timed beside it")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--assembly", default="device", choices=["device", "host"],
                    help="e2e leg: system / level-0 matrices assembled on the device at the hand-over (default) or handed over "
                         "assembled (always measured as well)")
    ap.add_argument("--coarse-levels", type=int, default=0,
                    help="SURVEY 8f N4 (not the reference's algorithm): multigrid levels below the base lattice; 0 = the "
                         "reference's hierarchy (the headline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the lexicographic-SSOR / coarse-level side legs")
    ap.add_argument("--cpu-budget-s", type=float, default=420.0)
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
        cfg["coarse"] = (f"NOT the reference's algorithm (SURVEY 8f N4): {args.coarse_levels} multigrid levels below the base "
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


def profile_traffic(kernel_id, world, n_rows):
    """dram__bytes_read + write per inner CG iteration from a committed `ncu --set full` capture of THIS kernel on THIS
    problem size (profiles/cg_traffic.json: list of records); None when no capture matches (never a pasted constant)."""
    p = os.path.join(ROOT, "profiles", "cg_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        recs = json.load(f)
    for r in recs:
        if r.get("kernel_id") == kernel_id and r.get("n_gpus") == world and r.get("level0_rows") == n_rows:
            return r
    return None


def golden_cycle(args):
    """The cluster log's printed values for the timed cycle (only the 64k-atom run has one)."""
    if args.atoms_n != 20 or args.cycles < 1 or args.cycles > 5:
        return None
    p = os.path.join(ROOT, "tests", "golden", "reference_goldens.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        g = json.load(f)["cluster_ssor_64k"][0]
    c = dict(g["cycles"][args.cycles - 1])
    c["source"] = f"{g['file']}:{c['line']}"
    return c


def write_atoms(args):
    lat = pkg().lattice
    pos, q = lat.nacl_lattice(args.atoms_n)
    d = tempfile.mkdtemp(prefix="gmg_bench_")
    path = os.path.join(d, f"atom_n{args.atoms_n}_{len(q)}.data")
    lat.write_lammps(path, pos, q)
    return path, pos, q


KERNEL_NAMES = {
    8: "gmg::cg_persistent_win2 XPAIR (coarse-level CG on the row-pattern matrix: TMA-filled shared-memory windows, two consecutive "
       "rows per lane in the dominant loop, tagged-word grid reductions, remainder rows on dedicated warps; h in global memory; "
       "one launch per V-cycle)",
    7: "gmg::cg_persistent_win2 XPAIR (coarse-level CG on the row-pattern matrix: TMA-filled shared-memory windows, two consecutive "
       "rows per lane in the dominant loop (LDS.128 / conflict-free LDS.64, 8 instead of 12 shared-memory wavefronts per run and "
       "64 rows), tagged-word grid reductions, remainder rows on dedicated warps, h = A d kept in shared memory; one launch per "
       "V-cycle)",
    6: "gmg::cg_persistent_win2 (coarse-level CG on the row-pattern matrix: TMA-filled shared-memory windows, tagged-word grid "
       "reductions, remainder rows on dedicated warps; h in global memory; one launch per V-cycle)",
    5: "gmg::cg_persistent_win2 (coarse-level CG on the row-pattern matrix: TMA-filled shared-memory windows, tagged-word grid "
       "reductions, remainder rows on dedicated warps, h = A d kept in shared memory; one launch per V-cycle)",
    4: "gmg::cg_persistent_win (coarse-level CG on the row-pattern matrix, TMA-filled shared-memory windows, row codes in "
       "global memory; one cooperative launch per V-cycle)",
    3: "gmg::cg_persistent_win (coarse-level CG on the row-pattern matrix, TMA-filled shared-memory windows; one cooperative "
       "launch per V-cycle)",
    2: "gmg::cg_persistent<512, PatView> (coarse-level CG on the row-pattern matrix, L1 gathers)",
    1: "gmg::cg_persistent<512, CsellView> (coarse-level CG, one cooperative launch per V-cycle)",
    0: "gmg::cg_persistent<512, SellView> (coarse-level CG, one cooperative launch per V-cycle)"}


# ===================================================================================== B200 arm
def run_b200(args):
    import numpy as np
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
    capi = P.capi
    atom_file, pos, q = write_atoms(args)
    warmup = max(args.warmup, 3)
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
            capi.connect_ranks(gmg, rank, world, all_gather_bytes)
            dist.barrier()
    B = P.hostapi.BenchProblem(prm, connect=connect)
    setup_s = time.time() - t0
    g = B.gmg
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    base_level = args.coarse_levels  # the level the base lattice sits on
    traffic = g.matrix_traffic(capi.GMG_LEVEL, 0)

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

    for _ in range(warmup):
        B.step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ms, outs, launches, prof = timed(B.step_device, args.steps)
    clocks = sampler.stop()
    its = [o[0] for o in outs]
    coarse_its = g.last_coarse_iterations()
    # N > 1: ONE problem row-partitioned over the GPUs (strong scaling): every rank times the same solve
    dofs_total = B.n_dofs * args.steps
    value = dofs_total / (ms * 1e-3)

    # ---- parity of the timed solve (every N): iteration counts and solution norms against the cluster log
    x_gpu = B.download_x()
    b_gpu = B.download_b() if world == 1 else None
    parity = {"outer_iterations": its[-1], "outer_iterations_all_steps_equal": len(set(its)) == 1,
              "coarse_iterations": coarse_its, "coarse_iterations_sum": int(sum(coarse_its)),
              "sol_l2": float(np.linalg.norm(x_gpu)), "sol_linf": float(np.abs(x_gpu).max()),
              "sol_l1": float(np.abs(x_gpu).sum()), "final_residual": outs[-1][1]}
    gold = golden_cycle(args)
    ok = True
    if gold and not args.coarse_levels:
        parity["golden"] = {k: gold[k] for k in ("source", "its", "sol_l1", "sol_l2", "sol_linf", "n_dofs", "n_dofs_level")}
        parity["golden"]["note"] = ("printed by the reference on 20 MPI ranks with processor-block SSOR; outer iterations "
                                    "must agree within +-2, norms to 2e-7 relative (the solve stops at 1e-8 |b|)")
        rel = {k: abs(parity[k] - gold[k]) / gold[k] for k in ("sol_l1", "sol_l2", "sol_linf")}
        parity["rel_err_vs_golden"] = rel
        ok = (max(rel.values()) <= 2e-7 and abs(its[-1] - gold["its"]) <= 2 and B.n_dofs == gold["n_dofs"]
              and B.level_n == gold["n_dofs_level"])
    parity["ok"] = bool(ok)

    # ---- end to end through the host-buffer entry points (the LaplaceProblem methods themselves)
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
    if full_e2e and args.assembly == "device" and not args.coarse_levels:
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
            B.step_host(full_e2e)

    peak, peak_src = measured_peak()
    # bytes per inner iteration: `stored` = what the format held on the device must move (pattern ids + table + remainder
    # + the CG's vector reads / writes); `csr` = the CSR-equivalent figure of SURVEY.md 8(d) (12 B per stored entry), kept
    # as a side note: the row-pattern format does not move those bytes, so it is NOT a roofline numerator
    cg_bytes, stored_bytes = traffic["csr_cg_iter_bytes"], traffic["cg_iter_bytes"]
    n_rows0 = B.level_n[0]
    if world > 1:
        t = torch.tensor([cg_bytes, stored_bytes], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        cg_bytes, stored_bytes = float(t[0].item()), float(t[1].item())  # all ranks together, per inner iteration
    per_s = prof["iterations"] / (prof["ms"] * 1e-3) / 1e9 if prof["ms"] > 0 else 0.0
    fmt = traffic.get("format", 0)
    kid = g.coarse_kernel(capi.GMG_LEVEL, 0)
    kernel = KERNEL_NAMES[kid]
    if world > 1:
        kernel = ("gmg::cg_persistent_dist<1024> (distributed coarse-level CG: halo rows and all-to-all tagged-word reductions over "
                  "NVLink peer memory inside the kernel, no grid.sync)")
    vec_mb = 4 * 8 * n_rows0 / 1e6
    l2_resident = fmt == 2 and vec_mb / world < 100
    tr = profile_traffic(kid if world == 1 else "dist", world, n_rows0)
    its_per_launch = prof["iterations"] / max(prof["launches"], 1)
    achieved = stored_bytes * per_s
    roofline = {
        "bound": "l2/latency" if l2_resident else "hbm", "kernel": kernel,
        "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
        "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
        "what": ("achieved = bytes the stored format must move per inner iteration x inner iterations / kernel time (CUDA events "
                 "around each launch in the timed region); peak = HBM copy peak.  " +
                 ("The working set (%d MB of vectors + 7 MB of pattern ids per GPU) is L2-resident: the bytes come from L2, "
                  "not HBM; the kernel is bound by shared-memory wavefronts, L2 latency and the grid barriers per iteration "
                  "(profiles/), so frac is a distance-to-HBM-roofline figure, not an HBM utilisation." % (vec_mb / world)
                  if l2_resident else "The vectors no longer fit L2: HBM-bound regime.")),
        "bytes_per_inner_iteration": stored_bytes,
        "format": {0: "SELL-32 (12 B/entry)", 1: "CSELL (4 B/entry)", 2: "row-pattern dictionary (4 B/row)"}[fmt],
        "us_per_inner_iteration": 1e3 * prof["ms"] / max(prof["iterations"], 1),
        "inner_iterations_per_launch": its_per_launch,
        "launches_in_timed_region": prof["launches"], "avg_launch_ms": prof["ms"] / max(prof["launches"], 1),
        "share_of_step": prof["ms"] / ms if ms > 0 else None,
        "traffic": (tr["dram_bytes_per_inner_iteration"] * its_per_launch) if tr else None,
        "traffic_source": (tr.get("source") if tr else "no ncu --set full capture of this kernel at this size / GPU count"),
        "csr_equivalent_side_note": {"bytes_per_inner_iteration": cg_bytes, "gbs": cg_bytes * per_s,
                                     "stored_nnz_level0": traffic["nnz"],
                                     "note": "SURVEY 8(d) CSR-equivalent bytes / time; exceeds the HBM peak because the format "
                                             "does not move them -- not a roofline fraction"},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "b200",
        "config": workload_config(args, {
            "n_dofs": B.n_dofs, "n_dofs_level": B.level_n, "n_active_cells": B.n_cells, "n_atoms": B.n_atoms,
            "cell_atom_pairs": B.n_pairs, "outer_iterations": its[-1], "parallelism": "1 GPU" if world == 1 else
            f"{world} GPUs: level 0 + system matrix row-partitioned in z-slabs, patch levels replicated, halos / all-reduces "
            f"over NVLink peer memory inside the kernels (no NCCL on the data path)",
            "v_cycle_ms": None, "setup_seconds_untimed": setup_s}),
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps,
                "ms_per_step_hierarchy_already_on_device": ms_e2 / e2e_steps,
                "includes_hierarchy_hand_over": bool(full_e2e),
                "what": ("compute_charge_densities + rhs assembly + solve() with host buffers: atoms/cells/CSR matrices/vectors H2D, "
                         "rhs/solution D2H") if full_e2e else
                        "compute_charge_densities + rhs assembly + gmg_pcg_solve with host buffers; EXCLUDES the hierarchy "
                        "hand-over (partitioned once at set-up), so compare with N=1's ms_per_step_hierarchy_already_on_device, "
                        "not with N=1's e2e value"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if e2e_dev:
        host_leg = dict(line["e2e"])
        line["e2e"].update({
            "value": B.n_dofs * e2e_steps / (e2e_dev["ms"] * 1e-3), "ms_per_step": e2e_dev["ms"] / e2e_steps,
            "h2d_bytes_per_step": e2e_dev["h2d"] // e2e_steps, "d2h_bytes_per_step": e2e_dev["d2h"] // e2e_steps,
            "what": "compute_charge_densities + rhs assembly + solve() with host buffers and `Matrix assembly = Device`: "
                    "atoms/cells/cell-dof maps/constraints/patch-level CSR/vectors H2D, system + level-0 matrices assembled on "
                    "the device (bit-identical CSR), rhs/solution D2H (the densities stay on the device: their consumers run there)",
            "host_assembled_matrices": {k: host_leg[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step",
                                                                 "d2h_bytes_per_step")}})
    # ---- V-cycle time: first (largest) and mean over the solve, from the coarse profile + one direct measurement
    src, dst = g.vec_alloc(B.n_dofs), g.vec_alloc(B.n_dofs)
    g.vec_upload(src, B.get("rhs"))
    for _ in range(3):
        g.vcycle_dev(src, dst)
    g.debug_vcycle_profile(True)
    msv, _, _, pv = timed(lambda: g.vcycle_dev(src, dst), 10)
    vp = g.debug_vcycle_profile(False)
    line["config"]["v_cycle_ms"] = msv / 10
    line["config"]["v_cycle_inner_iterations"] = pv["iterations"] / max(pv["launches"], 1)
    if world == 1:
        # ---- smoother: colour sweeps of the patch levels.  Algorithmic bytes of one SSOR application =
        # 2 SpMV + 16 n; a V-cycle applies it `steps` times before and after the coarse solve on every level >= 1.
        # Time = down + up parts of the V-cycle (CUDA events), which also hold the residual, restriction, prolongation
        # and copy kernels, so the fraction is a lower bound.
        sm_bytes = 0.0
        for l in range(1, B.n_levels):
            tl = g.matrix_traffic(capi.GMG_LEVEL, l)
            sm_bytes += 2 * 2 * (2 * tl["csr_spmv_bytes"] + 16 * B.level_n[l])
        sm_ms = (vp["down_ms"] + vp["up_ms"]) / max(vp["vcycles"], 1)
        if sm_ms > 0 and sm_bytes > 0:
            line["roofline_smoother"] = {
                "bound": "latency (patch levels of %s rows are L2-resident; one launch per colour)" % B.level_n[1:],
                "kernel": "gmg::sell_color_relax(_pdl) colour sweeps of levels >= 1 (+ residual / transfer kernels)",
                "achieved": sm_bytes / (sm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": sm_bytes / (sm_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_v_cycle": sm_bytes,
                "ms_per_v_cycle": sm_ms, "coarse_ms_per_v_cycle": vp["coarse_ms"] / max(vp["vcycles"], 1)}
        # ---- the one kernel of the step that still streams a matrix from HBM: the system-matrix SpMV of the outer PCG
        # (SELL, 12 B per entry; the refined mesh makes its rows too irregular for the row-pattern format)
        trs = g.matrix_traffic(capi.GMG_SYSTEM, 0)
        ms_s, _, _, _ = timed(lambda: g.spmv_dev(capi.GMG_SYSTEM, 0, src, dst), 20)
        gbs = trs["spmv_bytes"] * 20 / (ms_s * 1e-3) / 1e9
        line["roofline_level_spmv"] = {
            "bound": "hbm", "kernel": "gmg::sell_spmv<0, 0> (system matrix, %d rows, %d stored entries, SELL-32)" % (
                B.n_dofs, int(trs["nnz"])),
            "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_launch": trs["spmv_bytes"],
            "avg_launch_ms": ms_s / 20, "format": trs.get("format", 0),
            "note": "20 back-to-back launches; the matrix (%d MB) exceeds the 126 MB L2, x and y stay in it" % (
                trs["spmv_bytes"] // 10 ** 6)}
        # ---- RHS: fp64-ALU bound, reported as exp evaluations / s.  Separable Gaussians: 3 nq exponentials
        # per (cell, atom) pair instead of nq^3 (the reference's count is given beside it).
        bdev = g.vec_alloc(B.n_dofs)
        for _ in range(2):
            g.rhs_step_dev(bdev)
        ms_r, _, _, _ = timed(lambda: g.rhs_step_dev(bdev), 10)
        g.vec_free(bdev)
        nq = B.nq
        exp_sep, exp_ref = 3 * nq * B.n_pairs_active, nq ** 3 * B.n_pairs_active
        # per pair: 3 nq exp (~25 fp64 flops each incl. argument) + nq^3 (2 mul + 1 fma) products
        flops = B.n_pairs_active * (3 * nq * 25.0 + nq ** 3 * 4.0)
        line["roofline_rhs"] = {
            "bound": "fp64 ALU", "kernel": "gmg::density_reg_kernel / load_vector_kernel (densities + load vector, one RHS step)",
            "ms_per_rhs_step": ms_r / 10, "cell_atom_pairs_evaluated": B.n_pairs_active,
            "exp_evaluations_per_step": exp_sep, "exp_per_s": exp_sep / (ms_r / 10 * 1e-3),
            "reference_exp_evaluations_per_step": exp_ref, "reference_equivalent_exp_per_s": exp_ref / (ms_r / 10 * 1e-3),
            "fp64_flops_est": flops, "fp64_tflops_est": flops / (ms_r / 10 * 1e-3) / 1e12,
            "fp64_frac_of_peak_est": flops / (ms_r / 10 * 1e-3) / 1e12 / 37.0,
            "note": "fp64 peak taken as 37 TFLOP/s (B200 vector fp64, nominal); flop count is an estimate (25 flops per exp)"}
        # ---- binning: the reference's 6871 s loop (SSOR_64k_atoms.o876224:69), here gmg_bin_atoms through host buffers
        try:
            bt = [B.time_binning() for _ in range(3)]
            line["binning"] = {"ms": min(t[0] for t in bt), "cell_atom_pairs": bt[0][1],
                               "what": "rhs_assembly_optimization(): two gmg_bin_atoms calls (count, fill) with host buffers; "
                                       "lists bit-identical to the set-up's (checked); best of 3",
                               "reference_published_s": 6871.0 if args.atoms_n == 20 else None,
                               "reference_published_where": "20 MPI ranks, SSOR_64k_atoms.o876224:69"}
        except Exception as exc:
            print(f"bench.py: binning leg failed: {exc}", file=sys.stderr)
    g.vec_free(src)
    g.vec_free(dst)

    # ---- the same step with the reference's own smoother (lexicographic SSOR, level-scheduled on the device): the
    # parity-pinned path, timed beside the headline (multicolour SSOR is a different, colour-ordered SSOR operator)
    if args.smoother == "MulticolourSSOR" and not args.no_side_legs:
        try:
            g.set_smoother(capi.SMOOTHER_LEX_SSOR, 0.5, 2)
            g.setup()
            for _ in range(2):
                B.step_device()
            k_lex = max(2, min(args.steps, 5))
            ms_l, outs_l, launches_l, _ = timed(B.step_device, k_lex)
            x_lex = B.download_x()
            line["smoothers"] = {
                "MulticolourSSOR": {"ms_per_step": ms / args.steps, "value": value, "outer_iterations": its[-1],
                                    "role": "headline (north_star: multicolour SSOR stands in for the sequential SSOR)"},
                "SSOR": {"ms_per_step": ms_l / k_lex, "value": B.n_dofs * k_lex / (ms_l * 1e-3), "steps": k_lex,
                         "outer_iterations": outs_l[-1][0], "gpu_launches_per_step": int(launches_l // k_lex),
                         "rel_l2_solution_vs_multicolour": float(np.linalg.norm(x_lex - x_gpu) / np.linalg.norm(x_gpu)),
                         "role": "the reference's lexicographic SSOR(0.5) x 2 on one rank, level-scheduled wavefronts; "
                                 "reproduces the 1-rank golden iteration counts"},
                "why_counts_differ": "multicolour SSOR relaxes the rows colour by colour (8 vertex-parity colours) instead of in "
                                     "index order: same fixed point, slightly different error propagation; the reference's own "
                                     "count also moves with the rank count (processor-block SSOR: 6 its on 1 rank, 7 on 3)"}
        except Exception as exc:
            print(f"bench.py: lexicographic SSOR leg failed: {exc}", file=sys.stderr)
        finally:
            g.set_smoother(capi.SMOOTHER_MC_SSOR, 0.5, 2)
            g.setup()
        # ---- point Jacobi(0.5) x 2 (north_star smoother (2); the reference's test configuration): no colour launches at all
        try:
            g.set_smoother(capi.SMOOTHER_JACOBI, 0.5, 2)
            g.setup()
            for _ in range(2):
                B.step_device()
            k_j = max(2, min(args.steps, 5))
            ms_j, outs_j, launches_j, _ = timed(B.step_device, k_j)
            x_j = B.download_x()
            line.setdefault("smoothers", {})["Jacobi"] = {
                "ms_per_step": ms_j / k_j, "value": B.n_dofs * k_j / (ms_j * 1e-3), "steps": k_j,
                "outer_iterations": outs_j[-1][0], "gpu_launches_per_step": int(launches_j // k_j),
                "rel_l2_solution_vs_multicolour": float(np.linalg.norm(x_j - x_gpu) / np.linalg.norm(x_gpu)),
                "role": "damped point Jacobi(0.5) x 2 on every level (Ifpack's, as in the reference's Jacobi goldens)"}
        except Exception as exc:
            print(f"bench.py: Jacobi leg failed: {exc}", file=sys.stderr)
        finally:
            g.set_smoother(capi.SMOOTHER_MC_SSOR, 0.5, 2)
            g.setup()
        # ---- SURVEY 8f N4 (never the headline: a different hierarchy than the reference's): coarse levels below the base
        # mesh, so the coarse CG runs on (reps / 2^k + 1)^3 dofs; same meshes above the base lattice, same RHS, same solution
        if world == 1 and not args.coarse_levels:
            line["coarse_levels_below_base_mesh"] = {}
            for sm_name, k_lv in (("Jacobi", 3), ("MulticolourSSOR", 3)):
                B4 = None
                try:
                    prm4 = P.lattice.cluster_prm(atom_file, args.atoms_n, cycles=args.cycles, smoother=sm_name, device=local,
                                                 coarse_levels=k_lv)
                    B4 = P.hostapi.BenchProblem(prm4)
                    B4.gmg.set_stream(stream.cuda_stream)
                    for _ in range(3):
                        B4.step_device()
                    k4 = max(2, min(args.steps, 5))
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(stream):
                        e0.record(stream)
                        outs4 = [B4.step_device() for _ in range(k4)]
                        e1.record(stream)
                    barrier()
                    ms4 = e0.elapsed_time(e1)
                    x4 = B4.download_x()
                    line["coarse_levels_below_base_mesh"]["%s_k%d" % (sm_name, k_lv)] = {
                        "smoother": sm_name, "coarse_levels": k_lv, "ms_per_step": ms4 / k4,
                        "value": B4.n_dofs * k4 / (ms4 * 1e-3), "n_dofs_level": B4.level_n,
                        "outer_iterations": outs4[-1][0], "coarse_iterations": B4.gmg.last_coarse_iterations(),
                        "same_system_dofs": B4.n_dofs == B.n_dofs,
                        # (the deeper hierarchy numbers the dofs differently: the norms are what can be compared)
                        "rel_err_solution_norms_vs_headline": {
                            "l1": abs(float(np.abs(x4).sum()) - parity["sol_l1"]) / parity["sol_l1"],
                            "l2": abs(float(np.linalg.norm(x4)) - parity["sol_l2"]) / parity["sol_l2"],
                            "linf": abs(float(np.abs(x4).max()) - parity["sol_linf"]) / parity["sol_linf"]}}
                except Exception as exc:
                    print(f"bench.py: coarse-levels leg ({sm_name}, k = {k_lv}) failed: {exc}", file=sys.stderr)
                finally:
                    if B4 is not None:
                        B4.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], cmpv = cpu_baseline_from(B, pos, q, args, b_gpu, x_gpu)
        line["parity"].update(cmpv)
        line["parity"]["ok"] = bool(line["parity"]["ok"] and cmpv["rel_l2_rhs_vs_cpu_port"] <= 1e-12
                                    and cmpv["rel_l2_solution_vs_cpu_port"] <= 1e-6)
    B.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        _emit(line)
        if not line["parity"]["ok"]:
            print("bench.py: PARITY CHECK FAILED: " + json.dumps(line["parity"]), file=sys.stderr)
            sys.exit(3)


def host_cores():
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_cores():
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    from oracle import cport
    cport.set_threads(n)
    return n


def cpu_baseline_from(B, pos, q, args, b_gpu, x_gpu):
    """The oracle's C port timed on the host cores on the same last-cycle problem (one full step), and the distance of
    the GPU's load vector / solution from its."""
    import numpy as np
    cores = use_all_cores()
    from oracle import cpu_arm
    lo, H, reps = B.mesh()
    lists = (B.get("list_ptr"), B.get("list_atoms"))
    k = args.coarse_levels
    step = cpu_arm.CpuStep(B.get, B.n_levels, lo, H, pos, q, lists, B.nq, smoother="ssor", base_level=k,
                           reps_base=reps << k if k else None)
    r = step.run(B.get("x0"))
    cmpv = {"rel_l2_rhs_vs_cpu_port": float(np.linalg.norm(b_gpu - r["b"]) / np.linalg.norm(r["b"])),
            "rel_l2_solution_vs_cpu_port": float(np.linalg.norm(x_gpu - r["x"]) / np.linalg.norm(r["x"])),
            "cpu_port_outer_iterations": r["its"],
            "tolerances": "rhs 1e-12 (north_star); solution 1e-6 (both solves stop at 1e-8 |b| with different smoothers)"}
    return ({"value": B.n_dofs / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
             "smoother": f"processor-block lexicographic SSOR(0.5) x 2 with {step.n_blocks} blocks (the reference's Ifpack "
                         f"SSOR on {step.n_blocks} ranks); the GPU headline uses multicolour SSOR, its lexicographic-SSOR "
                         f"step is in `smoothers`",
             "sample": f"1 full step (densities + load vector {r['rhs_seconds']:.2f} s, MG-PCG {r['solve_seconds']:.2f} s, "
                       f"{r['its']} outer its) of the same last-cycle problem",
             "outer_iterations": r["its"]}, cmpv)


# ===================================================================================== reference (CPU) arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_cores()
    from oracle import cpu_arm
    P = pkg()  # (lattice + the pure-host ministep binding only: libgmg_b200.so is never mapped by this arm)
    pos, q = P.lattice.nacl_lattice(args.atoms_n)
    t0 = time.time()
    step, x0 = cpu_arm.adaptive_run_on_cpu(P.hostapi, pos, q, args.atoms_n, args.cycles, smoother="ssor",
                                           log=lambda s: print(s, file=sys.stderr, flush=True))
    setup_s = time.time() - t0
    # same warm-up and step counts as the GPU arm; shortened (and said so) only if the budget would be exceeded
    warmup = max(args.warmup, 3)
    t = time.perf_counter()
    r = step.run(x0)
    first = time.perf_counter() - t
    k, w = args.steps, warmup
    if first * (k + w) > args.cpu_budget_s:
        k = max(1, min(args.steps, int(args.cpu_budget_s / max(first, 1e-3)) - w))
        if k < 5:
            k, w = max(1, min(args.steps, 5)), max(1, int(args.cpu_budget_s / max(first, 1e-3)) - 5)
    for _ in range(w - 1):
        step.run(x0)
    t = time.perf_counter()
    for _ in range(k):
        r = step.run(x0)
    sec = time.perf_counter() - t
    value = step.n_dofs * k / sec
    with open("/proc/self/maps") as f:
        maps = f.read()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w,
        "ms_per_step": 1e3 * sec / k, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(args, {"n_dofs": step.n_dofs, "outer_iterations": r["its"], "smoother":
                                         f"processor-block SSOR(0.5) x 2, {step.n_blocks} blocks (= threads, as MPI ranks)",
                                         "setup_seconds_untimed": setup_s, "requested_steps": args.steps,
                                         "requested_warmup": args.warmup, "cuda_library_mapped": "libgmg_b200" in maps}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
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
