#!/usr/bin/env python
"""bench.py - HEOM ADO-steps/s (RK4, FP64) on N B200s, with HBM roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one RK4 step of the whole hierarchy.  ``value`` = ADOs x steps /
device time (CUDA events, state resident in HBM); ``e2e`` = the same metric
through ``DEOMSolver.run`` with host inputs and the trajectory copied back.
See DESIGN.md section "Measurement" for the definitions of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyqed_b200 import workloads as W  # noqa: E402

METRIC = "HEOM ADO-steps/sec (RK4)"
UNIT = "ADO-steps/s"

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the HBM-roofline target is quoted on;
    # 4 292 145 ADOs, 3.37 GB per array - fits one B200, far larger than L2
    "fmo7_K21_L8": lambda: W.fmo(lmax=8, n_matsubara=2),
    # documented fallback of SURVEY.md section 8d ("3alt")
    "fmo7_K14_L8": lambda: W.fmo(lmax=8, n_matsubara=1),
    "fmo7_K21_L6": lambda: W.fmo(lmax=6, n_matsubara=2),
    "fmo7_K21_L5": lambda: W.fmo(lmax=5, n_matsubara=2),
    "fmo7_K21_L4": lambda: W.fmo(lmax=4, n_matsubara=2),
    "fmo7_K21_L3": lambda: W.fmo(lmax=3, n_matsubara=2),
    # BASELINE.json configs[1]: 330 ADOs, 259 kB - cache resident, latency bound
    "fmo7_K7_L4": lambda: W.fmo(lmax=4, n_matsubara=0),
    "spin_boson_K2_L10": lambda: W.spin_boson(lmax=10),
    "polariton32_K4_L6": lambda: W.polariton(lmax=6),
    # SURVEY 8d's stress variant of config 4: dense random Hermitian H (every operator row full)
    "polariton32_dense_K4_L6": lambda: W.polariton(lmax=6, dense_h=True),
    "aggregate7_K6_L6": lambda: W.aggregate_2des(lmax=6),
}
DEFAULT_WORKLOAD = "fmo7_K21_L8"


def measured_traffic(workload, kernel, order):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/r02_traffic.json, r01_traffic.json), if it was taken for this workload/kernel/order."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                t = json.load(fh)
            if (t["workload"], t["kernel"]) == (workload, kernel) and t["storage_order"] == order:
                return t["dram_bytes_per_launch_avg"]
        except Exception:
            pass
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; here it is the thing being timed as
# the reference's CPU cost structure, never part of the GPU product path)
# ---------------------------------------------------------------------------
def cpu_reference_leg(workload_name, budget_s=12.0, steps=None):
    """Per-ADO-loop port of the reference's generate_dot_element/rk4
    (oracle.deom_oracle.rhs_loop) on a bounded sample of the workload: same N,
    K, operators and bath, hierarchy depth reduced until one RK4 step takes
    about a second."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.deom_oracle import DeomOracle
    w = WORKLOADS[workload_name]()
    lmax = w["lmax"]
    nind = len(w["expn"])
    from math import comb
    depth = lmax
    while depth > 1 and comb(depth + nind, depth) > 1500:
        depth -= 1
    o = DeomOracle(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"],
                   w["etal"], w["etar"], w["etaa"], w["mode"], depth,
                   w["pulse_system_func"], w["pulse_coupling_func"])
    rho = np.zeros((o.nmax, o.nsys, o.nsys), dtype=np.complex128)
    rho[0] = w["rho0"]
    rho = o.rk4_step(rho, w["dt"], 0.0, o.rhs_loop)  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        rho = o.rk4_step(rho, w["dt"], (done + 1) * w["dt"], o.rhs_loop)
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and el > budget_s):
            break
    return dict(value=o.nmax * done / el, unit=UNIT, cores=1, kind="port",
                sample=(f"{workload_name} operators and bath at depth {depth} ({o.nmax} ADOs) instead of "
                        f"{lmax}, {done} RK4 steps in {el:.1f} s, per-ADO NumPy loop restating "
                        f"generate_dot_element/rk4 (deom.py:641-766), 1 thread (the reference is serial)")), el, done


def cpu_native_leg(workload_name, budget_s=10.0, steps=None):
    """C/OpenMP restatement (oracle/heom_oracle.c) on every host core: the same
    dense per-ADO arithmetic as the reference, compiled and threaded.  This is
    the strongest faithful CPU implementation in the repo and the number the
    reference arm reports."""
    from math import comb
    from oracle import c_oracle
    w = WORKLOADS[workload_name]()
    nind, depth = len(w["expn"]), w["lmax"]
    # largest depth whose arrays are well out of the last-level cache but whose step still takes
    # well under a second (K = 21: depth 6, 296 010 ADOs, 232 MB per array)
    while depth > 1 and comb(depth + nind, depth) > 300000:
        depth -= 1
    nmax = comb(depth + nind, depth)
    # every core this process may run on (torchrun's OMP_NUM_THREADS=1 default is not a limit)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def go(nt):
        t0 = time.perf_counter()
        c_oracle.run(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"],
                     w["etal"], w["etar"], w["etaa"], w["mode"], depth, w["rho0"], w["dt"], nt,
                     w["pulse_system_func"], w["pulse_coupling_func"], threads=threads)
        return time.perf_counter() - t0
    t1 = go(1)  # warm-up (also builds the library on first use)
    t1 = go(1)
    nt = steps if steps is not None else max(10, min(400, int(budget_s / max(t1, 1e-4))))   # (each call first-touches ~1 GB: keep the run long enough)
    el = go(nt)
    full = comb(w["lmax"] + nind, w["lmax"])
    return dict(value=nmax * nt / el, unit=UNIT, cores=threads, kind="port", n_ado=nmax, depth=depth,
                sample=(f"{workload_name} operators and bath at depth {depth} ({nmax} ADOs, "
                        f"{nmax * w['system'].shape[0] ** 2 * 16 / 1e6:.0f} MB per array) instead of depth "
                        f"{w['lmax']} ({full} ADOs), {nt} RK4 steps in {el:.1f} s, C/OpenMP restatement of "
                        f"generate_dot_element/rk4 (deom.py:641-766), dense N x N products per term, "
                        f"{threads} threads; the work per ADO-step does not depend on the depth, so ADO-steps/s "
                        f"at full depth is expected to be the same or lower (more cache misses): "
                        f"{full / (nmax * nt / el):.1f} s per RK4 step extrapolated")), el, nt


def cpu_batched_leg(workload_name, budget_s=6.0):
    """Stronger CPU number: the batched-NumPy oracle (all ADOs per call)."""
    from oracle.deom_oracle import DeomOracle
    from math import comb
    w = WORKLOADS[workload_name]()
    nind, depth = len(w["expn"]), w["lmax"]
    while depth > 1 and comb(depth + nind, depth) > 70000:
        depth -= 1
    o = DeomOracle(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"],
                   w["etal"], w["etar"], w["etaa"], w["mode"], depth,
                   w["pulse_system_func"], w["pulse_coupling_func"])
    rho = np.zeros((o.nmax, o.nsys, o.nsys), dtype=np.complex128)
    rho[0] = w["rho0"]
    rho = o.rk4_step(rho, w["dt"], 0.0, o.rhs_batched)
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        rho = o.rk4_step(rho, w["dt"], 0.0, o.rhs_batched)
        done += 1
    el = time.perf_counter() - t0
    return dict(value=o.nmax * done / el, unit=UNIT,
                cores=int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)), kind="port",
                sample=f"batched-NumPy oracle at depth {depth} ({o.nmax} ADOs), {done} steps in {el:.1f} s")


def small_workload_leg(name, device, nt=2000):
    """Device-timed ADO-steps/s of a small configuration (one resident-kernel launch)."""
    import torch
    from pyqed_b200.heom import DEOMSolver, Bath
    w = WORKLOADS[name]()
    bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
    s = DEOMSolver(w["system"], w["system_dipole"], bath, w["coupling"], w["coupling_dipole"],
                   w["pulse_system_func"], w["pulse_coupling_func"], lmax=w["lmax"], device=device,
                   alias_rho0=False)
    s.run(w["rho0"].copy(), w["dt"], 10)
    plan = s._plan
    plan.set_state(w["rho0"][None])
    plan.propagate(w["dt"], 50, None, None, None, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.propagate(w["dt"], nt, None, None, None, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if plan.info("dataflow_tma_launches"):
        kernel, note = "stage_dataflow_tma_kernel", ("one cooperative launch, one CTA per ADO, state in registers / shared "
                                                     "memory, stage outputs exchanged through L2; latency bound")
    elif plan.info("dataflow_launches"):
        kernel, note = "stage_dataflow_kernel", "one cooperative launch, state in L2; latency bound"
    else:
        kernel = {0: "per-stage kernels", 4: "resident_cluster_kernel", 5: "resident_elem_kernel"}[
            plan.info("resident_kind") if plan.info("resident_launches") else 0]
        note = "state resident in distributed shared memory; an HBM fraction is not meaningful here"
    return {"value": plan.nmax * nt / (ms * 1e-3), "unit": UNIT, "n_ado": plan.nmax, "steps": nt,
            "us_per_step": 1e3 * ms / nt, "kernel": kernel, "note": note}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, el, done = cpu_native_leg(args.workload, steps=max(1, args.steps) if args.steps_given else None)
    py, _, _ = cpu_reference_leg(args.workload, budget_s=8.0)
    w = WORKLOADS[args.workload]()
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": done, "warmup": 1, "ms_per_step": 1e3 * el / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "nsys": int(w["system"].shape[0]),
                   "nind": int(len(w["expn"])), "lmax": int(w["lmax"]),
                   "sample_depth": cb["depth"], "sample_n_ado": cb["n_ado"],
                   "note": "CPU leg runs a bounded sample of the workload (same operators and bath, reduced "
                           "hierarchy depth), see cpu_baseline.sample"},
        "cpu_baseline": cb,
        "cpu_baseline_python_loop": py,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# correctness evidence carried by the bench line
# ---------------------------------------------------------------------------
RHO_REF = os.path.join(ROOT, "profiles", "r02_rho_sys_n1.json")


def fixture_parity(multi, transport, device, order, tuning, options, native):
    """The path that is being timed, run on the reference's own output: the fixture
    tests/golden/deom_fmo_K21_L3.npz (same operators and bath as the headline workload at depth
    3, trajectory and final ADOs written by the unmodified reference, tests/golden/make_golden.py).
    Returns max |rho_sys(t) - reference| over the trajectory (and over all final ADOs)."""
    from pyqed_b200.heom import DEOMSolver, Bath
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "deom_fmo_K21_L3.npz"), allow_pickle=False))
    dt, nt = float(g["dt"]), int(g["nt"])
    if not multi:
        bath = Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"])
        s = DEOMSolver(g["system"], g["system_dipole"], bath, g["coupling"], g["coupling_dipole"],
                       lmax=int(g["lmax"]), device=device, order=order, alias_rho0=False)
        s.tuning, s.options = tuning, dict(options, resident=0)
        _, traj = s.run(g["rho0"].copy(), dt, nt)
        ados = np.array(s.ddos)
    else:
        from pyqed_b200.heom.sharded import ShardedDEOM
        sh = ShardedDEOM(g["system"], g["system_dipole"], g["coupling"], g["coupling_dipole"], g["expn"],
                         g["etal"], g["etar"], g["etaa"], g["mode"], int(g["lmax"]), transport, device=device,
                         order=order, options=dict(options, resident=0), tuning=tuning, native=native)
        _, traj = sh.run(g["rho0"], dt, nt)
        ados = sh.gather_ados()
        sh.close()
    return {"fixture": "tests/golden/deom_fmo_K21_L3.npz (reference output, 2024 ADOs, %d steps)" % nt,
            "max_abs_diff_trajectory": float(np.max(np.abs(np.asarray(traj) - g["traj"]))),
            "max_abs_diff_all_ados": float(np.max(np.abs(ados - g["ados_final"])))}


def rho_reference(workload, steps):
    try:
        with open(RHO_REF) as fh:
            ref = json.load(fh)
        v = ref.get(workload, {}).get(str(steps))
        return None if v is None else np.array(v, dtype=np.float64).view(np.complex128).reshape(-1)
    except Exception:
        return None


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pyqed_b200.heom import DEOMSolver, Bath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    torch.cuda.set_device(local)
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    order = args.order if args.order >= 0 else 2   # blocked lexicographic: L2 locality of the neighbour rows

    w = WORKLOADS[args.workload]()
    n, nind, lmax = int(w["system"].shape[0]), int(len(w["expn"])), int(w["lmax"])
    K, Wm, dt = args.steps, args.warmup, w["dt"]
    tuning = dict(kernel=args.kernel, warps_per_cta=args.warps, use_graph=args.graph)
    options = {"qdiag": args.qdiag, "hermitian": args.herm, "resident": args.resident, "rk13": args.rk13,
               "prefetch": args.prefetch, "dynsched": args.dynsched, "packed": args.packed}
    in_bytes = sum(np.asarray(w[k]).nbytes for k in
                   ("rho0", "system", "system_dipole", "coupling", "coupling_dipole", "expn", "etal",
                    "etar", "etaa", "mode"))
    out_bytes = (K + 1) * n * n * 16
    fs = None
    if w["pulse_system_func"] is not None:
        from pyqed_b200.heom.deom import sample_pulse
        fs = sample_pulse(w["pulse_system_func"], dt, max(K, Wm))

    if args.batch > 1:
        return run_batch_arm(args, w, world, rank, local, multi, order, tuning, options)
    if not multi:
        bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
        solver = DEOMSolver(w["system"], w["system_dipole"], bath, w["coupling"], w["coupling_dipole"],
                            w["pulse_system_func"], w["pulse_coupling_func"], lmax=lmax, device=local,
                            order=order, alias_rho0=False)
        solver.tuning, solver.options = tuning, options
        # ---- e2e: the public call with host buffers (first call also builds the plan)
        t0 = time.perf_counter()
        solver.run(w["rho0"].copy(), dt, 1)
        setup_s = time.perf_counter() - t0
        plan = solver._plan
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, traj = solver.run(w["rho0"].copy(), dt, K)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        rho_end = np.asarray(traj[-1])
        owned = plan.nmax
        halo_bytes = 0

        def propagate(nsteps):
            plan.propagate(dt, nsteps, None if fs is None else fs[None, :nsteps], None, None, 0)
        plan.set_state(w["rho0"][None])
    else:
        from pyqed_b200.heom.sharded import ShardedDEOM, DistTransport
        t0 = time.perf_counter()
        sh = ShardedDEOM(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"],
                         w["expn"], w["etal"], w["etar"], w["etaa"], w["mode"], lmax,
                         DistTransport(), device=local, order=order, options=options, tuning=tuning,
                         peer_push={-1: None, 0: False, 1: True}[args.push],
                         fused_push={-1: None, 0: False, 1: True}[args.fused],
                         native={-1: None, 0: False, 1: True}[args.native],
                         rebalance={-1: None, 0: False, 1: True, 2: "force"}[args.rebalance])
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        sh.run(w["rho0"], dt, 1, w["pulse_system_func"], w["pulse_coupling_func"])
        setup_s = time.perf_counter() - t0
        plan = sh.plan
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, traj = sh.run(w["rho0"], dt, K, w["pulse_system_func"], w["pulse_coupling_func"])
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        rho_end = np.asarray(traj[-1])
        owned = sh.hi - sh.lo
        halo_bytes = sh.halo_bytes_per_stage()

        def propagate(nsteps):
            sh.propagate(dt, nsteps, None, None if fs is None else fs[:nsteps], None)
        sh.set_state(w["rho0"])
        if sh.native:
            sh.check_barriers()
    nmax = plan.nmax

    # ---- device-resident timing
    propagate(Wm)
    plan.synchronize()
    launches0 = plan.launch_count()
    plan.stage_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    propagate(K)
    ev1.record()
    torch.cuda.synchronize()
    if multi:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else {}
    stage_ms, stage_n = plan.stage_timing(False)
    if stage_n == 0 and plan.info("packed_steps") > 0:
        # kernel 7 runs the whole propagation in one call (no per-stage events): its stage time
        # is the step time / 4, pack and unpack passes included
        stage_ms, stage_n = ms, 4 * K
    launches = plan.launch_count() - launches0
    per_rank = None
    if multi:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt[0])
        mine = torch.tensor([owned, halo_bytes, stage_ms / max(stage_n, 1)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        # NVLink rate a rank's halo rows arrive with, averaged over a whole stage period (kernel +
        # barrier): bytes its pool receives per stage / (step time / 4)
        stage_s = ms * 1e-3 / K / 4.0
        per_rank = [dict(owned_ados=int(x[0]), halo_bytes_per_stage=int(x[1]), avg_stage_kernel_ms=float(x[2]),
                         nvlink_in_gbs=float(x[1]) / stage_s / 1e9)
                    for x in allr]

    tr = complex(np.trace(rho_end))
    # correctness evidence: (1) the timed path on the reference's own fixture; (2) rho_sys after
    # `steps` RK4 steps of the timed e2e call - at N > 1 against the committed 1-GPU values
    # (1-GPU-vs-N-GPU agreement at full depth); the trace alone would be blind to a
    # broken halo exchange (it is conserved whatever the neighbour rows hold)
    check = {"trace_rho_sys": [tr.real, tr.imag],
             "rho_sys_final": np.ascontiguousarray(rho_end.reshape(-1)).view(np.float64).reshape(-1, 2).tolist()}
    try:
        check["reference_fixture"] = fixture_parity(multi, DistTransport() if multi else None, local, order,
                                                    tuning, options, {-1: None, 0: False, 1: True}[args.native])
        check["reference_fixture"]["ok"] = bool(check["reference_fixture"]["max_abs_diff_trajectory"] < 1e-10
                                                and check["reference_fixture"]["max_abs_diff_all_ados"] < 1e-10)
    except Exception as exc:  # noqa: BLE001
        check["reference_fixture"] = {"error": repr(exc)[-300:], "ok": False}
    ref1 = rho_reference(args.workload, K)
    if ref1 is not None:
        check["max_abs_diff_vs_n1"] = float(np.max(np.abs(rho_end.reshape(-1) - ref1)))
        check["n1_reference"] = "profiles/r02_rho_sys_n1.json (1 GPU, same workload and steps)"
    else:
        check["max_abs_diff_vs_n1"] = None
    if args.save_rho_ref and rank == 0 and not multi:
        os.makedirs(os.path.dirname(args.save_rho_ref) or ".", exist_ok=True)
        try:
            with open(args.save_rho_ref) as fh:
                store = json.load(fh)
        except Exception:
            store = {}
        store.setdefault(args.workload, {})[str(K)] = check["rho_sys_final"]
        with open(args.save_rho_ref, "w") as fh:
            json.dump(store, fh)
    if rank == 0:
        peak, peak_src = peaks()
        bytes_per_step = 256.0 * n * n * nmax          # 16 array passes x 16 B x N^2
        value = nmax * K / (ms * 1e-3)
        avg_launch_ms = stage_ms / max(stage_n, 1)
        # dominant kernel on this rank: its share of the algorithmic bytes / its mean duration
        resident = plan.info("resident_launches") > 0 or plan.info("dataflow_launches") > 0
        bytes_per_launch = 256.0 * n * n * owned * (K if resident else 0.25)  # resident: one launch = K steps
        achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if stage_n else None
        state_mb = nmax * n * n * 16 / 1e6
        if plan.info("dataflow_dense_launches") > 0:
            kname = "stage_dataflow_tma_kernel<DENSE_H>"
        elif plan.info("dataflow_tma_launches") > 0:
            kname = "stage_dataflow_tma_kernel"
        elif plan.info("dataflow_launches") > 0:
            kname = "stage_dataflow_kernel"
        elif plan.info("packed_steps") > 0:
            kname = "stage_rows_sym_kernel<PACKED>"
        elif plan.info("sym_launches") > 0:
            kname = "stage_rows_sym_kernel"
        elif plan.info("resident_launches") > 0:
            kname = {4: "resident_cluster_kernel", 5: "resident_elem_kernel"}[plan.info("resident_kind")]
        else:
            k = args.kernel if args.kernel in (1, 2, 3) else (2 if n > 8 else (3 if plan.info("qdiag") else 1))
            kname = {1: "stage_rows_kernel", 2: "stage_generic_kernel", 3: "stage_rows_async_kernel"}[k]
        order_name = ["reference", "lexicographic", "blocked lexicographic"][order]
        fp64 = None
        if n > 8:
            # compute-shaped configuration (N = 32): algorithmic flops of SURVEY 8d's structured path,
            # 4 [16 N^3 + 16 N nnz_Q M_eff] per ADO-step, against the DFMA rate measured on this GPU
            # type (tools/fp64_peaks.cu; DMMA gives the same rate, so the commutator stays on DFMA)
            nnz_q = int(sum(np.count_nonzero(q) for q in np.asarray(w["coupling"])))
            flops_step = 4.0 * (16.0 * n ** 3 + 16.0 * n * nnz_q) * nmax
            peak_tf, src = 36.7, "fallback"
            try:
                with open(os.path.join(ROOT, "profiles", "r02_fp64_peaks.json")) as fh:
                    peak_tf, src = float(json.load(fh)["dfma_tflops"]), "measured (profiles/r02_fp64_peaks.json, DFMA; DMMA m8n8k4 / m16n8k8 give the same)"
            except Exception:
                pass
            ach = flops_step * K / (ms * 1e-3) / 1e12
            fp64 = {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "peak_source": src, "traffic": None, "kernel": kname,
                    "algorithmic_flops_per_step": flops_step,
                    "note": "H and Q_m are applied through their sparsity lists (and, in kernel 9, only W of "
                            "k = W + W^dagger is formed), so the executed flops are far below this dense-commutator "
                            "count; the run is bound by the flag/fetch latency between ADOs and by shared-memory "
                            "wavefronts (ncu: profiles/r02_kernel9_polariton32_ncu.txt; kernel 8: "
                            "profiles/r02_kernel8_polariton32_ncu.txt)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if multi else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": args.workload, "nsys": n, "nind": nind, "nmod": int(w["coupling"].shape[0]),
                "lmax": lmax, "n_ado": nmax, "dt": dt, "storage_order": ["reference", "lexicographic", "blocked lexicographic"][order],
                "state_mb_per_array": state_mb,
                "fast_paths": {"diagonal_Q": plan.info("qdiag"), "hermitian_ados": plan.info("hermitian"),
                               "real_H": plan.info("real_h"),
                               "rk4_array_passes_per_step": 13 if plan.info("rk_scheme") else 16},
                "l2": ("inputs larger than L2 (4 arrays of %.0f MB); no flush needed" % state_mb) if state_mb > 200
                      else "state is cache-resident by construction (time stepping re-reads its own output); no flush",
                "parallelism": "single GPU" if not multi else
                               (f"hierarchy sharded over {world} GPUs (contiguous ranges of the blocked lexicographic "
                                f"order, cost balanced); one exchange of neighbour rows per RK stage "
                                + ("(rank-local arrays = own ADOs + halo row pool; rows stored into the peers' pools "
                                   "over NVLink by the stage kernel's epilogue, CUDA-IPC mappings, flag barrier in "
                                   "peer memory, whole run in one C call)" if sh.native else
                                   (("(rows stored into peer memory over NVLink by the stage kernel's epilogue"
                                     if sh.fused else "(rows stored into peer memory over NVLink by a push kernel")
                                    + " + symmetric-memory barrier)"
                                    if sh.symm is not None else "(NCCL all_to_all_single)"))),
                "setup_s_first_call": setup_s,
                "setup_breakdown_s": (dict(getattr(sh, "timings", {}), construct_total=build_s) if multi else None),
                "state_mb_per_rank": (sh.state_nbytes / 1e6 if (multi and sh.native) else 4 * state_mb),
            },
            "clocks": clocks,
            "e2e": {"value": nmax * K / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": in_bytes / K, "d2h_bytes_per_step": out_bytes / K,
                    "call": ("DEOMSolver.run(rho0, dt, nt=steps)" if not multi else "ShardedDEOM.run(rho0, dt, nt=steps)")
                            + " with host arrays, plan cached"},
            "gpu_launches": launches,
            "roofline": fp64 if fp64 is not None else
                        {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": measured_traffic(args.workload, kname, order_name) if not multi else None,
                         "peak_source": peak_src, "kernel": kname,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": avg_launch_ms, "launches_timed": stage_n,
                         "whole_job_gbs": bytes_per_step * K / (ms * 1e-3) / 1e9,
                         "whole_job_frac_of_aggregate_peak": bytes_per_step * K / (ms * 1e-3) / 1e9 / (peak * world)},
            "check": check,
        }
        if per_rank:
            line["ranks"] = per_rank
        def aux(key, leg):
            """Auxiliary legs are reported beside the headline; a failure in one of them is
            recorded, never allowed to take the measured line down."""
            try:
                line[key] = leg()
            except Exception as exc:  # noqa: BLE001
                line[key] = {"error": repr(exc)[-300:]}
        if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_cpu:
            # BASELINE.json configs[1] (330 ADOs, cache resident) measured beside the
            # headline workload so that both readings of "the configuration" are on record
            aux("other_workloads", lambda: {"fmo7_K7_L4": small_workload_leg("fmo7_K7_L4", local),
                                            "polariton32_K4_L6": small_workload_leg("polariton32_K4_L6", local)})
        if world == 1 and not args.no_cpu:
            aux("cpu_baseline", lambda: cpu_native_leg(args.workload)[0])
            aux("cpu_baseline_python_loop", lambda: cpu_reference_leg(args.workload, budget_s=8.0)[0])
            aux("cpu_baseline_batched", lambda: cpu_batched_leg(args.workload))
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


def run_batch_arm(args, w, world, rank, local, multi, order, tuning, options):
    """Waiting-time scan (BASELINE configs[4]): ``--batch B`` trajectories of the same hierarchy that
    differ in their field tables, through ``DEOMSolver.run_batch``.  The reference runs them one
    after the other (B serial ``DEOMSolver.run`` calls, deom.py:1072).  Several GPUs: replicas only -
    the batch axis is split over the ranks, no communication."""
    import torch
    import torch.distributed as dist
    from pyqed_b200.heom import DEOMSolver, Bath
    from pyqed_b200.heom.deom import sample_pulse
    if args.workload != "aggregate7_K6_L6":
        raise SystemExit("--batch is defined for the waiting-time workload aggregate7_K6_L6")
    B = args.batch
    mine = list(range(rank, B, world))           # waiting indices of this rank
    n, nind, lmax = int(w["system"].shape[0]), int(len(w["expn"])), int(w["lmax"])
    K, Wm, dt = args.steps, args.warmup, w["dt"]
    fields = [W.aggregate_2des(lmax=lmax, waiting_index=b)["pulse_system_func"] for b in mine]
    bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
    solver = DEOMSolver(w["system"], w["system_dipole"], bath, w["coupling"], w["coupling_dipole"],
                        None, None, lmax=lmax, device=local, order=order, alias_rho0=False, shard=False)
    solver.tuning, solver.options = tuning, options
    rho0s = [w["rho0"]] * len(mine)
    t0 = time.perf_counter()
    solver.run_batch(rho0s, dt, 2, p1=w["observable"], pulse_system_funcs=fields)
    setup_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, sig = solver.run_batch(rho0s, dt, K, p1=w["observable"], pulse_system_funcs=fields)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    plan = solver._plan
    fs = np.stack([sample_pulse(f, dt, max(K, Wm)) for f in fields])
    plan.set_state(np.stack(rho0s))
    plan.propagate(dt, Wm, fs[:, :Wm], None, None, 0)
    plan.synchronize()
    launches0 = plan.launch_count()
    plan.stage_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    plan.propagate(dt, K, fs[:, :K], None, None, 0)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else {}
    stage_ms, stage_n = plan.stage_timing(False)
    launches = plan.launch_count() - launches0
    if multi:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt[0])
    if rank == 0:
        peak, peak_src = peaks()
        nmax = plan.nmax
        value = B * nmax * K / (ms * 1e-3)
        resident = plan.info("resident_launches") > 0
        kname = ({4: "resident_cluster_kernel", 5: "resident_elem_kernel"}[plan.info("resident_kind")] if resident
                 else "stage_rows_async_kernel")
        algo = 256.0 * n * n * nmax * len(mine) * K
        achieved = algo / (ms * 1e-3) / 1e9
        in_bytes = sum(np.asarray(w[k]).nbytes for k in ("system", "system_dipole", "coupling", "coupling_dipole",
                                                         "expn", "etal", "etar", "etaa", "mode")) + \
            B * (w["rho0"].nbytes + 3 * K * 8)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if multi else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "batch": B, "nsys": n, "nind": nind, "lmax": lmax, "n_ado": nmax,
                       "dt": dt, "waiting_times": "T_b = 0.05 b, b = 0..%d (pump/probe field tables)" % (B - 1),
                       "l2": "state is resident in distributed shared memory / L2 by construction; no flush",
                       "parallelism": "single GPU" if not multi else
                                      f"replicas only: the {B} trajectories are split over {world} GPUs, no communication",
                       "setup_s_first_call": setup_s},
            "clocks": clocks,
            "e2e": {"value": B * nmax * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": in_bytes / K,
                    "d2h_bytes_per_step": B * (K + 1) * 16 / K,
                    "call": "DEOMSolver.run_batch(rho0s, dt, nt=steps, p1=mu, pulse_system_funcs=[...]) with host arrays"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": kname,
                         "note": "the whole batch (%.0f MB of state) lives in distributed shared memory for all steps "
                                 "(one launch): the HBM fraction only says how far below the streaming bound the "
                                 "latency/issue-bound resident kernel runs" % (B * nmax * n * n * 16 * 4 / 1e6)
                                 if resident else "per-stage launches",
                         "launches_timed": stage_n},
            "check": {"signal_last": [float(np.real(sig[0][-1])), float(np.imag(sig[0][-1]))]},
        }
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--order", type=int, default=-1, help="storage order: 0 reference, 1 lexicographic; default 0 on one GPU, 1 when sharded")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--qdiag", type=int, default=-1)
    ap.add_argument("--herm", type=int, default=-1)
    ap.add_argument("--rk13", type=int, default=-1, help="0: accumulator RK4 (16 passes), default: difference form (13)")
    ap.add_argument("--resident", type=int, default=-1, help="0 off, 4 force kernel 4, default auto (kernel 5)")
    ap.add_argument("--fused", type=int, default=-1, help="multi-GPU: stage kernel stores halo rows itself")
    ap.add_argument("--push", type=int, default=-1, help="multi-GPU halo: 1 peer-memory stores, 0 NCCL all_to_all")
    ap.add_argument("--batch", type=int, default=1, help="aggregate7_K6_L6: number of waiting times (config 5: 64)")
    ap.add_argument("--native", type=int, default=-1, help="multi-GPU: 1 require / 0 forbid the rank-local layout")
    ap.add_argument("--rebalance", type=int, default=-1, help="multi-GPU: 0 = keep the static cut of the ranges, 2 = re-cut even if the ranks are balanced")
    ap.add_argument("--save-rho-ref", default=None, help="1 GPU: merge rho_sys after `steps` steps into this JSON")
    ap.add_argument("--prefetch", type=int, default=0, help="kernel 7: double-buffered tiles fetched one group ahead")
    ap.add_argument("--dynsched", type=int, default=-1, help="kernels 6/7: 0 = static group stride, default: global work counter")
    ap.add_argument("--packed", type=int, default=-1, help="0: never use packed Hermitian storage (kernel 7)")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 20
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
