"""Worker for the multi-rank tests (launched once per rank by test_sharded.py).

mode "cpu":  gloo, no GPU - halo bookkeeping and a sharded RK4 propagation in
             which the oracle's NumPy right-hand side stands in for the stage
             kernel (the oracle is test infrastructure; allowed here only).
mode "gpu":  the real ShardedDEOM on CUDA.  With --backend gloo several ranks may
             share one GPU (buffers are staged through the host).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pyqed_b200.heom import sharded as S  # noqa: E402


def cpu_mode(args):
    from oracle.deom_oracle import DeomOracle
    from pyqed_b200 import workloads as W
    tr = S.DistTransport()
    rank, world = tr.rank, tr.world
    w = W.fmo(lmax=3, n_matsubara=1, sites=4)          # K = 8, 495 ADOs, projector Q
    o = DeomOracle(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"],
                   w["etal"], w["etar"], w["etaa"], w["mode"], w["lmax"])
    n, K, nmax = o.nsys, o.nind, o.nmax
    slot_of_id = S._lex_rank(o.keys, o.lmax)
    id_of_slot = np.argsort(slot_of_id)
    # CSR links in slot order, same record format as the device tables
    nbr, meta, ptr = [], [], [0]
    for s in range(nmax):
        i = id_of_slot[s]
        for k in range(K):
            for tbl in (o.minus, o.plus):
                j = tbl[i, k]
                if j >= 0:
                    nbr.append(slot_of_id[j])
                    meta.append(int(o.mode[k]) << 16)   # projector on site mode[k]: row = mode
        ptr.append(len(nbr))
    nbr = torch.tensor(nbr, dtype=torch.int32)
    meta = torch.tensor(meta, dtype=torch.int32)
    bounds = S.cost_balanced_bounds(np.array(ptr), world)
    assert bounds[0] == 0 and bounds[-1] == nmax and all(b1 > b0 for b0, b1 in zip(bounds, bounds[1:]))
    lo, hi = bounds[rank], bounds[rank + 1]
    need = S.needed_items(nbr[ptr[lo]:ptr[hi]], meta[ptr[lo]:ptr[hi]], lo, hi, True)
    halo = S.HaloPlan(bounds, rank, world, need, True, tr)

    def exchange(arr):
        """arr: [nmax, n, n] complex in slot order; fills the needed foreign rows."""
        si = halo.send_items.numpy()
        send = arr[si >> 3, si & 7, :].copy()
        sendbuf = torch.from_numpy(send.view(np.float64).reshape(-1))
        recvbuf = torch.empty(len(halo.need) * 2 * n, dtype=torch.float64)
        tr.all_to_all(sendbuf, halo.send_counts, recvbuf, halo.recv_counts, 2 * n)
        ni = halo.need.numpy()
        arr[ni >> 3, ni & 7, :] = recvbuf.numpy().view(np.complex128).reshape(-1, n)

    # 1. bookkeeping: after one exchange every needed row holds the owner's values
    truth = (np.arange(nmax)[:, None, None] * 100 + np.arange(n)[None, :, None] * 10
             + np.arange(n)[None, None, :]).astype(np.complex128) * (1 + 0.5j)
    mine = np.full_like(truth, np.nan)
    mine[lo:hi] = truth[lo:hi]
    exchange(mine)
    ni = need.numpy()
    assert np.array_equal(mine[ni >> 3, ni & 7], truth[ni >> 3, ni & 7])
    assert np.isnan(mine[:lo]).sum() + np.isnan(mine[hi:]).sum() > 0   # only the needed rows moved

    # 2. sharded RK4: owner-computes with the oracle RHS, halo exchange per stage
    dt, nt = w["dt"], 5
    owned_ids = id_of_slot[lo:hi]

    def rhs_owned(y_slot):
        """k for the owned slots only; foreign rows that are not in the halo are
        NaN and must never contaminate the result."""
        y_id = y_slot[slot_of_id]                      # id order view for the oracle
        H, Q = o.operators_at(0.0)
        out = np.zeros((hi - lo, n, n), np.complex128)
        for a, i in enumerate(owned_ids):
            acc = -np.sum(o.keys[i] * o.expn) * y_id[i] - 1j * (H @ y_id[i] - y_id[i] @ H)
            for k in range(K):
                m = int(o.mode[k])
                for tbl, cl, cr in ((o.minus, 1j * np.sqrt(o.keys[i, k]) / np.sqrt(o.etaa[k]) * o.etal[k],
                                     1j * np.sqrt(o.keys[i, k]) / np.sqrt(o.etaa[k]) * o.etar[k]),
                                    (o.plus, 1j * np.sqrt(o.keys[i, k] + 1) * np.sqrt(o.etaa[k]),
                                     1j * np.sqrt(o.keys[i, k] + 1) * np.sqrt(o.etaa[k]))):
                    j = tbl[i, k]
                    if j < 0:
                        continue
                    row = y_id[j][m, :]                # the only part of the neighbour that is read
                    qa = np.zeros((n, n), np.complex128)
                    qa[m, :] = row                     # Q rho'
                    aq = np.zeros((n, n), np.complex128)
                    aq[:, m] = np.conj(row)            # rho' Q via Hermiticity of the ADOs
                    acc = acc - (cl * qa - cr * aq)
            out[a] = acc
        return out

    y = np.full((nmax, n, n), np.nan, dtype=np.complex128)
    y[:] = 0
    y[slot_of_id[0]] = w["rho0"]
    foreign = np.ones(nmax, bool)
    foreign[lo:hi] = False

    def stage_input(base, k, coef):
        s = np.full((nmax, n, n), np.nan, dtype=np.complex128)
        s[lo:hi] = base[lo:hi] + coef * k
        exchange(s)
        return s

    traj = [y[slot_of_id[0]].copy()]
    for _ in range(nt):
        k1 = rhs_owned(y)
        k2 = rhs_owned(stage_input(y, k1, dt / 2))
        k3 = rhs_owned(stage_input(y, k2, dt / 2))
        k4 = rhs_owned(stage_input(y, k3, dt))
        ynew = np.full((nmax, n, n), np.nan, dtype=np.complex128)
        ynew[lo:hi] = y[lo:hi] + (k1 + 2 * k2 + 2 * k3 + k4) * dt / 6
        exchange(ynew)
        y = ynew
        own0 = lo <= slot_of_id[0] < hi
        r0 = torch.from_numpy((y[slot_of_id[0]] if own0 else np.zeros((n, n), np.complex128)).view(np.float64).copy())
        dist.all_reduce(r0)
        traj.append(r0.numpy().view(np.complex128).reshape(n, n))
    _, ref = o.run(w["rho0"], dt, nt)
    err = max(np.max(np.abs(a - b)) for a, b in zip(traj, ref))
    assert err < 1e-12, err
    print(f"rank {rank}: cpu sharded ok, bounds {bounds}, halo items {len(need)}, err {err:.1e}", flush=True)


def gpu_mode(args):
    from conftest import golden
    tr = S.DistTransport()
    dev = int(os.environ.get("LOCAL_RANK", "0")) if args.backend == "nccl" else 0
    torch.cuda.set_device(dev)
    for name in args.cases.split(","):
        g = golden(name)
        sh = S.ShardedDEOM(g["system"], g["system_dipole"], g["coupling"], g["coupling_dipole"],
                           g["expn"], g["etal"], g["etar"], g["etaa"], g["mode"], int(g["lmax"]),
                           tr, device=dev, order=args.order,
                           tuning=dict(kernel=args.kernel, warps_per_cta=0, use_graph=0),
                           peer_push={-1: None, 0: False, 1: True}[args.push],
                           fused_push={-1: None, 0: False, 1: True}[args.fused],
                           native={-1: None, 0: False, 1: True}[args.native],
                           rebalance={0: None, 1: "force"}[args.rebalance])
        nt = int(g["nt"])
        from conftest import pulse_from_samples
        dt = float(g["dt"])
        ts, traj = sh.run(g["rho0"], dt, nt, pulse_from_samples(g["pulse_system"], dt),
                          pulse_from_samples(g["pulse_coupling"], dt))
        err = np.max(np.abs(traj - g["traj"]))
        assert err < 1e-12, (name, err)
        if "ados_final" in g:
            ados = sh.gather_ados()
            e2 = np.max(np.abs(ados - g["ados_final"]))
            assert e2 < 1e-12, (name, e2)
        if 0 < sh.hi - sh.lo < sh.nmax:
            assert sum(sh.halo.recv_counts) > 0   # a proper sub-range always has foreign neighbours
        if args.kernel == 6:   # did the stages really go through kernel 6 where it applies?
            print(f"rank {tr.rank}: {name} kernel6 stage launches {sh.plan.info('sym_launches')} of {4 * nt}", flush=True)
        items = sh.halo_bytes_per_stage() // (sh.elems * 8)
        if sh.native:
            sh.check_barriers()
        if args.rebalance and sh.native:
            print(f"rank {tr.rank}: {name} rebalanced={sh.timings['rebalance']['changed']} bounds {sh.bounds}", flush=True)
        print(f"rank {tr.rank}: {name} native={sh.native} packed={sh.plan.info('shard_packed')} "
              f"push={sh.symm is not None} fused={sh.fused} ok (owned {sh.hi - sh.lo} of {sh.nmax}, halo items "
              f"{items}, row items {sh.row_items}, err {err:.1e})", flush=True)
        sh.close()


def solver_mode(args):
    """``DEOMSolver.run`` itself under a multi-rank job: every rank calls it with the same arguments
    and gets the reference's trajectory, observable, keys and ADOs (drop-in API, sharded inside)."""
    from conftest import golden, pulse_from_samples
    from pyqed_b200.heom import DEOMSolver, Bath
    rank = dist.get_rank()
    dev = int(os.environ.get("LOCAL_RANK", "0")) if args.backend == "nccl" else 0
    torch.cuda.set_device(dev)
    for name in args.cases.split(","):
        g = golden(name)
        dt, nt = float(g["dt"]), int(g["nt"])
        bath = Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"])
        s = DEOMSolver(g["system"], g["system_dipole"], bath, g["coupling"], g["coupling_dipole"],
                       pulse_from_samples(g["pulse_system"], dt), pulse_from_samples(g["pulse_coupling"], dt),
                       lmax=int(g["lmax"]), device=dev, shard=True)
        p1 = g["p1"] if "p1" in g and g["p1"].ndim == 2 else None
        rho0 = g["rho0"].copy()
        t_save, out = s.run(rho0, dt, nt, p1)
        ref = g["traj"]
        err = np.max(np.abs(np.asarray(out) - ref))
        assert err < 1e-12, (name, err)
        assert np.allclose(t_save, np.arange(nt + 1) * dt)
        if "keys" in g:
            assert np.array_equal(s.keys, g["keys"])
        if "ados_final" in g:
            assert np.max(np.abs(s.ddos - g["ados_final"])) < 1e-12
            assert np.max(np.abs(rho0 - g["ados_final"][0])) < 1e-12     # rho0 aliasing of the reference
        sh = s._sharded[1]
        print(f"rank {rank}: {name} solver ok native={sh.native} p1={p1 is not None} err {err:.1e}", flush=True)
        sh.close()
    # a batch of trajectories is split over the ranks (replicas only) and gathered
    g = golden("deom_aggregate_L3_T0")
    g2 = golden("deom_aggregate_L3_T37")
    dt, nt = float(g["dt"]), int(g["nt"])
    bath = Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"])
    s = DEOMSolver(g["system"], g["system_dipole"], bath, g["coupling"], g["coupling_dipole"], lmax=int(g["lmax"]),
                   device=dev)
    fields = [pulse_from_samples(x["pulse_system"], dt) for x in (g, g2, g)]
    _, sig = s.run_batch([g["rho0"]] * 3, dt, nt, p1=g["p1"], pulse_system_funcs=fields)
    assert sig.shape == (3, nt + 1)
    assert max(np.max(np.abs(sig[0] - g["traj"])), np.max(np.abs(sig[1] - g2["traj"])),
               np.max(np.abs(sig[2] - g["traj"]))) < 1e-12
    print(f"rank {rank}: batch split ok", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["cpu", "gpu", "solver"])
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--cases", default="deom_fmo_K21_L2")
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--push", type=int, default=-1)
    ap.add_argument("--fused", type=int, default=-1)
    ap.add_argument("--kernel", type=int, default=0, help="stage kernel (tuning), e.g. 6")
    ap.add_argument("--rebalance", type=int, default=0, help="1: re-cut the ranges from measured stage times")
    ap.add_argument("--native", type=int, default=-1, help="rank-local arrays + fused peer stores: 1 require, 0 never")
    a = ap.parse_args()
    dist.init_process_group(a.backend)
    try:
        {"cpu": cpu_mode, "gpu": gpu_mode, "solver": solver_mode}[a.mode](a)
    finally:
        dist.destroy_process_group()
