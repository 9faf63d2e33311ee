"""Multi-rank path: halo bookkeeping + sharded RK4 (world_size 2 and 3)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "_dist_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _launch(world, args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_sharded_rk4_gloo_cpu(world):
    out = _launch(world, ["cpu"])
    assert out.count("cpu sharded ok") == world


def test_cost_balanced_bounds():
    import numpy as np
    from pyqed_b200.heom.sharded import cost_balanced_bounds
    lp = np.concatenate([[0], np.cumsum([20] * 10 + [4] * 90)])
    b = cost_balanced_bounds(lp, 4, base_cost=12)
    assert b[0] == 0 and b[-1] == 100 and b == sorted(b)
    cost = [sum(12 + (lp[i + 1] - lp[i]) for i in range(b[r], b[r + 1])) for r in range(4)]
    assert max(cost) - min(cost) <= 2 * 32
    assert cost_balanced_bounds(lp, 1) == [0, 100]


@pytest.mark.gpu
@pytest.mark.parametrize("world,order", [(2, 1), (3, 1), (2, 0), (2, 2), (3, 2)])
def test_sharded_gpu_ranks_sharing_one_device(world, order):
    """The real kernels and pack/unpack with several ranks on one GPU (gloo
    transport staged through the host); projector, sigma_z and dense couplings."""
    out = _launch(world, ["gpu", "--backend", "gloo", "--order", str(order), "--native", "0", "--cases",
                          "deom_fmo_K21_L2,deom_fmo_K7_L4,deom_spin_boson_L10,deom_random4_herm"])
    assert out.count(" ok (owned") == 4 * world
    assert out.count("native=False") == 4 * world


@pytest.mark.gpu
@pytest.mark.parametrize("world,kernel", [(2, 0), (3, 0), (2, 6), (4, 7)])
def test_rank_local_layout_ranks_sharing_one_device(world, kernel):
    """The rank-local layout (``csrc/heom_shard.cu``): arrays of own ADOs + halo row pool,
    localized link table, rows stored into the peers' pools by the stage kernel's epilogue
    through CUDA-IPC mappings.  Several ranks share the one GPU of the test box (gloo for the
    plumbing, a host barrier per stage instead of the device flag barrier); the peer stores, the
    push tables and the pool reads are the ones a multi-GPU run uses.  FMO cases take the layout
    (kernel 7 = packed storage unless kernel 6 is asked for), the others fall back."""
    out = _launch(world, ["gpu", "--backend", "gloo", "--order", "2", "--kernel", str(kernel), "--cases",
                          "deom_fmo_K21_L3,deom_fmo_K21_L2,deom_fmo_K7_L4,deom_spin_boson_L10"])
    assert out.count(" ok (owned") == 4 * world
    assert out.count("native=True") == 3 * world and out.count("native=False") == world
    assert out.count("native=True packed=%d" % (0 if kernel == 6 else 1)) == 3 * world


@pytest.mark.gpu
def test_sharded_nccl_two_gpus():
    """The general exchange on two GPUs: NCCL all_to_all halo (push=0), peer-memory stores by a
    separate kernel (push=1, fused=0) and by kernel 3's epilogue (fused=1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for push, fused in [(0, 0), (1, 0), (1, 1)]:
        out = _launch(2, ["gpu", "--backend", "nccl", "--order", "2", "--push", str(push), "--fused", str(fused),
                          "--native", "0",
                          "--cases", "deom_fmo_K21_L3,deom_fmo_K7_L4,deom_spin_boson_L10,deom_random4_herm"])
        assert out.count(" ok (owned") == 8
        assert out.count(f"push={bool(push)}") == 8
        # the dense-Q case cannot use the fused path (it needs the diagonal-Q kernel)
        assert out.count("fused=True") == (6 if fused else 0)


@pytest.mark.gpu
def test_rank_local_layout_rebalanced_ranges():
    """Ranges re-cut from measured stage times (what large hierarchies do once at set-up): the
    layout is torn down, the link table rebuilt and localized for the new ranges; results unchanged."""
    out = _launch(3, ["gpu", "--backend", "gloo", "--order", "2", "--native", "1", "--rebalance", "1", "--cases",
                      "deom_fmo_K21_L3,deom_fmo_K7_L4"])
    assert out.count(" ok (owned") == 6 and out.count("native=True") == 6
    assert out.count("rebalanced=True") >= 3


@pytest.mark.gpu
def test_rank_local_layout_two_gpus():
    """The rank-local layout on two GPUs: peer stores over NVLink and the device flag barrier
    (``pyqed_heom_shard_propagate``: no host round trip inside a run); packed (kernel 7) and full
    (kernel 6) storage."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for kernel in (0, 6):
        out = _launch(2, ["gpu", "--backend", "nccl", "--order", "2", "--kernel", str(kernel), "--native", "1",
                          "--cases", "deom_fmo_K21_L3,deom_fmo_K21_L2,deom_fmo_K7_L4"])
        assert out.count(" ok (owned") == 6 and out.count("native=True") == 6


@pytest.mark.gpu
def test_deomsolver_run_shards_under_torch_distributed():
    """The drop-in class itself: ``DEOMSolver(..., shard=True).run`` on every rank of a 2-rank job
    (ranks sharing the test box's GPU) returns the reference's results - rank-local layout for the
    FMO case, general exchange for the pulsed / sigma_z / p1 cases - and ``run_batch`` splits its
    trajectories over the ranks."""
    out = _launch(2, ["solver", "--backend", "gloo", "--cases",
                      "deom_fmo_K21_L2,deom_spin_boson_L10,deom_example_L10_p1,deom_aggregate_L3_T37"])
    assert out.count(" solver ok") == 8
    assert out.count("native=True") == 2 and out.count("p1=True") >= 2
    assert out.count("batch split ok") == 2
