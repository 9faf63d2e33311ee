"""Development probe: one rank's share of an N-rank sharded run, with all ranks on ONE GPU (gloo
plumbing, host barrier per stage), so that a single rank can be put under ncu:

    for r in 0..N-1:  RANK=r WORLD_SIZE=N MASTER_ADDR=127.0.0.1 MASTER_PORT=29600 python tools/shard_probe.py
    (rank P wrapped as: ncu --set full -k regex:stage_rows_sym ... python tools/shard_probe.py)

The profiled kernel is exactly the shard a rank of an N-GPU run processes (same ranges, push tables
and pool reads); its peer stores go to mappings of buffers on the same device instead of NVLink."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyqed_b200 import workloads as W  # noqa: E402
from pyqed_b200.heom.sharded import ShardedDEOM, DistTransport  # noqa: E402

steps = int(os.environ.get("PROBE_STEPS", "2"))
lmax = int(os.environ.get("PROBE_LMAX", "8"))
dist.init_process_group("gloo")
torch.cuda.set_device(0)
w = W.fmo(lmax=lmax, n_matsubara=2)
sh = ShardedDEOM(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"], w["etal"],
                 w["etar"], w["etaa"], w["mode"], lmax, DistTransport(), device=0, order=2, native=True,
                 rebalance=False)
sh.set_state(w["rho0"])
sh.propagate(w["dt"], steps)
sh.plan.synchronize()
print(f"rank {sh.rank}: owned {sh.hi - sh.lo}, pool rows {sh.need64.numel()}, pushed rows {sh.halo.send_items.numel()}", flush=True)
sh.close()
dist.destroy_process_group()
