"""Halo size per rank for contiguous partitions of a storage order (offline)."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests/tools")
from l2_order_sim import tables, order_of

K, L, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
names = sys.argv[4].split(',')
keys, minus, plus, tier = tables(K, L)
n = len(keys)
mode = np.repeat(np.arange(K // 3 if K % 3 == 0 else K), 3 if K % 3 == 0 else 1)
nl = (minus >= 0).sum(1) + (plus >= 0).sum(1)
for name in names:
    order = order_of(name, keys, tier, K, L)          # order[slot] = id
    slot_of_id = np.empty(n, np.int64); slot_of_id[order] = np.arange(n)
    # cost-balanced bounds: cost = 12 + links
    cost = (12 + nl[order]).cumsum()
    bounds = [0] + [int(np.searchsorted(cost, cost[-1] * r / world)) for r in range(1, world)] + [n]
    tot_items = 0; worst = 0; tot_ados = 0
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        ids = order[lo:hi]
        items = []
        for tbl in (minus, plus):
            nb = tbl[ids]                              # [cnt, K] neighbour ids
            sel = nb >= 0
            s = slot_of_id[nb[sel]]
            m = np.broadcast_to(mode[None, :], nb.shape)[sel]
            out = (s < lo) | (s >= hi)
            items.append(s[out] * 8 + m[out])
        items = np.unique(np.concatenate(items))
        ados = np.unique(items // 8)
        tot_items += len(items); tot_ados += len(ados); worst = max(worst, len(items) / (hi - lo))
        print(f'  {name} rank {r}: owned {hi-lo:8d} halo rows {len(items):9d} ({len(items)/(hi-lo):.2f}/owned) halo ADOs {len(ados):8d} ({len(ados)/(hi-lo):.2f}/owned)')
    print(f'{name}: total halo rows {tot_items} = {tot_items*112/1e6:.0f} MB/stage all ranks; per-rank worst {worst:.2f} rows/owned; full-ADO halo {tot_ados*784/1e6:.0f} MB')
