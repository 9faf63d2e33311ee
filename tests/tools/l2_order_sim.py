"""Offline estimate of gather locality for different ADO storage orders.

Models the stage kernel's sweep over the hierarchy in storage order and an L2
of given capacity with an LRU-by-age approximation (a line hits if fewer than
`cap` bytes were filled since its last touch).  Used to choose the storage
order before spending GPU time.  Development/test tooling: it uses the oracle's
index tables and is not part of the product path.
"""
import sys, time
import numpy as np
import numba as nb
sys.path.insert(0, '.')
from oracle import deom_oracle as DO


def tables(K, L):
    tab = DO.pascal_table(K, L)
    keys = DO.build_keys(K, L, tab).astype(np.int16)
    nmax = len(keys)
    minus = np.full((nmax, K), -1, np.int32)
    plus = np.full((nmax, K), -1, np.int32)
    tier = keys.sum(axis=1)
    run = np.cumsum(keys.astype(np.int64), axis=1)
    cols = np.arange(K)
    for k in range(K):
        # id(n -+ e_k): s_i changes by -+1 for i >= k
        sel = keys[:, k] > 0
        r = run[sel].copy(); r[:, k:] -= 1
        minus[sel, k] = tab[r + cols, cols + 1].sum(axis=1)
        sel = tier < L
        r = run[sel].copy(); r[:, k:] += 1
        plus[sel, k] = tab[r + cols, cols + 1].sum(axis=1)
    return keys, minus, plus, tier


@nb.njit(cache=True)
def simulate(order, slot_of_id, minus, plus, cap_bytes, own_bytes, nbr_bytes, stream_bytes):
    n = order.shape[0]
    K = minus.shape[1]
    last = np.full(n, -1e30)
    clock = 0.0
    miss_own = 0; miss_nbr = 0; acc_nbr = 0
    for s in range(n):
        i = order[s]
        if clock - last[i] >= cap_bytes:
            miss_own += 1; clock += own_bytes
        last[i] = clock
        for k in range(K):
            for tbl in range(2):
                j = minus[i, k] if tbl == 0 else plus[i, k]
                if j < 0: continue
                acc_nbr += 1
                if clock - last[j] >= cap_bytes:
                    miss_nbr += 1; clock += nbr_bytes
                last[j] = clock
        clock += stream_bytes
    return miss_own, miss_nbr, acc_nbr


def order_of(name, keys, tier, K, L):
    n = len(keys)
    if name == 'ref':
        return np.arange(n, dtype=np.int32)
    if name == 'lex':       # dim 0 most significant
        return np.lexsort(tuple(keys[:, k] for k in range(K - 1, -1, -1))).astype(np.int32)
    if name == 'lexrev':    # dim K-1 most significant
        return np.lexsort(tuple(keys[:, k] for k in range(K))).astype(np.int32)
    if name == 'tierlex':
        return np.lexsort(tuple(keys[:, k] for k in range(K - 1, -1, -1)) + (tier,)).astype(np.int32)
    if name.startswith('blk'):   # blkH: lex on the first H dims, then tier-major (ref id) inside the block
        H = int(name[3:])
        return np.lexsort((np.arange(n),) + tuple(keys[:, k] for k in range(H - 1, -1, -1))).astype(np.int32)
    if name.startswith('sumblk'):  # sumblkH: by per-mode occupation (sum over each mode's dims) of the first H modes
        H = int(name[6:]); per = K // 7 if K % 7 == 0 else 1
        ms = [keys[:, m*per:(m+1)*per].sum(axis=1) for m in range(H)]
        return np.lexsort((np.arange(n),) + tuple(ms[::-1])).astype(np.int32)
    raise KeyError(name)


if __name__ == '__main__':
    K, L = int(sys.argv[1]), int(sys.argv[2])
    names = sys.argv[3].split(',')
    t0 = time.time()
    keys, minus, plus, tier = tables(K, L)
    n = len(keys)
    print(f'K={K} L={L} nmax={n} links={(minus>=0).sum()+(plus>=0).sum()} tables {time.time()-t0:.1f}s', flush=True)
    for name in names:
        order = order_of(name, keys, tier, K, L)
        slot_of_id = np.empty(n, np.int32); slot_of_id[order] = np.arange(n, dtype=np.int32)
        for cap_mb, stream in [(100, 0.0), (100, 3136.0), (40, 0.0)]:
            mo, mn, an = simulate(order, slot_of_id, minus, plus, cap_mb * 1e6, 784.0, 352.0, stream)
            gb = (mo * 784 + mn * 352) / 1e9
            print(f'{name:10s} cap={cap_mb:4d}MB stream={stream:6.0f}: own miss {mo/n:.3f} nbr miss {mn/an:.3f} '
                  f'-> yin DRAM read {gb:.2f} GB (ideal {n*784/1e9:.2f})', flush=True)
