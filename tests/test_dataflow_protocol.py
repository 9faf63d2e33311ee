"""The synchronisation protocol of the persistent dataflow kernels (kernels 8 and 9,
pyqed_b200/csrc/heom_dataflow.cuh, heom_dataflow_tma.cuh), checked on the CPU as a randomised
interleaving model.

Every ADO is an agent that walks through the RK4 stages of all steps on its own.  Before stage s it waits
until each neighbour's publication counter shows the output of stage s-1, reads those outputs from one of
three buffers, and publishes its own output into another one:

    stage of a step :   0     1     2     3
    reads           :   P0    P1    P2    P1          (kernel 8: Y, SA, SB, SA)
    writes          :   P1    P2    P1    P0          (kernel 8: SA, SB, SA, Y)

with only THREE buffers for four stages.  The claim in the kernels' comments: an agent overwrites a buffer
only in a stage that it can enter after all its neighbours have finished the stage in which they read the
old contents (the neighbour relation is symmetric).  The model splits every stage into micro-steps (flag
checks, one read per neighbour, write, publish), lets a random scheduler interleave the agents, tags every
buffer with the publication number it holds, and asserts that every read finds exactly the tag it needs -
never a stale one, never one that was overwritten early.  The same argument shows that two buffers
(ping-pong) would do as far as the neighbours are concerned - kernel 8 needs the third because Y is also
its owner's start-of-step state, kernel 9 keeps the rotation of kernel 8 - and that one buffer does not:
the model must catch that."""
import random

import numpy as np
import pytest

from oracle.deom_oracle import build_keys, build_neighbours, pascal_table

READS = (0, 1, 2, 1)
WRITES = (1, 2, 1, 0)


def _graph(nind, lmax):
    tab = pascal_table(nind, lmax)
    keys = build_keys(nind, lmax, tab)
    minus, plus = build_neighbours(keys, lmax, tab)
    nb = []
    for n in range(len(keys)):
        s = {int(x) for x in minus[n] if x >= 0} | {int(x) for x in plus[n] if x >= 0}
        nb.append(sorted(s))
    for n, lst in enumerate(nb):           # symmetric, as the safety argument needs
        for m in lst:
            assert n in nb[m]
    return nb


def _run(nb, nstages, seed, reads=READS, writes=WRITES, nbuf=3):
    """Returns the number of reads checked; raises AssertionError on a protocol violation."""
    rng = random.Random(seed)
    n = len(nb)
    tag = np.zeros((n, nbuf), dtype=np.int64)      # publication number held by each buffer of each agent
    flag = np.zeros(n, dtype=np.int64)
    tag[:, reads[0]] = 1                            # the initial state, published by the prologue
    flag[:] = 1
    stage = [0] * n                                 # global stage index of the agent
    pc = [0] * n                                    # micro-step inside the stage: 0 wait, 1..len(nb) reads, then write, publish
    checked = 0
    live = list(range(n))
    while live:
        a = rng.choice(live)
        s = stage[a]
        if pc[a] == 0:                              # wait for the neighbours' flags
            if all(flag[m] >= s + 1 for m in nb[a]):
                pc[a] = 1
            continue
        k = pc[a] - 1
        if k < len(nb[a]):                          # read neighbour k's previous-stage output
            got = tag[nb[a][k], reads[s % 4]]
            assert got == s + 1, (a, nb[a][k], s, int(got))
            checked += 1
            pc[a] += 1
        elif k == len(nb[a]):                       # write the own output
            tag[a, writes[s % 4]] = s + 2
            pc[a] += 1
        else:                                       # publish
            flag[a] = s + 2
            stage[a] += 1
            pc[a] = 0
            if stage[a] == nstages:
                live.remove(a)
    return checked


@pytest.mark.parametrize("nind,lmax", [(2, 3), (3, 3), (4, 2)])
def test_three_buffer_rotation_is_safe_under_random_interleavings(nind, lmax):
    nb = _graph(nind, lmax)
    for seed in range(20):
        assert _run(nb, nstages=12, seed=seed) > 0


def test_ping_pong_would_be_safe_and_a_single_buffer_is_caught():
    nb = _graph(2, 3)
    for seed in range(20):
        assert _run(nb, nstages=12, seed=seed, reads=(0, 1, 0, 1), writes=(1, 0, 1, 0), nbuf=2) > 0
    # one buffer: an agent that is one stage ahead overwrites what a slower neighbour still has to read
    caught = 0
    for seed in range(20):
        try:
            _run(nb, nstages=12, seed=seed, reads=(0, 0, 0, 0), writes=(0, 0, 0, 0), nbuf=1)
        except AssertionError:
            caught += 1
    assert caught > 0
