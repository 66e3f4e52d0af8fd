"""The schedule of a planned launch (tb_onchip.cuh: plan_fill / plan_deal, the code plan_kernel runs on the device),
computed on the host through tb_plan_schedule: no GPU involved.  What the solver kernels rely on:

* every chain is run exactly once over [1, end): whole, or as one head [1, k) and one tail [k, end);
* a tail is the LAST segment of its machine and its head is the FIRST segment of a machine with a LOWER index
  (blocks are dispatched in index order, so whoever waits for a hand-over waits for a block that is running);
* consequently the wait-for graph has no cycle;
* head and tail are at least 8 iterations long, no machine gets more than its share (the larger of the mean load and
  the longest chain) plus 16 iterations — 149 equal chains on 148 machines included — and without usable estimates the chains are dealt out whole, round-robin."""
import ctypes as C

import numpy as np
import pytest

from thirring2d_b200.lib import load_library

INF = 0x7FFFFFFF
CONVERGED, MAXITER = 0, 1   # TB_CG_CONVERGED, anything else = no usable estimate


def schedule(est, status, machines):
    lib = load_library()
    est = np.ascontiguousarray(est, dtype=np.int32)
    status = np.ascontiguousarray(status, dtype=np.int32)
    n = len(est)
    segs = np.full((n + machines, 4), -7, dtype=np.int32)
    lo = np.full(machines, -7, dtype=np.int32)
    hi = np.full(machines, -7, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    rc = lib.tb_plan_schedule(est.ctypes.data_as(ip), status.ctypes.data_as(ip), n, machines,
                              segs.ctypes.data_as(ip), lo.ctypes.data_as(ip), hi.ctypes.data_as(ip))
    assert rc == 0
    return segs, lo, hi


def check(est, machines):
    est = np.asarray(est, dtype=np.int64)
    n = len(est)
    segs, lo, hi = schedule(est, np.zeros(n), machines)
    assert np.all(lo >= 0) and np.all(hi >= lo) and hi.max() <= n + machines - 1
    # the machines' segment ranges tile [0, nseg) without overlap (machine M-1 first: it is machine 0 of the fill)
    order = np.argsort(lo, kind="stable")
    nseg = int(hi.max())
    covered = np.zeros(nseg, dtype=int)
    for b in range(machines):
        covered[lo[b]:hi[b]] += 1
    assert np.all(covered == 1)
    heads, tails, whole = {}, {}, set()
    load = np.zeros(machines, dtype=np.int64)
    for b in range(machines):
        for pos, s in enumerate(range(lo[b], hi[b])):
            c, k0, k1, _ = (int(v) for v in segs[s])
            assert 0 <= c < n and k0 >= 1 and k1 > k0
            if k0 == 1 and k1 == INF:
                assert c not in whole and c not in heads and c not in tails
                whole.add(c)
                load[b] += est[c]
            elif k0 == 1:
                assert c not in heads and c not in whole
                assert pos == 0, "a head must be the first segment of its machine"
                heads[c] = (b, k1)
                load[b] += k1 - 1
            else:
                assert k1 == INF and c not in tails and c not in whole
                assert s == hi[b] - 1, "a tail must be the last segment of its machine"
                tails[c] = (b, k0)
                load[b] += est[c] - (k0 - 1)
    assert set(heads) == set(tails) and whole | set(heads) == set(range(n))
    waits_for = {}
    for c, (bt, k0) in tails.items():
        bh, k1 = heads[c]
        assert k1 == k0, "the tail resumes where the head stops"
        assert bh < bt, "the head runs on a block with a lower index than the block that waits for it"
        assert k0 - 1 >= 8 and est[c] - (k0 - 1) >= 8, "no hand-over for a handful of iterations"
        waits_for[bt] = bh
    for b in waits_for:   # no cycle (indices strictly decrease along the chain of waits)
        seen, x = set(), b
        while x in waits_for:
            assert x not in seen
            seen.add(x)
            x = waits_for[x]
    share = max(-(-int(est.sum()) // machines), int(est.max()))
    # no machine runs over its share by more than the "no hand-over for < 8 iterations" rule allows
    assert load.max() <= share + 16, (load.max(), share, int(np.argmax(load)))
    assert int(load.sum()) == int(est.sum())
    return segs, lo, hi, load, share


@pytest.mark.parametrize("n,machines", [(256, 148), (200, 148), (149, 148), (211, 148), (8, 7), (12, 7), (40, 33),
                                        (90, 74), (1000, 148), (5, 148), (3, 1), (300, 2)])
def test_schedule_invariants_uniform_estimates(n, machines):
    check(np.full(n, 277), machines)


@pytest.mark.parametrize("seed", range(12))
def test_schedule_invariants_ragged_estimates(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 700))
    machines = int(rng.choice([1, 2, 7, 15, 33, 74, 148]))
    est = np.exp(rng.uniform(np.log(9), np.log(4000), size=n)).astype(np.int64)
    if seed % 3 == 0:
        est[rng.integers(0, n)] = 60000          # one chain longer than the average share
    check(est, machines)


def test_no_usable_estimate_deals_the_chains_out_whole():
    n, machines = 211, 148
    est = np.full(n, 300)
    for status, bad_est in ((MAXITER, 300), (CONVERGED, 0)):
        st = np.zeros(n, dtype=np.int32)
        e = est.copy()
        st[5] = status
        e[5] = bad_est
        segs, lo, hi = schedule(e, st, machines)
        seen = []
        for b in range(machines):
            got = [tuple(int(v) for v in segs[s][:3]) for s in range(lo[b], hi[b])]
            assert got == [(c, 1, INF) for c in range(b, n, machines)]
            seen += [g[0] for g in got]
        assert sorted(seen) == list(range(n))


def test_rejects_bad_arguments():
    lib = load_library()
    assert lib.tb_plan_schedule(None, None, 4, 2, None, None, None) != 0


def simulate(est, machines, resident):
    """Event simulation of a planned launch: blocks are dispatched in index order onto `resident` slots, a block runs its
    segments in order, a tail cannot start before its chain's head has finished (it spins, holding its slot).  Returns
    the makespan in iterations; raises if the launch cannot finish."""
    est = np.asarray(est, dtype=np.int64)
    segs, lo, hi = schedule(est, np.zeros(len(est)), machines)
    head_done = {}                      # chain -> time its head finished
    need_head = {int(segs[s][0]) for b in range(machines) for s in range(lo[b], hi[b]) if segs[s][1] > 1}
    free_at = [0] * resident            # when each slot becomes free
    finish = 0
    for b in range(machines):           # index order: the hardware's dispatch order
        slot = int(np.argmin(free_at))
        t = free_at[slot]
        for s in range(lo[b], hi[b]):
            c, k0, k1, _ = (int(v) for v in segs[s])
            if k0 > 1:
                if c not in head_done:
                    raise AssertionError(f"block {b} waits for the head of chain {c}, which no earlier block runs")
                t = max(t, head_done[c])
                t += est[c] - (k0 - 1)
            elif k1 != INF:
                t += k1 - 1
                head_done[c] = t
            else:
                t += est[c]
        free_at[slot] = t
        finish = max(finish, t)
    assert need_head <= set(head_done)
    return finish


@pytest.mark.parametrize("n,machines", [(256, 148), (149, 148), (211, 148), (8, 7), (40, 33), (600, 148)])
def test_planned_launch_finishes_whatever_the_residency(n, machines):
    """A tail only ever waits for a block with a lower index whose FIRST job is the head: the launch finishes with all
    blocks resident (the real case: one block per SM) and also if only some of them were (dispatch in index order)."""
    rng = np.random.default_rng(n)
    est = rng.integers(250, 300, size=n)
    ideal = est.sum() / machines
    full = simulate(est, machines, machines)
    assert full <= max(ideal, est.max()) + 16 + 1, (full, ideal)          # nobody waits when every block is resident
    for resident in (1, 2, max(1, machines // 3), machines - 1):
        simulate(est, machines, resident)                                 # slower, but it ends
