"""Host-side logic of libmcsim_b200.so that needs no GPU: the fixed-order two-level feature reduction
(runMTimes so:0xc2a6e-0xc2b1f + runMultiThreads so:0xbca08-0xbca92) against the oracle's restatement."""
import ctypes as C
import random

from interactive_highway_simulation_b200 import _abi as abi


def random_statuses(rng, n, ok_groups, m):
    st = (abi.Status * n)()
    for i, s in enumerate(st):
        s.utility = rng.uniform(0.3, 1.1); s.comfort = rng.uniform(0.2, 1.0)
        s.utilityObjects = rng.uniform(0.5, 1.0); s.comfortObjects = rng.uniform(0.5, 1.0)
        s.finish = rng.random() < 0.7; s.fallBack = rng.random() < 0.2
        s.finishStep = rng.randrange(3, 40); s.ok = 1 if ok_groups[i // m] else 0
    return st


def test_reduce_features_matches_oracle(oracle, backend_lib):
    rng = random.Random(4)
    for trial in range(20):
        groups, m, steps = 8, rng.choice([1, 3, 20, 30]), rng.choice([25, 40, 100])
        ok_groups = [rng.random() < 0.9 for _ in range(groups)]
        if trial == 0:
            ok_groups = [True] * groups
        st = random_statuses(rng, groups * m, ok_groups, m)
        got = backend_lib.reduce_features(st, 1, groups, m, steps)[0]
        want = oracle.reduce(st, groups, m, steps)
        for f, _ in abi.Feature._fields_:
            a, b = getattr(got, f), getattr(want, f)
            assert a == b or (a != a and b != b), (trial, f, a, b)


def test_reduce_features_is_mean_of_group_means(backend_lib):
    rng = random.Random(9)
    groups, m = 8, 5
    st = random_statuses(rng, groups * m, [True] * groups, m)
    f = backend_lib.reduce_features(st, 1, groups, m, 40)[0]
    acc = 0.0
    for g in range(groups):
        u = 0.0
        for i in range(m):
            u = u + st[g * m + i].utility
        acc = acc + u / m
    assert f.utility == acc / groups
    assert f.fromNSims == groups * m


def test_reduce_features_rejects_bad_arguments(backend_lib):
    st = (abi.Status * 8)()
    feats = (abi.Feature * 1)()
    assert backend_lib.lib().mcs_reduce_features(st, 1, 0, 1, 40, feats) == abi.MCS_ERR_ARG
    assert backend_lib.lib().mcs_reduce_features(None, 1, 8, 1, 40, feats) == abi.MCS_ERR_ARG
