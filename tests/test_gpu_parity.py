"""Parity tests proper: the CUDA rollout path, called through the C ABI (libmcsim_b200.so), against
  (1) the oracle on the same seeded inputs,
  (2) the committed golden vectors the reference binary produced,
  (3) size-independent properties at BASELINE.json's full sizes.
Everything is bit-exact (==): statuses, draw counts, per-step trajectories, reduced features."""
import ctypes as C
import random

import pytest

import golden_io
from interactive_highway_simulation_b200 import _abi as abi
from scenes import ACTIONS, STATUS_KEYS, lane_change_cases, merge_exit_cases, probe_scene, random_scene, same

pytestmark = pytest.mark.gpu

ROLLOUT_CASES = golden_io.load_all("rollouts")


@pytest.fixture(scope="module")
def gpu(backend_lib):
    assert backend_lib.lib().mcs_device_count() > 0, "no CUDA device: the rollout path has no CPU fallback"
    return backend_lib


@pytest.fixture(autouse=True, params=["rollout_min_threads", "rollout_launch_sized", "bundle", "bundle_32_wide", "bundle_2_wide"])
def kernel_variant(request, monkeypatch):
    """Every test runs through both kernels and their launch shapes; results must not depend on any of it.
    rollout_kernel (one rollout's vehicles on the lanes of a warp; the only kernel for scenes above 64 vehicles): with the
    minimum block size of the scene (32 / 64 / 128 / 256 threads per rollout, so each instantiation and the shared-block
    path are exercised by the small scenes too) and with the library's own choice, which widens the rollouts of a small
    launch (capi.cu rollout_threads).  bundle_kernel (scenes up to 64 vehicles: the lanes of a warp are RW rollouts x 32/RW
    vehicles): the library's own width, 32 rollouts per warp (one vehicle per warp) and 2 rollouts per warp."""
    for name in ("MCS_ROLLOUT_THREADS", "MCS_KERNEL", "MCS_BUNDLE_RW_LOG2"):
        monkeypatch.delenv(name, raising=False)
    if request.param.startswith("rollout"):
        monkeypatch.setenv("MCS_KERNEL", "rollout")
        if request.param == "rollout_min_threads":
            monkeypatch.setenv("MCS_ROLLOUT_THREADS", "32")      # below every scene's minimum = the minimum
    else:
        monkeypatch.setenv("MCS_KERNEL", "bundle")               # whatever the launch size (the library's own rule: large merging launches)
        if request.param == "bundle_32_wide":
            monkeypatch.setenv("MCS_BUNDLE_RW_LOG2", "5")
        elif request.param == "bundle_2_wide":
            monkeypatch.setenv("MCS_BUNDLE_RW_LOG2", "1")
    return request.param


def status_dict(g):
    return dict(ok=g.ok, finish=g.finish, fallBack=g.fallBack, finishStep=g.finishStep, utility=g.utility, comfort=g.comfort,
                utilityObjects=g.utilityObjects, comfortObjects=g.comfortObjects, draws=g.draws, executedSteps=g.executedSteps)


def check_against_oracle(gpu, oracle, s, ego, dt, steps, min_steps, seed, count, tag, traj_rollouts=2, gap=-2, actions=ACTIONS):
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        for act in actions:
            cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS[act], gap, 0))
            sp = abi.SimParam(dt, steps, min_steps)
            st = ds.rollouts_range(cand, sp, seed, 0, count)
            o, _ = oracle.rollouts(s, ego, act, gap, dt, steps, min_steps, seed, 0, count, hist=True)
            for i in range(count):
                g = status_dict(st[i])
                if g["ok"] == 0 and o[i]["ok"] == 0:
                    continue
                for k in STATUS_KEYS + ("executedSteps",):
                    assert same(g[k], o[i][k]), (tag, act, "rollout", i, k, g[k], o[i][k])
            for r in range(min(traj_rollouts, count)):
                if st[r].ok == 0:
                    continue
                buf, ns, _ = ds.trajectories(cand[0], sp, seed, r)
                for ai, a in enumerate(s.agents):
                    ho = o[r]["hist"][a["id"]]
                    assert ns == len(ho)
                    for t in range(ns):
                        base = (ai * (steps + 1) + t) * 4
                        assert tuple(buf[base:base + 4]) == ho[t], (tag, act, "rollout", r, "agent", a["id"], "t", t)
    finally:
        ds.close()


def test_probe_scene_against_oracle(gpu, oracle):
    check_against_oracle(gpu, oracle, probe_scene(), 0, 0.3, 40, 30, 7, 32, "probe")


@pytest.mark.parametrize("case", lane_change_cases(10), ids=lambda c: c[0])
def test_random_lane_change_scenes_against_oracle(gpu, oracle, case):
    tag, s, ego, dt, steps, min_steps, seed = case
    check_against_oracle(gpu, oracle, s, ego, dt, steps, min_steps, seed, 12, tag)


@pytest.mark.parametrize("case", merge_exit_cases(24), ids=lambda c: c[0])
def test_merging_and_exit_scenes_against_oracle(gpu, oracle, case):
    """Ego `merging` / `exit` executing a gap (customizedMergingBehavior), non-ego `merging` vehicles choosing the closest
    gap, main-lane vehicles yielding to them (intention draws), gap re-targeting, behaviour switches on lane entry."""
    tag, s, ego, direction, gaps, seed = case
    for gap in gaps:
        check_against_oracle(gpu, oracle, s, ego, 0.3, 40, 10, seed, 8, tag, traj_rollouts=1, gap=gap, actions=(direction,))


def test_scene_errors_where_the_reference_is_undefined(gpu):
    """Unknown gap id / a non-ego vehicle without a usable behaviour: the reference reads invalid memory; we refuse."""
    rng = random.Random(3)
    s, ego = random_scene(rng, n_agents=8, lane_types=["merging", "main", "main"], ego_behavior="merging", ego_lane=0)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        sp = abi.SimParam(0.3, 40, 10)
        with pytest.raises(gpu.McsimError) as e:
            ds.rollouts_range((abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["left"], 9999, 0)), sp, 1, 0, 2)
        assert e.value.code == abi.MCS_ERR_SCENE
    finally:
        ds.close()
    s.agents[1 if s.agents[0]["id"] == ego else 0]["behavior"] = "exit"      # a non-ego `exit` vehicle on a non-exit lane
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        with pytest.raises(gpu.McsimError) as e:
            ds.rollouts_range((abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["left"], -1, 0)), sp, 1, 0, 2)
        assert e.value.code == abi.MCS_ERR_SCENE
    finally:
        ds.close()


import libm_check  # noqa: E402

LIBM_CASES = libm_check.load_cases()


@pytest.mark.parametrize("case", LIBM_CASES, ids=[c["tag"] for c in LIBM_CASES])
def test_against_the_stock_reference(gpu, case, record_property):
    """The CUDA path against the reference AS SHIPPED (its machine code with the platform libm's exp / pow, no pinned math
    interposed), at the shape of every BASELINE.json configuration: zero discrete mismatches (finish, fallBack,
    finishStep, draw count, state count, lane and lateral direction of every vehicle at every step) and every state of
    every vehicle within libm_check.STATE_TOL per step.  The measured deviation is recorded with the test."""
    s = golden_io.scene_of(case)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        cand = (abi.Candidate * 1)(abi.Candidate(case["ego"], abi.ACTIONS[case["action"]], case["gap"], 0))
        sp = abi.SimParam(case["dt"], case["steps"], case["min_steps"])
        st = ds.rollouts_range(cand, sp, case["seed"], case["first"], case["count"])
        statuses = [status_dict(st[i]) for i in range(case["count"])]
        hist = {}
        for i, r in enumerate(case["rollouts"]):
            if "hist" not in r:
                continue
            buf, ns, _ = ds.trajectories(cand[0], sp, case["seed"], case["first"] + i)
            stride = case["steps"] + 1
            hist[i] = {a["id"]: [tuple(buf[(ai * stride + t) * 4:(ai * stride + t) * 4 + 4]) for t in range(ns)]
                       for ai, a in enumerate(s.agents)}
        mismatches, dev_state, dev_stat = libm_check.compare(case, statuses, hist)
        record_property("max_state_deviation", dev_state)
        record_property("max_statistic_deviation", dev_stat)
        print("%s: discrete mismatches %d, max |state deviation| %.3g, max |statistic deviation| %.3g"
              % (case["tag"], len(mismatches), dev_state, dev_stat))
        assert mismatches == []
        assert dev_state <= libm_check.STATE_TOL and dev_stat <= libm_check.STAT_TOL, (dev_state, dev_stat)
    finally:
        ds.close()


def late_error_scene():
    """Two stacked merging lanes: the `merging` vehicle 1 moves left as soon as the gap in front of vehicle 2 is safe
    (which depends on vehicle 2's acceleration noise, i.e. on the rollout's draws), and from the upper lane there is no
    lane further left — mergingBehaviorClosestGap then reads a missing corridor in the reference (so:0xfe757 region).  With seed 5 and a
    14-step horizon rollouts 0..6 never get there, rollout 7 is the first that does (found with the oracle)."""
    from oracle.pyoracle import SceneSpec
    s = SceneSpec()
    s.add_corridor(1, "merging", 3.5, 0.0, 1.75, 27.78)
    s.add_corridor(2, "merging", 7.0, 3.5, 5.25, 27.78)
    s.add_agent(0, 300.0, 5.25, 25.0, 27.0, "idm")
    s.add_agent(1, 100.0, 1.75, 20.0, 22.0, "merging")
    s.add_agent(2, 88.0, 5.25, 26.0, 27.0, "idm")
    return s


def test_scene_error_in_a_later_rollout_is_reported(gpu, oracle):
    """The reference's behaviour can become undefined in ANY rollout (it depends on the trajectory, hence on the draws):
    every entry point must report it, not only when it happens in the first rollout of a candidate."""
    s = late_error_scene()
    cs, as_ = s.c_corridors(), s.c_agents()
    cand = (abi.Candidate * 1)(abi.Candidate(0, abi.ACTIONS["keep_lane"], -2, 0))
    sp = abi.SimParam(0.3, 14, 14)
    first_bad = None
    for r in range(32):
        one = (abi.Status * 1)()
        rc = oracle.lib.orc_rollouts(cs, len(cs), as_, len(as_), C.byref(cand[0]), C.byref(sp), 5, r, 1, one, None, None)
        if rc != 0:
            first_bad = r
            break
    assert first_bad == 7
    ds = gpu.DeviceScene(cs, as_)
    try:
        good = ds.rollouts_range(cand, sp, 5, 0, first_bad)                    # the rollouts before it are fine
        o, _ = oracle.rollouts(s, 0, "keep_lane", -2, 0.3, 14, 14, 5, 0, first_bad)
        for i in range(first_bad):
            for k in STATUS_KEYS:
                assert same(status_dict(good[i])[k], o[i][k]), (i, k)
        raw = (abi.Status * 32)()
        rc = gpu.lib().mcs_rollouts_range(ds._h, cand, 1, C.byref(sp), 5, 0, 32, raw)
        assert rc == abi.MCS_ERR_SCENE
        assert raw[0].overflow == 0 and raw[first_bad].overflow == 2           # not in the first record
        with pytest.raises(gpu.McsimError) as e:
            ds.rollouts(cand, sp, 4, 5)                                        # simulationMultiThread: statuses not requested
        assert e.value.code == abi.MCS_ERR_SCENE
        with pytest.raises(gpu.McsimError) as e:
            ds.rollouts_device(cand, sp, 5, 0, 32)
        assert e.value.code == abi.MCS_ERR_SCENE
        ds.rollouts_groups(cand, sp, 4, 5, 0, 1)                               # group 0 = rollouts 0..3: fine
        with pytest.raises(gpu.McsimError) as e:
            ds.rollouts_groups(cand, sp, 4, 5, 1, 1)                           # group 1 = rollouts 4..7
        assert e.value.code == abi.MCS_ERR_SCENE
        # the sharded path: the mark travels inside the partials, the combine refuses them on every rank
        parts = (abi.GroupPartial * 8)()
        rc = gpu.lib().mcs_rollouts_groups(ds._h, cand, 1, C.byref(sp), 4, 5, 0, 8, parts)
        assert rc == abi.MCS_ERR_SCENE and parts[0].reserved == 0 and parts[1].reserved == 2
        with pytest.raises(gpu.McsimError) as e:
            gpu.combine_groups(parts, 1, 8)
        assert e.value.code == abi.MCS_ERR_SCENE
    finally:
        ds.close()


@pytest.mark.parametrize("case", ROLLOUT_CASES, ids=golden_io.ids(ROLLOUT_CASES))
def test_against_reference_golden_vectors(gpu, case):
    """The CUDA path against what the reference's machine code returned (no oracle in between)."""
    s = golden_io.scene_of(case)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        cand = (abi.Candidate * 1)(abi.Candidate(case["ego"], abi.ACTIONS[case["action"]], case["gap"], 0))
        sp = abi.SimParam(case["dt"], case["steps"], case["min_steps"])
        st = ds.rollouts_range(cand, sp, case["seed"], case["first"], case["count"])
        for i, r in enumerate(case["rollouts"]):
            g = status_dict(st[i])
            if g["ok"] == 0 and r["ok"] == 0:
                continue
            for k in STATUS_KEYS:
                assert same(g[k], r[k]), (case["tag"], "rollout", i, k, g[k], r[k])
            if "hist" in r:
                buf, ns, _ = ds.trajectories(cand[0], sp, case["seed"], case["first"] + i)
                for ai, a in enumerate(s.agents):
                    hr = r["hist"][str(a["id"])]
                    assert ns == len(hr)
                    for t in range(ns):
                        base = (ai * (case["steps"] + 1) + t) * 4
                        assert tuple(buf[base:base + 4]) == tuple(hr[t]), (case["tag"], i, a["id"], t)
    finally:
        ds.close()


def test_features_two_level_mean(gpu, oracle):
    """simulationMultiThread semantics: 8 groups x n rollouts, mean of per-group means in fixed order; batched over
    candidates in one launch, and identical to the sharded path (mcs_rollouts_range + mcs_reduce_features)."""
    s = probe_scene()
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        cands = (abi.Candidate * 5)(*[abi.Candidate(0, abi.ACTIONS[a], -2, 0) for a in ACTIONS])
        sp = abi.SimParam(0.3, 40, 30)
        n = 20
        feats, st = ds.rollouts(cands, sp, n, 77, want_status=True)
        for c, act in enumerate(ACTIONS):
            o, ost = oracle.rollouts(s, 0, act, -2, 0.3, 40, 30, 77 + c, 0, 8 * n)
            want = oracle.reduce(ost, 8, n, 40)
            for f, _ in abi.Feature._fields_:
                assert same(getattr(feats[c], f), getattr(want, f)), (act, f)
        # sharded: two halves of the rollout range, gathered, reduced on the host
        half = 4 * n
        a = ds.rollouts_range(cands, sp, 77, 0, half)
        b = ds.rollouts_range(cands, sp, 77, half, half)
        allst = (abi.Status * (5 * 8 * n))()
        for c in range(5):
            for i in range(half):
                allst[c * 8 * n + i] = a[c * half + i]
                allst[c * 8 * n + half + i] = b[c * half + i]
        red = gpu.reduce_features(allst, 5, 8, n, 40)
        for c in range(5):
            for f, _ in abi.Feature._fields_:
                assert same(getattr(red[c], f), getattr(feats[c], f)), (c, f)
    finally:
        ds.close()


def test_invalid_candidates_yield_zero_sims(gpu):
    """Unknown ego id / command without a neighbour lane: the reference returns fromNSims = 0 (so:0xc2b80)."""
    s = probe_scene()
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        cands = (abi.Candidate * 3)(abi.Candidate(999, abi.ACTIONS["keep_lane"], -2, 0),
                                    abi.Candidate(5, abi.ACTIONS["left"], -2, 0),      # agent 5 drives in the leftmost lane
                                    abi.Candidate(0, abi.ACT_INVALID, -2, 0))
        feats, _ = ds.rollouts(cands, abi.SimParam(0.3, 40, 30), 2, 1)
        assert [f.fromNSims for f in feats] == [0, 0, 0]
    finally:
        ds.close()


def test_capacity_and_argument_errors(gpu):
    s = probe_scene()
    cs, as_ = s.c_corridors(), s.c_agents()
    h = C.c_void_p()
    assert gpu.lib().mcs_scene_create(cs, len(cs), as_, 0, 0, C.byref(h)) == abi.MCS_ERR_ARG          # empty agent list
    big = (abi.Agent * (abi.MCS_MAX_AGENTS + 1))()
    for i, a in enumerate(big):
        a.id = i
    assert gpu.lib().mcs_scene_create(cs, len(cs), big, len(big), 0, C.byref(h)) == abi.MCS_ERR_CAPACITY
    assert gpu.lib().mcs_scene_create(cs, len(cs), as_, len(as_), 99, C.byref(h)) == abi.MCS_ERR_CUDA  # no such device


def test_full_size_properties(gpu, oracle):
    """BASELINE.json configs[4] shape (256 vehicles x 100 steps): determinism, shard-independence, draw accounting and
    a spot check of a few rollouts against the oracle."""
    from interactive_highway_simulation_b200 import synthetic
    corridors, agents, ego = synthetic.lane_change_scene(256, 4, seed=0)
    ds = gpu.DeviceScene(corridors, agents)
    try:
        cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0))
        sp = abi.SimParam(0.3, 100, 100)
        R = 2048
        a = ds.rollouts_range(cand, sp, 2026, 0, R)
        b = ds.rollouts_range(cand, sp, 2026, 0, R)
        assert bytes(a) == bytes(b)                                      # idempotent / deterministic
        c = ds.rollouts_range(cand, sp, 2026, R - 64, 128)              # a shard that straddles the first batch
        for i in range(64):
            assert bytes(a[R - 64 + i]) == bytes(c[i])                   # a rollout depends only on (seed, index)
        for i in range(R):
            assert a[i].ok == 1 and a[i].executedSteps == 100 and a[i].finish == 1
            assert 256 + 255 * 100 <= a[i].draws <= 256 + 2 * 255 * 100   # ctor flags + noise (+ MOBIL) draws
        ms, vsteps = ds.rollouts_device(cand, sp, 2026, 0, R)
        assert vsteps == R * 256 * 100                                   # checksum of executed vehicle-steps
        ost = (abi.Status * 3)()
        rc = oracle.lib.orc_rollouts(corridors, len(corridors), agents, len(agents), C.byref(cand[0]), C.byref(sp), 2026,
                                     1000, 3, ost, None, None)
        assert rc == 0
        for i in range(3):
            assert bytes(a[1000 + i])[:48] == bytes(ost[i])[:48], i
    finally:
        ds.close()


DECISION_CASES = golden_io.load_all("decisions")


@pytest.mark.parametrize("case", DECISION_CASES, ids=golden_io.ids(DECISION_CASES))
def test_single_decisions_against_reference_golden_vectors(gpu, case):
    """mcs_decide_* (laneChangeMobilBehavior, fixedDirectionLaneChangeBehavior, fixedGapMergingBehavior,
    closestGapMergingBehavior) against what the reference binary's wrappers returned."""
    s = golden_io.scene_of(case)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        for call in case["calls"]:
            t = call["cmd"].split()
            ans = call["answer"]
            if t[0] == "mobil":
                out = ds.decide_mobil(int(t[1]), float(t[2]), float(t[3]), int(t[4]), case["key"])
                assert abi.COMMIT_NAMES[out.commit] == ans["commit"] and out.commit_keep_lane_step == int(ans["keep"]), call
                assert out.target_lane_id == int(ans["target"]), call
            elif t[0] == "fixed_dir":
                out = ds.decide_fixed_direction(int(t[1]), abi.ACTIONS[t[2]])
            elif t[0] == "fixed_gap":
                out = ds.decide_fixed_gap(int(t[1]), abi.ACTIONS[t[2]], int(t[3]))
            else:
                out = ds.decide_closest_gap(int(t[1]), abi.ACTIONS[t[2]])
            assert (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), call
    finally:
        ds.close()


def test_mobil_decisions_against_oracle(gpu, oracle):
    """laneChangeMobilBehavior with the commit state Python pokes into the agent (behaviors.py:75-80), both requireRss values."""
    rng = random.Random(12)
    commits = ["not_decided", "keep_lane", "left", "right"]
    checked = 0
    for sc in range(12):
        s, ego = random_scene(rng, rng.choice([2, 3, 4]), rng.choice([6, 12, 25]))
        for a in s.agents:
            a["commit"] = rng.choice(commits); a["keep_step"] = rng.randrange(0, 14)
            a["target_lane"] = rng.choice([-1, 1, 2, 3])
        cs, as_ = s.c_corridors(), s.c_agents()
        ds = gpu.DeviceScene(cs, as_)
        try:
            for a in rng.sample(s.agents, min(6, len(s.agents))):
                for req in (0, 1):
                    ax, vy = rng.uniform(-6, 1.5), rng.uniform(-0.3, 0.3)
                    if rng.random() < 0.2:
                        ax = a["rss"][2]
                    want = abi.ActionWithType()
                    act = abi.Action(ax, vy)
                    rc = oracle.lib.orc_decide_mobil(cs, len(cs), as_, len(as_), a["id"], C.byref(act), req, 900 + sc, C.byref(want))
                    assert rc == 0
                    got = ds.decide_mobil(a["id"], ax, vy, req, 900 + sc)
                    assert (got.ax, got.vy, got.commit, got.commit_keep_lane_step, got.target_lane_id) == \
                        (want.ax, want.vy, want.commit, want.commit_keep_lane_step, want.target_lane_id), (sc, a["id"], req)
                    checked += 1
        finally:
            ds.close()
    assert checked > 100


def test_decide_unknown_agent_is_an_error(gpu):
    s = probe_scene()
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        with pytest.raises(gpu.McsimError) as e:
            ds.decide_fixed_direction(4242, abi.ACTIONS["keep_lane"])
        assert e.value.code == abi.MCS_ERR_ARG
    finally:
        ds.close()


def test_sharded_groups_equal_single_gpu(gpu):
    """The multi-GPU exchange unit: per-group partials computed shard by shard (as 1, 2, 4, 8 ranks would) and combined
    in group order give exactly mcs_rollouts' features; also through a device buffer (the NCCL hand-off path)."""
    import torch
    from interactive_highway_simulation_b200 import sharding
    s = probe_scene()
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        cands = (abi.Candidate * 3)(*[abi.Candidate(0, abi.ACTIONS[a], -2, 0) for a in ("keep_lane", "left", "acc")])
        sp = abi.SimParam(0.3, 40, 30)
        n = 16
        want, _ = ds.rollouts(cands, sp, n, 31)
        for world in (1, 2, 4, 8):
            per = 8 // world
            chunks = []
            for rank in range(world):
                first, cnt = sharding.group_range(rank, world)
                t = torch.empty((3 * per, 8), dtype=torch.float64, device="cuda")
                ds.rollouts_groups(cands, sp, n, 31, first, cnt, d_partials_ptr=t.data_ptr())
                host = ds.rollouts_groups(cands, sp, n, 31, first, cnt)
                assert bytes(host) == t.cpu().numpy().tobytes()
                chunks.append(t)
            gathered = torch.cat(chunks, 0)
            feats = sharding.features_from_gathered(sharding.interleave_gathered(gathered, 3, world).cpu(), 3)
            for c in range(3):
                assert bytes(feats[c]) == bytes(want[c]), (world, c)
    finally:
        ds.close()


@pytest.mark.parametrize("n_agents", [1, 2, 31, 32, 33, 64, 65, 128, 129, 200, 256])
def test_block_size_boundaries_against_oracle(gpu, oracle, n_agents):
    """Every kernel instantiation (32 / 64 / 128 / 256 threads) at and around its capacity, incl. an ego without objects."""
    rng = random.Random(4000 + n_agents)
    s, ego = random_scene(rng, rng.choice([2, 3, 4]), n_agents, spacing=(3, 40))
    check_against_oracle(gpu, oracle, s, ego, 0.3, 30, 10, 900 + n_agents, 4, "n%d" % n_agents, traj_rollouts=1,
                         actions=("keep_lane", "left", "dcc"))


def test_eight_lanes_and_zero_steps(gpu, oracle):
    rng = random.Random(77)
    s, ego = random_scene(rng, 8, 120, spacing=(3, 50))                       # MCS_MAX_LANES corridors
    check_against_oracle(gpu, oracle, s, ego, 0.2, 25, 25, 5, 4, "8lanes", traj_rollouts=1, actions=("keep_lane", "right"))
    s, ego = random_scene(rng, 3, 12)
    check_against_oracle(gpu, oracle, s, ego, 0.3, 0, 0, 6, 2, "0steps", traj_rollouts=1, actions=("keep_lane",))   # comfort = 0/0
    check_against_oracle(gpu, oracle, s, ego, 0.3, 1, 1, 6, 2, "1step", traj_rollouts=1, actions=("acc",))


# MCS_FUZZ_BLOCKS=N widens the sweep to N x 10 scenes for an occasional long run (default 6: 60 scenes)
@pytest.mark.parametrize("block", range(int(__import__("os").environ.get("MCS_FUZZ_BLOCKS", "6"))))
def test_many_random_scenes_against_oracle(gpu, oracle, block):
    """Breadth: 60 more random scenes (lane change, merging, exit; sparse to dense) — statuses only."""
    for sc in range(10):
        rng = random.Random(50000 + 10 * block + sc)
        kind = rng.choice(["lc", "lc", "merge", "exit"])
        na = rng.choice([6, 11, 18, 27, 40, 70])
        spacing = rng.choice([(3, 60), (2, 20), (8, 100)])
        if kind == "lc":
            s, ego = random_scene(rng, rng.choice([2, 3, 4, 5]), na, spacing=spacing)
            actions, gap = (rng.choice(ACTIONS), rng.choice(ACTIONS)), -2
        elif kind == "merge":
            s, ego = random_scene(rng, n_agents=na, lane_types=["merging", "main"] + ["main"] * rng.choice([0, 1, 2]),
                                  ego_behavior="merging", ego_lane=0, spacing=spacing)
            actions, gap = ("left",), rng.choice([-1] + [a["id"] for a in s.agents])
        else:
            s, ego = random_scene(rng, n_agents=na, lane_types=["exit", "main"] + ["main"] * rng.choice([0, 1]),
                                  ego_behavior="exit", ego_lane=1, spacing=spacing)
            actions, gap = ("right",), rng.choice([-1] + [a["id"] for a in s.agents])
        check_against_oracle(gpu, oracle, s, ego, rng.choice([0.2, 0.3]), rng.choice([20, 40, 50]), rng.choice([5, 10, 30]),
                             7 + sc, 6, "%s_b%d_%d" % (kind, block, sc), traj_rollouts=0, gap=gap, actions=actions)


def test_sharded_entry_point_single_rank(gpu):
    """sharding.simulation_multi_thread_sharded without a process group (world 1): the device-tensor hand-off path that
    bench.py runs over NCCL gives mcs_rollouts' features."""
    import torch
    from interactive_highway_simulation_b200 import sharding, synthetic
    corridors, agents, ego, gaps = synthetic.merging_scene(40, 1)
    ds = gpu.DeviceScene(corridors, agents)
    try:
        torch.cuda.set_device(0)
        cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
        sp = abi.SimParam(0.3, 40, 10)
        got = sharding.simulation_multi_thread_sharded(ds, cands, sp, 8, 123)
        want, _ = ds.rollouts(cands, sp, 8, 123)
        for c in range(len(gaps)):
            assert bytes(got[c]) == bytes(want[c])
    finally:
        ds.close()


def test_full_size_exit_configuration(gpu, oracle):
    """BASELINE.json configs[3] shape: 200 vehicles with an exit lane, 4096 rollouts x 50-step horizon per candidate gap —
    determinism, shard independence, the two-level mean over the sharded statuses, and a spot check against the oracle."""
    from interactive_highway_simulation_b200 import synthetic
    corridors, agents, ego, gaps = synthetic.exit_scene(200, 0)
    ds = gpu.DeviceScene(corridors, agents)
    try:
        cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["right"], g, 0) for g in gaps])
        sp = abi.SimParam(0.3, 50, 10)
        n = 512                                                    # 8 groups x 512 = 4096 rollouts per candidate
        feats, st = ds.rollouts(cands, sp, n, 31, want_status=True)
        feats2, _ = ds.rollouts(cands, sp, n, 31)
        for c in range(len(gaps)):
            assert bytes(feats[c]) == bytes(feats2[c]) and feats[c].fromNSims == 4096
            assert 0.0 <= feats[c].successRate <= 1.0 and 0.0 <= feats[c].fallBackRate <= 1.0
        shard = ds.rollouts_range(cands, sp, 31, 3000, 200)         # rollouts 3000..3199 of every candidate again
        for c in range(len(gaps)):
            for i in range(200):
                assert bytes(shard[c * 200 + i]) == bytes(st[c * 4096 + 3000 + i])
        red = gpu.reduce_features(st, len(gaps), 8, n, sp.steps)
        for c in range(len(gaps)):
            assert bytes(red[c]) == bytes(feats[c])
        for c in (0, len(gaps) - 1):
            ost = (abi.Status * 2)()
            assert oracle.lib.orc_rollouts(corridors, len(corridors), agents, len(agents), C.byref(cands[c]), C.byref(sp), 31 + c,
                                           4000, 2, ost, None, None) == 0
            for i in range(2):
                assert bytes(st[c * 4096 + 4000 + i])[:48] == bytes(ost[i])[:48], (c, i)
    finally:
        ds.close()


def test_scene_lifecycle_does_not_leak(gpu):
    """One scene per decision is the reference's call pattern: create / launch / destroy 600 times, device memory stays flat."""
    import torch
    s = probe_scene()
    cs, as_ = s.c_corridors(), s.c_agents()
    cand = (abi.Candidate * 1)(abi.Candidate(0, abi.ACTIONS["keep_lane"], -2, 0))
    sp = abi.SimParam(0.3, 10, 10)
    def cycle(n):
        for i in range(n):
            ds = gpu.DeviceScene(cs, as_)
            ds.rollouts(cand, sp, 1, i)
            ds.close()
    cycle(50)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    cycle(600)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert abs(free0 - free1) < 64 << 20, (free0, free1)


ORDER_CASES = golden_io.load_all("order")


@pytest.mark.parametrize("case", ORDER_CASES, ids=golden_io.ids(ORDER_CASES))
def test_visit_order_matches_reference_binary(gpu, case):
    """The plan / random-draw order of the vehicles (iteration order of the reference's std::unordered_map<int, AgentPtr>),
    as recorded from the reference binary."""
    s = golden_io.scene_of(case)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        order = ds.visit_order()                                  # input indices in visit order
        assert [s.agents[i]["id"] for i in order] == case["agents"]
    finally:
        ds.close()


@pytest.mark.parametrize("kind", ["lane_change", "merging"])
def test_result_is_independent_of_threads_per_rollout(gpu, monkeypatch, kind):
    """The same rollouts on 32 / 64 / 128 / 256 threads each, own and shared blocks: byte-identical outcome records."""
    rng = random.Random(31337)
    if kind == "lane_change":
        s, ego = random_scene(rng, 4, 24, spacing=(3, 40))
        cands = [(ego, a, -2) for a in ("keep_lane", "left", "right")]
    else:
        s, ego = random_scene(rng, n_agents=40, lane_types=["merging", "main", "main"], ego_behavior="merging", ego_lane=0,
                              spacing=(3, 40))
        cands = [(ego, "left", -1)]
    cand = (abi.Candidate * len(cands))(*[abi.Candidate(e, abi.ACTIONS[a], g, 0) for e, a, g in cands])
    sp = abi.SimParam(0.3, 30, 10)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        ref = None
        for threads in (32, 64, 128, 256, 64, None):
            # 700 x candidates: enough rollouts for the shared-block path; 2000 merging rollouts: the 896-thread blocks
            for count in (3, 700) + ((2000,) if kind == "merging" and threads in (64, None) else ()):
                if threads is None:
                    monkeypatch.delenv("MCS_ROLLOUT_THREADS", raising=False)
                else:
                    monkeypatch.setenv("MCS_ROLLOUT_THREADS", str(threads))
                st = ds.rollouts_range(cand, sp, 99, 0, count)
                got = [bytes(st)[(c * count + i) * 64:(c * count + i + 1) * 64] for c in range(len(cands)) for i in range(3)]
                if ref is None:
                    ref = got
                    assert any(x != bytes(64) for x in ref)
                assert got == ref, (kind, threads, count)
    finally:
        ds.close()


@pytest.mark.parametrize("kind", ["lane_change", "merging", "exit"])
def test_bundle_kernel_equals_rollout_kernel(gpu, monkeypatch, kind):
    """The two kernels on the same rollouts — bundle_kernel at every width (1 .. 32 rollouts per warp) against
    rollout_kernel: byte-identical outcome records, in launches of 3, 700 and 2000 rollouts per candidate."""
    rng = random.Random(777)
    if kind == "lane_change":
        s, ego = random_scene(rng, 4, 30, spacing=(3, 40))
        cands = [(ego, a, -2) for a in ("keep_lane", "left", "right", "acc", "dcc")]
    elif kind == "merging":
        s, ego = random_scene(rng, n_agents=40, lane_types=["merging", "main", "main", "main"], ego_behavior="merging", ego_lane=0,
                              spacing=(3, 40))
        ids = [a["id"] for a in s.agents if 3.5 <= a["y"] < 7.0]
        cands = [(ego, "left", -1)] + [(ego, "left", g) for g in ids[:2]]
    else:
        s, ego = random_scene(rng, n_agents=33, lane_types=["exit", "main", "main"], ego_behavior="exit", ego_lane=1, spacing=(3, 40))
        cands = [(ego, "right", -1)]
    cand = (abi.Candidate * len(cands))(*[abi.Candidate(e, abi.ACTIONS[a], g, 0) for e, a, g in cands])
    sp = abi.SimParam(0.3, 40, 10)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        for count in (3, 700, 2000):
            monkeypatch.setenv("MCS_KERNEL", "rollout")
            monkeypatch.delenv("MCS_ROLLOUT_THREADS", raising=False)
            want = bytes(ds.rollouts_range(cand, sp, 4711, 0, count))
            assert any(want[i * 64:(i + 1) * 64] != want[:64] for i in range(1, count))
            monkeypatch.setenv("MCS_KERNEL", "bundle")
            for width in (None, 0, 1, 2, 3, 4, 5):
                if width is None:
                    monkeypatch.delenv("MCS_BUNDLE_RW_LOG2", raising=False)
                else:
                    monkeypatch.setenv("MCS_BUNDLE_RW_LOG2", str(width))
                got = bytes(ds.rollouts_range(cand, sp, 4711, 0, count))
                bad = [i for i in range(len(cands) * count) if got[i * 64:(i + 1) * 64] != want[i * 64:(i + 1) * 64]]
                assert not bad, (kind, count, width, len(bad), bad[:5])
    finally:
        ds.close()


@pytest.mark.parametrize("kind", ["merging", "exit", "exit_200"])
def test_helper_warp_equals_inline_gap_execution(gpu, monkeypatch, kind):
    """rollout_kernel<..., kHelper>: the ego's customizedMergingBehavior (and the re-targeting of its gap) on the rollout's spare
    warp, next to the other warps' plan phase, against the same kernel with the gap execution inline (MCS_NO_HELPER=1):
    byte-identical outcome records on 128- and 256-thread rollouts, own blocks and shared 768-thread blocks."""
    rng = random.Random(2024)
    if kind == "merging":
        s, ego = random_scene(rng, n_agents=40, lane_types=["merging", "main", "main", "main"], ego_behavior="merging", ego_lane=0,
                              spacing=(3, 40))
        ids = [a["id"] for a in s.agents if 3.5 <= a["y"] < 7.0]
        cands = [(ego, "left", -1)] + [(ego, "left", g) for g in ids[:3]]
    else:
        n = 200 if kind == "exit_200" else 60
        s, ego = random_scene(rng, n_agents=n, lane_types=["exit", "main", "main", "main"], ego_behavior="exit", ego_lane=1,
                              spacing=(3, 40))
        ids = [a["id"] for a in s.agents if a["y"] < 3.5]
        cands = [(ego, "right", -1)] + [(ego, "right", g) for g in ids[:3]]
    cand = (abi.Candidate * len(cands))(*[abi.Candidate(e, abi.ACTIONS[a], g, 0) for e, a, g in cands])
    sp = abi.SimParam(0.3, 50, 10)
    monkeypatch.setenv("MCS_KERNEL", "rollout")
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        for threads in ((256,) if kind == "exit_200" else (128, 256)):
            monkeypatch.setenv("MCS_ROLLOUT_THREADS", str(threads))
            for count in (3, 700):
                monkeypatch.setenv("MCS_NO_HELPER", "1")
                want = bytes(ds.rollouts_range(cand, sp, 515, 0, count))
                assert any(want[i * 64:(i + 1) * 64] != want[:64] for i in range(1, count))
                monkeypatch.delenv("MCS_NO_HELPER")
                got = bytes(ds.rollouts_range(cand, sp, 515, 0, count))
                bad = [i for i in range(len(cands) * count) if got[i * 64:(i + 1) * 64] != want[i * 64:(i + 1) * 64]]
                assert not bad, (kind, threads, count, len(bad), bad[:5])
    finally:
        ds.close()


@pytest.mark.parametrize("kind", ["lane_change", "merging"])
def test_long_horizon_in_a_large_launch(gpu, kind):
    """A long horizon makes the per-rollout shared-memory region large; a launch big enough to share thread blocks must
    fall back to fewer rollouts per block instead of failing, with the same outcome records as a small launch."""
    rng = random.Random(4242)
    if kind == "lane_change":
        s, ego = random_scene(rng, 3, 20, spacing=(3, 40))
        cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0))
    else:
        s, ego = random_scene(rng, n_agents=40, lane_types=["merging", "main", "main"], ego_behavior="merging", ego_lane=0,
                              spacing=(3, 40))
        cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["left"], -1, 0))
    sp = abi.SimParam(0.2, 900, 900)
    ds = gpu.DeviceScene(s.c_corridors(), s.c_agents())
    try:
        small = bytes(ds.rollouts_range(cand, sp, 5, 0, 2))
        big = bytes(ds.rollouts_range(cand, sp, 5, 0, 1500))
        assert big[:128] == small
        assert any(big[i * 64:(i + 1) * 64] != big[:64] for i in range(1, 1500))
    finally:
        ds.close()


def test_full_width_merging_rollouts_share_a_block_with_identical_results(gpu, monkeypatch):
    """129-256 vehicles with a merging / exit lane: in a launch that fills every SM three rollouts run in step in one
    768-thread block (instruction fetches shared); every outcome record equals the one from a block of its own."""
    from interactive_highway_simulation_b200 import synthetic
    corridors, agents, ego, gaps = synthetic.exit_scene(200, 0)
    ds = gpu.DeviceScene(corridors, agents)
    try:
        cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["right"], g, 0) for g in gaps])
        sp = abi.SimParam(0.3, 50, 10)
        count = 301                                                 # x 4 gaps: not a multiple of 3, last block partly filled
        shared = bytes(ds.rollouts_range(cands, sp, 77, 5, count))
        monkeypatch.setenv("MCS_SHARE_MIN_PER_SM", "1000000")       # every rollout in a block of its own
        own = bytes(ds.rollouts_range(cands, sp, 77, 5, count))
        assert shared == own
        assert len({shared[i * 64:(i + 1) * 64] for i in range(count * len(gaps))}) > count
    finally:
        ds.close()
