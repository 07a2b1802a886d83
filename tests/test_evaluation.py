"""The evaluation scripts' per-scene functions (evaluation_merging.py:53-106, evaluation_lane_change.py:58-101,
evaluation_exit.py:55-105) on the package's outer loop against the reference's own scripts
(tests/golden/evaluation_stats.json.gz, written by tests/golden/make_evaluation_golden.py): the scene is written in the
reference's yaml format, loaded, the ego gets its policy, the scene runs, and the statistics object and the console text
must come out as the reference's did.  CPU: host logic (oracle geometry, recorded inner answers); GPU: the real thing."""
import gzip
import json
import os

import pytest
import yaml

from outer_loop_cases import GOLDEN, OracleDeviceMap, RecordedInnerSimulator, seed_streams

with gzip.open(os.path.join(GOLDEN, "evaluation_stats.json.gz"), "rt") as f:
    RECORDS = json.load(f)
CASES = [(r, k) for r in RECORDS for k in range(len(r["evaluations"]))]
IDS = ["%s-%s-%s" % (r["kind"], r["evaluations"][k]["policy"], r["evaluations"][k]["fall_back"]) for r, k in CASES]


def write_scene(rec, tmp_path):
    d = tmp_path / "epoch0"
    d.mkdir()
    for name in ("agents", "map"):
        with open(d / (name + ".yml"), "w") as f:
            yaml.dump({int(k): v for k, v in rec[name].items()}, f, default_flow_style=False)
    return str(d)


def evaluate(ev_mod, rec, k, path, statistics):
    """the evaluations of one statistics object run in the recorded order up to evaluation k (they accumulate)"""
    e = rec["evaluations"][k]
    sim = ev_mod.Simulation(None, ev_mod.SimulationParameter(0.2, rec["steps"], render=False, write_gif=False, scenario=rec["kind"]))
    if rec["kind"] == "merging":
        return sim, ev_mod.evaluate_recorded_merging_scene(path, sim, e["policy"], 0, statistics, e["fall_back"])
    if rec["kind"] == "lane_change":
        return sim, ev_mod.evaluate_recorded_lc_scene(path, sim, e["policy"], 0, e["ego"], statistics)
    return sim, ev_mod.evaluate_recorded_exit_scene(path, sim, e["policy"], 0, e["ego"], statistics, e["fall_back"])


def check(ev_mod, rec, k, got_text, statistics, sim, tol):
    e = rec["evaluations"][k]
    assert sim.step + 1 == e["executed_steps"]
    for key, want in e["statistics"].items():
        got = getattr(statistics, key)
        if isinstance(want, list):
            assert len(got) == len(want), key
            for g, w in zip(got, want):
                assert g == w if isinstance(w, (list, int)) else abs(g - w) <= tol, (key, got, want)
        else:
            assert got == want, (key, got, want)
    assert got_text == e["text"] or tol > 0      # the console line rounds to 3 digits: identical when the numbers are


def fresh_statistics(ev_mod, kind):
    return ev_mod.LaneChangeStatistics() if kind == "lane_change" else ev_mod.MergingAndExitStatistics()


@pytest.mark.parametrize("rec,k", CASES, ids=IDS)
def test_evaluate_recorded_scene_host_logic(monkeypatch, tmp_path, rec, k):
    from interactive_highway_simulation_b200 import evaluation as ev_mod, mc_sim_highway as ms, outer_env
    monkeypatch.setattr(outer_env, "DeviceMap", OracleDeviceMap)
    e = rec["evaluations"][k]
    path = write_scene(rec, tmp_path)
    inner = RecordedInnerSimulator(ms, e["calls"])
    inner.install(monkeypatch)
    seed_streams(e["seed"])
    st = fresh_statistics(ev_mod, rec["kind"])
    sim, text = evaluate(ev_mod, rec, k, path, st)
    assert inner.k == len(e["calls"])
    check(ev_mod, rec, k, text, st, sim, 0.0)


def test_ego_candidates_and_generated_epochs(monkeypatch, tmp_path):
    """ego choice of the three scripts; generate_random_traffic writes epochs the loader reads back unchanged"""
    from interactive_highway_simulation_b200 import evaluation as ev_mod, outer_env
    monkeypatch.setattr(outer_env, "DeviceMap", OracleDeviceMap)
    for rec in RECORDS:
        sim = ev_mod.Simulation(None, ev_mod.SimulationParameter(0.2, 1, render=False, write_gif=False, scenario=rec["kind"]))
        sub = tmp_path / rec["kind"]
        sub.mkdir()
        sim.load_initial_scene(write_scene(rec, sub))
        got = ev_mod.ego_candidates(sim.environment, rec["kind"])
        if rec["kind"] == "merging":
            assert got == [a["id"] for a in rec["agents"].values() if "merging" in a["behavior_model"]]
        else:
            assert got == rec["ego_candidates"]
    seed_streams(5)
    sim = ev_mod.Simulation(None, ev_mod.SimulationParameter(0.2, 1, render=False, write_gif=False, scenario="exit"))
    sim.write_dir = str(tmp_path / "generated")
    ev_mod.generate_random_traffic("exit", 2, sim)
    for n in range(2):
        d = os.path.join(sim.write_dir, "epoch%d" % n)
        again = ev_mod.Simulation(None, sim.param)
        again.load_initial_scene(d)
        with open(os.path.join(d, "agents.yml")) as f:
            on_disk = yaml.safe_load(f)
        assert len(on_disk) > 20 and sorted(on_disk) == sorted(again.environment.agents)
        for i, a in again.environment.agents.items():
            assert (a.x, a.y, a.v, a.l, a.behavior_model, a.idm_param.vd) == \
                (on_disk[i]["x"], on_disk[i]["y"], on_disk[i]["v"], on_disk[i]["l"], on_disk[i]["behavior_model"], on_disk[i]["idm_p"]["vd"])
        with open(os.path.join(d, "map.yml")) as f:
            assert yaml.safe_load(f) == again.environment.map.to_dict()
    # vehicle ids across epochs: evaluation_merging.py never resets its generator's counter (ids run on from epoch to
    # epoch), evaluation_lane_change.py:56 / evaluation_exit.py:52 restart at 0 every epoch
    for kind, restarts in (("merging", False), ("lane_change", True)):
        seed_streams(6)
        sim = ev_mod.Simulation(None, ev_mod.SimulationParameter(0.2, 1, render=False, write_gif=False, scenario=kind))
        sim.write_dir = str(tmp_path / ("ids_" + kind))
        ev_mod.generate_random_traffic(kind, 2, sim)
        ids = []
        for n in range(2):
            with open(os.path.join(sim.write_dir, "epoch%d" % n, "agents.yml")) as f:
                ids.append(sorted(yaml.safe_load(f)))
        assert ids[0][0] == 0
        assert ids[1][0] == (0 if restarts else ids[0][-1] + 1), (kind, ids[0][-1], ids[1][0])


@pytest.mark.gpu
@pytest.mark.parametrize("rec,k", CASES, ids=IDS)
def test_evaluate_recorded_scene_on_the_device(backend_lib, tmp_path, rec, k):
    from interactive_highway_simulation_b200 import evaluation as ev_mod, mc_sim_highway as ms
    e = rec["evaluations"][k]
    path = write_scene(rec, tmp_path)
    seed_streams(e["seed"])
    ms.seed(e["seed"])
    st = fresh_statistics(ev_mod, rec["kind"])
    sim, text = evaluate(ev_mod, rec, k, path, st)
    check(ev_mod, rec, k, text, st, sim, 0.0)


def test_merge_statistics_concatenates_lists_and_sums_counters():
    from interactive_highway_simulation_b200 import evaluation as ev_mod
    a, b = ev_mod.LaneChangeStatistics(), ev_mod.LaneChangeStatistics()
    a.utility_ego += [0.5]; a.num_lane_change += 2; a.num_fallback += 1; a.lane_id_history += [3]
    b.utility_ego += [0.7, 0.9]; b.num_lane_change += 1; b.lane_id_history += [1, 2]
    total = ev_mod.merge_statistics("lane_change", [dict(a.__dict__), dict(b.__dict__)])
    assert total.utility_ego == [0.5, 0.7, 0.9] and total.num_lane_change == 3 and total.num_fallback == 1
    assert total.lane_id_history == [3, 1, 2]


@pytest.mark.gpu
def test_parallel_evaluations_equal_the_same_tasks_one_after_the_other(backend_lib, tmp_path):
    """evaluate_scenes_parallel: worker processes sharing the GPU, every task seeding its own random streams — texts and
    merged statistics identical to running the same tasks in this process in task order."""
    from interactive_highway_simulation_b200 import evaluation as ev_mod
    rec = [r for r in RECORDS if r["kind"] == "lane_change"][0]
    path = write_scene(rec, tmp_path)
    egos = rec["ego_candidates"][:3]
    tasks = [(path, policy, 0, ego, 1.0, 40 + i) for i, (ego, policy) in
             enumerate([(e, p) for e in egos for p in ("lane_change_learned", "idm_mobil_lane_change_safe")])]
    texts, total = ev_mod.evaluate_scenes_parallel("lane_change", tasks, workers=3, steps=25, devices=[0])
    want_texts, parts = [], []
    for t in tasks:
        text, part = ev_mod._evaluate_task(("lane_change",) + t[:5] + (t[5], 0.2, 25, 0))
        want_texts.append(text); parts.append(part)
    want = ev_mod.merge_statistics("lane_change", parts)
    assert texts == want_texts
    assert total.__dict__ == want.__dict__ and len(total.utility_ego) == len(tasks)
