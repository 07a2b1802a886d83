#!/usr/bin/env python
"""bench.py — MC rollout throughput (vehicle-steps/s) and decision latency of the B200 rollout path, next to the
reference's own CPU path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's own machine code on the host cores)

Throughput (the JSON line's `value`): a "step" is one launch of the rollout kernel over one batch of rollouts of
BASELINE.json's throughput configuration (configs[4]): 256 vehicles x 100 steps, every non-ego vehicle `idm_lc`, ego
`keep_lane`, minSteps = steps (no early exit), on a scene of the reference's own generator (synthetic.generated_scene:
RandomTrafficGenerator with np.random.seed(s), random.seed(s), s = --scene-seed; `scene_seeds` repeats it for s = 0..9);
`--rollouts` rollouts per step per GPU (default 65536 = 1/16 of the 1M sweep, so the default run ends in minutes).
Rollouts are independent, so ranks shard the rollout index range with no data-path collective (weak scaling: per-GPU
batch fixed).  `roofline` is against the FP64 multiply-add rate measured in the same run (the FP64 pipe,
not HBM, bounds this path); `roofline_hbm` keeps the HBM view.

Decision latency (`decision`): BASELINE.json configs[1] — merging lane + 3 main lanes, 40 vehicles, 512 rollouts per
candidate gap, 4 candidates, SimParameter(0.3, 40, 10) — through the public API: scene upload, one launch for all
candidates, (N>1: NCCL all-gather of the 64-byte group partials), fixed-order combine, learned model, final gap action.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "mc_rollout_vehicle_steps_per_sec"
UNIT = "vehicle-steps/s"
FLOP_PER_VEHICLE_STEP = 100.0     # 63 base + 380 per MOBIL evaluation at ~1/10 duty
SIM_DT, SIM_STEPS = 0.3, 100
N_VEHICLES, N_LANES = 256, 4
DEC_VEHICLES, DEC_N_PER_GROUP, DEC_STEPS, DEC_MIN_STEPS = 40, 64, 40, 10
N_SM = 148


def workload_config(args):
    return {"workload": "configs[4]: synthetic throughput sweep, %d vehicles x %d steps IDM/MOBIL/RSS, ego keep_lane, "
                        "minSteps=steps" % (N_VEHICLES, SIM_STEPS),
            "vehicles": N_VEHICLES, "lanes": N_LANES, "horizon_steps": SIM_STEPS, "dt": SIM_DT,
            "scene": "RandomTrafficGenerator (random_traffic_generator.py:77-136 restated, pinned on the reference's records) "
                     "on the lanes of evaluation_lane_change.py:18-44 made %d vehicles long; np.random.seed = random.seed = scene_seed"
                     % (N_VEHICLES // N_LANES),
            "scene_seed": args.scene_seed,
            "rollouts_per_step_per_gpu": args.rollouts,
            "reference_arm_rollouts_per_step": 8 * max(1, args.ref_rollouts // 8),
            "note": "both arms run this workload; the product arm times rollouts_per_step_per_gpu rollouts per step, the "
                    "reference arm (--impl reference, host CPU) a bounded sample of reference_arm_rollouts_per_step rollouts "
                    "per step; the metric is a rate (vehicle-steps/s)",
            "l2": "flushed between timed steps (256 MiB write); the 66 KB scene is re-read by every block by design"}


def throughput_scene(args, seed=None):
    from interactive_highway_simulation_b200 import synthetic
    return synthetic.generated_scene(args.scene_seed if seed is None else seed, N_VEHICLES, N_LANES)


def read_peaks():
    p = os.path.join(HERE, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# ------------------------------------------------------------------------------------------------ reference arm
def scene_spec(corridors, agents):
    from oracle.pyoracle import SceneSpec
    spec = SceneSpec()
    lane_names = {0: "main", 1: "merging", 2: "exit"}
    for c in corridors:
        spec.add_corridor(c.id, lane_names.get(c.type, "other"), c.left[1], c.right[1], c.center[1], c.v_limit,
                          x0=c.left[0], x1=c.left[2])
    beh = {0: "idm", 1: "idm_lc", 2: "merging", 3: "exit"}
    for a in agents:
        spec.add_agent(a.id, a.x, a.y, a.v, a.vd, beh[a.behavior], vy=a.vy, a=a.a, l=a.l, w=a.w, idm=tuple(a.idm),
                       rss=tuple(a.rss), yp=tuple(a.yp))
    return spec


def reference_time(spec, dt, steps, min_steps, mode, m, ego, action, gap, reps=1):
    """Wall time of the reference's own machine code (oracle/_ref): runMultiThreads (`mt`: its own 8 std::threads x m)
    or runMTimes (`m`: one thread, m rollouts) on this box's host cores, platform libm, lock-free interposed rand()."""
    from oracle.pyoracle import RefHarness
    ref = RefHarness(math="libm")     # stock path: the platform libm, as the reference ships
    ref.load(spec, dt, steps, min_steps)
    ref._send("rng counter 12345")
    line, _ = ref.raw("bench %s %d %d %s %d %d" % (mode, m, ego, action, gap, reps), tag="bench")
    ref.close()
    kv = dict(t.split("=", 1) for t in line.split()[1:] if "=" in t)
    return int(kv["rollouts"]), float(kv["best_s"])


def reference_throughput(args, m, mode):
    corridors, agents, ego = throughput_scene(args)
    return reference_time(scene_spec(corridors, agents), SIM_DT, SIM_STEPS, SIM_STEPS, mode, m, ego, "keep_lane", -2)


def port_throughput(args, seconds=12.0):
    """BASELINE.md §3 item 3, the fair best-CPU number: the C++ restatement (oracle/mcsim_oracle.cpp — scalar, no
    shared_ptr / strings, one counter stream per rollout) on every host core (std::thread x nproc), same scene, same
    workload, bounded sample; best of the passes that fit the time budget."""
    from oracle.pyoracle import Oracle, ORACLE_SO
    from interactive_highway_simulation_b200 import _abi as abi
    if not os.path.exists(ORACLE_SO):
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "oracle"), "libmcsim_oracle.so"], check=True)
    corridors, agents, ego = throughput_scene(args)
    orc = Oracle(0)
    cores = os.cpu_count() or 1
    cand = abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0)
    sp = abi.SimParam(SIM_DT, SIM_STEPS, SIM_STEPS)
    count = 64 * cores
    st = (abi.Status * count)()
    best, passes, t_all = 1e30, 0, time.perf_counter()
    while passes < 5 and (passes == 0 or time.perf_counter() - t_all < seconds):
        t0 = time.perf_counter()
        rc = orc.lib.orc_rollouts_mt(corridors, len(corridors), agents, len(agents), C.byref(cand), C.byref(sp), 2026,
                                     passes * count, count, cores, st)
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError("oracle error %d" % rc)
        best = min(best, dt); passes += 1
    units = sum(s_.executedSteps for s_ in st) * N_VEHICLES
    return {"value": units / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d rollouts of the same scene per pass, %d passes (best), oracle/mcsim_oracle.cpp on %d std::threads, "
                      "one counter stream per rollout" % (count, passes, cores)}


def reference_decision_ms(args):
    """The reference's decision on the configs[1] scene: one runMultiThreads call per candidate gap, sequentially
    (behaviors.py:125-126), 8 x DEC_N_PER_GROUP rollouts each."""
    from interactive_highway_simulation_b200 import synthetic
    corridors, agents, ego, gaps = synthetic.merging_scene(DEC_VEHICLES, args.scene_seed)
    spec = scene_spec(corridors, agents)
    total = 0.0
    for g in gaps:
        _, secs = reference_time(spec, 0.3, DEC_STEPS, DEC_MIN_STEPS, "mt", DEC_N_PER_GROUP, ego, "left", g)
        total += secs
    return 1e3 * total, len(gaps)


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    m = max(1, args.ref_rollouts // 8)
    times, units = [], 0
    for i in range(args.warmup + args.steps):
        rollouts, secs = reference_throughput(args, m, "mt")
        if i >= args.warmup:
            times.append(secs); units += rollouts * N_VEHICLES * SIM_STEPS
    total = sum(times)
    value = units / total
    sample = "%d rollouts per step through the reference's runMultiThreads (8 std::threads x %d), platform libm, " \
             "lock-free interposed rand()" % (8 * m, m)
    dec_ms, n_gaps = reference_decision_ms(args)
    try:
        port = port_throughput(args)
    except Exception as e:
        port = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %s" % e}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(8, cores), "kind": "reference", "sample": sample},
           "cpu_baseline_port": port,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "decision": {"latency_ms": dec_ms, "candidates": n_gaps, "rollouts_per_candidate": 8 * DEC_N_PER_GROUP,
                        "vehicles": DEC_VEHICLES, "note": "sum of one runMultiThreads call per candidate gap"},
           "gpu_launches": 0}
    print(json.dumps(out), file=RESULT_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ FP64 peak probe
def fp64_dgemm_tflops(torch, device):
    """FP64 rate of a cuBLAS DGEMM (tensor-core DMMA) — a note beside the roofline, NOT its denominator: the rollout kernels
    issue scalar FP64 instructions, whose ceiling is the multiply-add stream measured by mcs_fp64_peak."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# ------------------------------------------------------------------------------------------------ decision latency
def decision_latency(torch, dist, world, device, args):
    """p50 wall time of one full learned merging decision through the public API (host buffers in, action out)."""
    from interactive_highway_simulation_b200 import backend, synthetic, sharding, learned_model, _abi as abi
    corridors, agents, ego, gaps = synthetic.merging_scene(DEC_VEHICLES, args.scene_seed)
    cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
    sp = abi.SimParam(0.3, DEC_STEPS, DEC_MIN_STEPS)
    model = learned_model.LearnedMergingModel()
    times = []
    decision = None
    for i in range(args.decisions + 5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        scene = backend.DeviceScene(corridors, agents, device=device)                    # H2D: the marshalled scene
        feats = sharding.simulation_multi_thread_sharded(scene, cands, sp, DEC_N_PER_GROUP, 1000 + i)
        idx, _ = model.inference([learned_model.feature_vector(f) for f in feats])
        decision = gaps[idx]
        act = scene.decide_fixed_gap(ego, abi.ACTIONS["left"], decision)                 # D2H: the chosen action
        t1 = time.perf_counter()
        scene.close()
        if i >= 5:
            times.append(t1 - t0)
    t = torch.tensor(times, dtype=torch.float64, device=device)
    parity = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the N-GPU features against the 1-GPU ones, byte for byte: every rank simulates all 8 groups on its own GPU once
        # and compares with what the sharded path (its 8/N groups + the all-gather) gave it; the verdicts are AND-ed
        scene = backend.DeviceScene(corridors, agents, device=device)
        sharded = sharding.simulation_multi_thread_sharded(scene, cands, sp, DEC_N_PER_GROUP, 4242)
        single, _ = scene.rollouts(cands, sp, DEC_N_PER_GROUP, 4242)
        scene.close()
        same = all(bytes(sharded[c]) == bytes(single[c]) for c in range(len(gaps)))
        flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = bool(int(flag[0]))
        if not parity:
            raise SystemExit("bench.py: features on %d GPUs differ from the 1-GPU features" % world)
    ts = sorted(t.tolist())
    return {"parity_vs_1gpu": parity, "latency_ms_p50": 1e3 * ts[len(ts) // 2], "latency_ms_p90": 1e3 * ts[int(len(ts) * 0.9)], "decisions": len(ts),
            "candidates": len(gaps), "rollouts_per_candidate": 8 * DEC_N_PER_GROUP, "vehicles": DEC_VEHICLES,
            "horizon_steps": DEC_STEPS, "last_decision_gap_id": decision, "last_action": [act.ax, act.vy],
            "h2d_bytes": C.sizeof(corridors) + C.sizeof(agents) + C.sizeof(cands), "d2h_bytes": 64 * 8 * len(gaps) + 16,
            "exchange": "none (1 GPU)" if world == 1 else "NCCL all_gather_into_tensor of %d B per rank" % (64 * 8 // world * len(gaps))}


# ------------------------------------------------------------------------------------------------ product arm
def outer_env_workload(n_scenes=512, n_vehicles=200, seed=0):
    """configs[3]-sized outer scenes on the map of evaluation_exit.py:18-43 (exit lane + 3 main lanes): positions drawn
    over the whole map, 2 % of the slots off the road."""
    import numpy as np
    from interactive_highway_simulation_b200 import traffic_generator as tg
    spec, _, _, _ = tg.evaluation_map("exit", v_limit=100 / 3.6)
    corridors = spec
    rng = np.random.default_rng(seed)
    x = rng.uniform(-200.0, 500.0, (n_scenes, n_vehicles))
    y = rng.uniform(0.0, 14.0, (n_scenes, n_vehicles))
    y[rng.random((n_scenes, n_vehicles)) < 0.02] = 30.0
    return corridors, x, y


def outer_env_block(args, reps=10):
    """Secondary measurement (SURVEY.md §8 rows a9/a10 outer twins, a13, a14): mce_env_match — lane match, per-lane
    ranking and the 10 neighbours of every vehicle — on 512 scenes x 200 vehicles through the C ABI with host buffers."""
    import numpy as np
    from interactive_highway_simulation_b200 import outer_env
    corridors, x, y = outer_env_workload()
    dm = outer_env.DeviceMap(corridors, 0)
    n_scenes, n = x.shape
    for _ in range(3):
        t = dm.match(x, y, n_scenes)
    k_ms, wall = [0.0, 0.0], 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        t = dm.match(x, y, n_scenes)
        wall += time.perf_counter() - t0
        ms = dm.last_kernel_ms()
        k_ms[0] += ms[0]; k_ms[1] += ms[1]
    k_ms = [v / reps for v in k_ms]
    # the same call with fewer tables fetched: lanes + arc lengths only, and nothing (everything stays resident on the device
    # for the neighbour queries and the builders)
    walls = {}
    for name, fetch in (("lanes_and_arcs", ("lane", "arc")), ("resident_only", ())):
        dm.match(x, y, n_scenes, fetch=fetch)
        t0 = time.perf_counter()
        for _ in range(reps):
            dm.match(x, y, n_scenes, fetch=fetch)
        walls[name] = (time.perf_counter() - t0) / reps * 1e3
    vehicles = n_scenes * n
    alg_bytes = vehicles * (16 + 4 + 8 + 4 + 10 * 12) + n_scenes * 36      # x, y in; lane, arc, order, 10 x (slot, dis) out
    peak_gbs, peak_src = read_peaks()
    out = {"workload": "%d scenes x %d vehicles, exit map of evaluation_exit.py (4 lanes)" % (n_scenes, n),
           "matched_fraction": float((t.lane >= 0).mean()),
           "kernel_ms": {"env_match_kernel": k_ms[0], "env_neighbors_kernel": k_ms[1]},
           "vehicles_per_s_kernels": vehicles / ((k_ms[0] + k_ms[1]) * 1e-3),
           "e2e": {"vehicles_per_s": vehicles * reps / wall, "ms_per_call": wall / reps * 1e3,
                   "h2d_bytes_per_call": vehicles * 16, "d2h_bytes_per_call": alg_bytes - vehicles * 16,
                   "ms_per_call_lanes_and_arcs_fetched": walls["lanes_and_arcs"], "ms_per_call_resident_only": walls["resident_only"],
                   "note": "pageable host buffers; the full call reads 13.9 MB of tables back, which is what it costs — a caller "
                           "fetches the tables it reads (any output may be NULL) and the rest stays resident for the queries"},
           "roofline": {"bound": "hbm", "achieved": alg_bytes / ((k_ms[0] + k_ms[1]) * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                        "frac": alg_bytes / ((k_ms[0] + k_ms[1]) * 1e-3) / 1e9 / peak_gbs, "peak_source": peak_src,
                        "algorithmic_bytes_per_vehicle": 152,
                        "note": "152 B per vehicle in and out; the kernels are compute/latency bound (point-in-polygon and "
                                "projection per corridor, O(n) rank per vehicle), not HBM bound"}}
    if not args.no_cpu_baseline:
        from oracle import env_oracle as eo                       # the checker, timed as the CPU baseline only
        pos = {c["id"]: k for k, c in enumerate(corridors)}
        pm = eo.PreparedMap([dict(c, left_index=pos[c["left_id"]] if c["left_id"] is not None else -1,
                                  right_index=pos[c["right_id"]] if c["right_id"] is not None else -1) for c in corridors])
        t0 = time.perf_counter()
        done = 0
        while done < 4 and time.perf_counter() - t0 < 10.0:
            xs, ys = list(x[done]), list(y[done])
            lane, arc, order, start = eo.match(pm, xs, ys)
            for i in range(n):
                eo.neighbors(pm, lane, arc, order, start, i, xs[i], ys[i])
            done += 1
        out["cpu_baseline"] = {"value": done * n / (time.perf_counter() - t0), "unit": "vehicles/s", "cores": 1, "kind": "port",
                               "sample": "%d scenes of %d vehicles through the Python restatement of environment.py (the reference's "
                                         "own path is Python + shapely per vehicle)" % (done, n)}
    dm.close()
    return out


def decision_configs_block(reps=10):
    """One decision at the size of each remaining BASELINE.json configuration (configs[1] is the `decision` block): kernel
    time (CUDA events inside mcs_rollouts_device) and end-to-end time of scene upload + all candidates' rollouts + feature
    read-back through the C ABI with host buffers."""
    from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
    acts = ["keep_lane", "dcc", "acc", "left", "right"]
    out = []

    def measure(name, corridors, agents, cands, sp, n_per_group, seed):
        sc = backend.DeviceScene(corridors, agents)
        R = 8 * n_per_group
        k_ms, units = 1e9, 0
        for _ in range(3):
            ms, units = sc.rollouts_device(cands, sp, seed, 0, R)
            k_ms = min(k_ms, ms)
        sc.close()
        walls = []
        for i in range(reps + 2):
            t0 = time.perf_counter()
            sc = backend.DeviceScene(corridors, agents)
            sc.rollouts(cands, sp, n_per_group, seed + i)
            sc.close()
            walls.append((time.perf_counter() - t0) * 1e3)
        walls = sorted(walls[2:])
        out.append({"config": name, "candidates": len(cands), "rollouts_per_candidate": R, "vehicles": len(agents), "horizon_steps": sp.steps,
                    "kernel_ms": k_ms, "e2e_ms_p50": walls[len(walls) // 2], "vehicle_steps_per_s_kernel": units / k_ms * 1e3})

    c, a, ego = synthetic.lane_change_scene(100, 4, seed=1)
    measure("configs[2] lane_change_learned, 100 vehicles, 5 actions x 160 rollouts x 40 steps", c, a,
            (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[x], -2, 0) for x in acts]), abi.SimParam(0.3, 40, 30), 20, 2)
    c, a, ego, gaps = synthetic.exit_scene(200, 0)
    measure("configs[3] exit_learned, 200 vehicles, %d gaps x 4096 rollouts x 50 steps" % len(gaps), c, a,
            (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["right"], g, 0) for g in gaps]), abi.SimParam(0.3, 50, 10), 512, 3)
    return out


def outer_loop_block(steps=100):
    """configs[0]: simulation_multilane on the shipped test_scene (21 vehicles, merging lane + 2 main lanes), dt 0.2, 100
    steps, headless — every step is the sequential plan-and-move loop (device neighbour tables / border distances, the MOBIL
    and closest-gap decision kernels, device step_along_path) and one re-match.  Wall clock of the public call."""
    import random
    import numpy as np
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, simulation_multilane as sm
    scene = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "test_scene_epoch0")
    out = {}
    for rep in range(2):                                       # the first pass warms the context / allocator
        random.seed(0); np.random.seed(0); ms.seed(1000)
        sim = sm.Simulation(None, sm.SimulationParameter(0.2, steps, render=False, write_gif=False))
        sim.load_initial_scene(scene)
        n0 = len(sim.environment.agents)
        t0 = time.perf_counter()
        sim.run()
        wall = time.perf_counter() - t0
        hist = sum(len(a.state_history) - 1 for a in list(sim.environment.agents.values()) + list(sim.environment.removed_agents.values()))
        out = {"workload": "test_scene epoch0, %d vehicles, dt 0.2, %d steps, render=False" % (n0, steps),
               "wall_s": wall, "ms_per_step": wall / steps * 1e3, "vehicle_steps": hist, "vehicle_steps_per_s": hist / wall,
               "vehicles_left_on_map": len(sim.environment.agents),
               "note": "sequential outer loop of the reference (environment.py:316-323): one small launch per query, latency bound"}
    rec = os.path.join(HERE, "profiles", "r02_reference_outer_loop_cpu.json")
    if os.path.exists(rec):
        out["cpu_baseline"] = json.load(open(rec))     # scripts/time_reference_outer_loop.py (cannot run on the GPU box)
    return out


def evaluation_block(steps=100):
    """the evaluation scripts' per-scene functions (evaluation_merging.py:53-106,
    evaluation_lane_change.py:58-101, evaluation_exit.py:55-105) with the LEARNED ego policy through the public API — yaml scene
    on disk -> headless outer loop on the device (lane match, builders, decisions, rollouts) -> statistics — on the epochs of
    tests/golden/evaluation_stats.json.gz, for the scripts' own horizon (dt 0.2, 100 steps); beside it the reference's own
    functions on the same epochs (recorded in the build container, scripts/time_reference_evaluation.py)."""
    import gzip
    import random
    import tempfile
    import numpy as np
    import yaml
    from interactive_highway_simulation_b200 import evaluation as ev, mc_sim_highway as ms
    with gzip.open(os.path.join(HERE, "tests", "golden", "evaluation_stats.json.gz"), "rt") as f:
        records = json.load(f)
    learned = {"merging": "merging_learned", "lane_change": "lane_change_learned", "exit": "exit_learned"}
    tmp = tempfile.mkdtemp(prefix="bench_eval")
    out = {}
    for rec in records:
        kind = rec["kind"]
        path = os.path.join(tmp, kind, "epoch0")
        os.makedirs(path)
        for name in ("agents", "map"):
            with open(os.path.join(path, name + ".yml"), "w") as f:
                yaml.dump({int(k): v for k, v in rec[name].items()}, f, default_flow_style=False)
        best, info = 1e30, None
        for rep in range(2):                     # the first pass warms the context
            sim = ev.Simulation(None, ev.SimulationParameter(0.2, steps, render=False, write_gif=False, scenario=kind))
            st = ev.LaneChangeStatistics() if kind == "lane_change" else ev.MergingAndExitStatistics()
            random.seed(7); np.random.seed(7); ms.seed(7)
            t0 = time.perf_counter()
            if kind == "merging":
                ev.evaluate_recorded_merging_scene(path, sim, learned[kind], 0, st, 1.0)
            elif kind == "lane_change":
                ev.evaluate_recorded_lc_scene(path, sim, learned[kind], 0, rec["ego_candidates"][0], st)
            else:
                ev.evaluate_recorded_exit_scene(path, sim, learned[kind], 0, rec["ego_candidates"][0], st, 1.0)
            wall = time.perf_counter() - t0
            if wall < best:
                best, info = wall, {"policy": learned[kind], "vehicles": len(rec["agents"]), "executed_steps": sim.step + 1,
                                    "wall_s": wall, "evaluations_per_s": 1.0 / wall}
        out[kind] = info
    res = {"horizon": "dt 0.2, %d steps" % steps, "kinds": out,
           "note": "one evaluation = load the yaml epoch, give the ego its learned policy, run the scene, accumulate the statistics; "
                   "every step the ego makes one learned decision (4-5 candidates x 160-240 rollouts in one launch)"}
    # the scripts' 500-epoch loops are independent evaluations: they shard over the GPUs of the box (evaluation.py
    # evaluate_scenes_parallel: one worker process per device; on a one-GPU box this only shows the cost of the worker pool)
    try:
        import torch
        devices = list(range(torch.cuda.device_count()))
        rec = [r for r in records if r["kind"] == "lane_change"][0]
        path = os.path.join(tmp, "lane_change", "epoch0")
        egos = rec["ego_candidates"]
        tasks = [(path, learned["lane_change"], 0, egos[i % len(egos)], 1.0, 100 + i) for i in range(4 * len(devices))]
        pool = ev.EvaluationPool(len(devices), devices)
        try:
            ev.evaluate_scenes_parallel("lane_change", tasks[:len(devices)], pool=pool, steps=steps, devices=devices)   # contexts, caches
            t0 = time.perf_counter()
            _, total = ev.evaluate_scenes_parallel("lane_change", tasks, pool=pool, steps=steps, devices=devices)
            wall = time.perf_counter() - t0
        finally:
            pool.close()
        res["sharded"] = {"kind": "lane_change", "policy": learned["lane_change"], "devices": len(devices), "workers": len(devices),
                          "evaluations": len(tasks), "wall_s": wall, "evaluations_per_s": len(tasks) / wall,
                          "statistics_entries": len(total.utility_ego),
                          "note": "independent (epoch, ego) evaluations dealt to one worker process per visible GPU, each task seeding "
                                  "its own random streams; worker processes sharing ONE GPU do not overlap (measured 5.5 -> 5.8 "
                                  "evaluations/s from 1 to 15 workers), so the shard unit is the GPU"}
    except Exception as exc:      # a secondary block must not cost the bench line
        res["sharded"] = {"error": repr(exc)}
    recf = os.path.join(HERE, "profiles", "r02_reference_evaluation_cpu.json")
    if os.path.exists(recf):
        res["cpu_baseline"] = json.load(open(recf))
    return res


def run_product(args):
    import torch
    from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the rollout path has no CPU fallback")
    device = local if world > 1 else 0
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    corridors, agents, ego = throughput_scene(args)
    scene = backend.DeviceScene(corridors, agents, device=device)
    cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0))
    sp = abi.SimParam(SIM_DT, SIM_STEPS, SIM_STEPS)
    R = args.rollouts
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    seed = 2026

    def one_step(i):
        flush.zero_()
        torch.cuda.synchronize(device)
        first = (i * world + rank) * R        # disjoint rollout index ranges per rank and step
        return scene.rollouts_device(cand, sp, seed, first, R)     # CUDA events on the kernel's own stream

    for i in range(args.warmup):
        one_step(i)
    barrier()
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    kernel_ms, vsteps = 0.0, 0
    for i in range(args.steps):
        ms, vs = one_step(args.warmup + i)
        kernel_ms += ms; vsteps += vs
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end through the reference-facing call (simulationMultiThread semantics): host buffers in, every step
    # uploads the scene (mcs_scene_create), launches, reduces in fixed order and reads the features back
    e2e_R = max(8, (R // 8) * 8)
    h2d = C.sizeof(corridors) + C.sizeof(agents) + C.sizeof(cand)
    d2h = C.sizeof(abi.Feature)
    e2e_steps = args.steps             # as many timed end-to-end steps as kernel-timed ones
    sc2 = backend.DeviceScene(corridors, agents, device=device); sc2.rollouts(cand, sp, 8, seed); sc2.close()   # warm-up
    barrier()
    e2e_t0 = time.perf_counter()
    e2e_units = 0
    for i in range(e2e_steps):
        sc2 = backend.DeviceScene(corridors, agents, device=device)
        feats, _ = sc2.rollouts(cand, sp, e2e_R // 8, seed + 17 + i)
        e2e_units += e2e_R * N_VEHICLES * SIM_STEPS
        sc2.close()
    torch.cuda.synchronize(device)
    e2e_s = time.perf_counter() - e2e_t0

    decision = decision_latency(torch, dist, world, device, args) if args.decisions > 0 else None

    # ---- max over ranks, whole-job aggregate
    t = torch.tensor([kernel_ms, float(vsteps), e2e_s, float(e2e_units)], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tsum = t
    kernel_ms_max, units_total = float(tmax[0]), float(tsum[1])
    e2e_s_max, e2e_units_total = float(tmax[2]), float(tsum[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = units_total / (kernel_ms_max * 1e-3)
    # ---- roofline of the rollout kernel (one launch = one step).  SURVEY.md §8(d): the binding resource is the FP64 pipe
    # (and instruction issue), not HBM and not the tensor cores.  achieved = algorithmic flop per launch (100 per
    # vehicle-step x the vehicle-steps the launch executed, counted on the device) / the launch's average duration (CUDA
    # events on the kernel's stream inside the timed region); peak = a stream of independent FP64 fused multiply-adds
    # measured by libmcsim_b200's own probe kernel in this run (MEASURED_PEAKS.json holds no FP64 figure).
    peak_gbs, peak_src = read_peaks()
    launch_s = kernel_ms / args.steps * 1e-3
    fp64_peak, fp64_ms = backend.fp64_peak_tflops(device)
    fp64_achieved = (vsteps / args.steps) * FLOP_PER_VEHICLE_STEP / launch_s / 1e12
    traffic = None
    prof = os.path.join(HERE, "profiles", "r02_rollout_kernel_dram.json")
    if not os.path.exists(prof):
        prof = os.path.join(HERE, "profiles", "r01_rollout_kernel_dram.json")
    try:
        pj = json.load(open(prof))
        traffic = pj.get("dram_bytes_per_rollout", 0) * R or None
    except Exception:
        pj = {}
    roofline = {"bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
                "traffic": traffic,
                "flop_per_vehicle_step": FLOP_PER_VEHICLE_STEP,
                "peak_source": "measured in this run: mcs_fp64_peak, 8 independent DFMA chains per thread, 2048 threads per SM "
                               "(%.2f ms); the CUDA-core FP64 rate, which is what a scalar FP64 kernel can reach" % fp64_ms,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel "
                                  "(%s), scaled to this launch's rollouts; bytes, not flop — see roofline_hbm" % os.path.basename(prof),
                "fp64_tensor_dgemm_tflops": fp64_dgemm_tflops(torch, device),
                "note": "100 algorithmic FP64 flop per vehicle-step (63 base IDM + RSS + lateral + integration, + 380 per MOBIL "
                        "evaluation at ~1/10 duty); the kernel's own limiter is instruction issue and dependent-instruction "
                        "latency of FP64 / control code (profiles/: issue slots, FP64 pipe, stall breakdown)"}
    scene_bytes = C.sizeof(corridors) + 30 * 8 * 256 + 8 * 4 * 256 + 256     # SoA scene read once (L2-shared)
    alg_bytes = R * 64 + scene_bytes
    achieved_gbs = alg_bytes / launch_s / 1e9
    roofline_hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                    "traffic": traffic, "peak_source": peak_src,
                    "note": "algorithmic HBM traffic is 64 B per rollout + one scene read (state lives in registers / shared "
                            "memory for the whole rollout): by design the kernel is nowhere near the HBM roof"}

    # ---- the same workload on the scenes of generator seeds 0..9: one warm-up + two timed launches each
    scene_seeds = None
    if world == 1 and args.seed_sweep:
        per_seed = []
        Rs = min(R, 16384)
        for sd in range(10):
            c2, a2, ego2 = throughput_scene(args, seed=sd)
            sc = backend.DeviceScene(c2, a2, device=device)
            cd = (abi.Candidate * 1)(abi.Candidate(ego2, abi.ACTIONS["keep_lane"], -2, 0))
            sc.rollouts_device(cd, sp, seed, 0, Rs)
            ms_sum, vs_sum = 0.0, 0
            for rep in range(2):
                flush.zero_(); torch.cuda.synchronize(device)
                ms, vs = sc.rollouts_device(cd, sp, seed, (rep + 1) * Rs, Rs)
                ms_sum += ms; vs_sum += vs
            sc.close()
            per_seed.append(vs_sum / (ms_sum * 1e-3))
        srt = sorted(per_seed)
        scene_seeds = {"seeds": list(range(10)), "rollouts_per_launch": Rs, "vehicle_steps_per_s": per_seed,
                       "median": 0.5 * (srt[4] + srt[5]), "min": srt[0], "max": srt[-1]}

    # ---- CPU baseline: the reference's own machine code on this box's host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            m = max(1, args.ref_rollouts // 8)
            rollouts, secs = reference_throughput(args, m, "mt")
            r1, s1 = reference_throughput(args, max(1, m), "m")
            cpu = {"value": rollouts * N_VEHICLES * SIM_STEPS / secs, "unit": UNIT, "cores": min(8, os.cpu_count() or 1),
                   "kind": "reference",
                   "sample": "%d rollouts of the same scene through the reference's runMultiThreads (its hard-wired 8 threads; "
                             "this box has %d cores), platform libm; single-thread runMTimes on %d rollouts: %.3g vehicle-steps/s"
                             % (rollouts, os.cpu_count() or 1, r1, r1 * N_VEHICLES * SIM_STEPS / s1)}
            if decision is not None:
                dec_ms, _ = reference_decision_ms(args)
                decision["cpu_reference_latency_ms"] = dec_ms
        except Exception as e:    # oracle/_ref is built only where /root/reference exists
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    cpu_port = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_port = port_throughput(args)
        except Exception as e:
            cpu_port = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %s" % e}

    outer_env = None
    if world == 1:
        try:
            outer_env = outer_env_block(args)
        except Exception as e:
            outer_env = {"error": str(e)}
    outer_loop = None
    if world == 1:
        try:
            outer_loop = outer_loop_block()
        except Exception as e:
            outer_loop = {"error": str(e)}
    decision_configs = None
    if world == 1:
        try:
            decision_configs = decision_configs_block()
        except Exception as e:
            decision_configs = {"error": str(e)}
    evaluation = None
    if world == 1:
        try:
            evaluation = evaluation_block()
        except Exception as e:
            evaluation = {"error": str(e)}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args),
           "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "cpu_baseline_port": cpu_port,
           "scene_seeds": scene_seeds, "clocks": clocks,
           "value_wall": {"value": units_total / wall, "unit": UNIT,
                          "note": "the same vehicle-steps over the wall clock of the timed region (barrier to barrier: L2 flush, "
                                  "candidate upload, launch and read-back overheads included); `value` divides by the CUDA-event "
                                  "time of the kernels alone; compare CPU baselines (wall clock) with this or with `e2e`"},
           "e2e": {"value": e2e_units_total / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h,
                   "note": "mcs_scene_create + mcs_rollouts (8 groups x %d) + feature read-back per step, host buffers" % (e2e_R // 8)},
           "decision": decision, "decision_configs": decision_configs, "outer_env": outer_env, "outer_loop": outer_loop,
           "evaluation": evaluation,
           "gpu_launches": args.steps, "wall_s_timed_region": wall,
           "rollouts_per_s": units_total / (N_VEHICLES * SIM_STEPS) / (kernel_ms_max * 1e-3)}
    print(json.dumps(out), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rollouts", type=int, default=65536, help="rollouts per step per GPU")
    ap.add_argument("--ref-rollouts", type=int, default=192, help="rollouts per CPU-reference sample")
    ap.add_argument("--decisions", type=int, default=40, help="timed decisions for the latency figure (0 = skip)")
    ap.add_argument("--scene-seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-seed-sweep", dest="seed_sweep", action="store_false", help="skip the scene seeds 0..9 block")
    args = ap.parse_args()
    # exactly ONE line on stdout: whatever libraries write to file descriptor 1 (NCCL prints its version banner there when
    # NCCL_DEBUG is set in the environment) goes to stderr; the JSON line goes to the saved descriptor
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_product(args)


if __name__ == "__main__":
    sys.exit(main())
