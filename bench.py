#!/usr/bin/env python
"""bench.py — MC rollout throughput (vehicle-steps/s) of the B200 rollout path, next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's own machine code on the host cores)

A "step" is one launch of the rollout kernel over one batch of rollouts of BASELINE.json's throughput
configuration (configs[4]): 256 vehicles x 100 steps, every non-ego vehicle `idm_lc`, ego `keep_lane`,
minSteps = steps (no early exit); `--rollouts` rollouts per step per GPU (default 65536 = 1/16 of the 1M sweep,
so the default run ends in minutes).  Rollouts are independent, so ranks shard the rollout index range with no
data-path collective (weak scaling: per-GPU batch fixed).
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


def workload_config(args):
    return {"workload": "configs[4]: synthetic throughput sweep, %d vehicles x %d steps IDM/MOBIL/RSS, ego keep_lane, "
                        "minSteps=steps" % (N_VEHICLES, SIM_STEPS),
            "vehicles": N_VEHICLES, "lanes": N_LANES, "horizon_steps": SIM_STEPS, "dt": SIM_DT,
            "rollouts_per_step_per_gpu": args.rollouts, "scene_seed": args.scene_seed,
            "l2": "flushed between timed steps (256 MiB write); the 66 KB scene is re-read by every block by design"}


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


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------ reference arm
def reference_rollouts_per_s(args, sample_m, threads_mode):
    """Time the reference's own machine code (oracle/_ref) on this box's host cores: runMultiThreads (its own
    8-thread fan-out) or runMTimes (one thread) over a bounded sample of the same scene."""
    from oracle.pyoracle import RefHarness, SceneSpec
    from interactive_highway_simulation_b200 import synthetic
    corridors, agents, ego = synthetic.lane_change_scene(N_VEHICLES, N_LANES, seed=args.scene_seed)
    spec = SceneSpec()
    lane_names = {0: "main", 1: "merging", 2: "exit"}
    for c in corridors:
        spec.add_corridor(c.id, lane_names.get(c.type, "other"), c.left[1], c.right[1], c.center[1], c.v_limit,
                          x0=c.left[0], x1=c.left[2])
    beh = {0: "idm", 1: "idm_lc", 2: "merging", 3: "exit"}
    for a in agents:
        spec.add_agent(a.id, a.x, a.y, a.v, a.vd, beh[a.behavior], vy=a.vy, a=a.a, l=a.l, w=a.w, idm=tuple(a.idm),
                       rss=tuple(a.rss), yp=tuple(a.yp))
    ref = RefHarness(math="libm")     # stock path: the platform libm, as the reference ships
    ref.load(spec, SIM_DT, SIM_STEPS, SIM_STEPS)
    ref._send("rng counter 12345")
    line, _ = ref.raw("bench %s %d %d keep_lane -2 1" % (threads_mode, sample_m, ego), tag="bench")
    ref.close()
    kv = dict(t.split("=", 1) for t in line.split()[1:] if "=" in t)
    rollouts = int(kv["rollouts"]); secs = float(kv["best_s"])
    return rollouts, secs


def run_reference(args):
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    m = max(1, args.ref_rollouts // 8)
    times, units = [], 0
    for i in range(args.warmup + args.steps):
        rollouts, secs = reference_rollouts_per_s(args, m, "mt")
        if i >= args.warmup:
            times.append(secs); units += rollouts * N_VEHICLES * SIM_STEPS
    total = sum(times)
    value = units / total
    cfg = workload_config(args)
    sample = "%d rollouts per step through the reference's runMultiThreads (8 std::threads x %d), platform libm, " \
             "lock-free interposed rand()" % (8 * m, m)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(8, cores), "kind": "reference", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ FP64 peak probe
def fp64_peak_tflops(torch, device):
    """Measured FP64 FMA rate of this GPU via a torch DGEMM (cuBLAS); the denominator of the FP64 view of the roofline."""
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


# ------------------------------------------------------------------------------------------------ product arm
def run_product(args):
    import torch
    from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi

    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the rollout path has no CPU fallback")
    device = local if world > 1 else 0
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    corridors, agents, ego = synthetic.lane_change_scene(N_VEHICLES, N_LANES, seed=args.scene_seed)
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
        return scene.rollouts_device(cand, sp, seed, first, R)

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

    # ---- end-to-end through the reference-facing call (simulationMultiThread semantics): host buffers in,
    # scene upload + launch + fixed-order reduction + feature read-back inside the timed region
    e2e_R = max(8, (R // 8) * 8)
    h2d = C.sizeof(corridors) + C.sizeof(agents) + C.sizeof(cand)
    d2h = C.sizeof(abi.Feature)
    e2e_t0 = time.perf_counter()
    e2e_units = 0
    e2e_steps = max(1, min(args.steps, 3))
    for i in range(e2e_steps):
        sc2 = backend.DeviceScene(corridors, agents, device=device)
        feats, _ = sc2.rollouts(cand, sp, e2e_R // 8, seed + 17 + i)
        e2e_units += e2e_R * N_VEHICLES * SIM_STEPS
        sc2.close()
    torch.cuda.synchronize(device)
    e2e_s = time.perf_counter() - e2e_t0

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
    # roofline of the rollout kernel (one launch = one step)
    peak_gbs, peak_src = read_peaks()
    launch_s = kernel_ms / args.steps * 1e-3
    scene_bytes = C.sizeof(corridors) + 28 * 8 * 256 + 8 * 4 * 256 + 256     # SoA scene read once (L2-shared)
    alg_bytes = R * 64 + scene_bytes
    achieved_gbs = alg_bytes / launch_s / 1e9
    fp64_peak = fp64_peak_tflops(torch, device)
    fp64_achieved = (vsteps / args.steps) * FLOP_PER_VEHICLE_STEP / launch_s / 1e12
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                "traffic": None, "peak_source": peak_src,
                "note": "algorithmic HBM traffic is 64 B per rollout + one scene read: the kernel is FP64 issue/latency "
                        "bound, see fp64"}
    prof = os.path.join(HERE, "profiles", "r01_rollout_kernel_dram.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    fp64 = {"achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
            "flop_per_vehicle_step": FLOP_PER_VEHICLE_STEP, "peak_source": "cuBLAS DGEMM 4096^3 measured in this run"}

    # ---- CPU baseline: the reference's own machine code on this box's host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            m = max(1, args.ref_rollouts // 8)
            rollouts, secs = reference_rollouts_per_s(args, m, "mt")
            r1, s1 = reference_rollouts_per_s(args, max(1, m), "m")
            cpu = {"value": rollouts * N_VEHICLES * SIM_STEPS / secs, "unit": UNIT, "cores": min(8, os.cpu_count() or 1),
                   "kind": "reference",
                   "sample": "%d rollouts of the same scene through the reference's runMultiThreads (8 threads), platform "
                             "libm; single-thread runMTimes on %d rollouts: %.3g vehicle-steps/s"
                             % (rollouts, r1, r1 * N_VEHICLES * SIM_STEPS / s1)}
        except Exception as e:    # oracle/_ref is built only where /root/reference exists
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args),
           "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu, "clocks": clocks,
           "e2e": {"value": e2e_units_total / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h,
                   "note": "mcs_scene_create + mcs_rollouts (8 groups x %d) + feature read-back per step, host buffers" % (e2e_R // 8)},
           "gpu_launches": args.steps, "wall_s_timed_region": wall,
           "rollouts_per_s": units_total / (N_VEHICLES * SIM_STEPS) / (kernel_ms_max * 1e-3)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rollouts", type=int, default=65536, help="rollouts per step per GPU")
    ap.add_argument("--ref-rollouts", type=int, default=192, help="rollouts per CPU-reference sample")
    ap.add_argument("--scene-seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_product(args)


if __name__ == "__main__":
    sys.exit(main())
