"""Throughput of independent evaluations in worker processes sharing the GPU (diagnostic).  usage: eval_parallel.py [workers...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
if __name__ == "__main__":
    import gzip, tempfile, yaml
    from interactive_highway_simulation_b200 import evaluation as ev
    HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with gzip.open(os.path.join(HERE, "tests", "golden", "evaluation_stats.json.gz"), "rt") as f:
        records = json.load(f)
    tmp = tempfile.mkdtemp(prefix="evalpar")
    for kind, policy in (("lane_change", "lane_change_learned"),):
        rec = [r for r in records if r["kind"] == kind][0]
        path = os.path.join(tmp, kind, "epoch0"); os.makedirs(path)
        for name in ("agents", "map"):
            with open(os.path.join(path, name + ".yml"), "w") as f:
                yaml.dump({int(k): v for k, v in rec[name].items()}, f, default_flow_style=False)
        egos = rec.get("ego_candidates") or [0]
        for workers in [int(a) for a in sys.argv[1:]] or [1, 4, 8, 15]:
            tasks = [(path, policy, 0, egos[i % len(egos)], 1.0, 100 + i) for i in range(4 * workers)]
            import torch
            devices = list(range(torch.cuda.device_count())) if os.environ.get("EVAL_ALL_GPUS") else [0]
            pool = ev.EvaluationPool(workers, devices)
            ev.evaluate_scenes_parallel(kind, tasks[:workers], pool=pool, devices=devices)
            t0 = time.perf_counter()
            ev.evaluate_scenes_parallel(kind, tasks, pool=pool, devices=devices)
            wall = time.perf_counter() - t0
            pool.close()
            print("%-12s %2d workers: %3d evaluations in %.2f s = %.1f evaluations/s" % (kind, workers, len(tasks), wall, len(tasks) / wall), flush=True)
