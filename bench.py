#!/usr/bin/env python
"""bench.py -- Msamples/s (paths x bounces) of the guided path tracer on CBOX, one JSON line.

A "step" is one complete guided render (all training iterations + the final iteration, tree maintenance
included) of the workload below through the C ABI of libppg_b200.so.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                          the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[1]): CBOX 1024x1024, default Mueller'17 parameters (sppPerPass 4, maxDepth 10,
rrDepth 10, strictNormals, nearest/nearest filters, sampleCombination automatic, sTreeThreshold 12000).  The
reference quotes it on a 60 s budget; Msamples/s is a rate, so a step uses an equal-spp budget instead
(budgetType=spp, default 252 spp = 63 passes = iterations of 1,2,4,8,16,32 passes) so that a step takes about a
second and equal-spp image comparisons stay meaningful.  --budget-seconds runs the literal 60 s configuration.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200"))

import numpy as np  # noqa: E402


def load_scene(size):
    from ppg_b200.scene import SceneDesc
    return SceneDesc.load(os.path.join(ROOT, "scenes", "cbox.npz")).with_film(size, size)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, torch copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm

def cpu_run(kind_pref, size, budget, nthreads):
    """Times the CPU oracle (the reference algorithm: restated tracer; SD-tree compiled verbatim from the reference when
    oracle/_ref/libppg_oracle_ref.so travelled) on all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    kind = "ref" if (kind_pref == "ref" and O.have_ref()) else "port"
    sc = load_scene(size)
    props = dict(sc.integrator, budget=str(budget))
    o = O.Oracle(O.params_from_xml(props), sc, nthreads=nthreads, kind=kind)
    t = time.perf_counter()
    img, st = o.render()
    dt = time.perf_counter() - t
    o.close()
    return {"seconds": dt, "vertices": st["total_vertices"], "paths": st["total_paths"], "msamples": st["total_vertices"] / dt / 1e6,
            "kind": kind, "image": img, "stats": st}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    size, budget = args.cpu_size, args.cpu_budget
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_run("ref", size, budget, cores)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["msamples"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals]) * 1e3)
    kind = "reference" if vals[0]["kind"] == "ref" else "port"
    sample = (f"CBOX {size}x{size}, default params, budget {budget} spp ({vals[0]['paths']} paths, {vals[0]['vertices']} vertices per step); "
              f"CPU tracer restated from guided_path.cpp, SD-tree {'compiled verbatim from the reference' if kind == 'reference' else 'restated'}; OpenMP over 32x32 blocks")
    line = {
        "impl": "reference", "metric": "Msamples/sec (paths x bounces)", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, bounded=f"CPU arm runs a bounded sample of it: {size}x{size}, {budget} spp"),
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port" if kind == "port" else "reference", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, bounded=None):
    c = {"workload": f"CBOX {args.size}x{args.size}, default Mueller'17 params (sppPerPass=4, maxDepth=10, rrDepth=10, nearest/nearest, automatic, sTreeThreshold=12000), "
                     + (f"budgetType=seconds budget={args.budget_seconds}" if args.budget_seconds else f"budgetType=spp budget={args.budget} (equal-spp stand-in for the 60 s budget)"),
         "scene": "scenes/cbox.npz (flat-array form of the reference's scenes/cbox/cbox.xml)", "sharding": f"32x32 image blocks interleaved over {args.gpus} rank(s), one tree allreduce per training iteration",
         "l2": "inputs larger than L2: path state + vertex records of one pass-batch are ~1 GB, every kernel streams them once"}
    if bounded:
        c["bounded_sample"] = bounded
    return c


# ------------------------------------------------------------------------------------------- GPU arm

def gpu_arm(args):
    import torch
    from ppg_b200 import capi
    from ppg_b200.integrator import GuidedPathTracer, torch_allreduce
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = load_scene(args.size)
    props = dict(sc.integrator)
    if args.budget_seconds:
        props.update(budgetType="seconds", budget=str(args.budget_seconds))
    else:
        props.update(budgetType="spp", budget=str(args.budget))
    g = GuidedPathTracer(props, device=local)
    g.set_scene(sc)
    if world > 1:
        g.set_shard(rank, world)
        g.set_allreduce(torch_allreduce())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- `value`: inputs resident in HBM, film left in HBM
    for _ in range(args.warmup):
        g.render_device()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    t0 = time.perf_counter()
    stats = []
    for _ in range(args.steps):
        _, st = g.render_device()
        stats.append(st)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = sum(s["render_device_ms"] for s in stats)          # CUDA events on the library's launching stream
    verts = sum(s["total_vertices"] for s in stats); paths = sum(s["total_paths"] for s in stats)
    launches = sum(s["kernel_launches"] for s in stats)
    if dist is not None:
        t = torch.tensor([dev_ms, wall * 1e3], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
        c = torch.tensor([verts, paths, launches], device="cuda", dtype=torch.float64); dist.all_reduce(c, op=dist.ReduceOp.SUM)
        verts, paths, launches = (int(x) for x in c.tolist())
    else:
        wall_ms = wall * 1e3
    value = verts / (dev_ms * 1e-3) / 1e6

    # ---- `e2e`: host buffers through the public call; scene upload (H2D) and film download (D2H) inside the timed region
    arrays = g._scene_arrays
    h2d = int(sum(a.nbytes for a in (arrays.positions, arrays.normals, arrays.uvs, arrays.indices, arrays.triangle_shape, arrays.shapes, arrays.bsdfs, arrays.radiance)))
    d2h = args.size * args.size * 3 * 4
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter(); ev = 0
    for _ in range(e2e_steps):
        g.set_scene(sc)
        if world > 1:
            g.set_shard(rank, world)
        img, st = g.render()
        ev += st["total_vertices"]
    barrier()
    e2e_wall = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_wall], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_wall = float(t[0])
        c = torch.tensor([ev], device="cuda", dtype=torch.float64); dist.all_reduce(c, op=dist.ReduceOp.SUM); ev = int(c[0])
    e2e_value = ev / e2e_wall / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (bounce: raygen/intersect/shade/guide/compact), SURVEY 8d
    peak, peak_src = peaks()
    train_v = sum(i["vertices"] for s in stats for i in s["iterations"] if not i["is_final"])
    final_v = sum(i["vertices"] for s in stats for i in s["iterations"] if i["is_final"])
    rec_w = [(i["recorded_vertices"], i["s_tree_depth_avg"]) for s in stats for i in s["iterations"] if i["recorded_vertices"]]
    d_s = float(sum(n * d for n, d in rec_w) / max(1, sum(n for n, _ in rec_w))) if rec_w else 11.0
    dd = [i["depth_avg"] for s in stats for i in s["iterations"] if i["depth_avg"] > 0]
    d_d = float(np.mean(dd)) if dd else 5.5
    b_bounce_train = 368 + 16 * d_s + 24 * d_d        # path state 160 + hit record 32 + triangle 96 + vertex record write 80 + tree descents
    b_bounce_final = 288 + 16 * d_s + 24 * d_d
    bounce_ms = sum(s["kernel_ms"]["bounce"] for s in stats); bounce_n = sum(s["kernel_count"]["bounce"] for s in stats)
    my_train = sum(i["vertices"] for i in stats[0]["iterations"] if not i["is_final"]) * len(stats)   # rank 0's share when sharded
    my_final = sum(i["vertices"] for i in stats[0]["iterations"] if i["is_final"]) * len(stats)
    alg_bytes = my_train * b_bounce_train + my_final * b_bounce_final
    achieved = alg_bytes / (bounce_ms * 1e-3) / 1e9 if bounce_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("bounce_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "bounce_kernel (ray generation + BVH intersect + shade + S/D-tree guide + compaction; one launch per path depth)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_vertex": {"training": b_bounce_train, "final": b_bounce_final, "d_S": d_s, "d_D": d_d},
                "launches": int(bounce_n), "avg_launch_ms": bounce_ms / max(1, bounce_n),
                "kernel_ms_share": {k: sum(s["kernel_ms"][k] for s in stats) / max(1e-9, dev_ms) for k in capi.KERNEL_CLASSES},
                "pipeline_frac": (train_v * (468 + 16 * d_s + 48 * d_d) + final_v * (296 + 16 * d_s + 24 * d_d)) / (dev_ms * 1e-3) / 1e9 / peak / max(1, world)}

    # ---- CPU baseline on the box's host cores: bounded sample of the same workload
    cores = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_run("ref", args.cpu_size, args.cpu_budget, cores)
        cpu = {"value": r["msamples"], "unit": "Msamples/s", "cores": cores, "kind": "reference" if r["kind"] == "ref" else "port",
               "sample": f"CBOX {args.cpu_size}x{args.cpu_size}, default params, budget {args.cpu_budget} spp: {r['paths']} paths / {r['vertices']} vertices in {r['seconds']:.1f} s "
                         f"(CPU tracer restated from guided_path.cpp; SD-tree {'compiled verbatim from the reference' if r['kind'] == 'ref' else 'restated'})"}
    line = {
        "metric": "Msamples/sec (paths x bounces)", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms / args.steps, "kernel_ms_per_step": sum(sum(s["kernel_ms"].values()) for s in stats) / args.steps, "mpaths_per_s": paths / (dev_ms * 1e-3) / 1e6,
        "final_variance": stats[-1]["final_variance"], "iterations": stats[-1]["n_iterations"], "total_passes": stats[-1]["total_passes"],
    }
    if args.verbose:
        for it in stats[-1]["iterations"]:
            print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in it.items() if k in ("iteration", "passes", "seconds", "reset_seconds", "build_seconds", "variance", "s_tree_leaves", "vertices", "nodes_avg", "depth_avg", "s_tree_depth_avg")}, file=sys.stderr)
        print({"render_device_ms": stats[-1]["render_device_ms"], "render_seconds": stats[-1]["render_seconds"], "kernel_ms": stats[-1]["kernel_ms"]}, file=sys.stderr)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--budget", type=int, default=252, help="spp budget of one step")
    ap.add_argument("--budget-seconds", type=float, default=0.0, help="run the literal time-budget configuration instead")
    ap.add_argument("--cpu-size", type=int, default=512)
    ap.add_argument("--cpu-budget", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
