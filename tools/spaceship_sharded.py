#!/usr/bin/env python
"""BASELINE config 4 on the GPU box: spaceship-improved.xml's own settings (inversevar / stochastic / box / kl, sTreeThreshold 4000,
sppPerPass 1) at 1920x1080, image blocks sharded over the ranks of one box with the per-iteration tree allreduce (+ Adam replica
averaging).  Run under torchrun (or alone for 1 GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29620 tools/spaceship_sharded.py [W H spp]
Rank 0 prints one JSON line (Msamples/s over all ranks, max-over-ranks device time) and saves the image as .npy under gpurun_out/."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from common import load_fixture_scene
from ppg_b200.integrator import GuidedPathTracer, torch_allreduce

W, H, spp = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]) if len(sys.argv) > 3 else (1920, 1080, "255")
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = load_fixture_scene("spaceship-improved").with_film(W, H)
props = dict(sc.integrator, budget=spp)
g = GuidedPathTracer(props, device=local); g.set_scene(sc)
if world > 1:
    g.set_shard(rank, world); g.set_allreduce(torch_allreduce())
torch.cuda.synchronize()
if dist is not None: dist.barrier()
t0 = time.perf_counter()
img, st = g.render()
torch.cuda.synchronize()
if dist is not None: dist.barrier()
wall = time.perf_counter() - t0
dev_ms, verts = st["render_device_ms"], st["total_vertices"]
if dist is not None:
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_ms = float(t[0])
    c = torch.tensor([verts], device="cuda", dtype=torch.float64); dist.all_reduce(c, op=dist.ReduceOp.SUM); verts = int(c[0])
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", f"spaceship_{W}x{H}_{spp}spp_n{world}.npy"), img[::4, ::4].astype(np.float16))
    print(json.dumps({"workload": f"SPACESHIP {W}x{H}, spaceship-improved.xml settings (inversevar/stochastic/box/kl, sTreeThreshold=4000, sppPerPass=1), budget {spp} spp",
                      "n_gpus": world, "value": verts / dev_ms / 1e3, "unit": "Msamples/s", "device_ms": dev_ms, "wall_s": wall, "vertices": verts,
                      "iterations": [{"passes": i["passes"], "s_tree_leaves": i["s_tree_leaves"], "variance": i["variance"]} for i in st["iterations"]],
                      "kernel_ms_rank0": st["kernel_ms"], "image_mean": float(img.mean()), "finite": bool(np.isfinite(img).all())}), flush=True)
if dist is not None:
    dist.destroy_process_group()
