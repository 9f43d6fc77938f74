"""Worker for the multi-rank tests (spawned by test_sharded_*.py): one rank of a tile-sharded render whose per-iteration
tree exchange goes through torch.distributed (gloo on CPU tensors).  mode 'oracle': CPU oracle; mode 'gpu': the CUDA
library on cuda:0 (all ranks share the device; the exchange buffer is staged through the host for gloo); mode 'nccl': one GPU
per rank, the library's own NCCL communicator (ppg_nccl_init), gloo only carries the unique id."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    mode, rank, world, port, size, budget, outdir = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), sys.argv[6], sys.argv[7]
    extra = dict(kv.split("=", 1) for kv in sys.argv[8:])
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import load_cbox
    sc = load_cbox(size)
    props = dict(sc.integrator, budget=budget, **extra)
    if mode == "oracle":
        import oracle_lib as O
        o = O.Oracle(O.params_from_xml(props), sc, nthreads=2, kind="port")
        o.set_shard(rank, world)

        def red(a):
            t = torch.from_numpy(a)
            dist.all_reduce(t)
        o.set_allreduce(red)
        img, st = o.render()
        e = o.export(0)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), img=img, weights=[i["weight_avg"] * i["s_tree_leaves"] for i in st["iterations"]],
                 leaves=[i["s_tree_leaves"] for i in st["iterations"]], paths=st["total_paths"], sums=e["sums"], children=e["children"], s_children=e["s_children"])
    else:
        from ppg_b200.integrator import GuidedPathTracer

        class _Ptr:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

        def red(ptr, n):
            t = torch.as_tensor(_Ptr(ptr, n), device="cuda:0")
            c = t.cpu()
            dist.all_reduce(c)
            t.copy_(c)
            torch.cuda.synchronize()
        if mode == "nccl":           # one GPU per rank; the library's own NCCL communicator (bootstrap of the unique id over gloo)
            g = GuidedPathTracer(props, device=rank)
            g.set_scene(sc); g.init_nccl()
        else:
            g = GuidedPathTracer(props, device=0)
            g.set_scene(sc); g.set_shard(rank, world); g.set_allreduce(red)
        img, st = g.render()
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), img=img, weights=[i["weight_avg"] * i["s_tree_leaves"] for i in st["iterations"]],
                 leaves=[i["s_tree_leaves"] for i in st["iterations"]], paths=st["total_paths"], variance=[i["variance"] for i in st["iterations"]])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
