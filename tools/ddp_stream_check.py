"""Two-rank check of the streaming job (run under torchrun, NCCL): the ranks deal the frame batches of ONE
trajectory between them and must select exactly the rows a single rank selects; then the same with the
mini-batch slices + packed all-reduce exchange, which has to give a complete selection too.
Driven by tests/test_streaming_gpu.py::test_two_rank_job_equals_single_rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mddatasetbuilder_b200 import synth  # noqa: E402
from mddatasetbuilder_b200.engine import Engine  # noqa: E402
from mddatasetbuilder_b200.streaming import ResidentSource, StreamingSelector  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    s = synth.make_system(4000, 0.05, seed=synth.BASE_SEED + 77)
    F, K = 9, 50
    eng = Engine(local, s.atomname)
    pos = torch.as_tensor(synth.frame_positions(s, F), device=dev)
    rowptr, col, bo = synth.bond_table_device(s, F, dev)

    def job(r, w, **kw):
        sel = StreamingSelector(eng, s.types, n_clusters=K, n_each=2, seed=3, periodic=True, pool_rows=400,
                                want_molecules=True, **kw)
        return sel.run(ResidentSource(pos, s.cell(), 2, rowptr=rowptr, col=col, bo=bo, rank=r, world=w))

    single = job(0, 1)                      # before the process group exists: a plain one-rank job
    dist.init_process_group("nccl", device_id=dev)
    multi = job(rank, world)
    ok = single.keys == multi.keys and single.counts == multi.counts and single.chosen == multi.chosen
    # one rank without any frames (fewer batches than ranks): 1 batch of 9 frames
    sel = StreamingSelector(eng, s.types, n_clusters=K, n_each=2, seed=3, periodic=True, pool_rows=400)
    lone = sel.run(ResidentSource(pos, s.cell(), 9, rowptr=rowptr, col=col, bo=bo, rank=rank, world=world))
    sel1 = job(rank, world, kmeans_kwargs={"exchange": "allreduce"})
    full = all(len(v) == len(single.chosen[k]) for k, v in sel1.chosen.items())
    t = torch.tensor([int(ok), int(lone.chosen == single.chosen), int(full)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("two-rank == one-rank:", bool(t[0]), "| idle rank:", bool(t[1]), "| allreduce exchange complete:", bool(t[2]),
              "| fit comm", sel1.stats.get("fit_comm"))
        if int(t.min()) == 1:
            print("STREAM-DDP-OK")
    dist.destroy_process_group()
    return 0 if int(t.min()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
