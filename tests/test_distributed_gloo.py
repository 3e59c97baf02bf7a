"""world_size-2 CPU (gloo) test of the multi-GPU plumbing: contiguous sharding, gather of reconstructions, fp64 metric
all-reduce.  The data path itself needs no collective, so this is all the N > 1 logic there is."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clip_neural_image_conpression_b200 import parallel


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_total: int, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = parallel.init_from_env(backend="gloo")
    lo, hi = parallel.shard_bounds(n_total, r, w)
    # "reconstruction" of image i is a tensor filled with i; "psnr" of image i is float(i)
    local = torch.stack([torch.full((3, 4, 4), float(i)) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 3, 4, 4)
    full = parallel.gather_shards(local, n_total)
    sums = parallel.reduce_sums([float(sum(range(lo, hi))), hi - lo], "cpu")
    mx = parallel.max_over_ranks(float(rank + 1), "cpu")
    parallel.barrier()
    q.put((rank, full[:, 0, 0, 0].tolist(), sums.tolist(), mx))
    dist.destroy_process_group()


def test_two_rank_shard_gather_reduce():
    ctx = mp.get_context("spawn")
    for n_total in (5, 8):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = [q.get(timeout=120) for _ in range(2)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for rank, ids, sums, mx in got:
            assert ids == [float(i) for i in range(n_total)]                  # every rank holds all images, in order
            assert sums == [float(sum(range(n_total))), float(n_total)]       # metric sum / count over all ranks
            assert mx == 2.0
