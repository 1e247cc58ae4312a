"""N > 1 path of bench.py on CPU: two gloo ranks each take their slice of the seeded stream; the slices
tile the single-process workload exactly (pairs are independent, so sharding is a pure partition) and the
whole-job reduction is max-of-times / sum-of-cells."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    b = bench.rank_batch("cfg3_edit_100_300", 500, rank)
    dev_ms, e2e_ms, cells = bench.reduce_over_ranks(10.0 + rank, 20.0 - rank, float(b.cells()), world, "cpu")
    digest = int(np.sum(b.residues.astype(np.uint64) * (np.arange(b.residues.size, dtype=np.uint64) % 65521)))
    q.put((rank, b.n_pairs, int(b.residues.size), digest, b.cells(), dev_ms, e2e_ms, cells))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    sys.path.insert(0, ROOT)
    import bench
    from biogarden_b200 import synth
    world, port = 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synth.make("cfg3_edit_100_300", n_pairs=1000)
    split = int(full.seq_off[1000])
    parts = [full.residues[:split], full.residues[split:]]
    for (rank, n_pairs, nres, digest, cells, dev_ms, e2e_ms, cells_tot) in got:
        assert n_pairs == 500 and nres == parts[rank].size
        assert digest == int(np.sum(parts[rank].astype(np.uint64) * (np.arange(nres, dtype=np.uint64) % 65521)))
        assert dev_ms == 11.0 and e2e_ms == 20.0            # max over ranks
        assert cells_tot == float(full.cells())             # sum over ranks = the whole workload
    assert bench.shard_range(10, 0, 3) == (0, 3) and bench.shard_range(10, 2, 3) == (6, 10)
