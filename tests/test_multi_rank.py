"""N > 1 host logic on CPU: world_size-2 gloo.  Instances shard by contiguous ranges (no data-path collective);
the only collective is the final gather of per-instance summaries.  The per-rank compute is stood in
for by the CPU oracle (test infrastructure) -- what is under test is sharding, seed assignment and the gather."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_every_instance_once():
    from abides_b200 import runtime
    for n in (0, 1, 5, 8, 1024, 4099):
        for world in (1, 2, 3, 8):
            spans = [runtime.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import oracle
    from abides_b200 import runtime, lowering
    from abides_b200.engine import INSTANCE_RESULT_DTYPE
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = runtime.parallel_seeds(3, n_total)
    lo, hi = runtime.shard_bounds(n_total, world, rank)
    specs = [s for s, _ in runtime.build_instances("sparse_zi_100", [["-s", s] for s in seeds[lo:hi]])]
    hb = lowering.HostBatch(specs)
    res, _ = oracle.run_batch(hb, n_threads=2)
    local = np.frombuffer(res, dtype=INSTANCE_RESULT_DTYPE)
    allres = runtime.gather_summaries(local, dist)
    if rank == 0:
        q.put(allres["ttl_messages"].tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_process():
    import torch.multiprocessing as mp
    import oracle
    from abides_b200 import runtime, lowering
    n_total, world, port = 5, 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seeds = runtime.parallel_seeds(3, n_total)
    specs = [s for s, _ in runtime.build_instances("sparse_zi_100", [["-s", s] for s in seeds])]
    res, _ = oracle.run_batch(lowering.HostBatch(specs), n_threads=2)
    want = [res[i].ttl_messages for i in range(n_total)]
    assert got == want and len(set(want)) > 1
