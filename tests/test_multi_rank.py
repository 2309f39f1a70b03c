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


class _HostEngine:
    """Stands in for abides_b200.engine.Engine in the CPU suite: same calls, computed by the host build of the lane
    engine (tests/hostsim).  What is under test is the sweep driver around it: chunking, the ticket counter, merging."""
    calls = []

    def __init__(self, device):
        self.device, self.hb, self.out = device, None, None

    def set_engine(self, name):
        pass

    def load(self, hb):
        self.hb = hb

    def run(self):
        import hostsim
        res, ares, _, _, _ = hostsim.run(self.hb)
        self.out = (res, ares)
        _HostEngine.calls.append((self.device, self.hb.n_instances))

    def results_arrays(self):
        return self.out

    def last_run_ms(self):
        return 1.0, 1


def test_multi_device_sweep_driver_returns_single_batch_results(monkeypatch):
    """runtime.run_sweep over two devices (chunks handed out by a ticket counter, one worker thread per device) returns,
    in seed order, exactly what one batch on one device returns."""
    import hostsim
    from abides_b200 import runtime, sweep
    engines = {}
    monkeypatch.setattr(runtime, "engine", lambda dev=0: engines.setdefault(dev, _HostEngine(dev)))
    _HostEngine.calls = []
    seeds = runtime.parallel_seeds(9, 11)
    out = runtime.run_sweep("sparse_zi_100", seeds, devices=(0, 1), chunk_instances=3)
    assert sum(n for _, n in _HostEngine.calls) == 11 and len(_HostEngine.calls) == 4
    assert [p[1:3] for p in out.placement] == [(0, 3), (3, 3), (6, 3), (9, 2)]
    res1, ares1, _, _, _ = hostsim.run(sweep.build_sweep("sparse_zi_100", seeds))
    assert np.array_equal(out.res["ttl_messages"], res1["ttl_messages"]) and len(set(out.res["ttl_messages"])) > 1
    assert np.array_equal(out.ares["cash"], ares1["cash"])
    assert np.array_equal(out.agents(7)["holdings"], ares1["holdings"][7 * 101:8 * 101])
    # the config-built path (any config) through the same driver
    specs = [s for s, _ in runtime.build_instances("sparse_zi_100", [["-s", s] for s in seeds[:5]])]
    out2 = runtime.run_specs(specs, devices=[0, 1], chunk_instances=2)
    assert np.array_equal(out2.res["ttl_messages"], res1["ttl_messages"][:5])
