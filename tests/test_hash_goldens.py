"""Many more seeds of the BASELINE configs against the unmodified reference (tests/golden/sweep_hashes.json, made by
tests/golden/gen_hash_goldens.py): per seed ttl_messages, deliveries, last trade, final times and a SHA-256 of every
agent's final cash / holdings / marked-to-market.  CPU: the oracle and the host build of the lane engine, on the
config-built instances AND on the vectorised sweep batch.  GPU: both device engines on the seeded sweep batch."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
import hostsim
from abides_b200 import _abi, lowering, runtime, sweep, engine as eng_mod
from parity_util import GOLDEN, device_run

HASHES = json.load(open(os.path.join(GOLDEN, "sweep_hashes.json")))
CPU_SEEDS = {"sparse_zi_100": 64, "sparse_zi_1000": 12, "rmsc03_1000": 12}     # the CPU suite has minutes, not hours


def finals_digest(cash, holdings, mtm):
    h = hashlib.sha256()
    for a in (cash, holdings, mtm):
        h.update(np.ascontiguousarray(a, np.int64).tobytes())
    return h.hexdigest()


def seeds_of(name, limit=None):
    """All hashed seeds, or the `limit` smallest plus every seed on which the reference itself raised."""
    s = sorted(HASHES[name]["seeds"], key=int)
    if limit:
        s = s[:limit] + [x for x in s[limit:] if "reference_raises" in HASHES[name]["seeds"][x]]
    return [int(x) for x in s]


def check(name, seeds, res, ares, n_agents):
    n_ok = 0
    for i, seed in enumerate(seeds):
        want = HASHES[name]["seeds"][str(seed)]
        if "reference_raises" in want:                       # the reference crashed on this seed: we stop it and say so
            assert res["status"][i] == _abi.ESTATE, (name, seed, want)
            continue
        assert res["status"][i] == 0, (name, seed)
        for f in ("ttl_messages", "delivered", "last_trade", "final_time", "slowest_agent_finish_time"):
            assert res[f][i] == want[f], (name, seed, f)
        sl = slice(i * n_agents, (i + 1) * n_agents)
        assert finals_digest(ares["cash"][sl], ares["holdings"][sl], ares["marked_to_market"][sl]) == want["finals_sha256"], (name, seed)
        n_ok += 1
    assert n_ok >= len(seeds) // 2


@pytest.mark.parametrize("name", sorted(HASHES))
def test_oracle_and_lane_logic_match_reference_hashes(name):
    sw = HASHES[name]
    seeds = seeds_of(name, CPU_SEEDS[name])
    specs = [s for s, _ in runtime.build_instances(sw["config"], [list(sw["args"]) + ["-s", str(s)] for s in seeds])]
    hb = lowering.HostBatch(specs)
    n = specs[0].n_agents
    ores, oares = oracle.run_batch(hb)
    check(name, seeds, np.frombuffer(ores, dtype=eng_mod.INSTANCE_RESULT_DTYPE), np.frombuffer(oares, dtype=eng_mod.AGENT_RESULT_DTYPE), n)
    res, ares, _, _, _ = hostsim.run(hb)
    check(name, seeds, res, ares, n)
    res, ares, _, _, _ = hostsim.run(sweep.build_sweep(sw["config"], seeds, sw["args"]))
    check(name, seeds, res, ares, n)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["lane", "warp"])
@pytest.mark.parametrize("name", sorted(HASHES))
def test_device_engines_match_reference_hashes(name, engine):
    """All hashed seeds per config, untraced (the lean kernel variants), through the seeded sweep batch."""
    sw = HASHES[name]
    seeds = seeds_of(name)
    sb = sweep.build_sweep(sw["config"], seeds, sw["args"])
    eng, res, ares = device_run(sb, engine=engine)
    check(name, seeds, res, ares, sb.groups[0].n_agents)
