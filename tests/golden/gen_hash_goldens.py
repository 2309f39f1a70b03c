#!/usr/bin/env python
"""Hash-only goldens: many more seeds of the BASELINE configs through the UNMODIFIED reference, one process per run
(gen_golden.py), keeping per seed only ttl_messages, deliveries, last trade, final times and a SHA-256 of every
agent's final cash / holdings / marked-to-market -- so real numpy / glibc sits in the loop on far more runs than the
full fixtures cover.  Build container only (the reference is not present on the GPU box):

    python tests/golden/gen_hash_goldens.py --jobs 6        # writes tests/golden/sweep_hashes.json
"""
import argparse
import concurrent.futures as cf
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SWEEPS = {
    "sparse_zi_100": dict(config="sparse_zi_100", ref_args=["-c", "sparse_zi_100"], our_args=[], n=64),
    "sparse_zi_1000": dict(config="sparse_zi_1000", ref_args=["-c", "sparse_zi_1000"], our_args=[], n=64),
    "rmsc03_1000": dict(config="rmsc03", ref_args=["-c", "rmsc03", "-t", "ABM", "-d", "20200603", "--end-time", "10:00:00"],
                        our_args=["-t", "ABM", "-d", "20200603", "--end-time", "10:00:00"], n=64),
}


def seeds_for(name, n):
    """config/parallel.py:10-11 seeds, one stream per sweep."""
    rs = np.random.RandomState(int(hashlib.sha256(name.encode()).hexdigest()[:8], 16))
    return [int(s) for s in rs.randint(low=0, high=2 ** 32, size=n)]


def finals_digest(cash, holdings, mtm):
    h = hashlib.sha256()
    for a in (cash, holdings, mtm):
        h.update(np.ascontiguousarray(a, np.int64).tobytes())
    return h.hexdigest()


def one(job):
    name, seed, ref_args = job
    out = tempfile.mkdtemp(prefix="abides_hash_")
    tag = "%s_%d" % (name, seed)
    run = subprocess.run([sys.executable, os.path.join(HERE, "gen_golden.py"), "--name", tag, "--out", out, "--"] + ref_args +
                         ["-s", str(seed)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if run.returncode != 0:
        # the reference itself raised (e.g. AdaptiveMarketMakerAgent.receiveMessage before the book has ever been
        # two-sided: UnboundLocalError, AdaptiveMarketMakerAgent.py:141-176): the engines must stop such an instance
        # with ABIDES_ESTATE
        last = [ln for ln in run.stderr.strip().splitlines() if ln.strip()][-1]
        return name, seed, dict(reference_raises=last.split(":")[0].strip())
    g = np.load(os.path.join(out, tag + ".golden.npz"))
    meta = json.loads(str(g["meta"]))
    rec = dict(ttl_messages=meta["ttl_messages"], delivered=meta["delivered"], last_trade=meta["last_trade"],
               final_time=meta["final_time"], slowest_agent_finish_time=meta["slowest_agent_finish_time"],
               finals_sha256=finals_digest(g["cash"], g["holdings"], g["mtm"]))
    os.remove(os.path.join(out, tag + ".golden.npz"))
    return name, seed, rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=os.path.join(HERE, "sweep_hashes.json"))
    args = ap.parse_args()
    result = json.load(open(args.out)) if os.path.exists(args.out) else {}
    jobs = []
    for name, sw in SWEEPS.items():
        if args.only and name != args.only:
            continue
        result.setdefault(name, dict(config=sw["config"], args=sw["our_args"], seeds={}))
        for seed in seeds_for(name, sw["n"]):
            if str(seed) not in result[name]["seeds"]:
                jobs.append((name, seed, sw["ref_args"]))
    with cf.ThreadPoolExecutor(args.jobs) as ex:
        for k, (name, seed, rec) in enumerate(ex.map(one, jobs)):
            result[name]["seeds"][str(seed)] = rec
            if k % 8 == 7:
                json.dump(result, open(args.out, "w"), indent=0, sort_keys=True)
    json.dump(result, open(args.out, "w"), indent=0, sort_keys=True)
    print({k: len(v["seeds"]) for k, v in result.items()})


if __name__ == "__main__":
    main()
