"""Shared helpers of the parity tests: load golden fixtures, run oracle / device, diff traces."""
import json
import os

import numpy as np

from abides_b200 import lowering, _abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACE_FUNDAMENTAL = 64


def load_golden(name, directory=GOLDEN):
    g = np.load(os.path.join(directory, name + ".golden.npz"))
    meta = json.loads(str(g["meta"]))
    return g, meta


def load_spec(name, directory=GOLDEN):
    return lowering.load_spec(os.path.join(directory, name + ".spec.npz"))


def first_trace_mismatch(a, b):
    """Index of the first differing record of two int64[n,7] traces, or -1."""
    n = min(len(a), len(b))
    bad = np.nonzero((a[:n] != b[:n]).any(axis=1))[0]
    if len(bad):
        return int(bad[0])
    return -1 if len(a) == len(b) else n


def describe_mismatch(a, b, i, names=("got", "want")):
    lines = []
    for j in range(max(0, i - 3), min(max(len(a), len(b)), i + 3)):
        ra = a[j].tolist() if j < len(a) else None
        rb = b[j].tolist() if j < len(b) else None
        lines.append("%s %d %s=%s %s=%s" % ("->" if j == i else "  ", j, names[0], ra, names[1], rb))
    return "\n".join(lines)


def split_device_trace(tr):
    """Device traces interleave oracle steps (code 64); return (dispatch records, fundamental records)."""
    f = tr[:, 2] == TRACE_FUNDAMENTAL
    return tr[~f], tr[f]


def device_run(hb, device=0, engine=None):
    """Load + run a HostBatch through the C ABI; engine: None (the loader's choice), "warp" or "lane"."""
    from abides_b200.engine import Engine
    eng = Engine(device)
    if engine is not None:
        eng.set_engine(engine)
    eng.load(hb)
    eng.run()
    res, ares = eng.results_arrays()
    return eng, res.copy(), ares.copy()


def agent_slices(hb, inst):
    ic = hb.instances[inst]
    return slice(ic.agent_offset, ic.agent_offset + ic.n_agents)


def random_book_rows(seed, n_books=512, n_ops=300):
    """Shallow random books (limit orders only) for the replay harnesses: rows int64[n, 8], offsets int64[n_books + 1]."""
    rng = np.random.default_rng(seed)
    rows, off = [], [0]
    for b in range(n_books):
        t = 0
        for i in range(n_ops):
            t += 1000
            buy = int(rng.integers(0, 2))
            price = 100000 + (-1 if buy else 1) * int(rng.integers(-30, 200))
            rows.append((7, int(rng.integers(1, 50)), i + 1, price, int(rng.integers(1, 100)), buy, 0, t))
        off.append(len(rows))
    return np.array(rows, np.int64), np.array(off, np.int64)


def fuzz_book_rows(seed, n_books=48):
    """Deep, wide books with every message type (limit / cancel of live, dead and repeated ids / modify / market)."""
    rng = np.random.default_rng(seed)
    rows, off = [], [0]
    for b in range(n_books):
        t, next_id, placed = 0, 1, []
        width = int(rng.choice([5, 40, 400, 3000]))
        n_ops = int(rng.choice([400, 1500, 4000]))
        for i in range(n_ops):
            t += int(rng.integers(1, 1000))
            u = rng.random()
            agent = int(rng.integers(1, 60))
            if u < 0.62 or not placed:                      # limit order; mostly passive, sometimes crossing
                buy = int(rng.integers(0, 2))
                off_px = int(rng.integers(-width // 10 - 1, width))
                price = 100000 - off_px if buy else 100000 + off_px
                qty = int(rng.integers(1, 400))
                rows.append((7, agent, next_id, price, qty, buy, 0, t))
                placed.append((next_id, price, buy, agent))
                next_id += 1
            elif u < 0.88:                                  # cancel (some already gone, some twice)
                oid, price, buy, ag = placed[int(rng.integers(0, len(placed)))]
                rows.append((9, ag, oid, price, 1, buy, 0, t))
            elif u < 0.93:                                  # modify
                oid, price, buy, ag = placed[int(rng.integers(0, len(placed)))]
                rows.append((10, ag, oid, price, 1, buy, int(rng.integers(1, 500)), t))
            else:                                           # market order
                rows.append((8, agent, next_id, 0, int(rng.integers(1, 3000)), int(rng.integers(0, 2)), 0, t))
                next_id += 1
        off.append(len(rows))
    return np.array(rows, np.int64), np.array(off, np.int64)
