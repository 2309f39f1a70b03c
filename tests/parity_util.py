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


def device_run(hb, device=0):
    from abides_b200.engine import Engine
    eng = Engine(device)
    eng.load(hb)
    eng.run()
    res, ares = eng.results_arrays()
    return eng, res.copy(), ares.copy()


def agent_slices(hb, inst):
    ic = hb.instances[inst]
    return slice(ic.agent_offset, ic.agent_offset + ic.n_agents)
