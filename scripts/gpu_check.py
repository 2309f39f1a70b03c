"""Bring-up diagnostics: run golden specs through the CUDA engine and the oracle, print first mismatches."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from abides_b200 import lowering
from abides_b200.engine import Engine
from parity_util import load_golden, load_spec, first_trace_mismatch, describe_mismatch, split_device_trace

name = sys.argv[1] if len(sys.argv) > 1 else "zi100_s123456789"
g, meta = load_golden(name)
spec = load_spec(name)
cap = meta["delivered"] + len(g["f_time"]) + 64
hb = lowering.HostBatch([spec], trace_capacity=cap)
eng = Engine(0)
t = time.time(); eng.load(hb); print("load s", time.time() - t)
t = time.time(); eng.run(); print("run s", time.time() - t, "kernel ms", eng.last_run_ms())
res, ares = eng.results_arrays()
print({k: int(res[k][0]) for k in res.dtype.names if k != "reserved"})
print("want ttl", meta["ttl_messages"], "delivered", meta["delivered"])
tr, n = eng.trace(0)
disp, fund = split_device_trace(tr)
print("trace n", n, "disp", len(disp), "fund", len(fund), "golden", len(g["trace"]))
i = first_trace_mismatch(disp, g["trace"])
print("first mismatch", i)
if i >= 0:
    print(describe_mismatch(disp, g["trace"], i, ("device", "reference")))
for k, gk in (("cash", "cash"), ("holdings", "holdings"), ("marked_to_market", "mtm")):
    print(k, "mismatches", int((ares[k] != g[gk]).sum()))
