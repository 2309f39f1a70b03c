#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in-process.

Only runs in the build container (the reference is not present on the GPU box); its outputs
are committed under tests/golden/.  One reference run per process (the reference keeps
process-global counters: Message.uniq, Order.order_id).

    python tests/golden/gen_golden.py --name zi100_s123456789 --spec -- -c sparse_zi_100 -s 123456789

What is recorded, per run:
  * the lowered instance spec (the exact inputs the reference's Kernel.runner received, incl.
    every RandomState) -- `<name>.spec.npz` when --spec is given, else only its sha256;
  * the dispatch trace: one record per agent.wakeup()/receiveMessage() call with the message
    payload (see include/abides_b200.h: abides_trace_rec) -- full when --trace, plus its sha256;
  * ttl_messages (Kernel.py:204), per-agent final CASH / holdings / marked-to-market /
    FINAL_VALUATION, "Mean ending value by agent type", the oracle's f_log.
Compat shims: pandas.io.json.json_normalize alias; Timedelta.delta (= .value, removed in pandas 2); post-run order-book archival no-op.
"""
import argparse
import hashlib
import io
import json
import os
import re
import sys
import tempfile
from contextlib import redirect_stdout

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", required=True)
    ap.add_argument("--spec", action="store_true", help="store the full lowered spec")
    ap.add_argument("--trace", action="store_true", help="store the full dispatch trace")
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--config-path", default=None,
                    help="run this config FILE (written against the reference API) instead of config.<name>")
    ap.add_argument("ref_args", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    ref_args = [a for a in args.ref_args if a != "--"]

    sys.dont_write_bytecode = True
    sys.path.insert(0, REPO)
    sys.path.insert(0, REF)
    import numpy as np
    import pandas as pd
    import pandas.io.json
    pandas.io.json.json_normalize = pd.json_normalize
    pd.Timedelta.delta = property(lambda self: self.value)     # removed in pandas 2 (ExchangeAgent.py:309)

    from abides_b200 import lowering, _abi

    import Kernel as RK
    import agent.ExchangeAgent as EA
    EA.ExchangeAgent.logOrderBookSnapshots = lambda self, symbol: None

    trace = []
    captured = {}

    def payload(code_name, body, to_exchange):
        o = body.get("order")
        if to_exchange:
            p = [body.get("sender", -1), 0, 0, 0]
            if o is not None:
                p[1] = o.order_id
                p[2] = getattr(o, "limit_price", 0)
                p[3] = o.quantity if o.is_buy_order else -o.quantity
                if code_name == "CANCEL_ORDER":
                    # the message aliases the sender's live order object, whose quantity the sender keeps
                    # mutating on partial fills; only the side is meaningful to the book (OrderBook.py:300-348)
                    p[3] = 1 if o.is_buy_order else -1
            elif code_name == "QUERY_SPREAD":
                p[1] = body["depth"]
            elif code_name == "MARKET_DATA_SUBSCRIPTION_REQUEST":
                p[1] = body["levels"]
            return p
        if code_name in ("ORDER_ACCEPTED", "ORDER_EXECUTED", "ORDER_CANCELLED"):
            return [o.order_id, o.limit_price, o.quantity if o.is_buy_order else -o.quantity,
                    -1 if o.fill_price is None else o.fill_price]
        if code_name == "ORDER_MODIFIED":
            o = body["new_order"]
            return [o.order_id, o.limit_price, o.quantity if o.is_buy_order else -o.quantity, -1]
        if code_name in ("QUERY_SPREAD", "MARKET_DATA"):
            b, a = body["bids"], body["asks"]
            return [b[0][0] if b else -1, b[0][1] if b else 0, a[0][0] if a else -1, a[0][1] if a else 0]
        if code_name in ("WHEN_MKT_OPEN", "WHEN_MKT_CLOSE"):
            return [body["data"].value, 0, 0, 0]
        if code_name == "QUERY_TRANSACTED_VOLUME":
            return [int(body["transacted_volume"]), 0, 0, 0]
        if code_name == "QUERY_LAST_TRADE":
            return [-1 if body["data"] is None else int(body["data"]), 0, 0, 0]
        return [0, 0, 0, 0]

    def wrap(agent, is_exchange):
        ow, orm = agent.wakeup, agent.receiveMessage

        def wakeup(currentTime):
            trace.append((currentTime.value, agent.id, 0, 0, 0, 0, 0))
            return ow(currentTime)

        def receiveMessage(currentTime, msg):
            name = msg.body["msg"]
            code = _abi.MSG_CODES[name]
            if not is_exchange and code < 11:
                code |= _abi.MSG_REPLY
            p = payload(name, msg.body, is_exchange)
            trace.append((currentTime.value, agent.id, code, int(p[0]), int(p[1]), int(p[2]), int(p[3])))
            return orm(currentTime, msg)

        agent.wakeup, agent.receiveMessage = wakeup, receiveMessage

    orig_runner = RK.Kernel.runner

    def runner(self, agents=[], startTime=None, stopTime=None, num_simulations=1,
               defaultComputationDelay=1, defaultLatency=1, agentLatency=None, latencyNoise=[1.0],
               agentLatencyModel=None, skip_log=False, seed=None, oracle=None, log_dir=None):
        captured["spec"] = lowering.lower_instance(
            agents, startTime, stopTime, defaultComputationDelay=defaultComputationDelay,
            defaultLatency=defaultLatency, agentLatency=agentLatency, latencyNoise=latencyNoise,
            agentLatencyModel=agentLatencyModel, oracle=oracle, kernel_random_state=self.random_state,
            global_state=np.random.get_state())
        captured["agents"], captured["oracle"], captured["kernel"] = agents, oracle, self
        for a in agents:
            wrap(a, a.id == 0)
        return orig_runner(self, agents=agents, startTime=startTime, stopTime=stopTime,
                           num_simulations=num_simulations, defaultComputationDelay=defaultComputationDelay,
                           defaultLatency=defaultLatency, agentLatency=agentLatency,
                           latencyNoise=latencyNoise, agentLatencyModel=agentLatencyModel,
                           skip_log=True, seed=seed, oracle=oracle, log_dir=log_dir)

    RK.Kernel.runner = runner

    cfg = ref_args[ref_args.index("-c") + 1]
    sys.argv = ["abides.py"] + ref_args
    work = tempfile.mkdtemp(prefix="abides_golden_")
    os.chdir(work)
    buf = io.StringIO()
    import importlib
    with redirect_stdout(buf):
        if args.config_path:
            import runpy
            runpy.run_path(os.path.join(REPO, args.config_path) if not os.path.isabs(args.config_path)
                           else args.config_path, run_name="__abides_config__")
        else:
            importlib.import_module("config." + cfg)
    out = buf.getvalue()
    m = re.search(r"messages: (\d+), messages per second", out)
    ttl = int(m.group(1))
    means = {}
    sec = out.split("Mean ending value by agent type:")[1].split("Simulation ending!")[0]
    for line in sec.strip().splitlines():
        k, v = line.rsplit(":", 1)
        means[k.strip()] = int(v)

    agents, oracle, kernel, spec = captured["agents"], captured["oracle"], captured["kernel"], captured["spec"]
    n = len(agents)
    cash = np.zeros(n, np.int64)
    hold = np.zeros(n, np.int64)
    mtm = np.zeros(n, np.int64)
    fval = np.full(n, np.nan)
    sym = next(iter(agents[0].order_books))
    for a in agents[1:]:
        cash[a.id] = a.holdings["CASH"]
        hold[a.id] = a.holdings.get(sym, 0)
        for ev in a.log:
            if ev["EventType"] == "MARKED_TO_MARKET":
                mtm[a.id] = ev["Event"]
            if ev["EventType"] == "FINAL_VALUATION":
                fval[a.id] = float(ev["Event"])
            if ev["EventType"] == "ARRIVAL_MID":        # TWAP / VWAP agents: abides_agent_result.final_valuation
                fval[a.id] = float(ev["Event"])         # carries their arrival mid price (include/abides_b200.h)
    f_log = oracle.f_log[sym] if oracle is not None else []       # config/marketreplay.py runs without an oracle
    f_t = np.array([x["FundamentalTime"].value for x in f_log], np.int64)
    f_v = np.array([x["FundamentalValue"] for x in f_log], np.float64)
    tr = np.array(trace, dtype=np.int64).reshape(-1, 7)
    tr_sha = hashlib.sha256(tr.tobytes()).hexdigest()
    spec_sha = spec_digest(spec)
    book = agents[0].order_books[sym]
    meta = dict(config=cfg, ref_args=ref_args, ttl_messages=ttl, delivered=int(len(tr)), means=means,
                trace_sha256=tr_sha, spec_sha256=spec_sha, n_agents=n,
                last_trade=-1 if book.last_trade is None else int(book.last_trade), final_time=int(kernel.currentTime.value),
                slowest_agent_finish_time=int(max(kernel.agentCurrentTimes).value),
                numpy=np.__version__, pandas=pd.__version__)
    os.makedirs(args.out, exist_ok=True)
    base = os.path.join(args.out, args.name)
    save = dict(meta=json.dumps(meta), cash=cash, holdings=hold, mtm=mtm, final_valuation=fval,
                f_time=f_t, f_value=f_v)
    if args.trace:
        save["trace"] = tr
    np.savez_compressed(base + ".golden.npz", **save)
    if args.spec:
        lowering.save_spec(base + ".spec.npz", spec)
    print(json.dumps(meta, indent=1))


def spec_digest(spec):
    import numpy as np
    h = hashlib.sha256()
    h.update(json.dumps(spec.cfg, sort_keys=True).encode())
    # columns added after the first fixtures were frozen (exec_*: POVExecutionAgent trade=True; mm_*, hbl_L: MarketMaker
    # / HBL agents) only enter the digest of the rows that use them, so the older fixtures keep their spec_sha256
    late = ("exec_", "mm_", "hbl_", "sub")
    types = [{k: v for k, v in row.items()
              if not (k.startswith(late) and not (row.get("exec_trade") if k.startswith("exec_") else v))}
             for row in spec.types]
    h.update(json.dumps(types, sort_keys=True).encode())
    for a in (spec.agent_type, spec.lat_to, spec.lat_from, spec.size, spec.wakeup_time, spec.agent_theta,
              spec.theta, spec.mt_key, spec.mt_pos, spec.mt_has_gauss, spec.mt_gauss):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    main()
