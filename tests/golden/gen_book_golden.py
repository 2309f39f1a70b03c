#!/usr/bin/env python
"""Golden vectors for the order-book replay harness: drive the reference's REAL util.OrderBook.OrderBook
(/root/reference) with a stub owner and record, per message, the notifications it sends and the inside
quotes / last trade / level counts afterwards.  Container-only; output: tests/golden/book_replay.golden.npz

Tapes: (0) data/sample_orders_file.csv interpreted as agent/examples/MarketReplayAgent.py:47-62 does
(a MODIFY keeps the resting order's price: the C ABI carries only the new quantity); (1..) random tapes
mixing limit orders, cancels (live and stale ids, wrong price), market orders and modifies.
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
import numpy as np
import pandas as pd
import pandas.io.json
pandas.io.json.json_normalize = pd.json_normalize

from util.OrderBook import OrderBook          # noqa: E402
from util.order.LimitOrder import LimitOrder  # noqa: E402
from util.order.MarketOrder import MarketOrder  # noqa: E402
from util.order import Order as OrderMod      # noqa: E402

LIMIT, MARKET, CANCEL, MODIFY = 7, 8, 9, 10
KIND = {"ORDER_ACCEPTED": 11, "ORDER_EXECUTED": 12, "ORDER_CANCELLED": 13, "ORDER_MODIFIED": 14}
STREAM_HISTORY = 5


class Owner:
    def __init__(self):
        self.currentTime = pd.Timestamp("2019-06-03 09:30:00")
        self.stream_history = STREAM_HISTORY
        self.book_freq = None
        self.sent = []

    def sendMessage(self, recipient, msg):
        o = msg.body.get("order", msg.body.get("new_order"))
        self.sent.append((KIND[msg.body["msg"]], recipient, int(o.is_buy_order), o.order_id, o.limit_price, o.quantity,
                          -1 if o.fill_price is None else o.fill_price))

    def logEvent(self, *a, **k):
        pass


def replay(tape):
    """tape rows: (kind, agent, order_id, price, qty, is_buy, new_qty, time_ns)"""
    own = Owner()
    book = OrderBook(own, "X")
    OrderMod.Order.order_id = 1 << 30
    OrderMod.Order._order_ids = set()
    events, counts, states = [], [], []
    for i, (kind, agent, oid, price, qty, is_buy, new_qty, t) in enumerate(tape):
        own.currentTime = pd.Timestamp(t)
        own.sent = []
        if kind == LIMIT:
            book.handleLimitOrder(LimitOrder(agent, own.currentTime, "X", qty, bool(is_buy), price, order_id=oid))
        elif kind == MARKET:
            book.handleMarketOrder(MarketOrder(agent, own.currentTime, "X", qty, bool(is_buy), order_id=oid))
        elif kind == CANCEL:
            book.cancelOrder(LimitOrder(agent, own.currentTime, "X", qty, bool(is_buy), price, order_id=oid))
        elif kind == MODIFY:
            book.modifyOrder(LimitOrder(agent, own.currentTime, "X", qty, bool(is_buy), price, order_id=oid),
                             LimitOrder(agent, own.currentTime, "X", new_qty, bool(is_buy), price, order_id=oid))
        for e in own.sent:
            events.append((i,) + e)
        counts.append(len(own.sent))
        b, a = book.getInsideBids(1), book.getInsideAsks(1)
        states.append((b[0][0] if b else -1, b[0][1] if b else 0, a[0][0] if a else -1, a[0][1] if a else 0,
                       -1 if book.last_trade is None else book.last_trade, len(book.bids), len(book.asks)))
    return events, counts, states


def sample_file_tape():
    df = pd.read_csv(os.path.join(REF, "data", "sample_orders_file.csv"))
    tape, live = [], {}
    for _, r in df.iterrows():
        ts = str(r["TIMESTAMP"])
        t = pd.Timestamp(pd.to_datetime("%.9f" % r["TIMESTAMP"], format="%Y%m%d%H%M%S.%f")).value
        oid, price, size, buy = int(r["ORDER_ID"]), int(r["PRICE"]), int(r["SIZE"]), int(r["BUY_SELL_FLAG"] == "BUY")
        ex = live.get(oid)
        if ex is None and size > 0:
            tape.append((LIMIT, 1, oid, price, size, buy, 0, t))
            live[oid] = (price, size, buy)
        elif ex is not None and size == 0:
            tape.append((CANCEL, 1, oid, ex[0], ex[1], ex[2], 0, t))
        elif ex is not None:
            tape.append((MODIFY, 1, oid, ex[0], ex[1], ex[2], size, t))
    return tape


def random_tape(rng, n, wide):
    tape, live, t, next_id = [], [], pd.Timestamp("2019-06-03 09:30:00").value, 1
    for _ in range(n):
        t += int(rng.integers(0, 3)) * 1000
        u = rng.random()
        if u < 0.62 or not live:
            buy = int(rng.integers(0, 2))
            mid = 100000
            off = int(rng.integers(-40, 400 if wide else 25))
            price = mid - off if buy else mid + off
            qty = int(rng.integers(1, 200))
            agent = int(rng.integers(1, 40))
            tape.append((LIMIT, agent, next_id, price, qty, buy, 0, t))
            live.append((next_id, agent, price, qty, buy))
            next_id += 1
        elif u < 0.82:
            k = int(rng.integers(0, len(live)))
            oid, agent, price, qty, buy = live[k]
            if rng.random() < 0.1:
                price += 1                      # wrong price level: silently ignored
            elif rng.random() < 0.9:
                live.pop(k)
            tape.append((CANCEL, agent, oid, price, qty, buy, 0, t))
        elif u < 0.92:
            k = int(rng.integers(0, len(live)))
            oid, agent, price, qty, buy = live[k]
            tape.append((MODIFY, agent, oid, price, qty, buy, int(rng.integers(1, 300)), t))
        else:
            tape.append((MARKET, int(rng.integers(1, 40)), (1 << 29) + len(tape), 0, int(rng.integers(1, 1500)), int(rng.integers(0, 2)), 0, t))
    return tape


def main():
    rng = np.random.default_rng(2026)
    tapes = [sample_file_tape()] + [random_tape(rng, n, wide) for n, wide in
                                   ((400, False), (1500, False), (1500, True), (3000, True), (60, False))] + [[]]
    ops, off, ev, cnt, st = [], [0], [], [], []
    for tp in tapes:
        e, c, s = replay(tp)
        base = len(ops)
        ops.extend(tp)
        ev.extend([(base + x[0],) + x[1:] for x in e])
        cnt.extend(c)
        st.extend(s)
        off.append(len(ops))
    out = os.path.join(REPO, "tests", "golden", "book_replay.golden.npz")
    np.savez_compressed(out, ops=np.array(ops, np.int64).reshape(-1, 8), op_offset=np.array(off, np.int64),
                        events=np.array(ev, np.int64).reshape(-1, 8), counts=np.array(cnt, np.int32),
                        states=np.array(st, np.int64).reshape(-1, 7), stream_history=STREAM_HISTORY)
    print("tapes", len(tapes), "ops", len(ops), "events", len(ev), "->", out)


if __name__ == "__main__":
    main()
