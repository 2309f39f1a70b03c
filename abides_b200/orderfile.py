"""Order files as simulation input (reference: agent/examples/MarketReplayAgent.py:12-125, config/marketreplay.py).

The reference's MarketReplayAgent wakes at every timestamp of a historical L3 order file and turns each row into the
message its `placeOrder` (:47-62) would send: an unseen order id with a positive size is a LIMIT order, a known id with
size 0 a CANCEL of the resting order, a known id with another size a MODIFY.  `parse_orders_file` does that translation
(its own parser: the reference's `L3OrdersProcessor` does not parse the sample file it ships under current pandas) and
`replay` runs the resulting tape through the device order book (abides_replay_book): every notification the exchange
would send and the top of book after every message."""
import csv
import datetime as dt

import numpy as np

from . import _abi

LIMIT, CANCEL, MODIFY = _abi.MSG_CODES["LIMIT_ORDER"], _abi.MSG_CODES["CANCEL_ORDER"], _abi.MSG_CODES["MODIFY_ORDER"]
_EPOCH = dt.datetime(1970, 1, 1)


def _timestamp_ns(text):
    """'20190603093000.028558000' -> ns since the epoch (the reference trims the digits strptime cannot take)."""
    whole, _, frac = text.strip().partition(".")
    t = dt.datetime.strptime(whole, "%Y%m%d%H%M%S")
    ns = int((frac + "000000000")[:9]) if frac else 0
    return int((t - _EPOCH).total_seconds()) * 10 ** 9 + ns


def read_rows(path, price_scale=1, start_ns=None, end_ns=None):
    """Rows TIMESTAMP, ORDER_ID, PRICE, SIZE, BUY_SELL_FLAG (',' or '|' separated; flag BUY / SELL or 0 / 1) of an order
    file inside [start_ns, end_ns) -> list of (time_ns, order_id, price, size, is_buy)."""
    with open(path, newline="") as f:
        head = f.readline()
        sep = "|" if "|" in head else ","
        cols = [c.strip() for c in head.strip().split(sep)]
        rows = list(csv.reader(f, delimiter=sep))
    ix = {c: cols.index(c) for c in ("TIMESTAMP", "ORDER_ID", "PRICE", "SIZE", "BUY_SELL_FLAG")}
    out = []
    for r in rows:
        if len(r) < len(cols):
            continue
        t = _timestamp_ns(r[ix["TIMESTAMP"]])
        if (start_ns is not None and t < start_ns) or (end_ns is not None and t >= end_ns):
            continue
        flag = r[ix["BUY_SELL_FLAG"]].strip().upper()
        # PRICE: float * scale, truncated (MarketReplayAgent.py:113-114: .astype(float) * 100, .astype(int))
        out.append((t, int(r[ix["ORDER_ID"]]), int(float(r[ix["PRICE"]]) * price_scale), int(r[ix["SIZE"]]),
                    1 if flag in ("BUY", "0") else 0))
    return out


def parse_orders_file(path, agent=1, price_scale=1, start_ns=None, end_ns=None):
    """An order file -> int64[n, 8] tape rows (kind, agent, order_id, price, qty, is_buy, new_qty, time_ns) for
    engine.replay_book: the message MarketReplayAgent.placeOrder (:47-62) sends for every row."""
    tape, live = [], {}
    for t, oid, price, size, buy in read_rows(path, price_scale, start_ns, end_ns):
        ex = live.get(oid)
        if ex is None and size > 0:
            tape.append((LIMIT, agent, oid, price, size, buy, 0, t))
            live[oid] = (price, size, buy)
        elif ex is not None and size == 0:
            tape.append((CANCEL, agent, oid, ex[0], ex[1], ex[2], 0, t))
        elif ex is not None:
            tape.append((MODIFY, agent, oid, ex[0], ex[1], ex[2], size, t))
    return np.array(tape, np.int64).reshape(-1, 8)


def replay(engine, path, stream_history=25000, max_events_per_op=64, **parse_kw):
    """Replay an order file through the device book.  Returns (tape, events int64[m, 8], counts, states int64[n, 7])."""
    from .engine import replay_book
    tape = parse_orders_file(path, **parse_kw)
    events, counts, states = replay_book(engine, np.array([0, len(tape)], np.int64), tape, stream_history, max_events_per_op)
    return tape, events, counts, states
