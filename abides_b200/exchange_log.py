"""The reference's per-agent log archives, rebuilt from the device trace.

What the reference writes at kernelTerminating (agent/Agent.py:87-98 -> Kernel.writeLog, Kernel.py:487-510), one
bz2-pickled DataFrame each, index `EventTime`, columns `EventType`, `Event`:

* the exchange's log `<exchange name>.bz2` (EXCHANGE_AGENT.bz2 in the rmsc configs): every message it received before the
  close (order messages as `order.to_dict()` when `log_orders`, the others as the sender id; agent/ExchangeAgent.py:
  126-150), every ORDER_ACCEPTED / _EXECUTED / _CANCELLED it sent (:404-408) and, after each handleLimitOrder, BEST_BID /
  BEST_ASK / LAST_TRADE (util/OrderBook.py:108-132);
* `ORDERBOOK_<symbol>_FULL.bz2` when `book_freq == 0` and `wide_book`: the full depth after every handleLimitOrder
  (OrderBook.py:138-156, :492-520; ExchangeAgent.py:319-396), a sparse frame times x quotes, bids negative;
* one log per trading agent with `log_to_file` (agent/TradingAgent.py:106-168, :390-500, :575-603).

Everything here is derived from the device's trace of one instance: the dispatched events plus the exchange-log records
(`ABIDES_TRACE_EXCHANGE_LOG`, include/abides_b200.h).  The book depth is an ACCOUNT of the device's own notifications
(accepted - executed - cancelled per price), not a second matching engine; the device's top-of-book record cross-checks
it at every snapshot.

The BID_DEPTH / ASK_DEPTH rows of agents that ask for more than the best level (market makers: five levels, execution
agents: the whole book) come from the same account: the levels it shows when the exchange handles the QUERY_SPREAD are
what the reply carries; the reply's top of book, which the trace does carry, cross-checks it.

Agents with `log_orders=True` also log what they send: ORDER_SUBMITTED from the placement records, CANCEL_SUBMITTED at the
send time of each cancellation, which under the deterministic latency model is the exchange's receive time minus the
computation delay and the agent's latency (needs the lowered instance, `spec`).

Not reproduced: `book_freq` sampling other than 0 (every update), the long-format (`wide_book=False`) archive (the
reference itself fails there: pd.SparseDataFrame, ExchangeAgent.py:384), and the logs of agents with `log_orders=True`
that place orders under a random latency model or without `spec` -- those agents are skipped and reported."""
import os
from array import array

import numpy as np

from abides_b200 import _abi

M = _abi.MSG_CODES
ORDER_MSGS = (M["LIMIT_ORDER"], M["MARKET_ORDER"], M["CANCEL_ORDER"], M["MODIFY_ORDER"])
QUERY_MSGS = (M["QUERY_LAST_TRADE"], M["QUERY_SPREAD"], M["QUERY_ORDER_STREAM"], M["QUERY_TRANSACTED_VOLUME"])
SENT = _abi.TRACE_EXCH_SENT
NAT = np.iinfo(np.int64).min


def _iso(ns):
    import pandas as pd
    return pd.Timestamp(int(ns)).isoformat()


def _order_dict(agent, placed_ns, symbol, qty, is_buy, oid, fill, limit):
    """Order.to_dict() (util/order/Order.py:52-55) of a LimitOrder (limit is not None) or a MarketOrder."""
    d = {'agent_id': int(agent), 'time_placed': _iso(placed_ns), 'symbol': symbol, 'quantity': int(qty),
         'is_buy_order': bool(is_buy), 'order_id': int(oid), 'fill_price': None if fill is None or fill < 0 else int(fill),
         'tag': None}
    if limit is not None:
        d['limit_price'] = int(limit)
    return d


class ExchangeArchive:
    """Walks one instance's trace once; afterwards `exchange_rows`, `book_rows` and `agent_rows` hold the archives."""

    def __init__(self, kernel, trace, spec=None):
        self.kernel = kernel
        # agents with log_orders=True also log what they SEND; a cancellation's send time is not in the trace, but under
        # the deterministic latency model it is the exchange's receive time minus computation delay and latency -- which
        # needs the lowered instance (`spec`: lat_to, default_computation_delay)
        self.spec = spec if spec is not None and spec.cfg.get("latency_kind") == _abi.LAT_DETERMINISTIC else None
        self.exch = kernel.agents[0]
        self.symbol = next(iter(self.exch.order_books))
        self.mkt_close = _ns(self.exch.mkt_close)
        self.log_orders = bool(self.exch.log_orders)
        self.snapshots = self.exch.book_freq is not None
        self.exchange_rows = [(NAT, 'AGENT_TYPE', self.exch.type)]
        # book snapshots as compact columns: one time per snapshot, its non-zero cells as (snapshot, quote, signed volume)
        self.book_times = array('q')
        self.book_cells = (array('l'), array('q'), array('q'))
        self.quotes_seen = set()
        self.agent_rows = {}
        self.skipped_agents = []
        self._walk(np.asarray(trace, np.int64))

    # ---- agents -------------------------------------------------------------------------------------------------
    def _agent_logged(self, a):
        return bool(getattr(a, "log_to_file", False)) and a.id != 0

    def _walk(self, tr):
        sym = self.symbol
        logged = {a.id: a for a in self.kernel.agents if self._agent_logged(a)}
        own_orders = {i for i, a in logged.items() if getattr(a, "log_orders", False)}     # these log what they send, too
        # an agent's rows, one block per dispatched event: [time, trace index, rows logged on arrival, CANCEL_SUBMITTED
        # rows, ORDER_SUBMITTED rows] -- every supported strategy cancels before it places inside one handler
        blocks = {i: [[NAT, -1, [(NAT, 'AGENT_TYPE', a.type), (NAT, 'STARTING_CASH', a.starting_cash)], [], []]]
                  for i, a in logged.items()}
        copy_hist = {}         # id of an agent's own copy -> [(trace index, quantity shown from then on), ...]
        sent_tp = {}           # order id in a notification -> time_placed it carries
        hold = {i: {'CASH': a.starting_cash} for i, a in logged.items()}
        woke, last_time = set(), {}
        placed = {}            # order id in the agent's message -> (time_placed, agent, price, signed qty, is_market)
        copy_qty = {}          # id of the agent's own copy of an order -> [quantity the copy shows, message id]
        book_time = {}         # order id in the book -> time_placed
        modified_at = {}       # order id -> time_placed of the new order of its latest MODIFY_ORDER
        levels = {}            # price -> [side (+1 bid / -1 ask), [[order id, quantity], ...] oldest first]
        depth = {}             # price -> signed resting volume (bids negative, as the archive stores them)
        cur = None             # the order message the exchange is handling: [time_placed, message id, row to patch]
        n_exec = 0             # ORDER_EXECUTED notifications of the current exchange event: aggressor, resting, aggressor, ...
        spread_replies = {}    # agent -> the (bids, asks, traded) of its QUERY_SPREAD replies still in flight
        traded = False         # a trade has happened: the book's last_trade is an int (before: the oracle's opening price)
        open_price_last = {}   # agent -> its known last trade is still that opening price (a float in most configs)
        ex = self.exchange_rows
        for ti, (time, agent, code, p0, p1, p2, p3) in enumerate(tr.tolist()):
            if code == _abi.TRACE_FUNDAMENTAL:
                continue
            if code == _abi.TRACE_ORDER_PLACED:
                if p3 == 2:                                        # the new order of a modification: only its time matters
                    modified_at[p0] = time
                    continue
                placed[p0] = (time, agent, p1, p2, p3)
                # the agent keeps a deepcopy; the copy of order id 0 takes a fresh id (Order.py:29): the next one
                copy_qty[p0 if p0 != 0 else 1] = [abs(p2), p0]
                copy_hist[p0 if p0 != 0 else 1] = [(ti, abs(p2))]
                if agent in own_orders and agent in logged:
                    if self.spec is None:                      # CANCEL_SUBMITTED times would be guesses: no log at all
                        self.skipped_agents.append(logged.pop(agent).name)
                    else:                                      # ORDER_SUBMITTED: order.to_dict(), whose copy of id 0 is id 2
                        blocks[agent][-1][4].append((time, 'ORDER_SUBMITTED', _order_dict(
                            agent, time, sym, abs(p2), p2 > 0, p0 if p0 != 0 else 2, None, None if p3 == 1 else p1)))
                continue
            if code >= SENT:
                msg = code - SENT
                if cur is not None and cur[2] is not None:        # the logged copy of order id 0 took the id before the book's
                    r = ex[cur[2]]
                    r[2]['order_id'] = p0 - 1
                    cur[2] = None
                tp = book_time.get(p0, cur[0] if cur is not None else time)
                sent_tp[p0] = tp
                if self.log_orders and msg != M["ORDER_MODIFIED"]:   # ORDER_MODIFIED is sent but not logged (ExchangeAgent.py:404-411)
                    ex.append([time, _abi.MSG_NAMES[msg], _order_dict(agent, tp, sym, abs(p2), p2 > 0, p0, p3, p1)])
                # the depth account: price levels as the FIFO lists the book keeps (OrderBook.py:272-298)
                lv = levels.get(p1)
                if msg == M["ORDER_ACCEPTED"]:
                    book_time[p0] = tp
                    if lv is None:
                        lv = levels[p1] = [1 if p2 > 0 else -1, []]
                    lv[1].append([p0, abs(p2)])
                elif msg == M["ORDER_EXECUTED"]:
                    n_exec += 1
                    if n_exec % 2 == 0:                            # the resting side: always the oldest order of its level
                        if lv is None or lv[1][0][0] != p0:
                            raise RuntimeError("order-book account: execution of %d is not at the front of level %d" % (p0, p1))
                        lv[1][0][1] -= abs(p2)
                        if lv[1][0][1] <= 0:
                            lv[1].pop(0)
                elif msg == M["ORDER_CANCELLED"]:                  # the first order of that id at the level (:319-346)
                    k = next((i for i, e in enumerate(lv[1]) if e[0] == p0), None) if lv else None
                    if k is None:
                        raise RuntimeError("order-book account: cancellation of an unknown order %d" % p0)
                    lv[1].pop(k)
                elif msg == M["ORDER_MODIFIED"]:                   # book[i][0] = new_order (:359): the level's FIRST order
                    if lv is None or not lv[1]:
                        raise RuntimeError("order-book account: modification at an empty level %d" % p1)
                    lv[1][0] = [p0, abs(p2)]
                    book_time[p0] = modified_at.get(p0, time)
                if lv is not None:
                    v = sum(e[1] for e in lv[1])
                    if v:
                        depth[p1] = -lv[0] * v
                    else:
                        depth.pop(p1, None)
                        if not lv[1]:
                            del levels[p1]
                continue
            if code == _abi.TRACE_BOOK_TOP:
                if p0 >= 0:
                    ex.append([time, 'BEST_BID', "{},{},{}".format(sym, p0, p1)])
                if p2 >= 0:
                    ex.append([time, 'BEST_ASK', "{},{},{}".format(sym, p2, p3)])
                if self.snapshots:
                    bids = [q for q, v in depth.items() if v < 0]
                    asks = [q for q, v in depth.items() if v > 0]
                    best_bid, best_ask = (max(bids) if bids else -1), (min(asks) if asks else -1)
                    if (best_bid, best_ask) != (p0, p2) or (bids and -depth[best_bid] != p1) or (asks and depth[best_ask] != p3):
                        raise RuntimeError("order-book account disagrees with the device's top of book at %d" % time)
                    self.quotes_seen.update(depth)
                    row = len(self.book_times)
                    self.book_times.append(time)
                    for q, v in depth.items():
                        self.book_cells[0].append(row)
                        self.book_cells[1].append(q)
                        self.book_cells[2].append(v)
                continue
            if code == _abi.TRACE_LAST_TRADE:
                ex.append([time, 'LAST_TRADE', "{},${:0.4f}".format(p0, p1)])
                traded = True
                continue
            # ---- a dispatched event ----
            base = code & ~_abi.MSG_REPLY
            if agent == 0:
                cur = None
                n_exec = 0
                if code == 0:
                    continue
                if base == M["CANCEL_ORDER"] and p0 in own_orders and p0 in logged and p1 in copy_qty:
                    # CANCEL_SUBMITTED (TradingAgent.py:366-374): logged by the handler that sent this message, whenever
                    # the exchange gets it (also after the close)
                    sent = time - int(self.spec.cfg["default_computation_delay"]) - int(self.spec.lat_to[p0])
                    blk = next((b for b in reversed(blocks[p0]) if b[0] == sent), None)
                    if blk is None:
                        raise RuntimeError("no handler of agent %d at %d sent the CANCEL_ORDER received at %d" % (p0, sent, time))
                    q = [v for i, v in copy_hist[p1] if i <= blk[1]][-1]
                    blk[3].append((sent, 'CANCEL_SUBMITTED', _order_dict(p0, placed[copy_qty[p1][1]][0], sym, q, p3 > 0, p1, None, p2)))
                closed = time > self.mkt_close
                if closed and base not in QUERY_MSGS:
                    continue                                       # answered MKT_CLOSED before any logging (:126-144)
                if base in ORDER_MSGS:
                    if base == M["CANCEL_ORDER"]:
                        c = copy_qty.get(p1)
                        pl = placed.get(c[1]) if c else None
                        if self.log_orders:
                            if pl is None:
                                raise RuntimeError("CANCEL_ORDER for an order the trace never saw placed (id %d)" % p1)
                            ex.append([time, 'CANCEL_ORDER', _order_dict(p0, pl[0], sym, c[0], p3 > 0, p1, None, p2)])
                        continue
                    if base == M["MODIFY_ORDER"]:                  # logs the sender's own copy of the order (:148)
                        c = copy_qty.get(p1)
                        pl = placed.get(c[1]) if c else None
                        if self.log_orders:
                            if pl is None:
                                raise RuntimeError("MODIFY_ORDER for an order the trace never saw placed (id %d)" % p1)
                            ex.append([time, 'MODIFY_ORDER', _order_dict(p0, pl[0], sym, c[0], p3 > 0, p1, None, p2)])
                        cur = [pl[0] if pl else time, p1, None]
                        continue
                    pl = placed.get(p1)
                    tp = pl[0] if pl else time
                    cur = [tp, p1, None]
                    if self.log_orders:
                        lim = p2 if base == M["LIMIT_ORDER"] else None
                        ex.append([time, _abi.MSG_NAMES[base], _order_dict(p0, tp, sym, abs(p3), p3 > 0, p1, None, lim)])
                        if p1 == 0:
                            cur[2] = len(ex) - 1
                else:
                    ex.append([time, _abi.MSG_NAMES[base], int(p0)])
                    if base == M["QUERY_SPREAD"] and p0 in logged:   # getInsideBids / getInsideAsks(depth) right now (:189-204)
                        bids = sorted(((q, -v) for q, v in depth.items() if v < 0), reverse=True)[:max(0, p1)]
                        asks = sorted((q, v) for q, v in depth.items() if v > 0)[:max(0, p1)]
                        spread_replies.setdefault(p0, []).append((bids, asks, traded))
                continue
            # delivered to a trading agent
            last_time[agent] = time
            if base == M["ORDER_EXECUTED"]:
                c = copy_qty.get(p0)
                if c is not None and abs(p2) < c[0]:               # orderExecuted (:416-421): a full fill only drops the copy
                    c[0] -= abs(p2)
                    copy_hist[p0].append((ti, c[0]))
            if agent not in logged:
                continue
            blocks[agent].append([time, ti, [], [], []])
            r = blocks[agent][-1][2]
            if agent in own_orders and base in (M["ORDER_EXECUTED"], M["ORDER_ACCEPTED"], M["ORDER_CANCELLED"]):
                r.append((time, _abi.MSG_NAMES[base], _order_dict(agent, sent_tp.get(p0, time), sym, abs(p2), p2 > 0, p0, p3, p1)))
            if code == 0:
                if agent not in woke:
                    woke.add(agent)
                    r.append((time, 'HOLDINGS_UPDATED', dict(hold[agent])))
            elif base == M["QUERY_SPREAD"]:
                top = ((int(p0), int(p1)) if p0 >= 0 else None, (int(p2), int(p3)) if p2 >= 0 else None)
                q = spread_replies.get(agent, [])
                k = next((i for i, (b, a, _) in enumerate(q) if ((b[0] if b else None), (a[0] if a else None)) == top), None)
                if k is None:
                    raise RuntimeError("order-book account disagrees with the QUERY_SPREAD reply to agent %d at %d" % (agent, time))
                bids, asks, had_trade = q.pop(k)
                open_price_last[agent] = not had_trade
                r.append((time, 'BID_DEPTH', bids))
                r.append((time, 'ASK_DEPTH', asks))
                r.append((time, 'IMBALANCE', [sum(x[1] for x in bids), sum(x[1] for x in asks)]))
            elif base == M["ORDER_EXECUTED"]:
                h = hold[agent]
                h[sym] = h.get(sym, 0) + p2
                if h[sym] == 0:
                    del h[sym]
                h['CASH'] -= p2 * p3
                r.append((time, 'HOLDINGS_UPDATED', dict(h)))
            elif base == M["MKT_CLOSED"]:
                r.append((time, 'MKT_CLOSED', ''))
        rows = {i: [row for b in blocks[i] for part in b[2:] for row in part] for i in logged}
        for i, a in logged.items():                                # kernelStopping (TradingAgent.py:122-133, subclasses)
            t = last_time.get(i, NAT)
            r = rows[i]
            h = dict(a.holdings)
            r.append((t, 'FINAL_HOLDINGS', a.fmtHoldings(h)))
            r.append((t, 'FINAL_CASH_POSITION', h['CASH']))
            for s, shares in h.items():
                if s == 'CASH':
                    continue
                lt = (a.marked_to_market - h['CASH']) // shares   # one symbol: marked_to_market = CASH + last_trade * shares
                if open_price_last.get(i) and getattr(self.kernel.oracle, "symbols", None):
                    # the agent only ever saw the opening price, which the exchange took from the oracle as it is
                    # configured -- 1e5, a float, in the stock configs (ExchangeAgent.py:84-86, TradingAgent.py:590-601)
                    lt = type(self.kernel.oracle.symbols[s]['r_bar'])(lt)
                r.append((t, 'MARK_TO_MARKET', "{} {} @ {} == {}".format(shares, s, lt, lt * shares)))
            r.append((t, 'MARKED_TO_MARKET', a.marked_to_market))
            r.append((t, 'ENDING_CASH', a.marked_to_market))
            for row in summary_tail(a):
                r.append((t,) + row)
        self.agent_rows = {logged[i].name: r for i, r in rows.items() if i in logged}

    # ---- frames -------------------------------------------------------------------------------------------------
    def exchange_frame(self):
        return _event_frame(self.exchange_rows)

    def agent_frames(self):
        return {name: _event_frame(r) for name, r in self.agent_rows.items()}

    @property
    def n_snapshots(self):
        return len(self.book_times)

    def book_frame(self):
        """OrderBook.book_log_to_df + ExchangeAgent.logOrderBookSnapshots for book_freq == 0, wide_book: a sparse frame
        snapshot times x every quote seen, of several snapshots in one nanosecond the last, times ascending."""
        import pandas as pd
        from scipy.sparse import coo_matrix
        if not len(self.book_times):
            return None
        t = np.frombuffer(self.book_times, np.int64)
        keep = np.ones(len(t), bool)
        keep[:-1] = t[1:] != t[:-1]                     # snapshots are taken in time order: duplicates are adjacent
        new_row = np.cumsum(keep) - 1
        r = np.frombuffer(self.book_cells[0], np.int64) if len(self.book_cells[0]) else np.zeros(0, np.int64)
        q = np.frombuffer(self.book_cells[1], np.int64) if len(self.book_cells[1]) else np.zeros(0, np.int64)
        v = np.frombuffer(self.book_cells[2], np.int64) if len(self.book_cells[2]) else np.zeros(0, np.int64)
        quotes = np.array(sorted(self.quotes_seen), np.int64)
        sel = keep[r]
        S = coo_matrix((v[sel], (new_row[r[sel]], np.searchsorted(quotes, q[sel]))), shape=(int(keep.sum()), len(quotes)),
                       dtype=np.int64).tocsc()
        df = pd.DataFrame.sparse.from_spmatrix(S, columns=[int(x) for x in quotes])
        df.index = pd.DatetimeIndex(pd.to_datetime(t[keep]), name='QuoteTime')
        return df


def summary_tail(a):
    """The (type, value) rows an agent logs after ENDING_CASH, all of them also summary-log rows: FINAL_VALUATION
    (ZeroIntelligenceAgent.py:93, ValueAgent.py:75, NoiseAgent.py:68) and a trading ExecutionAgent's report
    (agent/execution/ExecutionAgent.py:32-42).  Needs runtime.write_back() to have run."""
    out = []
    if a.final_valuation is not None:
        # ZI / HBL report an integer surplus, Value / Noise agents a ratio to the starting cash
        zi = any(k.__name__ == "ZeroIntelligenceAgent" for k in type(a).__mro__)
        out.append(('FINAL_VALUATION', int(a.final_valuation) if zi else a.final_valuation))
    if hasattr(a, "execution_time_horizon") and a.trade:
        avg, mid = a.average_transaction_price, a.arrival_price
        if avg is None or mid is None:
            raise RuntimeError("the reference raises in ExecutionAgent.kernelStopping: no fill or no arrival price")
        out += [('DIRECTION', a.direction), ('TOTAL_QTY', a.quantity), ('REM_QTY', a.rem_quantity), ('ARRIVAL_MID', mid),
                ('AVG_TXN_PRICE', avg), ('SLIPPAGE', avg - mid if a.direction == 'BUY' else mid - avg)]
    return out


def _event_frame(rows):
    import pandas as pd
    t = np.array([r[0] for r in rows], np.int64)
    df = pd.DataFrame({'EventTime': pd.to_datetime(t).where(t != NAT, pd.NaT), 'EventType': [r[1] for r in rows],
                       'Event': pd.Series([r[2] for r in rows], dtype=object)})
    return df.set_index('EventTime')


def _ns(t):
    import pandas as pd
    return int(pd.Timestamp(t).value)


def write_archives(kernel, trace, path, spec=None):
    """Write the exchange's log, the order-book archive and the agents' logs under `path` with the reference's file
    names (Kernel.writeLog: the agent's name without blanks).  Returns the ExchangeArchive."""
    ar = ExchangeArchive(kernel, trace, spec)
    os.makedirs(path, exist_ok=True)

    def put(df, name):
        df.to_pickle(os.path.join(path, "{}.bz2".format(name.replace(" ", ""))), compression='bz2')

    if getattr(ar.exch, "log_to_file", True):
        put(ar.exchange_frame(), ar.exch.name)
    if ar.snapshots and str(ar.exch.book_freq).isdigit() and int(ar.exch.book_freq) == 0 and ar.exch.wide_book:
        bf = ar.book_frame()
        if bf is not None:
            put(bf, "ORDERBOOK_{}_FULL".format(ar.symbol))
    for name, df in ar.agent_frames().items():
        put(df, name)
    return ar
