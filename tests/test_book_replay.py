"""Order-book replay parity: identical order streams into util/OrderBook.py (reference,
golden), the C oracle and the CUDA book; fills, notifications and the book state after EVERY message are
bit-exact (integer cents)."""
import os

import numpy as np
import pytest

import oracle
from abides_b200 import engine as eng_mod
from parity_util import GOLDEN


def golden():
    g = np.load(os.path.join(GOLDEN, "book_replay.golden.npz"))
    return g, int(g["stream_history"])


def check(events, counts, states, g):
    # the reference's market-order children take ids from its global generator; both restatements use 1<<30 + k
    assert np.array_equal(counts, g["counts"])
    assert np.array_equal(states, g["states"])
    ge = g["events"]
    assert events.shape == ge.shape
    assert np.array_equal(events[:, [0, 1, 2, 3, 5, 6, 7]], ge[:, [0, 1, 2, 3, 5, 6, 7]])
    assert np.array_equal(events[:, 4], ge[:, 4])


def test_oracle_book_replay_matches_reference_orderbook():
    g, sh = golden()
    ops = eng_mod.make_book_ops(g["ops"])
    n = int(g["op_offset"][-1])
    ev, cnt, st = oracle.replay_book(g["op_offset"], ops, sh, 64)
    events, counts, states = eng_mod.unpack_replay(ev, cnt[:n], st, n, 64)
    check(events, counts, states, g)


@pytest.mark.gpu
def test_device_book_replay_matches_reference_orderbook():
    g, sh = golden()
    e = eng_mod.Engine(0)
    events, counts, states = eng_mod.replay_book(e, g["op_offset"], g["ops"], sh, 64)
    check(events, counts, states, g)


@pytest.mark.gpu
def test_device_book_replay_many_books_property():
    """Size-independent properties at scale: 512 random books; shares conserved, book never crossed."""
    rng = np.random.default_rng(7)
    rows, off = [], [0]
    for b in range(512):
        t = 0
        for i in range(300):
            t += 1000
            buy = int(rng.integers(0, 2))
            price = 100000 + (-1 if buy else 1) * int(rng.integers(-30, 200))
            rows.append((7, int(rng.integers(1, 50)), i + 1, price, int(rng.integers(1, 100)), buy, 0, t))
        off.append(len(rows))
    rows = np.array(rows, np.int64)
    e = eng_mod.Engine(0)
    events, counts, states = eng_mod.replay_book(e, np.array(off, np.int64), rows, 10, 64)
    ok = (states[:, 0] < 0) | (states[:, 2] < 0) | (states[:, 0] < states[:, 2])
    assert ok.all()                                   # best bid < best ask after every message
    ex = events[events[:, 1] == 12]
    assert len(ex) % 2 == 0
    buys = ex[ex[:, 3] == 1][:, 6].sum()
    sells = ex[ex[:, 3] == 0][:, 6].sum()
    assert buys == sells                              # every fill has both sides
    ev2, cnt2, st2 = oracle.replay_book(np.array(off, np.int64), eng_mod.make_book_ops(rows), 10, 64)
    events_o, counts_o, states_o = eng_mod.unpack_replay(ev2, cnt2[:len(rows)], st2, len(rows), 64)
    assert np.array_equal(states, states_o)
    assert np.array_equal(events, events_o)


@pytest.mark.gpu
def test_device_deep_book_fuzz_matches_oracle():
    """Deep, wide books with every message type: thousands of resting orders per book (far past the 64 per side kept
    in shared memory, so the pool, its slot reuse, refills and level sums that run into the pool are all exercised),
    cancels of live / dead / never-seen orders, modifies (incl. the reference's index-0 overwrite) and market orders
    that sweep several levels.  Device == oracle on every notification and on the book state after every message."""
    rng = np.random.default_rng(20261020)
    rows, off = [], [0]
    for b in range(48):
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
    rows = np.array(rows, np.int64)
    offs = np.array(off, np.int64)
    e = eng_mod.Engine(0)
    events, counts, states = eng_mod.replay_book(e, offs, rows, 25, 96)
    ev2, cnt2, st2 = oracle.replay_book(offs, eng_mod.make_book_ops(rows), 25, 96)
    events_o, counts_o, states_o = eng_mod.unpack_replay(ev2, cnt2[:len(rows)], st2, len(rows), 96)
    assert (counts <= 96).all() and np.array_equal(counts, counts_o)
    assert np.array_equal(states, states_o)
    # market-order children take fresh ids from 1 << 30 in both restatements
    assert np.array_equal(events, events_o)
    assert states[:, 5].max() > 100 and states[:, 6].max() > 100      # books really were deep (levels per side; 64 orders per side fit in shared memory)


ORDERS_FILE = os.path.join(GOLDEN, "sample_orders_file.csv")     # the reference's data/sample_orders_file.csv (10 rows)


def test_order_file_parser_reproduces_the_market_replay_tape():
    """abides_b200.orderfile.parse_orders_file turns the order file into the messages MarketReplayAgent.placeOrder
    (:47-62) sends: tape 0 of the book-replay golden was built from the same file by the generator."""
    from abides_b200 import orderfile
    g, _ = golden()
    lo, hi = int(g["op_offset"][0]), int(g["op_offset"][1])
    tape = orderfile.parse_orders_file(ORDERS_FILE)
    assert np.array_equal(tape[:, :7], g["ops"][lo:hi, :7])
    # the generator read the timestamps through pandas' float64 column (sub-millisecond digits lost); the parser keeps them
    assert np.abs(tape[:, 7] - g["ops"][lo:hi, 7]).max() < 4 * 10 ** 6 and tape[0, 7] == 1559554200028558000
    assert {int(k) for k in tape[:, 0]} == {7, 9, 10}           # limit orders, a cancel and a modify


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["warp", "lane"])
def test_order_file_replay_on_device_matches_reference_orderbook(engine):
    from abides_b200 import orderfile
    g, sh = golden()
    lo, hi = int(g["op_offset"][0]), int(g["op_offset"][1])
    e = eng_mod.Engine(0)
    e.set_engine(engine)
    tape, events, counts, states = orderfile.replay(e, ORDERS_FILE, stream_history=sh)
    assert np.array_equal(counts, g["counts"][lo:hi]) and np.array_equal(states, g["states"][lo:hi])
    ge = g["events"][(g["events"][:, 0] >= lo) & (g["events"][:, 0] < hi)]
    assert np.array_equal(events, ge)
