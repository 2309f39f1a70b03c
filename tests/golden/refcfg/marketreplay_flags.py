"""Golden-vector generator input (runs on the UNMODIFIED reference only): config/marketreplay.py -- one ExchangeAgent
and the reference's MarketReplayAgent, no oracle, zero latency, zero computation delay -- with the order file as a flag.
The reference's L3OrdersProcessor cannot read a comma-separated file (and drops the first data row of a '|'-separated
one), so the tape is handed to it the way its own cache does: as the processed-orders pickle it looks for first
(agent/examples/MarketReplayAgent.py:89-92), written here from the CSV with the reference's price rule (dollars * 100,
truncated, :113-114).  book_freq=0 / wide_book instead of the stock 'all', on which the reference's archival fails."""
import argparse
import csv
import datetime as dt
import os
import pickle
import tempfile

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.examples.MarketReplayAgent import MarketReplayAgent
from util import util
from util.order import LimitOrder

ap = argparse.ArgumentParser()
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-t', '--ticker', required=True)
ap.add_argument('-d', '--date', required=True)
ap.add_argument('-l', '--log_dir', default=None)
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('--orders-file', required=True)
ap.add_argument('--start-time', default='09:00:00')
ap.add_argument('--end-time', default='16:00:00')
args, _ = ap.parse_known_args()
np.random.seed(args.seed)
util.silent_mode = True
LimitOrder.silent_mode = True

day = pd.to_datetime(args.date)
symbol = args.ticker
mkt_open, mkt_close = day + pd.to_timedelta(args.start_time), day + pd.to_timedelta(args.end_time)

orders = {}
with open(args.orders_file, newline="") as f:
    for r in csv.DictReader(f):
        whole, _, frac = r['TIMESTAMP'].partition('.')
        t = pd.Timestamp(dt.datetime.strptime(whole, '%Y%m%d%H%M%S')) + pd.Timedelta(int((frac + '0' * 9)[:9]), 'ns')
        if not (mkt_open <= t < mkt_close):
            continue
        orders.setdefault(t, []).append({'ORDER_ID': int(r['ORDER_ID']), 'PRICE': int(float(r['PRICE']) * 100),
                                         'SIZE': int(r['SIZE']), 'BUY_SELL_FLAG': r['BUY_SELL_FLAG']})
cache = tempfile.mkdtemp(prefix='abides_replay_') + os.sep
with open('{}marketreplay_{}_{}.pkl'.format(cache, symbol, day.date()), 'wb') as h:
    pickle.dump(dict(sorted(orders.items())), h)


def rs():
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, dtype='uint64'))


agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[symbol], log_orders=True, pipeline_delay=0, computation_delay=0, stream_history=10,
                        book_freq=0, wide_book=True, random_state=rs())]
agents.append(MarketReplayAgent(id=1, name="MARKET_REPLAY_AGENT", type='MarketReplayAgent', symbol=symbol, log_orders=False,
                                date=day, start_time=mkt_open, end_time=mkt_close, orders_file_path=args.orders_file,
                                processed_orders_folder_path=cache, starting_cash=0, random_state=rs()))
kernel = Kernel("Market Replay Kernel", random_state=rs())
kernel.runner(agents=agents, startTime=day, stopTime=day + pd.to_timedelta('17:00:00'),
              agentLatency=np.zeros((len(agents), len(agents))), latencyNoise=[0.0], defaultComputationDelay=0,
              defaultLatency=0, oracle=None, log_dir=args.log_dir)
