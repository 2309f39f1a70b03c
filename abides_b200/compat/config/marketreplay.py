"""marketreplay: one exchange and one MarketReplayAgent replaying a historical L3 order file, no oracle, zero latency and
zero computation delay (reference: config/marketreplay.py:14-148).  Same flags; the reference hard-codes the file as
/efs/data/DOW30/<ticker>/<ticker>.<date> with its cache under /efs/data/marketreplay/ -- `--orders-file` and
`--processed-orders-folder` (extensions) point them elsewhere.  The reference's exchange is built with book_freq='all',
which makes its own archival step fail at the end of the run (`resample('all')`, agent/ExchangeAgent.py:353); here the
order book archive is written per update (book_freq=0, wide) instead."""
import argparse
import sys

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.examples.MarketReplayAgent import MarketReplayAgent
from util import util
from util.order import LimitOrder

ap = argparse.ArgumentParser(description='Detailed options for market replay config.')
ap.add_argument('-c', '--config', required=True, help='Name of config file to execute')
ap.add_argument('-t', '--ticker', required=True, help='Name of the stock/symbol')
ap.add_argument('-d', '--date', required=True, help='Historical date')
ap.add_argument('-l', '--log_dir', default=None, help='Log directory name (default: unix timestamp at program start)')
ap.add_argument('-s', '--seed', type=int, default=None, help='numpy.random.seed() for simulation')
ap.add_argument('-v', '--verbose', action='store_true', help='Maximum verbosity!')
ap.add_argument('--config_help', action='store_true', help='Print argument options for this config file')
ap.add_argument('--orders-file', default=None, help='order file (default: /efs/data/DOW30/<ticker>/<ticker>.<date>)')
ap.add_argument('--processed-orders-folder', default='/efs/data/marketreplay/')
ap.add_argument('--start-time', default='09:00:00')
ap.add_argument('--end-time', default='16:00:00')
args, _ = ap.parse_known_args()
if args.config_help:
    ap.print_help()
    sys.exit()

seed = args.seed
if not seed:
    seed = int(pd.Timestamp.now().timestamp() * 1000000) % (2 ** 32 - 1)
np.random.seed(seed)
util.silent_mode = LimitOrder.silent_mode = not args.verbose
print("Configuration seed: {}".format(seed))
print("Log Directory: {}".format(args.log_dir))

day = pd.to_datetime(args.date)
symbol = args.ticker
print("Symbol: {}".format(symbol))
print("Date: {}\n".format(args.date))
mkt_open, mkt_close = day + pd.to_timedelta(args.start_time), day + pd.to_timedelta(args.end_time)
print("Market Open : {}".format(mkt_open))
print("Market Close: {}".format(mkt_close))


def fresh_state():
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, dtype='uint64'))


agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[symbol], log_orders=True, pipeline_delay=0, computation_delay=0, stream_history=10,
                        book_freq=0, wide_book=True, random_state=fresh_state())]
agents.append(MarketReplayAgent(id=1, name="MARKET_REPLAY_AGENT", type='MarketReplayAgent', symbol=symbol, log_orders=False,
                                date=day, start_time=mkt_open, end_time=mkt_close,
                                orders_file_path=args.orders_file or '/efs/data/DOW30/{0}/{0}.{1}'.format(symbol, args.date),
                                processed_orders_folder_path=args.processed_orders_folder, starting_cash=0,
                                random_state=fresh_state()))
kernel = Kernel("Market Replay Kernel", random_state=fresh_state())
kernel.runner(agents=agents, startTime=day, stopTime=day + pd.to_timedelta('17:00:00'),
              agentLatency=np.zeros((len(agents), len(agents))), latencyNoise=[0.0], defaultComputationDelay=0,
              defaultLatency=0, oracle=None, log_dir=args.log_dir)
