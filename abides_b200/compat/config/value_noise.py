"""value_noise: exchange + 100 noise + 50 value agents for one market hour (09:30-10:30), legacy symmetric latency
matrix with per-message noise, one-second computation delay (reference: config/value_noise.py).  Same flags and
the same order of draws from the global numpy stream."""
import argparse
import sys

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.NoiseAgent import NoiseAgent
from agent.ValueAgent import ValueAgent
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder
from util.util import ExchangeStarLatency

ap = argparse.ArgumentParser(description='Detailed options for sparse_zi config.')
ap.add_argument('-b', '--book_freq', default=None)
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-l', '--log_dir', default=None)
ap.add_argument('-n', '--obs_noise', type=float, default=1000000)
ap.add_argument('-o', '--log_orders', action='store_true')
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('-v', '--verbose', action='store_true')
ap.add_argument('--config_help', action='store_true')
args, _ = ap.parse_known_args()
if args.config_help:
    ap.print_help()
    sys.exit()

seed = args.seed or int(pd.Timestamp.now().timestamp() * 1000000) % (2 ** 32 - 1)
np.random.seed(seed)
util.silent_mode = LimitOrder.silent_mode = not args.verbose
print("Silent mode: {}".format(util.silent_mode))
print("Logging orders: {}".format(args.log_orders))
print("Book freq: {}".format(args.book_freq))
print("ZeroIntelligenceAgent noise: {:0.4f}".format(args.obs_noise))
print("Configuration seed: {}\n".format(seed))


def fresh_state():
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, dtype='uint64'))


day = pd.to_datetime('2019-06-28')
sym = 'JPM'
fund = {'r_bar': 1e5, 'kappa': 1.67e-12, 'agent_kappa': 1.67e-15, 'sigma_s': 0, 'fund_vol': 1e-8,
        'megashock_lambda_a': 2.77778e-13, 'megashock_mean': 1e3, 'megashock_var': 5e4, 'random_state': fresh_state()}
kernel = Kernel("Base Kernel", random_state=fresh_state())
mkt_open, mkt_close = day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta('10:30:00')
oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, {sym: fund})

agents = [ExchangeAgent(0, "Exchange Agent 0", "ExchangeAgent", mkt_open, mkt_close, [sym],
                        log_orders=args.log_orders, book_freq=args.book_freq, pipeline_delay=0, computation_delay=0,
                        stream_history=10, random_state=fresh_state())]
for _ in range(100):
    j = len(agents)
    rs = fresh_state()                                              # the agent's seed first, then its wake time
    agents.append(NoiseAgent(j, "NoiseAgent {}".format(j), "NoiseAgent", random_state=rs,
                             log_orders=args.log_orders, symbol=sym, starting_cash=10000000,
                             wakeup_time=mkt_open + np.random.rand() * (mkt_close - mkt_open)))
for _ in range(50):
    j = len(agents)
    agents.append(ValueAgent(j, "Value Agent {}".format(j), "ValueAgent {}".format(j), random_state=fresh_state(),
                             log_orders=args.log_orders, symbol=sym, sigma_n=args.obs_noise, r_bar=fund['r_bar'],
                             kappa=fund['agent_kappa'], sigma_s=fund['fund_vol'], lambda_a=1e-12))

# N x N uniform draw as the reference; it mirrors the upper triangle down and sets the diagonal to 20000, and only
# the exchange row / column is ever read on this path
n = len(agents)
m = np.random.uniform(low=21000, high=13000000, size=(n, n))
row0 = m[0, :].copy()
row0[0] = 20000
kernel.runner(agents=agents, startTime=day, stopTime=day + pd.to_timedelta('17:00:00'),
              agentLatency=ExchangeStarLatency(row0, row0), latencyNoise=[0.25, 0.25, 0.20, 0.15, 0.10, 0.05],
              defaultComputationDelay=1000000000, oracle=oracle, log_dir=args.log_dir)
