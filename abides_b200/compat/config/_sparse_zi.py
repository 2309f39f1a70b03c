"""Shared body of the sparse_zi_* populations: one exchange + zero-intelligence agents in seven
(R_min, R_max, eta) groups over a sparse mean-reverting fundamental.  Draw order follows the reference
builders (config/sparse_zi_100.py:69-194, config/sparse_zi_1000.py) so a seed yields the same agents."""
import argparse
import sys

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.ZeroIntelligenceAgent import ZeroIntelligenceAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder
from util.util import ExchangeStarLatency

GROUP_RULES = [(0, 250, 1), (0, 500, 1), (0, 1000, 0.8), (0, 1000, 1), (0, 2000, 0.8), (250, 500, 0.8), (250, 500, 1)]


def _fresh_state(**kw):
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, **kw))


def run(group_sizes, new_latency_model):
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

    day = pd.to_datetime('2019-06-28')
    mkt_open, mkt_close = day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta('16:00:00')
    sym = 'JPM'
    fund = {'r_bar': 1e5, 'kappa': 1.67e-12, 'agent_kappa': 1.67e-15, 'sigma_s': 0, 'fund_vol': 1e-8,
            'megashock_lambda_a': 2.77778e-13, 'megashock_mean': 1e3, 'megashock_var': 5e4,
            'random_state': _fresh_state(dtype='uint64')}
    symbols = {sym: fund}
    kernel = Kernel("Base Kernel", random_state=_fresh_state(dtype='uint64'))
    latency_rstate = _fresh_state() if new_latency_model else None
    oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, symbols)

    agents = [ExchangeAgent(0, "Exchange Agent 0", "ExchangeAgent", mkt_open, mkt_close, [sym],
                            log_orders=args.log_orders, book_freq=args.book_freq, pipeline_delay=0,
                            computation_delay=0, stream_history=10, random_state=_fresh_state(dtype='uint64'))]
    for g, (count, (r_lo, r_hi, eta)) in enumerate(zip(group_sizes, GROUP_RULES)):
        label = "Type {} [{} <= R <= {}, eta={}]".format(g + 1, r_lo, r_hi, eta)
        for _ in range(count):
            j = len(agents)
            agents.append(ZeroIntelligenceAgent(
                j, "ZI Agent {} {}".format(j, label), "ZeroIntelligenceAgent {}".format(label),
                random_state=_fresh_state(dtype='uint64'), log_orders=args.log_orders, symbol=sym,
                starting_cash=10000000, sigma_n=args.obs_noise, r_bar=fund['r_bar'], kappa=fund['agent_kappa'],
                sigma_s=fund['fund_vol'], q_max=10, sigma_pv=5e6, R_min=r_lo, R_max=r_hi, eta=eta, lambda_a=1e-12))

    n = len(agents)
    latency, noise, model = None, None, None
    if new_latency_model:
        # N x N uniform draw as the reference; only the exchange row / column is ever read
        m = np.random.uniform(low=21000, high=100000, size=(n, n))
        model = LatencyModel(latency_model='cubic', random_state=latency_rstate,
                             kwargs={'connected': True, 'min_latency': ExchangeStarLatency(m[:, 0].copy(), m[0, :].copy()),
                                     'jitter': 0.3, 'jitter_clip': 0.05, 'jitter_unit': 5})
    else:
        # legacy symmetric matrix (upper triangle mirrored down, diagonal 20000) + per-message noise
        m = np.random.uniform(low=21000, high=13000000, size=(n, n))
        row0 = m[0, :].copy()
        row0[0] = 20000
        latency = ExchangeStarLatency(row0, row0)
        noise = [0.25, 0.25, 0.20, 0.15, 0.10, 0.05]

    kernel.runner(agents=agents, startTime=day, stopTime=day + pd.to_timedelta('17:00:00'),
                  agentLatencyModel=model, agentLatency=latency, latencyNoise=noise,
                  defaultComputationDelay=1000000000, oracle=oracle, log_dir=args.log_dir)
