"""Golden-vector generator input (runs on the UNMODIFIED reference only): the rmsc01 population (config/rmsc01.py:
1 MarketMakerAgent + 50 ZI + 25 HBL + 24 momentum) with the agent counts and the closing time as flags, written
against the reference's own API in the draw order of config/rmsc01.py, so that
`abides_b200/compat/config/rmsc01.py --num-zi .. --num-hbl .. --num-momentum .. --end-time ..` lowers to the same
inputs (spec_sha256)."""
import argparse

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.HeuristicBeliefLearningAgent import HeuristicBeliefLearningAgent
from agent.ZeroIntelligenceAgent import ZeroIntelligenceAgent
from agent.examples.MomentumAgent import MomentumAgent
from agent.market_makers.MarketMakerAgent import MarketMakerAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder

ap = argparse.ArgumentParser()
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('--num-zi', type=int, default=50)
ap.add_argument('--num-hbl', type=int, default=25)
ap.add_argument('--num-momentum', type=int, default=24)
ap.add_argument('--end-time', default='16:00:00')
args, _ = ap.parse_known_args()
np.random.seed(args.seed)
util.silent_mode = True
LimitOrder.silent_mode = True


def rs(**kw):
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, **kw))


day = pd.to_datetime('2019-06-28')
mkt_open, mkt_close = day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta(args.end_time)
sym, cash = 'JPM', 10000000
agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[sym], log_orders=False, pipeline_delay=0, computation_delay=0, stream_history=10,
                        book_freq=0, random_state=rs(dtype='uint64'))]
agents.append(MarketMakerAgent(id=1, name="MARKET_MAKER_AGENT_1", type='MarketMakerAgent', symbol=sym, starting_cash=cash,
                               min_size=500, max_size=1000, log_orders=False, random_state=rs(dtype='uint64')))
symbols = {sym: {'r_bar': 1e5, 'kappa': 1.67e-12, 'agent_kappa': 1.67e-15, 'sigma_s': 0, 'fund_vol': 1e-8,
                 'megashock_lambda_a': 2.77778e-13, 'megashock_mean': 1e3, 'megashock_var': 5e4,
                 'random_state': rs(dtype='uint64')}}
oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, symbols)
kw = dict(symbol=sym, starting_cash=cash, sigma_n=10000, sigma_s=symbols[sym]['fund_vol'], kappa=symbols[sym]['agent_kappa'],
          r_bar=symbols[sym]['r_bar'], q_max=10, sigma_pv=5e4, R_min=0, R_max=100, eta=1, lambda_a=1e-12, log_orders=False)
for _ in range(args.num_zi):
    j = len(agents)
    agents.append(ZeroIntelligenceAgent(id=j, name="ZI_AGENT_{}".format(j), type="ZeroIntelligenceAgent",
                                        random_state=rs(dtype='uint64'), **kw))
for _ in range(args.num_hbl):
    j = len(agents)
    agents.append(HeuristicBeliefLearningAgent(id=j, name="HBL_AGENT_{}".format(j), type="HeuristicBeliefLearningAgent", L=2,
                                               random_state=rs(dtype='uint64'), **kw))
for _ in range(args.num_momentum):
    j = len(agents)
    agents.append(MomentumAgent(id=j, name="MOMENTUM_AGENT_{}".format(j), type="MomentumAgent", symbol=sym,
                                starting_cash=cash, min_size=1, max_size=10, log_orders=False,
                                random_state=rs(dtype='uint64')))
kernel = Kernel("RMSC01 Kernel", random_state=rs(dtype='uint64'))
lat_rs = rs()
lat = util.meters_to_light_ns(util.generate_uniform_random_pairwise_dist_on_line(0.0, 3866660, len(agents), random_state=lat_rs))
model = LatencyModel(latency_model='deterministic', random_state=lat_rs, kwargs={'connected': True, 'min_latency': lat})
kernel.runner(agents=agents, startTime=day, stopTime=mkt_close + pd.to_timedelta('00:01:00'), agentLatencyModel=model,
              defaultComputationDelay=50, oracle=oracle, log_dir=None)
