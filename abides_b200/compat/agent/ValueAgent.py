"""ValueAgent descriptor (reference: agent/ValueAgent.py:11-45).  Order size comes from the
process-global numpy stream, as in the reference."""
import numpy as np

from agent.TradingAgent import TradingAgent


class ValueAgent(TradingAgent):
    def __init__(self, id, name, type, symbol='IBM', starting_cash=100000, sigma_n=10000, r_bar=100000, kappa=0.05,
                 sigma_s=100000, lambda_a=0.005, log_orders=False, log_to_file=True, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders,
                         log_to_file=log_to_file, random_state=random_state)
        self.symbol = symbol
        self.sigma_n, self.r_bar, self.kappa, self.sigma_s, self.lambda_a = sigma_n, r_bar, kappa, sigma_s, lambda_a
        self.percent_aggr = 0.1
        self.size = np.random.randint(20, 50)
        self.depth_spread = 2
