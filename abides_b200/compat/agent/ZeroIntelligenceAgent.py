"""ZeroIntelligenceAgent descriptor (reference: agent/ZeroIntelligenceAgent.py:11-51).  The private
values theta are drawn here, from the agent's own stream, exactly as the reference constructor does."""
from math import sqrt

import numpy as np

from agent.TradingAgent import TradingAgent


class ZeroIntelligenceAgent(TradingAgent):
    def __init__(self, id, name, type, symbol='IBM', starting_cash=100000, sigma_n=1000, r_bar=100000, kappa=0.05,
                 sigma_s=100000, q_max=10, sigma_pv=5000000, R_min=0, R_max=250, eta=1.0, lambda_a=0.005,
                 log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.sigma_n, self.r_bar, self.kappa, self.sigma_s = sigma_n, r_bar, kappa, sigma_s
        self.q_max, self.sigma_pv = q_max, sigma_pv
        self.R_min, self.R_max, self.eta, self.lambda_a = R_min, R_max, eta, lambda_a
        draws = self.random_state.normal(loc=0, scale=sqrt(sigma_pv), size=(q_max * 2))
        self.theta = [int(v) for v in sorted(np.round(draws).tolist(), reverse=True)]
