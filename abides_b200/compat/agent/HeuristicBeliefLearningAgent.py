"""HeuristicBeliefLearningAgent descriptor (reference: agent/HeuristicBeliefLearningAgent.py:13-27): a
ZeroIntelligenceAgent that, once the exchange's order stream holds L trade-delimited buckets, prices its order
at the limit price with the highest expected surplus under the empirical fill probabilities of those buckets."""
from agent.ZeroIntelligenceAgent import ZeroIntelligenceAgent


class HeuristicBeliefLearningAgent(ZeroIntelligenceAgent):
    def __init__(self, id, name, type, symbol='IBM', starting_cash=100000, sigma_n=1000, r_bar=100000, kappa=0.05,
                 sigma_s=100000, q_max=10, sigma_pv=5000000, R_min=0, R_max=250, eta=1.0, lambda_a=0.005, L=8,
                 log_orders=False, random_state=None):
        super().__init__(id, name, type, symbol=symbol, starting_cash=starting_cash, sigma_n=sigma_n, r_bar=r_bar,
                         kappa=kappa, sigma_s=sigma_s, q_max=q_max, sigma_pv=sigma_pv, R_min=R_min, R_max=R_max,
                         eta=eta, lambda_a=lambda_a, log_orders=log_orders, random_state=random_state)
        self.L = L
