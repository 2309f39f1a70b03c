"""NoiseAgent descriptor (reference: agent/NoiseAgent.py:11-34)."""
import numpy as np

from agent.TradingAgent import TradingAgent


class NoiseAgent(TradingAgent):
    def __init__(self, id, name, type, symbol='IBM', starting_cash=100000, log_orders=False, log_to_file=True,
                 random_state=None, wakeup_time=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders,
                         log_to_file=log_to_file, random_state=random_state)
        self.wakeup_time = (wakeup_time,)        # the reference stores a 1-tuple (NoiseAgent.py:18)
        self.symbol = symbol
        self.size = np.random.randint(20, 50)
