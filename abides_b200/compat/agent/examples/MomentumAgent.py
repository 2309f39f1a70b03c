"""MomentumAgent descriptor (reference: agent/examples/MomentumAgent.py:13-27)."""
from agent.TradingAgent import TradingAgent


class MomentumAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, starting_cash, min_size, max_size, wake_up_freq='60s',
                 subscribe=False, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.min_size, self.max_size = min_size, max_size
        self.size = self.random_state.randint(self.min_size, self.max_size)
        self.wake_up_freq = wake_up_freq
        self.subscribe = subscribe
