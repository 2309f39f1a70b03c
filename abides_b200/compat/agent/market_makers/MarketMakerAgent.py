"""MarketMakerAgent descriptor (reference: agent/market_makers/MarketMakerAgent.py:27-50), polling mode only.
Every wake-up it cancels its orders, polls the spread and quotes 2 * subscribe_num_levels prices per side
around the mid with a freshly drawn size per price.  The constructor consumes one draw of the agent's stream,
as the reference's does."""
from agent.TradingAgent import TradingAgent


class MarketMakerAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, starting_cash, min_size, max_size, wake_up_freq='1s',
                 subscribe=False, subscribe_freq=10e9, subscribe_num_levels=5, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.min_size, self.max_size = min_size, max_size
        self.size = round(self.random_state.randint(self.min_size, self.max_size) / 2)
        self.wake_up_freq = wake_up_freq
        self.subscribe, self.subscribe_freq, self.subscribe_num_levels = subscribe, subscribe_freq, subscribe_num_levels
        self.last_spread = 10
