"""SubscriptionAgent descriptor (reference: agent/examples/SubscriptionAgent.py:7-55): at its first wake it asks the
exchange for `levels` book levels at most every `freq` ns, then only listens (the reference prints each update)."""
from agent.TradingAgent import TradingAgent


class SubscriptionAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, starting_cash, levels, freq, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.levels, self.freq = levels, freq
        self.subscribe = True
        self.subscription_requested = False
        self.state = 'AWAITING_MARKET_DATA'
