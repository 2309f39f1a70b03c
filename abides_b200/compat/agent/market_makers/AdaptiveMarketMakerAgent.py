"""AdaptiveMarketMakerAgent descriptor (reference: agent/market_makers/AdaptiveMarketMakerAgent.py:29-110)."""
from math import ceil

from agent.TradingAgent import TradingAgent

ANCHOR_TOP_STR, ANCHOR_BOTTOM_STR, ANCHOR_MIDDLE_STR = 'top', 'bottom', 'middle'
ADAPTIVE_SPREAD_STR = 'adaptive'
INITIAL_SPREAD_VALUE = 50


class AdaptiveMarketMakerAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, starting_cash, pov=0.05, min_order_size=20, window_size=5,
                 anchor=ANCHOR_MIDDLE_STR, num_ticks=20, level_spacing=0.5, wake_up_freq='1s', subscribe=False,
                 subscribe_freq=10e9, subscribe_num_levels=1, cancel_limit_delay=50, skew_beta=0, spread_alpha=0.85,
                 backstop_quantity=None, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        if anchor not in (ANCHOR_TOP_STR, ANCHOR_BOTTOM_STR, ANCHOR_MIDDLE_STR):
            raise ValueError("Variable anchor must take the value `bottom`, `middle` or `top`")
        self.is_adaptive = False
        self.symbol = symbol
        self.pov = pov
        self.min_order_size = min_order_size
        self.anchor = anchor
        try:
            self.window_size = int(window_size)
        except (TypeError, ValueError):
            if str(window_size).lower() != ADAPTIVE_SPREAD_STR:
                raise ValueError("Variable window_size must be of type int or string adaptive.")
            self.is_adaptive = True
            self.anchor = ANCHOR_MIDDLE_STR
            self.window_size = None
        self.num_ticks = num_ticks
        self.level_spacing = level_spacing
        self.wake_up_freq = wake_up_freq
        self.subscribe = subscribe
        self.subscribe_freq = subscribe_freq
        self.subscribe_num_levels = subscribe_num_levels
        self.cancel_limit_delay = cancel_limit_delay
        self.skew_beta = skew_beta
        self.spread_alpha = spread_alpha
        self.backstop_quantity = backstop_quantity
        self.last_spread = INITIAL_SPREAD_VALUE
        self.tick_size = None if self.is_adaptive else ceil(self.last_spread * self.level_spacing)
