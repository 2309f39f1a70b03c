"""POVExecutionAgent descriptor (reference: agent/execution/POVExecutionAgent.py:12-72).  With trade=False (every
BASELINE population) it is a one-minute heartbeat that must still exist so agent ids, latency rows and message
counts match; with trade=True (rmsc03 `-e`) it buys / sells `pov` of the last minute's transacted volume with a
MARKET order on every wake between start_time and end_time."""
import pandas as pd

from agent.TradingAgent import TradingAgent


class POVExecutionAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, starting_cash, direction, quantity, pov, start_time, freq,
                 lookback_period, end_time=None, trade=True, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.direction, self.quantity, self.rem_quantity, self.pov = direction, quantity, quantity, pov
        self.start_time = start_time
        self.end_time = end_time if end_time is not None else start_time + pd.to_timedelta('24 hours')
        self.freq = freq
        self.look_back_period = lookback_period
        self.trade = trade
