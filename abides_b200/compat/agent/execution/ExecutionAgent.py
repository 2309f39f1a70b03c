"""ExecutionAgent descriptor (reference: agent/execution/ExecutionAgent.py:8-105): the common part of the TWAP and
VWAP agents.  On the device the agent wakes on every point of `execution_time_horizon`, asks for the spread, and on
the reply cancels its resting orders and places schedule[slot] shares at the far touch -- or, on the second-to-last
point, a MARKET order for whatever is left.  Subclasses fill `schedule` ({pd.Interval: shares})."""
from agent.TradingAgent import TradingAgent


class ExecutionAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, starting_cash, direction, quantity, execution_time_horizon,
                 trade=True, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.direction, self.quantity, self.rem_quantity = direction, quantity, quantity
        self.execution_time_horizon = execution_time_horizon
        self.start_time, self.end_time = execution_time_horizon[0], execution_time_horizon[-1]
        self.schedule = None
        self.arrival_price = None
        self.accepted_orders = []
        self.trade = trade
        self.state = 'AWAITING_WAKEUP'
