"""TWAPExecutionAgent descriptor (reference: agent/execution/TWAPExecutionAgent.py:6-44): the parent order split
evenly -- int(quantity / len(horizon)) shares in every `freq` bin between the first and last horizon point."""
import pandas as pd

from agent.execution.ExecutionAgent import ExecutionAgent


class TWAPExecutionAgent(ExecutionAgent):
    def __init__(self, id, name, type, symbol, starting_cash, direction, quantity, execution_time_horizon, freq,
                 trade=True, log_orders=False, random_state=None):
        super().__init__(id, name, type, symbol, starting_cash, direction=direction, quantity=quantity,
                         execution_time_horizon=execution_time_horizon, trade=trade, log_orders=log_orders,
                         random_state=random_state)
        self.freq = freq
        self.schedule = self.generate_schedule()

    def generate_schedule(self):
        child = int(self.quantity / len(self.execution_time_horizon))
        return {b: child for b in pd.interval_range(start=self.start_time, end=self.end_time, freq=self.freq)}
