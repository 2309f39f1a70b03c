"""VWAPExecutionAgent descriptor (reference: agent/execution/VWAPExecutionAgent.py:6-60): every `freq` bin gets
round(profile[bin start] * quantity) shares, the profile being a pickled {time: fraction} mapping or the
reference's synthetic U-shaped day (x^2 + 2x + 2 over the minutes of 09:30-16:00, normalised)."""
import pandas as pd

from agent.execution.ExecutionAgent import ExecutionAgent


class VWAPExecutionAgent(ExecutionAgent):
    def __init__(self, id, name, type, symbol, starting_cash, direction, quantity, execution_time_horizon, freq,
                 volume_profile_path, trade=True, log_orders=False, random_state=None):
        super().__init__(id, name, type, symbol, starting_cash, direction=direction, quantity=quantity,
                         execution_time_horizon=execution_time_horizon, trade=trade, log_orders=log_orders,
                         random_state=random_state)
        self.freq = freq
        self.volume_profile_path = volume_profile_path
        self.schedule = self.generate_schedule()

    def generate_schedule(self):
        if self.volume_profile_path is None:
            profile = VWAPExecutionAgent.synthetic_volume_profile(self.start_time, self.freq)
        else:
            profile = pd.read_pickle(self.volume_profile_path).to_dict()
        return {b: round(profile[b.left] * self.quantity)
                for b in pd.interval_range(start=self.start_time, end=self.end_time, freq=self.freq)}

    @staticmethod
    def synthetic_volume_profile(date, freq):
        day = pd.to_datetime(date.date())
        grid = pd.date_range(day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta('16:00:00'), freq=freq)
        half = int(len(grid) / 2)
        weights = {t: x ** 2 + 2 * x + 2 for t, x in zip(grid, range(int(-len(grid) / 2), half, 1))}
        scale = 1.0 / sum(weights.values())
        return {t: w * scale for t, w in weights.items()}
