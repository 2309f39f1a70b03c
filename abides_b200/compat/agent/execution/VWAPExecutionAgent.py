"""VWAPExecutionAgent descriptor (reference: agent/execution/VWAPExecutionAgent.py:6-60).

Each `freq` bin of the horizon is assigned round(profile[bin start] * quantity) shares.  The profile is either a
pickled {time: fraction of the day's volume} mapping or, with `volume_profile_path=None`, the reference's synthetic
U-shaped day: weight x^2 + 2x + 2 for x = -n/2 .. n/2 - 1 over the n grid points of 09:30-16:00, normalised to 1
(for an odd n the last grid point gets no weight, as in the reference, whose zip() stops one short)."""
import numpy as np
import pandas as pd

from agent.execution.ExecutionAgent import ExecutionAgent

_SESSION = (pd.to_timedelta('09:30:00'), pd.to_timedelta('16:00:00'))


def _u_shaped_day(day_start, freq):
    grid = pd.date_range(day_start + _SESSION[0], day_start + _SESSION[1], freq=freq)
    half = int(len(grid) / 2)
    xs = np.arange(int(-len(grid) / 2), half, dtype=np.int64)
    weights = [int(w) for w in xs * xs + 2 * xs + 2]
    unit = 1.0 / sum(weights)
    return {when: w * unit for when, w in zip(grid, weights)}


class VWAPExecutionAgent(ExecutionAgent):
    def __init__(self, id, name, type, symbol, starting_cash, direction, quantity, execution_time_horizon, freq,
                 volume_profile_path, trade=True, log_orders=False, random_state=None):
        ExecutionAgent.__init__(self, id, name, type, symbol, starting_cash, direction=direction, quantity=quantity,
                                execution_time_horizon=execution_time_horizon, trade=trade, log_orders=log_orders,
                                random_state=random_state)
        self.freq, self.volume_profile_path = freq, volume_profile_path
        self.schedule = self.generate_schedule()

    def generate_schedule(self):
        path = self.volume_profile_path
        share = pd.read_pickle(path).to_dict() if path is not None else self.synthetic_volume_profile(self.start_time, self.freq)
        bins = pd.interval_range(start=self.start_time, end=self.end_time, freq=self.freq)
        return {b: round(share[b.left] * self.quantity) for b in bins}

    @staticmethod
    def synthetic_volume_profile(date, freq):
        return _u_shaped_day(pd.to_datetime(date.date()), freq)
