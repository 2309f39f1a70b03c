"""Host-side helpers with the reference's names (reference: util/util.py).  Config-time sampling stays
on the host so that a given seed yields the same population as the reference; results are passed to
the device as data."""
import numpy as np
import pandas as pd

silent_mode = False


def log_print(fmt, *args):
    if not silent_mode:
        print(fmt.format(*args))


def be_silent():
    return silent_mode


def delist(list_of_lists):
    return [x for sub in list_of_lists for x in sub]


def _real_cbrt(v):
    return -((-v) ** (1.0 / 3.0)) if v < 0 else v ** (1.0 / 3.0)


def get_wake_time(open_time, close_time, a=0, b=1):
    """One U-quadratic draw on [open_time, close_time] from the global numpy stream (util/util.py:35-57)."""
    y = np.random.rand()
    alpha = 12 / ((b - a) ** 3)
    beta = (b + a) / 2
    mult = _real_cbrt((3 / alpha) * y - (beta - a) ** 3) + beta
    return open_time + mult * (close_time - open_time)


class ExchangeStarLatency:
    """Latency description holding only what this path reads of an N x N matrix: column 0 (agent ->
    exchange) and row 0 (exchange -> agent).  Accepted by LatencyModel(min_latency=...)."""

    def __init__(self, to_exchange, from_exchange):
        self.to_exchange = np.asarray(to_exchange)
        self.from_exchange = np.asarray(from_exchange)
        self.ndim = 2
        self.shape = (len(self.to_exchange), len(self.to_exchange))


def generate_uniform_random_pairwise_dist_on_line(left, right, num_points, random_state=None):
    """Full pairwise |x_i - x_j| matrix, as the reference returns (util/util.py:106-121)."""
    x = random_state.uniform(low=left, high=right, size=num_points)
    return np.abs(x[:, None] - x[None, :])


def exchange_distances_on_line(left, right, num_points, random_state=None):
    """Same draws as generate_uniform_random_pairwise_dist_on_line, but only the distances to point 0."""
    x = random_state.uniform(low=left, high=right, size=num_points)
    d = np.abs(x - x[0])
    return ExchangeStarLatency(d, d)


def meters_to_light_ns(x):
    if isinstance(x, ExchangeStarLatency):
        return ExchangeStarLatency((x.to_exchange / 299792458e-9).astype(int),
                                   (x.from_exchange / 299792458e-9).astype(int))
    return (x / 299792458e-9).astype(int)


def validate_window_size(s):
    try:
        return int(s)
    except ValueError:
        if s.lower() == 'adaptive':
            return s.lower()
        raise ValueError(f'String {s} must be integer or string "adaptive".')


def sigmoid(x, beta):
    if x >= 0:
        return 1 / (1 + np.exp(-beta * x))
    z = np.exp(beta * x)
    return z / (1 + z)


def numeric(s):
    s = s.rstrip(',')
    for cast in (int, float):
        try:
            return cast(s)
        except ValueError:
            pass
    return s
