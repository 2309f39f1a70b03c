"""MeanRevertingOracle (reference: util/oracle/MeanRevertingOracle.py): the dense OU fundamental, one value per
nanosecond step between mkt_open and mkt_close, generated at construction from the process-global np.random stream.
The series is computed on the GPU (abides_dense_fundamental, one thread per symbol) and the global stream is left
where the reference's own draw would have left it, so everything constructed afterwards draws the same values."""
from math import sqrt

import numpy as np
import pandas as pd


class MeanRevertingOracle:

    def __init__(self, mkt_open, mkt_close, symbols):
        from abides_b200 import runtime
        self.mkt_open, self.mkt_close, self.symbols = mkt_open, mkt_close, symbols
        self.r = {}
        n_steps = int(pd.Timestamp(mkt_close).value - pd.Timestamp(mkt_open).value)
        if n_steps <= 0:
            raise ValueError("mkt_close must be after mkt_open")
        index = pd.date_range(mkt_open, periods=n_steps, freq='ns')
        for name, s in symbols.items():                       # the reference draws symbol after symbol from one stream
            series, (state,) = runtime.engine().dense_fundamental([np.random.get_state()], s['r_bar'], s['kappa'],
                                                                 s['sigma_s'], n_steps)
            np.random.set_state(state)
            self.r[name] = pd.Series(series[0], index=index)

    def getDailyOpenPrice(self, symbol, mkt_open=None):
        if mkt_open is not None and self.mkt_open is None:
            self.mkt_open = mkt_open
        return self.r[symbol].loc[self.mkt_open]

    def observePrice(self, symbol, currentTime, sigma_n=1000, random_state=None):
        if currentTime >= self.mkt_close:
            r_t = self.r[symbol].loc[self.mkt_close - pd.Timedelta('1ns')]
        else:
            r_t = self.r[symbol].loc[currentTime]
        if sigma_n == 0:
            return r_t
        return int(round(random_state.normal(loc=r_t, scale=sqrt(sigma_n))))
