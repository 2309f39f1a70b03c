"""Descriptor of the sparse mean-reverting (OU + megashock) fundamental
(reference: util/oracle/SparseMeanRevertingOracle.py).  The constructor performs the reference's
config-time draws (first megashock, :58-72); observation / advance run on the device."""
from math import sqrt

import numpy as np
import pandas as pd


class SparseMeanRevertingOracle:
    def __init__(self, mkt_open, mkt_close, symbols):
        self.mkt_open = mkt_open
        self.mkt_close = mkt_close
        self.symbols = symbols
        self.f_log = {}
        self.r = {}
        self.megashocks = {}
        for name, s in symbols.items():
            self.r[name] = (mkt_open, s['r_bar'])
            self.f_log[name] = [{'FundamentalTime': mkt_open, 'FundamentalValue': s['r_bar']}]
            gap = np.random.exponential(scale=1.0 / s['megashock_lambda_a'])       # global stream
            when = mkt_open + pd.Timedelta(gap, unit='ns')
            size = s['random_state'].normal(loc=s['megashock_mean'], scale=sqrt(s['megashock_var']))
            if s['random_state'].randint(2) != 0:
                size = -size
            self.megashocks[name] = [{'MegashockTime': when, 'MegashockValue': size}]

    def getDailyOpenPrice(self, symbol, mkt_open=None):
        return self.symbols[symbol]['r_bar']

    def observePrice(self, symbol, currentTime, sigma_n=1000, random_state=None):
        raise RuntimeError("SparseMeanRevertingOracle.observePrice runs on the device inside Kernel.runner")
