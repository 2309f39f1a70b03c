"""MarketReplayAgent descriptor (reference: agent/examples/MarketReplayAgent.py:12-125).  The agent replays a historical
L3 order file: it wakes at every timestamp of the file and turns each row into a limit order (unseen ORDER_ID, SIZE > 0),
a cancellation (known ORDER_ID, SIZE 0) or a modification (known ORDER_ID, another SIZE).  On the device it is agent class
ABIDES_MARKET_REPLAY (lane engine); `historical_orders` keeps the reference's attribute names (`orders_dict`,
`wakeup_times`, `first_wakeup`) -- that is what the lowering reads.

The file is read by `abides_b200.orderfile` ('|' or ',' separated, BUY / SELL or 0 / 1 flags): the reference's
L3OrdersProcessor (:72-125) drops the first data row (`.iloc[1:]`), needs '|' and 0 / 1 flags and fails on the comma-
separated sample it ships (data/sample_orders_file.csv); as there, PRICE is dollars and becomes integer cents (:113-114).
A processed-orders pickle `<folder>marketreplay_<symbol>_<date>.pkl`, if present, is used instead, exactly as the
reference does (:89-92)."""
import os
import pickle

import pandas as pd

from agent.TradingAgent import TradingAgent


class MarketReplayAgent(TradingAgent):
    def __init__(self, id, name, type, symbol, date, start_time, end_time, orders_file_path, processed_orders_folder_path,
                 starting_cash, log_orders=False, random_state=None):
        super().__init__(id, name, type, starting_cash=starting_cash, log_orders=log_orders, random_state=random_state)
        self.symbol = symbol
        self.date = date
        self.log_orders = log_orders
        self.executed_trades = dict()
        self.state = 'AWAITING_WAKEUP'
        self.historical_orders = L3OrdersProcessor(self.symbol, self.date, start_time, end_time, orders_file_path,
                                                   processed_orders_folder_path)
        self.wakeup_times = self.historical_orders.wakeup_times


class L3OrdersProcessor:
    COLUMNS = ['TIMESTAMP', 'ORDER_ID', 'PRICE', 'SIZE', 'BUY_SELL_FLAG']

    def __init__(self, symbol, date, start_time, end_time, orders_file_path, processed_orders_folder_path):
        self.symbol, self.date = symbol, date
        self.start_time, self.end_time = start_time, end_time
        self.orders_file_path = orders_file_path
        self.processed_orders_folder_path = processed_orders_folder_path
        self.orders_dict = self.processOrders()
        self.wakeup_times = [*self.orders_dict]
        self.first_wakeup = self.wakeup_times[0]          # IndexError on an empty tape, as in the reference (:84)

    def processOrders(self):
        from abides_b200 import orderfile
        processed = '{}marketreplay_{}_{}.pkl'.format(self.processed_orders_folder_path, self.symbol,
                                                      pd.Timestamp(self.date).date())
        if os.path.isfile(processed):
            with open(processed, 'rb') as handle:
                return pickle.load(handle)
        rows = orderfile.read_rows(self.orders_file_path, price_scale=100, start_ns=pd.Timestamp(self.start_time).value,
                                   end_ns=pd.Timestamp(self.end_time).value)
        out = {}
        for t, oid, price, size, buy in rows:               # groupby(level=0): ascending time, file order inside a group
            out.setdefault(pd.Timestamp(t), []).append({'ORDER_ID': oid, 'PRICE': price, 'SIZE': size,
                                                        'BUY_SELL_FLAG': 'BUY' if buy else 'SELL'})
        return dict(sorted(out.items()))
