"""ExchangeAgent descriptor (reference: agent/ExchangeAgent.py:26-69).  The order books it owns live
in shared memory on the device; `order_books` only records which symbols are listed."""
from agent.FinancialAgent import FinancialAgent


class _DeviceOrderBook:
    def __init__(self, symbol):
        self.symbol = symbol
        self.last_trade = None


class ExchangeAgent(FinancialAgent):
    def __init__(self, id, name, type, mkt_open, mkt_close, symbols, book_freq='S', wide_book=False,
                 pipeline_delay=40000, computation_delay=1, stream_history=0, log_orders=False, random_state=None):
        super().__init__(id, name, type, random_state)
        self.mkt_open = mkt_open
        self.mkt_close = mkt_close
        self.pipeline_delay = pipeline_delay
        self.computation_delay = computation_delay
        self.stream_history = stream_history
        self.log_orders = log_orders
        self.book_freq = book_freq
        self.wide_book = wide_book
        self.order_books = {s: _DeviceOrderBook(s) for s in symbols}
        self.subscription_dict = {}
