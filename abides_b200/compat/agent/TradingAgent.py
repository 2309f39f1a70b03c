"""TradingAgent descriptor (reference: agent/TradingAgent.py).  Holds the constructor state the
engine lowers (starting cash, log flags) and receives the end-of-run portfolio back from the device
(holdings, last trade, marked-to-market) so `kernelStopping`-style reporting works unchanged."""
import sys

from agent.FinancialAgent import FinancialAgent


class TradingAgent(FinancialAgent):
    def __init__(self, id, name, type, random_state=None, starting_cash=100000, log_orders=False, log_to_file=True):
        FinancialAgent.__init__(self, id, name, type, random_state)
        self.log_to_file = log_to_file
        self.mkt_open = None
        self.mkt_close = None
        self.log_orders = log_orders
        if log_orders is None:
            self.log_orders = False
            self.log_to_file = False
        self.starting_cash = starting_cash
        self.MKT_BUY = sys.maxsize
        self.MKT_SELL = 0
        self.holdings = {'CASH': starting_cash}
        self.orders = {}
        self.last_trade = {}
        self.daily_close_price = {}
        self.known_bids = {}
        self.known_asks = {}
        self.first_wake = True
        self.mkt_closed = False
        self.marked_to_market = None
        self.final_valuation = None

    def getHoldings(self, symbol):
        return self.holdings.get(symbol, 0)

    def fmtHoldings(self, holdings):
        parts = ["{}: {}".format(k, v) for k, v in sorted(holdings.items()) if k != 'CASH']
        parts.append("CASH: {}".format(holdings['CASH']))
        return '{ ' + ', '.join(parts) + ' }'
