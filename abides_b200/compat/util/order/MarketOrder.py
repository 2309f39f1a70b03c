"""reference: util/order/MarketOrder.py"""
from util.order.Order import Order


class MarketOrder(Order):
    def __init__(self, agent_id, time_placed, symbol, quantity, is_buy_order, order_id=None, tag=None):
        super().__init__(agent_id, time_placed, symbol, quantity, is_buy_order, order_id=order_id, tag=tag)

    def __str__(self):
        side = "BUY" if self.is_buy_order else "SELL"
        return "(Agent {} @ {}) : MKT {} {} {}".format(self.agent_id, self.time_placed, side, self.quantity, self.symbol)

    __repr__ = __str__
