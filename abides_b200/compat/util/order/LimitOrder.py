"""reference: util/order/LimitOrder.py"""
from util.order.Order import Order

silent_mode = False


class LimitOrder(Order):
    def __init__(self, agent_id, time_placed, symbol, quantity, is_buy_order, limit_price, order_id=None, tag=None):
        super().__init__(agent_id, time_placed, symbol, quantity, is_buy_order, order_id, tag=tag)
        self.limit_price = limit_price

    def __str__(self):
        if silent_mode:
            return ''
        side = "BUY" if self.is_buy_order else "SELL"
        return "(Agent {} @ {}) : {} {} {} @ {}".format(self.agent_id, self.time_placed, side, self.quantity,
                                                       self.symbol, self.limit_price)

    __repr__ = __str__
