"""Order value objects (reference: util/order/Order.py).  Device orders are 16-byte records; these
classes mirror the fields for user code and logs."""


class Order:
    order_id = 0

    def __init__(self, agent_id, time_placed, symbol, quantity, is_buy_order, order_id=None, tag=None):
        self.agent_id = agent_id
        self.time_placed = time_placed
        self.symbol = symbol
        self.quantity = quantity
        self.is_buy_order = is_buy_order
        if order_id:
            self.order_id = order_id
        else:
            self.order_id = Order.order_id
            Order.order_id += 1
        self.fill_price = None
        self.tag = tag

    def to_dict(self):
        d = dict(self.__dict__)
        d['time_placed'] = self.time_placed.isoformat() if hasattr(self.time_placed, 'isoformat') else self.time_placed
        return d
