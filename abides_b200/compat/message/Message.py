"""Message envelope kept for API compatibility (reference: message/Message.py).  On the B200 engine
messages are 48-byte device records; these classes only exist so user code can still construct them."""
from enum import Enum


class MessageType(Enum):
    MESSAGE = 1
    WAKEUP = 2

    def __lt__(self, other):
        return self.value < other.value


class Message:
    uniq = 0

    def __init__(self, body=None):
        self.body = body
        self.uniq = Message.uniq
        Message.uniq += 1

    def __lt__(self, other):
        return self.uniq < other.uniq

    def __str__(self):
        return str(self.body)
