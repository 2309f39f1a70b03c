"""reference: agent/FinancialAgent.py"""
from agent.Agent import Agent


class FinancialAgent(Agent):
    def __init__(self, id, name, type, random_state):
        super().__init__(id, name, type, random_state)

    def dollarize(self, cents):
        return dollarize(cents)


def dollarize(cents):
    if isinstance(cents, list):
        return [dollarize(x) for x in cents]
    if isinstance(cents, int):
        return "${:0.2f}".format(cents / 100)
    raise TypeError("dollarize(cents) called without int or list of ints: {}".format(cents))
