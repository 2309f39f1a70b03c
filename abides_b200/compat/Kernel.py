"""Kernel with the reference's constructor and `runner` signature (reference: Kernel.py:12, :51-55).

`runner` does not iterate a Python priority queue: it lowers the agents, oracle and latency model it
was handed to the C-ABI tables, runs the event loop on the GPU (abides_b200.engine.Engine -- there is
no CPU path), writes the end-of-run portfolios back into the agent objects and prints the reference's
summary lines.  Under `abides_b200.runtime.capture()` it only records the lowered instance, which is
how seed sweeps are batched into one launch."""
import sys

import numpy as np
import pandas as pd

from abides_b200 import lowering, runtime


class Kernel:
    def __init__(self, kernel_name, random_state=None):
        self.name = kernel_name
        self.random_state = random_state
        if not random_state:
            raise ValueError("A valid, seeded np.random.RandomState object is required for the Kernel", self.name)
        self.currentTime = None
        self.kernelWallClockStart = pd.Timestamp('now')
        self.meanResultByAgentType = {}
        self.agentCountByType = {}
        self.summaryLog = []
        self.custom_state = {}

    def runner(self, agents=[], startTime=None, stopTime=None, num_simulations=1, defaultComputationDelay=1,
               defaultLatency=1, agentLatency=None, latencyNoise=[1.0], agentLatencyModel=None, skip_log=False,
               seed=None, oracle=None, log_dir=None):
        self.agents = agents
        self.startTime, self.stopTime = startTime, stopTime
        self.seed, self.skip_log, self.oracle = seed, skip_log, oracle
        self.log_dir = log_dir if log_dir else str(int(self.kernelWallClockStart.timestamp()))
        if num_simulations != 1:
            raise NotImplementedError("num_simulations > 1: batch instances with abides_b200.runtime instead")
        spec = lowering.lower_instance(
            agents, startTime, stopTime, defaultComputationDelay=defaultComputationDelay,
            defaultLatency=defaultLatency, agentLatency=agentLatency, latencyNoise=latencyNoise,
            agentLatencyModel=agentLatencyModel, oracle=oracle, kernel_random_state=self.random_state,
            global_state=np.random.get_state(), seed=seed or 0)
        for a in agents:
            a.kernel = self
        if runtime.capturing():
            runtime.captured(spec, self)
            return self.custom_state
        trace = None
        if skip_log:
            out = runtime.run_specs([spec])
        else:
            out, trace = runtime.run_spec_logged(spec)
        runtime.write_back(self, out, 0)
        runtime.write_logs(self, trace, exchange_log=trace is not None, spec=spec)
        runtime.report(self, out, 0)
        return self.custom_state

    # services some user configs call on the kernel object
    def appendSummaryLog(self, sender, eventType, event):
        self.summaryLog.append({'AgentID': sender, 'AgentStrategy': self.agents[sender].type,
                                'EventType': eventType, 'Event': event})

    def findAgentByType(self, type=None):
        for agent in self.agents:
            if isinstance(agent, type):
                return agent.id

    @staticmethod
    def fmtTime(simulationTime):
        return simulationTime
