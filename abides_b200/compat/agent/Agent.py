"""Agent descriptor base (reference: agent/Agent.py:8-185).  Same constructor contract; the lifecycle
callbacks and message handlers of the supported classes execute on the device, so the Python methods
that would drive a CPU simulation are not provided."""


class Agent:
    def __init__(self, id, name, type, random_state, log_to_file=True):
        if not random_state:
            raise ValueError("A valid, seeded np.random.RandomState object is required for every agent.Agent", name)
        self.id = id
        self.name = name
        self.type = type
        self.log_to_file = log_to_file
        self.random_state = random_state
        self.kernel = None
        self.currentTime = None
        self.log = []

    def logEvent(self, eventType, event='', appendSummaryLog=False):
        self.log.append({'EventTime': self.currentTime, 'EventType': eventType, 'Event': event})
        if appendSummaryLog and self.kernel is not None:
            self.kernel.appendSummaryLog(self.id, eventType, event)

    def __lt__(self, other):
        return str(self.id) < str(other.id)
