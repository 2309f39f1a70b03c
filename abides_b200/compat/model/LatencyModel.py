"""Descriptor of the latency model (reference: model/LatencyModel.py).  Only the agent<->exchange
row/column of `min_latency` is shipped to the device; the per-message cubic jitter draw happens there."""
import sys

import numpy as np


class LatencyModel:
    def __init__(self, latency_model='cubic', random_state=None, **kwargs):
        self.latency_model = latency_model.lower()
        self.random_state = random_state
        if 'kwargs' in kwargs:
            kwargs = kwargs['kwargs']
        if self.latency_model not in ('cubic', 'deterministic'):
            print(f"Config error: unknown latency model requested ({self.latency_model})")
            sys.exit()
        if 'min_latency' not in kwargs:
            print(f"Config error: {self.latency_model} latency model requires parameter 'min_latency' as 2-D ndarray.")
            sys.exit()
        if self.latency_model == 'cubic':
            kwargs.setdefault('connected', True)
            kwargs.setdefault('jitter', 0.5)
            kwargs.setdefault('jitter_clip', 0.1)
            kwargs.setdefault('jitter_unit', 10.0)
        self.kwargs = kwargs

    def get_latency(self, sender_id=None, recipient_id=None):
        """Host-side sampler with the reference's semantics (used by tests; the engine samples on device)."""
        kw = self.kwargs
        base = self._extract(kw['min_latency'], sender_id, recipient_id)
        if self.latency_model == 'deterministic':
            return base
        if not self._extract(kw['connected'], sender_id, recipient_id):
            return -1
        a = self._extract(kw['jitter'], sender_id, recipient_id)
        clip = self._extract(kw['jitter_clip'], sender_id, recipient_id)
        unit = self._extract(kw['jitter_unit'], sender_id, recipient_id)
        x = self.random_state.uniform(low=clip, high=1.0)
        return base + ((a / x ** 3) * (base / unit))

    @staticmethod
    def _extract(param, sid, rid):
        if np.isscalar(param):
            return param
        if hasattr(param, 'to_exchange'):
            return param.to_exchange[sid] if rid == 0 else param.from_exchange[rid]
        if isinstance(param, np.ndarray):
            return param[sid] if param.ndim == 1 else param[sid, rid]
        print("Config error: LatencyModel parameter is not scalar, 1-D ndarray, or 2-D ndarray.")
        sys.exit()
