"""sparse_zi_100: 1 exchange + 100 ZI agents, cubic latency model (reference: config/sparse_zi_100.py)."""
from config import _sparse_zi

_sparse_zi.run([15, 15, 14, 14, 14, 14, 14], new_latency_model=True)
