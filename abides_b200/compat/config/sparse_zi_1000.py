"""sparse_zi_1000: 1 exchange + 1000 ZI agents, legacy latency matrix + noise (reference: config/sparse_zi_1000.py)."""
from config import _sparse_zi

_sparse_zi.run([143, 143, 143, 143, 143, 143, 142], new_latency_model=False)
