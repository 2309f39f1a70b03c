"""CPU oracle of the ABIDES hot path -- TEST INFRASTRUCTURE ONLY (see abides_oracle.c).

Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by the
product package abides_b200.
"""
from .pyoracle import build, lib, run_instance, run_batch, replay_book, math_probe, dense_fundamental  # noqa: F401
