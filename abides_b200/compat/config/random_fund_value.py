"""random_fund_value: exchange + 5000 noise + 100 value agents (reference: config/random_fund_value.py)."""
from config._random_fund import run

run("random_fund_value", diverse=False)
