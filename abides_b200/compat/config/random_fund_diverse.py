"""random_fund_diverse: random_fund_value plus one MarketMakerAgent and 25 momentum agents
(reference: config/random_fund_diverse.py)."""
from config._random_fund import run

run("random_fund_diverse", diverse=True)
