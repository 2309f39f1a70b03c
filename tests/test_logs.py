"""Log files: summary_log.bz2 and fundamental_<symbol>.bz2 as the reference writes them
(Kernel.py:514-532, ExchangeAgent.py:91-101), checked against files the unmodified reference produced
(tests/golden/gen_log_golden.py)."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from abides_b200 import engine as eng_mod, lowering, runtime
from gen_log_golden import plain_repr
from parity_util import GOLDEN, TRACE_FUNDAMENTAL, load_golden

NAME, CONFIG, SEED = "zi100_s123456789", "sparse_zi_100", 123456789
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_files(path, name=NAME):
    want = np.load(os.path.join(GOLDEN, name + ".logs.npz"))
    s = pd.read_pickle(os.path.join(path, "summary_log.bz2"), compression="bz2")
    assert list(s.columns) == ["AgentID", "AgentStrategy", "EventType", "Event"]
    assert np.array_equal(s.AgentID.to_numpy(np.int64), want["summary_agent"])
    assert list(s.AgentStrategy) == list(want["summary_strategy"])
    assert list(s.EventType) == list(want["summary_type"])
    # numbers by value: the reference's TWAP/VWAP market order carries a float quantity (12e5 - executed), which turns
    # the cash of whoever trades against it into a Python float ('10001230.0'); the engine reports integers
    got = [plain_repr(x) for x in s.Event]
    assert len(got) == len(want["summary_event"])
    for a, b in zip(got, want["summary_event"]):
        assert a == b or (a[0] != "'" and b[0] != "'" and float(a) == float(b)), (a, b)
    f = pd.read_pickle(os.path.join(path, str(want["fundamental_file"])), compression="bz2")
    assert f.index.name == "FundamentalTime" and list(f.columns) == ["FundamentalValue"]
    assert str(f.index.dtype) == "datetime64[ns]" and f.FundamentalValue.dtype == np.float64
    assert np.array_equal(f.index.to_numpy("datetime64[ns]").astype(np.int64), want["fundamental_time"])
    assert np.allclose(f.FundamentalValue.to_numpy(), want["fundamental_value"], rtol=1e-9, atol=0)


@pytest.mark.parametrize("name", [NAME, "exec_twap_s6"])
def test_log_writers_reproduce_reference_files(tmp_path, name):
    """Host side only: results and trace come from the reference golden, the files from runtime.write_logs.
    exec_twap_s6 adds Value / Noise FINAL_VALUATION ratios and the execution agent's DIRECTION ... SLIPPAGE rows."""
    from test_oracle_golden import CASES
    g, meta = load_golden(name)
    cfg, args = CASES[name]
    ((spec, kernel),) = runtime.build_instances(cfg, [list(args) + ["-l", "x"]])
    hb = lowering.HostBatch([spec])
    res = np.zeros(1, eng_mod.INSTANCE_RESULT_DTYPE)
    for k in ("ttl_messages", "last_trade", "slowest_agent_finish_time"):
        res[k][0] = meta[k]
    ares = np.zeros(meta["n_agents"], eng_mod.AGENT_RESULT_DTYPE)
    ares["cash"], ares["holdings"], ares["marked_to_market"] = g["cash"], g["holdings"], g["mtm"]
    ares["final_valuation"] = g["final_valuation"]
    out = runtime.RunOutput(hb, res, ares, 0.0, 1.0)
    runtime.write_back(kernel, out, 0)
    n = len(g["f_time"]) - 1
    fund = np.zeros((n, 7), np.int64)
    fund[:, 0], fund[:, 1], fund[:, 2], fund[:, 3] = g["f_time"][1:], -1, TRACE_FUNDAMENTAL, g["f_value"][1:]
    trace = np.concatenate([g["trace"], fund]) if "trace" in g.files else fund
    path = runtime.write_logs(kernel, trace, root=str(tmp_path))
    assert path == os.path.join(str(tmp_path), "log", "x")
    _check_files(path, name)


@pytest.mark.gpu
def test_cli_run_writes_reference_logs(tmp_path):
    """`python -m abides_b200 -c sparse_zi_100 -s ... -l run` (abides.py) end to end on the device."""
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "abides_b200", "-c", CONFIG, "-s", str(SEED), "-l", "run"],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Simulation ending!" in r.stdout
    _check_files(os.path.join(str(tmp_path), "log", "run"))
    # ... and the exchange's log and the 100 agents' logs (tests/test_exchange_log.py)
    from test_exchange_log import check_directory
    check_directory(os.path.join(str(tmp_path), "log", "run"), NAME)


def _load_parallel():
    import importlib.util
    spec = importlib.util.spec_from_file_location("abides_parallel", os.path.join(runtime.COMPAT_DIR, "config", "parallel.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_sweep_driver_draws_the_reference_seeds():
    """config/parallel.py:10,:56: np.random.seed(seed) then np.random.randint(0, 2 ** 32, n)."""
    m = _load_parallel()
    np.random.seed(11)
    want = [int(x) for x in np.random.randint(0, 2 ** 32, 5)]
    assert m.sweep_seeds(11, 5) == want == runtime.parallel_seeds(11, 5)


@pytest.mark.gpu
def test_sweep_driver_runs_one_batch_and_writes_per_seed_logs(tmp_path, monkeypatch, capsys):
    """`parallel.py --config sparse_zi_100 --num_simulations 3 --seed 123456789` equals three single runs; the first
    global seed of that stream is not the golden's, so the check is against the single-instance path."""
    m = _load_parallel()
    monkeypatch.chdir(tmp_path)
    out, kernels = m.main(["--config", CONFIG, "--num_simulations", "3", "--seed", "7", "--log_folder", "sw"])
    seeds = m.sweep_seeds(7, 3)
    text = capsys.readouterr().out
    assert text.count("Simulation ending!") == 3
    for i, s in enumerate(seeds):
        f = os.path.join(str(tmp_path), "log", "sw_seed_%d" % s, "summary_log.bz2")
        df = pd.read_pickle(f, compression="bz2")
        assert len(df) == 400 and list(df.EventType[:1]) == ["STARTING_CASH"]
        single, (k1,) = runtime.run_config_batch(CONFIG, [["-s", s]])
        assert single.res["ttl_messages"][0] == out.res["ttl_messages"][i]
        assert [a.holdings for a in k1.agents[1:]] == [a.holdings for a in kernels[i].agents[1:]]
