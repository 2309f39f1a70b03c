#!/usr/bin/env python
"""Golden for the log FILES the reference writes: run the UNMODIFIED reference (/root/reference) with logging on,
in a scratch directory, and store what its summary_log.bz2 and fundamental_<symbol>.bz2 hold as plain arrays
(summary Events as the repr() of their Python values: ints, floats and strings occur).

    python tests/golden/gen_log_golden.py --name zi100_s123456789 -- -c sparse_zi_100 -s 123456789

Only runs in the build container.  Compat shims: pandas.io.json.json_normalize alias, Timedelta.delta (= .value).
"""
import argparse
import contextlib
import glob
import importlib
import io
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")


def plain_repr(x):
    """repr of the Python value of a summary-log Event (ints, floats and strings occur)."""
    return repr(x.item() if hasattr(x, "item") else x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", required=True)
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--config-path", default=None,
                    help="run this config FILE (written against the reference API) instead of config.<name>")
    ap.add_argument("ref_args", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    ref_args = [a for a in args.ref_args if a != "--"]
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import numpy as np
    import pandas as pd
    import pandas.io.json
    pandas.io.json.json_normalize = pd.json_normalize
    pd.Timedelta.delta = property(lambda self: self.value)     # removed in pandas 2 (ExchangeAgent.py:309)

    work = tempfile.mkdtemp(prefix="abides_loggolden_")
    os.chdir(work)
    sys.argv = ["abides.py"] + ref_args + ["-l", "golden"]
    cfg = ref_args[ref_args.index("-c") + 1]
    with contextlib.redirect_stdout(io.StringIO()):
        if args.config_path:
            import runpy
            runpy.run_path(args.config_path if os.path.isabs(args.config_path)
                           else os.path.join(REPO, args.config_path), run_name="__abides_config__")
        else:
            importlib.import_module("config." + cfg)
    d = os.path.join(work, "log", "golden")
    s = pd.read_pickle(os.path.join(d, "summary_log.bz2"), compression="bz2")
    (ff,) = glob.glob(os.path.join(d, "fundamental_*.bz2"))
    f = pd.read_pickle(ff, compression="bz2")
    assert list(s.columns) == ["AgentID", "AgentStrategy", "EventType", "Event"]
    assert f.index.name == "FundamentalTime" and list(f.columns) == ["FundamentalValue"]
    np.savez_compressed(
        os.path.join(args.out, args.name + ".logs.npz"),
        summary_agent=s.AgentID.to_numpy(np.int64), summary_strategy=s.AgentStrategy.to_numpy(str),
        summary_type=s.EventType.to_numpy(str), summary_event=np.array([plain_repr(x) for x in s.Event], dtype=str),
        fundamental_file=os.path.basename(ff), fundamental_time=f.index.to_numpy("datetime64[ns]").astype(np.int64),
        fundamental_value=f.FundamentalValue.to_numpy(np.float64),
        files=np.array(sorted(os.listdir(d)), dtype=str))
    print(len(s), "summary rows,", len(f), "fundamental rows ->", args.name + ".logs.npz")


if __name__ == "__main__":
    main()
