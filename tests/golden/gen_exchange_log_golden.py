#!/usr/bin/env python
"""Golden for the exchange's and the agents' log FILES: run the UNMODIFIED reference (/root/reference)
with logging on in a scratch directory and store what it wrote -- EXCHANGE_AGENT.bz2 (agent/ExchangeAgent.py:147-150,
:404-408, util/OrderBook.py:111-132), ORDERBOOK_<symbol>_FULL.bz2 (ExchangeAgent.py:319-396, OrderBook.py:138-156,
:492-520) and every per-agent log (agent/Agent.py:87-113) -- as plain arrays:

    <name>.exlog.npz   files            names of the files in the log directory
                       log_name[i], log_time[i] (ns, NaT = INT64_MIN), log_type[i], log_event[i] (canonical text)
                       log_start[k] .. log_start[k+1]: rows of file log_files[k]
                       book_file, book_time, book_quotes, book_row / book_col / book_val (non-zero cells)

    python tests/golden/gen_exchange_log_golden.py --name rmsc03_s1234_0940 -- -c rmsc03 -t ABM -d 20200603 -s 1234 --end-time 09:40:00

Only runs in the build container.  Compat shims: pandas.io.json.json_normalize alias, Timedelta.delta."""
import argparse
import contextlib
import importlib
import io
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")


def canon(x):
    """Canonical text of a logged Event: Python ints / floats / strings / None, dicts and lists of them; numpy scalars
    and bools by VALUE (the reference logs `is_buy_order` as whatever the strategy passed: bool, int or numpy int)."""
    if hasattr(x, "item") and not isinstance(x, (list, tuple, dict, str)):
        x = x.item()
    if isinstance(x, bool):
        return str(int(x))
    if isinstance(x, int):
        return str(x)
    if isinstance(x, float):
        return str(int(x)) if x == int(x) and abs(x) < 2 ** 62 else repr(x)
    if x is None:
        return "None"
    if isinstance(x, str):
        return "'" + x + "'"
    if isinstance(x, dict):
        return "{" + ", ".join("%s: %s" % (canon(k), canon(v)) for k, v in x.items()) + "}"
    if isinstance(x, (list, tuple)):
        return "[" + ", ".join(canon(v) for v in x) + "]"
    return repr(x)


def digest(log_time, log_type, log_event, book=None):
    """SHA-256 of one event log (times, types, canonical events) and of the order-book archive's cells: what a hash-only
    fixture (exlog_hashes.json) keeps of a run whose archives are too big to commit."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(log_time, np.int64).tobytes())
    h.update("\n".join(str(x) for x in log_type).encode())
    h.update("\n".join(str(x) for x in log_event).encode())
    out = {"rows": int(len(log_time)), "sha256": h.hexdigest()}
    if book is not None:
        hb = hashlib.sha256()
        for k in ("book_time", "book_quotes", "book_row", "book_col", "book_val"):
            hb.update(np.ascontiguousarray(book[k], np.int64).tobytes())
        out.update(book_shape=[int(len(book["book_time"])), int(len(book["book_quotes"]))], book_sha256=hb.hexdigest())
    return out


def hash_entry(files, log_files, log_start, log_time, log_type, log_event, book, ref_args):
    logs = {}
    for k, f in enumerate(log_files):
        lo, hi = int(log_start[k]), int(log_start[k + 1])
        logs[str(f)] = digest(log_time[lo:hi], log_type[lo:hi], log_event[lo:hi])
    out = dict(files=[str(f) for f in files], logs=logs, ref_args=list(ref_args))
    if book is not None:
        b = digest([], [], [], book)
        out.update(book_shape=b["book_shape"], book_sha256=b["book_sha256"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", required=True)
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--config-path", default=None,
                    help="run this config FILE (written against the reference API) instead of config.<name>")
    ap.add_argument("--no-book-archive", action="store_true",
                    help="compat shim: skip ExchangeAgent.logOrderBookSnapshots (a long-format archive, wide_book=False, "
                         "fails in the reference itself under current pandas: pd.SparseDataFrame, ExchangeAgent.py:384)")
    ap.add_argument("--hash-only", action="store_true",
                    help="keep only the digests of the log files and of the order-book archive, in exlog_hashes.json")
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

    if args.no_book_archive:
        import agent.ExchangeAgent as EA
        EA.ExchangeAgent.logOrderBookSnapshots = lambda self, symbol: None
    work = tempfile.mkdtemp(prefix="abides_exlog_")
    os.chdir(work)
    sys.argv = ["abides.py"] + ref_args + ["-l", "golden"]
    cfg = ref_args[ref_args.index("-c") + 1]
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        if args.config_path:
            import runpy
            runpy.run_path(args.config_path if os.path.isabs(args.config_path) else os.path.join(REPO, args.config_path),
                           run_name="__abides_config__")
        else:
            importlib.import_module("config." + cfg)
    d = os.path.join(work, "log", "golden")
    if not os.path.isdir(d):                                   # a config without a log-dir flag: the run's only directory
        (only,) = os.listdir(os.path.join(work, "log"))
        d = os.path.join(work, "log", only)
    files = sorted(os.listdir(d))
    nat = np.iinfo(np.int64).min
    log_files, log_start, times, types, events = [], [0], [], [], []
    book = {}
    for f in files:
        if f.startswith("summary_log") or f.startswith("fundamental_"):
            continue
        df = pd.read_pickle(os.path.join(d, f), compression="bz2")
        if f.startswith("ORDERBOOK_"):
            assert df.index.name == "QuoteTime"
            m = df.sparse.to_coo() if hasattr(df, "sparse") else None
            dense = np.asarray(m.todense()) if m is not None else df.to_numpy()
            r, c = np.nonzero(dense)
            book = dict(book_file=f, book_time=df.index.to_numpy("datetime64[ns]").astype(np.int64),
                        book_quotes=np.array([int(q) for q in df.columns], np.int64),
                        book_row=r.astype(np.int32), book_col=c.astype(np.int32), book_val=dense[r, c].astype(np.int64))
            continue
        assert df.index.name == "EventTime" and list(df.columns) == ["EventType", "Event"], (f, df.columns)
        t = df.index.to_numpy("datetime64[ns]").astype(np.int64)
        t[np.isnat(df.index.to_numpy("datetime64[ns]"))] = nat
        log_files.append(f)
        times.append(t)
        types += list(df.EventType)
        events += [canon(x) for x in df.Event]
        log_start.append(log_start[-1] + len(df))
    if args.hash_only:
        import json
        path = os.path.join(args.out, "exlog_hashes.json")
        table = json.load(open(path)) if os.path.exists(path) else {}
        table[args.name] = hash_entry(files, log_files, log_start, np.concatenate(times), types, events, book or None, ref_args)
        json.dump(table, open(path, "w"), indent=1, sort_keys=True)
        print(args.name, "->", path, log_start[-1], "rows in", len(log_files), "logs")
        return
    np.savez_compressed(
        os.path.join(args.out, args.name + ".exlog.npz"), files=np.array(files, dtype=str),
        log_files=np.array(log_files, dtype=str), log_start=np.array(log_start, np.int64),
        log_time=np.concatenate(times), log_type=np.array(types, dtype=str), log_event=np.array(events, dtype=str), **book)
    print(len(log_files), "logs,", log_start[-1], "rows,", len(book.get("book_time", [])), "book snapshots ->",
          args.name + ".exlog.npz")


if __name__ == "__main__":
    main()
