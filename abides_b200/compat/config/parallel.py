"""Seed sweep driver with the reference's command line (reference: config/parallel.py:9-60):

    python abides_b200/compat/config/parallel.py --config sparse_zi_100 --num_simulations 1024 --seed 1 --log_folder sweep

The reference starts one `abides.py` process per seed, `--num_parallel` at a time.  Here the same seeds (the same
draws from np.random after `np.random.seed(seed)`) become ONE batch: every seed's config body is lowered on the host,
all instances run in a single launch on the GPU, and each simulation then gets what its own process would have left
behind: the end-of-run summary lines and ./log/<log_folder>_seed_<seed>/summary_log.bz2.  Flags this driver does not
know (e.g. `-t ABM -d 20200603` for rmsc03) are passed on to the config, which the reference's driver cannot do;
`--num_parallel` is accepted and ignored (the batch is the parallelism); `--gpus N` spreads the batch over N GPUs of
the box (one worker thread per GPU taking chunks of instances from a ticket counter)."""
import argparse
import datetime as dt
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from abides_b200 import _abi, runtime  # noqa: E402


def sweep_seeds(seed, num_simulations):
    """np.random.seed(seed); np.random.randint(0, 2 ** 32, num_simulations)  (parallel.py:10, :56)."""
    np.random.seed(seed)
    return [int(s) for s in np.random.randint(0, 2 ** 32, num_simulations)]


def main(argv=None):
    ap = argparse.ArgumentParser(description='Main config to run multiple ABIDES simulations in parallel')
    ap.add_argument('--seed', type=int, default=None, help='Seed controlling the generated global seeds')
    ap.add_argument('--num_simulations', type=int, default=1, help='Total number of simulations to run')
    ap.add_argument('--num_parallel', type=int, default=None, help='(ignored: all simulations form one GPU batch)')
    ap.add_argument('--config', required=True, help='Name of config file to execute')
    ap.add_argument('--log_folder', required=True, help='Log directory name')
    ap.add_argument('--gpus', type=int, default=1, help='GPUs of this box to spread the simulations over')
    ap.add_argument('-v', '--verbose', action='store_true', help='Maximum verbosity!')
    args, passthrough = ap.parse_known_args(argv)
    t0 = dt.datetime.now()
    seeds = sweep_seeds(args.seed, args.num_simulations)
    print(f'Total number of simulation: {args.num_simulations}')
    print(f'Configuration: {args.config}')
    print(f'Global Seeds: {seeds}')
    arg_lists = [passthrough + ["-s", s, "-l", "{}_seed_{}".format(args.log_folder, s)] + (["-v"] if args.verbose else [])
                 for s in seeds]
    out, kernels = runtime.run_config_batch(args.config, arg_lists, allow_status=(_abi.ESTATE,),
                                            devices=list(range(args.gpus)) if args.gpus > 1 else None)
    for i, (seed, kernel) in enumerate(zip(seeds, kernels)):
        print(f'--- seed {seed} ---')
        if out.res["status"][i] == _abi.ESTATE:
            print("stopped: an agent reached a state in which the reference raises (ABIDES_ESTATE); no logs written")
            continue
        runtime.write_logs(kernel)
        runtime.report(kernel, out, i, per_agent=args.verbose)
    msgs = int(out.res["ttl_messages"].sum())
    print(f'Batch of {len(seeds)} simulations: {msgs} messages, device time {out.kernel_ms / 1e3:.3f} s, '
          f'{msgs / max(out.kernel_ms, 1e-9) * 1e3:.0f} messages per second')
    print(f'Total time taken to run in parallel: {dt.datetime.now() - t0}')
    return out, kernels


if __name__ == "__main__":
    main()
