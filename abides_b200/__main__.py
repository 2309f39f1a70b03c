"""`python -m abides_b200 -c <config> [config flags]` -- the reference's abides.py (abides.py:18-30)
on the B200 engine: the config module body is the program."""
import argparse
import runpy
import sys

from . import runtime


def main():
    ap = argparse.ArgumentParser(description='Simulation configuration.')
    ap.add_argument('-c', '--config', required=True, help='Name of config file to execute')
    ap.add_argument('--config-help', action='store_true', help='Print argument options for the specific config file.')
    args, _ = ap.parse_known_args()
    path = runtime.find_config(args.config)
    sys.path.insert(0, runtime.COMPAT_DIR)
    runpy.run_path(path, run_name="__abides_config__")


if __name__ == "__main__":
    main()
