import sys

from . import run

if len(sys.argv) < 2:
    sys.exit("usage: python -m ouzelum_b200.compat <reference_script.py> [script args...]")
run(sys.argv[1], sys.argv[2:])
