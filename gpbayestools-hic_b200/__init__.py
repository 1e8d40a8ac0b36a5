"""gpbt_b200: B200 (sm_100a) drop-in for GPBayesTools-HIC's batched emulator / log-posterior path.

Public surface mirrors the reference modules (paths under /root/reference):
    gpbt_b200.emulator.Emulator          <- src/emulator.py Emulator
    gpbt_b200.mcmc.Chain, mvn_loglike    <- src/mcmc.py
    gpbt_b200.workdir, parse_model_parameter_file  <- src/__init__.py
The arithmetic runs in hand-written CUDA behind the C ABI in include/gpbt.h; there is no CPU path.
"""
import os
from pathlib import Path

__version__ = "0.1.0"

# env WORKDIR as in the reference (src/__init__.py:15); nothing is created at import time
workdir = Path(os.getenv("WORKDIR", "."))


def parse_model_parameter_file(parfile):
    """`name: label, min, max` lines, `#` starts a comment (src/__init__.py:21-32).
    Returns {name: [label, min, max]} in file order."""
    table = {}
    with open(parfile, "r") as fh:
        for raw in fh:
            body = raw.split("#", 1)[0]
            if not body.strip():
                continue
            name, _, rest = body.partition(":")
            fields = [f.strip() for f in rest.split(",")]
            table[name] = [fields[0], float(fields[1]), float(fields[2])] + fields[3:]
    return table
