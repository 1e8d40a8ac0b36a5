"""A short device-sampler run for ncu:
    python tools/profile_sampler.py [c1|c2] [walkers] [steps] [graph|eager]
e.g. ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python tools/profile_sampler.py c1 128 12 eager
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpbt_b200  # noqa: E402,F401
from gpbt_b200.device import DeviceChain  # noqa: E402
from gpbt_b200.sampler import DeviceEnsembleSampler  # noqa: E402
from gpbt_b200.state import EmulatorState  # noqa: E402
from tests import goldens  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
graph = (sys.argv[4] if len(sys.argv) > 4 else "eager") == "graph"
g = goldens.load(cfg + "_rbf")
sts = goldens.oracle_states(g)
states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"],
                                    s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
dc = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
rng = np.random.default_rng(5)
x0 = 0.5 * (g["lo"] + g["hi"]) + 0.25 * (g["hi"] - g["lo"]) * rng.uniform(-1, 1, (nw, len(g["lo"])))
s = DeviceEnsembleSampler(nw, x0.shape[1], dc, seed=1, use_graph=graph)
s.set_state(x0)
s.advance(steps)
print("acceptance", s.acceptance_fraction.mean())
