"""Small run for compute-sanitizer (racecheck / memcheck): 64 Laikago environments, a few full steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tds_b200, tds_b200.workloads as wl
n = 64
sim = tds_b200.laikago_sim(n)
w = wl.laikago(n)
sim.env_set_state(w["q"], w["qd"])
act = sim.alloc(12)
for _ in range(3):
    sim.env_step_device(act)
q, qd = sim.env_get_state() if hasattr(sim, "env_get_state") else (None, None)
print("ok", None if q is None else float(np.abs(q).max()))
