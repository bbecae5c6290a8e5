"""Workload for ncu: Laikago x4096, settle until the toes are in contact, then a few env-steps.
ncu --set full -k regex:tds_step_kernel -s 260 -c 2 python scripts/profile_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tds_b200, tds_b200.workloads as wl
n = int(os.environ.get("TDS_ENVS", "4096"))
sim = tds_b200.laikago_sim(n, auto_reset=True)
w = wl.laikago(n)
sim.env_set_state(w["q"], w["qd"])
g = torch.Generator().manual_seed(0)
acts = [(torch.rand((12, sim.n_stride), generator=g) * 0.8 - 0.4).cuda() for _ in range(8)]
for i in range(270):
    sim.env_step_device(acts[i % 8])
torch.cuda.synchronize()
q, _ = sim.env_get_state()
print("z mean", q[:, 2].mean())
