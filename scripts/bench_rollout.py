"""Throughput of the device-resident ARS rollout (reset + policy + step + bookkeeping, nothing leaves the GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tds_b200
n, horizon = 4096, 400
sim = tds_b200.laikago_sim(n)
rng = np.random.default_rng(0)
pol = torch.tensor(np.concatenate([0.05 * rng.standard_normal((12 * 36, sim.n_stride)), 0.05 * rng.standard_normal((12, sim.n_stride))]),
                   dtype=torch.float32, device="cuda")
tot = torch.zeros(n, device="cuda"); steps = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(st):
    sim.env_reset_device(seed=1, stream=st)
    sim.env_rollout_device(pol, 20, 0.0, tot, steps, stream=st)   # warm-up
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        sim.env_reset_device(seed=2, stream=st)
        sim.env_rollout_device(pol, horizon, 0.0, tot, steps, stream=st)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); st.synchronize()
    ev0.record(st); g.replay(); ev1.record(st); st.synchronize()
ms = ev0.elapsed_time(ev1)
print("rollout: %d envs x (10 settle + %d policy steps) in %.3f ms -> %.4g env-steps/s incl. reset, policy, bookkeeping; kernel: %s; "
      "mean steps alive %.1f, mean return %.4f" % (n, horizon, ms, n * (horizon + 10) / (ms * 1e-3), sim.kernel_name(),
                                                    steps.float().mean().item(), tot.mean().item()))
