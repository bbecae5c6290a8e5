"""Device-resident env-step throughput of the Ant environment (4096 envs), specialised vs table-driven kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tds_b200
n, K = 4096, 500
for kern in ("spec", "team"):
    os.environ["TDS_B200_KERNEL"] = kern
    sim = tds_b200.ant_sim(n, auto_reset=True)
    sim.env_reset_device(seed=3)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    acts = [torch.rand((8, sim.n_stride), device="cuda") * 0.8 - 0.4 for _ in range(16)]
    with torch.cuda.stream(st):
        for i in range(20):
            sim.env_step_device(acts[i % 16], stream=st)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(K):
                sim.env_step_device(acts[i % 16], stream=st)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay(); st.synchronize()
        ev0.record(st); g.replay(); ev1.record(st); st.synchronize()
    ms = ev0.elapsed_time(ev1)
    print("ant x%d, %s: %.2f us/step -> %.4g env-steps/s  [%s]" % (n, kern, ms * 1e3 / K, n * K / (ms * 1e-3), sim.kernel_name()))
