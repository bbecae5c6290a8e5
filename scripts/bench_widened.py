"""Device-resident throughput of the paths widened in round 2 after the GPU budget was spent - NOT YET RUN ON A GPU:
worlds of several multibodies (tds_stepw.cu) and the RigidBody path of World::step (tds_rigid.cu).
    python scripts/bench_widened.py [n_worlds]
CUDA-event timing on the launching stream, K launches captured into one graph, one untimed replay before the timed one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import tds_b200
import tds_b200.workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = 200


def timed(launch, st):
    with torch.cuda.stream(st):
        for _ in range(5):
            launch()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(K):
                launch()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay(); st.synchronize()
        ev0.record(st); g.replay(); ev1.record(st); st.synchronize()
    return ev0.elapsed_time(ev1) * 1e3 / K     # us per launch


for kind in wl.MULTIBODY_WORLDS:
    w = wl.multibody_world(kind, n)
    for prec, lab in ((tds_b200.PREC_MIXED, "mixed"), (tds_b200.PREC_F64, "f64")):
        sim = tds_b200.BatchSim(w["model"], n, precision=prec, **w["params"])
        q, qd, tau = sim.alloc(sim.n_q), sim.alloc(sim.n_qd), sim.alloc(sim.n_tau)
        q[:, :n] = torch.tensor(w["q"].T, dtype=torch.float32); qd[:, :n] = torch.tensor(w["qd"].T, dtype=torch.float32)
        tau[:, :n] = torch.tensor(w["tau"].T, dtype=torch.float32)
        q2, qd2 = sim.alloc(sim.n_q), sim.alloc(sim.n_qd)
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        us = timed(lambda: sim.step_device(tds_b200.MODE_FULL, q, qd, tau, q_out=q2, qd_out=qd2, stream=st), st)
        print("world %-16s x%d %-5s %8.2f us/step -> %.4g world-steps/s  [%s]" % (kind, n, lab, us, n / (us * 1e-6), sim.kernel_name()))

for kind in wl.RIGID_WORLDS:
    w = wl.rigid_world(kind, n)
    world = tds_b200.RigidWorld(w["bodies"], n, **w["params"])
    nb = world.n_bodies
    s = torch.zeros((13 * nb, world.n_stride), dtype=torch.float64, device="cuda")
    s[:, :n] = torch.tensor(w["state"].reshape(n, 13 * nb).T)
    s[6::13, n:] = 1.0
    f = torch.zeros((3 * nb, world.n_stride), dtype=torch.float64, device="cuda")
    f[:, :n] = torch.tensor(w["force"].reshape(n, 3 * nb).T)
    o = torch.zeros_like(s)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    us = timed(lambda: world.step_device(s, o, f, steps=1, stream=st), st)
    print("rigid %-16s x%d f64   %8.2f us/step -> %.4g world-steps/s  (%d bodies, %d solver iterations)" %
          (kind, n, us, n / (us * 1e-6), nb, int(w["params"]["num_solver_iterations"])))
