import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tds_b200, tds_b200.workloads as wl
from tds_b200.model import load_model, fixture_path
from oracle import port

def rel(a, r): return np.abs(a - r) / np.maximum(1, np.abs(r))

g = np.load("tests/golden/laikago.npz")
n = 64
sim = tds_b200.laikago_sim(n)
raw = sim.step_host(2, g["q_in"], g["qd_in"], g["tau"])
pd = sim.step_host(2, g["q_in"], g["qd_in"], g["action"], use_pd=True)
print("raw vs golden qd", rel(raw["qd"], g["qd_out"]).max(), "pd vs golden qd", rel(pd["qd"], g["qd_out"]).max())
print("pd vs raw max abs", np.abs(pd["qd"] - raw["qd"]).max(), "nan?", np.isnan(pd["qd"]).any())
bad = np.argwhere(rel(pd["qd"], g["qd_out"]) > 1e-5)
print("bad entries (env, dof):", bad[:20].tolist())
# FD only comparison of qdd with pd vs raw
a = sim.step_host(0, g["q_in"], g["qd_in"], g["tau"])
b = sim.step_host(0, g["q_in"], g["qd_in"], g["action"], use_pd=True)
print("FD qdd pd-vs-raw max abs", np.abs(a["qdd"] - b["qdd"]).max())
# env path
sim.env_set_state(g["q_in"], g["qd_in"])
obs = np.zeros((n, 36), dtype=np.float32); rew = np.zeros(n, dtype=np.float32); done = np.zeros(n, dtype=np.float32)
sim.env_step_host(g["action"].astype(np.float32), obs, rew, done)
ref = g["env_output_templated"]
print("env obs err", rel(obs.astype(np.float64), ref[:, :36]).max(), "done eq", np.array_equal(done, g["env_done"]), "rew err", np.abs(rew - g["env_reward"]).max())
print("done", done[:10], g["env_done"][:10], "rew", rew[:4], g["env_reward"][:4])
# v1
m = tds_b200.CudaModelV1()
m.allocate(n)
out = np.full((n, 411), 123.0)
m.forward_zero(g["env_input"], out)
print("v1 nan count", np.isnan(out).sum(), "first row", out[0, :8], "ref", ref[0, :8])
print("v1 q/qd err", np.nanmax(rel(out[:, :36], ref[:, :36])), "vis err", np.nanmax(np.abs(out[:, 36:156] - ref[:, 36:156])))
m.deallocate()
# humanoid in f64
for prec in (0, 1):
    gh = np.load("tests/golden/humanoid.npz")
    mh = load_model(fixture_path("humanoid"))
    s = tds_b200.BatchSim(mh, 64, friction=1.0, keep_all_points=False, precision=prec)
    o = s.step_host(2, gh["q_in"], gh["qd_in"], gh["tau"], want_contacts=True)
    e = rel(o["qd"], gh["qd_out"]).max(axis=1)
    pen = np.array([(d < 0).sum() for d in gh["contact_dist"]])
    print("humanoid prec", prec, "err by env (pen count, err):", [(int(p), float(f"{x:.1e}")) for p, x in zip(pen, e)][:64:4])
# pendulum f64
gp = np.load("tests/golden/pendulum5.npz")
s = tds_b200.BatchSim(load_model(fixture_path("pendulum5")), 64, precision=1)
o = s.step_host(0, gp["q_in"], gp["qd_in"], gp["tau"])
print("pendulum f64 qdd err", rel(o["qdd"], gp["qdd"]).max())
# fresh laikago
w = wl.laikago_perturbed(512, seed=777)
s = tds_b200.laikago_sim(512)
o = s.step_host(2, w["q"], w["qd"], w["action"], use_pd=True)
P = port.make_params(friction=1.0, keep_all_points=True)
x = np.zeros((512, 51)); x[:, :18], x[:, 18:36], x[:, 36:48], x[:, 48:] = w["q"], w["qd"], w["action"], [100.0, 2.0, 50.0]
r = port.locomotion_step(s.model, P, tds_b200.envs.LAIKAGO_INITIAL_POSES, 6, x, 411)
eq = rel(o["q"], r[:, :18]); eqd = rel(o["qd"], r[:, 18:36])
print("fresh laikago q err", eq.max(), "qd err", eqd.max(), "n envs over tol", (eqd.max(axis=1) > 1e-5).sum(), "worst env", eqd.max(axis=1).argmax())
