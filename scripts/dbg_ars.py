import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, tds_b200
n, horizon = 64, 12
n_params = 12 * 36 + 12
g = torch.Generator().manual_seed(3)
w = (0.01 * torch.randn(n_params, generator=g)).cuda()
def mk():
    return tds_b200.laikago_sim(n)
sim = mk()
deltas = torch.zeros((n_params, sim.n_stride)); deltas[:, :n] = torch.randn((n_params, n), generator=g); deltas = deltas.cuda()
params = (w[:, None] + 0.03 * deltas).contiguous()
st = torch.cuda.current_stream()
def roll(s, p):
    tot = torch.zeros(s.n_stride, device="cuda"); steps = torch.zeros(s.n_stride, dtype=torch.int32, device="cuda")
    s.env_reset_device(seed=11, settle_steps=10, stream=st)
    q0, _ = (None, None)
    s.env_rollout_device(p, horizon, 0.0, tot, steps, stream=st)
    torch.cuda.synchronize()
    return tot[:n].cpu().numpy(), steps[:n].cpu().numpy()
a, sa = roll(sim, params)
b, sb = roll(sim, params)
sim2 = mk()
c, sc = roll(sim2, params)
print("same sim twice equal:", np.array_equal(a, b), "other sim equal:", np.array_equal(a, c), np.abs(a - c).max())
# perturb kernel vs torch
import ctypes
L = sim._L
p2 = torch.empty_like(params)
L.tds_b200_ars_perturb_device(sim._h, ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(deltas.data_ptr()), ctypes.c_float(0.03), ctypes.c_void_p(p2.data_ptr()), n_params, ctypes.c_void_p(st.cuda_stream))
torch.cuda.synchronize()
print("perturb max diff vs torch:", float((p2[:, :n] - params[:, :n]).abs().max()))
d, sd = roll(sim, p2)
print("rollout with kernel-perturbed params equal to torch params:", np.array_equal(a, d), np.abs(a - d).max())
# state after reset identical?
sim.env_reset_device(seed=11, settle_steps=10, stream=st); torch.cuda.synchronize(); qa, _ = sim.env_get_state()
sim2.env_reset_device(seed=11, settle_steps=10, stream=st); torch.cuda.synchronize(); qb, _ = sim2.env_get_state()
print("reset states equal:", np.array_equal(qa, qb))
w2 = w.clone()
rp, rn = sim.ars_train_step(w2, deltas, horizon, delta_std=0.03, step_size=0.02, shift=0.0, seed=11)
torch.cuda.synchronize()
print("ars r_pos equal to roll:", np.array_equal(rp[:n].cpu().numpy(), a), np.abs(rp[:n].cpu().numpy() - a).max())
