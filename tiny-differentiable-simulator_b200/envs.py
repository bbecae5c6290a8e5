"""Environment-level mirrors of the reference's vectorized Laikago environment and of the dlopen'd
"C-ABI v1" model library.

VectorizedLaikagoEnv follows pytinydiffsim.VectorizedLaikagoEnv
(python/pytinydiffsim_includes.h:153-227, bound at python/pytinydiffsim.inl:1155-1175) which wraps
VectorizedEnvironment<Algebra, LaikagoContactSimulation>::{reset,step}
(examples/ars/ars_vectorized_environment.h:163-291).
"""
import ctypes

import numpy as np

from . import _lib
from .model import fixture_path, load_model
from .sim import BatchSim, MODE_FULL

LAIKAGO_INITIAL_POSES = np.array([0.2, 0.0, -0.7] * 4)   # laikago_environment2.h:50-61
LAIKAGO_KP, LAIKAGO_KD, LAIKAGO_MAX_FORCE = 100.0, 2.0, 50.0  # laikago_environment2.h:43-45
LAIKAGO_START_Z = 0.48


def laikago_reset_pose():
    q = np.zeros(18)
    q[2] = LAIKAGO_START_Z
    q[6:18] = LAIKAGO_INITIAL_POSES
    return q


def laikago_sim(n_envs, device=0, model=None, auto_reset=False, **kw):
    """BatchSim configured like LaikagoContactSimulation (fixed-base emulation, friction 1,
    keep_all_points, dt 1e-3; locomotion_contact_simulation.h:131-135, laikago_environment2.h:36-47)."""
    if model is None:
        model = load_model(fixture_path("laikago"))
    sim = BatchSim(model, n_envs, device=device, dt=1e-3, friction=1.0, keep_all_points=True, **kw)
    sim.set_env(LAIKAGO_INITIAL_POSES, start_link=6, kp=LAIKAGO_KP, kd=LAIKAGO_KD, max_force=LAIKAGO_MAX_FORCE,
                action_limit=0.4, reward_kind=1)
    sim.set_auto_reset(bool(auto_reset), laikago_reset_pose())   # the pose is also what env_reset_device starts from
    return sim


class VectorizedLaikagoEnvOutput:
    def __init__(self, obs, rewards, dones, visual_world_transforms=None):
        self.obs = obs
        self.rewards = rewards
        self.dones = dones
        self.visual_world_transforms = visual_world_transforms


class VectorizedLaikagoEnv:
    def __init__(self, num_envs, auto_reset_when_done=True, device=0, seed=12345, model=None, with_visual_transforms=False):
        """with_visual_transforms: also return the reference's `visual_world_transforms` rows from step() (off by default: the
        plain step moves 200 B per environment over PCIe, the rows 1.6 KB)."""
        self.num_envs = num_envs
        self.auto_reset = auto_reset_when_done
        self.with_visual_transforms = bool(with_visual_transforms)
        self._output_dim = 411     # cuda_model_laikago_forward_zero_meta().output_dim
        self.sim = laikago_sim(num_envs, device=device, model=model)
        self.rng = np.random.default_rng(seed)
        self._obs = np.zeros((num_envs, self.obs_dim()), dtype=np.float32)
        self._rew = np.zeros(num_envs, dtype=np.float32)
        self._done = np.zeros(num_envs, dtype=np.float32)
        self._was_done = np.zeros(num_envs, dtype=bool)

    def action_dim(self):
        return 12

    def obs_dim(self):
        return self.sim.n_q + self.sim.n_qd

    def urdf_filename(self):
        return "laikago/laikago_toes_zup_xyz_xyzrot.urdf"

    def _initial_state(self, n):
        # LaikagoContactSimulation::reset, laikago_environment2.h:63-116 (fixed-base branch)
        q = np.zeros((n, self.sim.n_q))
        q[:, 2] = LAIKAGO_START_Z
        q[:, 6:18] = LAIKAGO_INITIAL_POSES + 0.05 * (self.rng.random((n, 12)) - 0.5) * 2.0
        return q, np.zeros((n, self.sim.n_qd))

    def _settle(self, steps=10):
        zero = np.zeros((self.num_envs, self.action_dim()), dtype=np.float32)
        for _ in range(steps):
            self.sim.env_step_host(zero, self._obs, self._rew, self._done)

    def reset(self):
        q, qd = self._initial_state(self.num_envs)
        self.sim.env_set_state(q, qd)
        self._settle()
        self._was_done = np.zeros(self.num_envs, dtype=bool)
        return self._obs.copy()

    def rollouts(self, policies, rollout_length, shift=0.0, noise=None):
        """ARSVectorizedWorker::rollouts (examples/ars/ars_vectorized_worker.h:51-141) on the device: reset with joint
        noise (given [n][12], else drawn from this env's generator), 10 settle steps, then rollout_length steps of the
        per-environment linear policies [n][12*36 + 12].  Returns (total_rewards, steps)."""
        if noise is None:
            noise = 0.05 * (self.rng.random((self.num_envs, self.action_dim())) - 0.5) * 2.0
        return self.sim.env_rollout_host(policies, rollout_length, shift=shift, noise=noise)

    def output_dim(self):
        """LocomotionContactSimulation::output_dim(): q | qd | one (pos3, quat4) record per link x visual slot | up.z
        (locomotion_contact_simulation.h:75-77; 411 for Laikago, 155 for Ant - only the first n_visuals records are written)."""
        return self._output_dim

    def _step_with_visuals(self, a):
        """The step through tds_b200_env_step_visual_device: same physics, plus the per-visual world transforms of the step
        (locomotion_contact_simulation.h:281-299) - returned as the reference's `visual_world_transforms` rows
        q | qd | n_visuals x (pos3, quat xyzw) | up.z (= 1: fixed-base emulation, :131) | unwritten tail (zeros)."""
        import torch
        sim, n = self.sim, self.num_envs
        nv = sim.num_visuals()
        if not hasattr(self, "_vis"):
            dev = f"cuda:{sim.device}"
            self._vis = dict(act=sim.alloc(self.action_dim()), rew=sim.alloc(1), done=sim.alloc(1),
                             pos=torch.zeros((n * nv, 4), device=dev), quat=torch.zeros((n * nv, 4), device=dev))
        v = self._vis
        v["act"][:, :n] = torch.from_numpy(np.ascontiguousarray(a.T)).to(v["act"].device)
        torch.cuda.synchronize(v["act"].device)
        sim.env_step_visual_device(v["act"], v["pos"], v["quat"], reward=v["rew"], done=v["done"])
        torch.cuda.synchronize(v["act"].device)
        q, qd = sim.env_get_state()
        self._obs[:, :sim.n_q], self._obs[:, sim.n_q:] = q, qd
        self._rew[:] = v["rew"][0, :n].cpu().numpy()
        self._done[:] = v["done"][0, :n].cpu().numpy()
        out = np.zeros((n, self._output_dim), dtype=np.float32)
        nq = sim.n_q + sim.n_qd
        out[:, :nq] = self._obs
        rec = np.concatenate([v["pos"].cpu().numpy().reshape(n, nv, 4)[:, :, :3], v["quat"].cpu().numpy().reshape(n, nv, 4)], axis=2)
        out[:, nq:nq + nv * 7] = rec.reshape(n, nv * 7)
        out[:, nq + nv * 7] = 1.0
        return out

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        assert a.shape == (self.num_envs, self.action_dim())
        visuals = None
        if self.with_visual_transforms:
            visuals = self._step_with_visuals(a)
        else:
            self.sim.env_step_host(a, self._obs, self._rew, self._done)
        obs = self._obs.copy()
        rewards, dones = self._rew.copy(), self._done.copy()
        if self.auto_reset and dones.any():
            # ars_vectorized_environment.h:261-272: contact_sim.reset() of the finished environments = reset pose + joint noise,
            # zero velocities, then 10 settle steps (laikago_environment2.h:63-116); the observation is the settled state.
            # tds_b200_env_reset_device does exactly that under a mask, the others are left untouched.
            import torch
            idx = np.nonzero(dones)[0]
            dev = f"cuda:{self.sim.device}"
            noise = np.zeros((self.action_dim(), self.sim.n_stride), dtype=np.float32)
            noise[:, idx] = (0.05 * (self.rng.random((idx.size, self.action_dim())) - 0.5) * 2.0).T
            mask = torch.from_numpy(np.ascontiguousarray(dones, dtype=np.float32)).to(dev)
            noise_t = torch.from_numpy(noise).to(dev)
            torch.cuda.synchronize(dev)     # produced on torch's stream, consumed on the simulator's own
            self.sim.env_reset_device(mask=mask, noise=noise_t, settle_steps=10)
            q, qd = self.sim.env_get_state()
            obs[idx, :self.sim.n_q] = q[idx]
            obs[idx, self.sim.n_q:] = qd[idx]
        elif not self.auto_reset:
            # :252-277: an environment that was already done keeps stepping but reports reward 0 and stays done
            rewards[self._was_done] = 0.0
            dones[self._was_done] = 1.0
            self._was_done = dones > 0
        obs[:, 0] = 0.0  # ars_vectorized_environment.h:285-287
        obs[:, 1] = 0.0
        return VectorizedLaikagoEnvOutput(obs, rewards, dones, visuals)


ANT_INITIAL_POSES = np.array([0.0, -0.5] * 4)        # ant_environment2.h:43-54
ANT_KP, ANT_KD, ANT_MAX_FORCE, ANT_DT = 15.0, 0.3, 3.0, 0.01   # ant_environment2.h:62-66
ANT_START_Z = 0.48


def ant_reset_pose():
    q = np.zeros(14)
    q[2] = ANT_START_Z
    q[6:14] = ANT_INITIAL_POSES
    return q


def ant_sim(n_envs, device=0, model=None, auto_reset=False, **kw):
    """BatchSim configured like AntContactSimulation2 (gym/ant_org_xyz_xyzrot.urdf, fixed-base emulation, dt 0.01,
    friction 1, keep_all_points; ant_environment2.h:28-70): reward = forward velocity, done = torso below 0.26."""
    if model is None:
        model = load_model(fixture_path("ant"))
    sim = BatchSim(model, n_envs, device=device, dt=ANT_DT, friction=1.0, keep_all_points=True, **kw)
    sim.set_env(ANT_INITIAL_POSES, start_link=6, kp=ANT_KP, kd=ANT_KD, max_force=ANT_MAX_FORCE, action_limit=0.4, reward_kind=3)
    sim.set_auto_reset(bool(auto_reset), ant_reset_pose())
    return sim


class VectorizedAntEnv(VectorizedLaikagoEnv):
    """pytinydiffsim.VectorizedAntEnv (python/pytinydiffsim_includes.h:58-141): same vectorized environment template
    as the Laikago one, over AntContactSimulation2 (8 actions, 28 observations)."""

    def __init__(self, num_envs, auto_reset_when_done=True, device=0, seed=12345, model=None, with_visual_transforms=False):
        self.num_envs = num_envs
        self.auto_reset = auto_reset_when_done
        self.with_visual_transforms = bool(with_visual_transforms)
        self._output_dim = 155     # cuda_model_ant_forward_zero_meta().output_dim
        self.sim = ant_sim(num_envs, device=device, model=model)
        self.rng = np.random.default_rng(seed)
        self._obs = np.zeros((num_envs, self.obs_dim()), dtype=np.float32)
        self._rew = np.zeros(num_envs, dtype=np.float32)
        self._done = np.zeros(num_envs, dtype=np.float32)
        self._was_done = np.zeros(num_envs, dtype=bool)

    def action_dim(self):
        return 8

    def urdf_filename(self):
        return "gym/ant_org_xyz_xyzrot.urdf"

    def _initial_state(self, n):
        q = np.zeros((n, self.sim.n_q))
        q[:, 2] = ANT_START_Z
        q[:, 6:14] = ANT_INITIAL_POSES + 0.05 * (self.rng.random((n, 8)) - 0.5) * 2.0   # ant_environment2.h:124-134
        return q, np.zeros((n, self.sim.n_qd))


class CudaModelV1:
    """What tds::CudaModel / ars_train_policy_cuda's loader do with the dlopen'd library
    (examples/ars/ars_train_policy_cuda.cpp:183-308): meta, allocate, forward_zero, deallocate."""

    def __init__(self, model_name="cuda_model_laikago"):
        self._L = _lib.lib()
        self._name = model_name
        self._fz = getattr(self._L, model_name + "_forward_zero")
        self._meta = getattr(self._L, model_name + "_forward_zero_meta")
        self._alloc = getattr(self._L, model_name + "_forward_zero_allocate")
        self._dealloc = getattr(self._L, model_name + "_forward_zero_deallocate")
        m = self._meta()
        self.input_dim, self.output_dim, self.global_dim = m.input_dim, m.output_dim, m.global_dim
        self._n = 0

    def allocate(self, num_total_threads):
        self._alloc(int(num_total_threads))
        self._n = num_total_threads

    def deallocate(self):
        self._dealloc()
        self._n = 0

    def forward_zero(self, inputs, outputs=None, num_threads_per_block=64):
        x = np.ascontiguousarray(inputs, dtype=np.float64)
        n = x.shape[0]
        assert x.shape[1] == self.input_dim and n <= self._n
        if outputs is None:
            outputs = np.zeros((n, self.output_dim))
        dp = ctypes.POINTER(ctypes.c_double)
        blocks = (n + num_threads_per_block - 1) // num_threads_per_block
        self._fz(n, blocks, num_threads_per_block, outputs.ctypes.data_as(dp), x.ctypes.data_as(dp))
        return outputs
