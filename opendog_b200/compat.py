"""Drop-in environment objects with the reference's own surfaces, backed by the batched CUDA simulator.

* `BatchedQuadrupedEnv`  — `QuadrupedEnv` of sim2real/train.py:151-411 for `num_envs` environments on CUDA tensors.
* `QuadrupedEnv`         — the same class name / constructor / numpy 4-tuple API for ONE environment, so
                           `sim2real/train.py` (and its JSON exporter, :600-636) runs unmodified with
                           `from opendog_b200.compat import QuadrupedEnv`.
* `WalkVecEnv`           — `ScaleActionWrapper(WalkEnvironmentV0)` behind the stable-baselines3 `VecEnv` protocol
                           that `train/train.py:63-87,117-130` consumes (numpy in / numpy out, list of info dicts,
                           auto-reset with `terminal_observation`), replacing `SubprocVecEnv`.
Everything computes on the GPU; without the CUDA library or a device these raise (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .env import BatchedWalkEnv, _ptr

REASONS = ("max_steps", "mj_error", "orientation_limit", "too_much_backward")        # sim2real/train.py:393-402
ACTUATOR_NAMES_ORDERED = ["FR_tigh_actuator", "FR_knee_actuator", "FL_tigh_actuator", "FL_knee_actuator",
                          "BR_tigh_actuator", "BR_knee_actuator", "BL_tigh_actuator", "BL_knee_actuator"]


class BatchedQuadrupedEnv:
    state_dim = 3 + 8 + 8 + 1 + 2          # sim2real/train.py:164
    action_dim = 4

    def __init__(self, num_envs: int, model: str = "our_robot", device=None, auto_reset: bool = True, max_steps: int = 250,
                 **sim_config):
        # POLICY_DECISION_DT / timestep = 0.10 / 0.002 = 50 mj_step per policy step (sim2real/train.py:156)
        self.sim = BatchedWalkEnv(num_envs, model=model, device=device, info_keys=None, frame_skip=50, scale_actions=0,
                                  auto_reset=0, **sim_config)
        self.L, self.device, self.num_envs = self.sim.L, self.sim.device, self.sim.num_envs
        cfg = _lib.OdgS2RConfig()
        self.L.odg_s2r_default_config(C.byref(cfg))
        cfg.auto_reset = 1 if auto_reset else 0
        cfg.max_steps = int(max_steps)     # MAX_STEPS_PER_EPISODE (sim2real/train.py:68,539); enforced only with auto_reset
        h = C.c_void_p()
        _lib.check(self.L.odg_s2r_create(self.sim._h, C.byref(self.sim._model), C.byref(cfg), C.byref(h)), "odg_s2r_create")
        self._h = h
        N, dev = num_envs, self.device
        self.obs = torch.empty(N, self.state_dim, device=dev)
        self.reward = torch.empty(N, device=dev)
        self.done = torch.empty(N, dtype=torch.uint8, device=dev)
        self.reason = torch.empty(N, dtype=torch.uint8, device=dev)
        self.sim_target_rad = torch.empty(N, 8, device=dev)
        self.terminal_obs = torch.empty(N, self.state_dim, device=dev)

    def close(self):
        if getattr(self, "_h", None):
            self.L.odg_s2r_destroy(self._h)
            self._h = None
        self.sim.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return self.sim._stream()

    def reset(self, mask=None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(self.L.odg_s2r_reset(self._h, _ptr(m), _ptr(self.obs), self._stream()), "odg_s2r_reset")
        return self.obs

    def step(self, action: torch.Tensor):
        a = action.to(device=self.device, dtype=torch.float32).contiguous()
        if a.shape != (self.num_envs, 4):
            raise ValueError(f"action must be [{self.num_envs}, 4]")
        _lib.check(self.L.odg_s2r_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), _ptr(self.reason),
                                       _ptr(self.sim_target_rad), _ptr(self.terminal_obs), self._stream()), "odg_s2r_step")
        info = {"sim_target_rad": self.sim_target_rad, "termination_reason": self.reason, "terminal_obs": self.terminal_obs}
        return self.obs, self.reward, self.done.bool(), info

    def set_bookkeeping(self, counter=None, prev_x=None, cum_pos=None, cum_neg=None, prev_net=None, last_cmd=None):
        def f(t, dt):
            return None if t is None else torch.as_tensor(t).to(device=self.device, dtype=dt).contiguous()
        args = [f(counter, torch.int32), f(prev_x, torch.float64), f(cum_pos, torch.float64), f(cum_neg, torch.float64),
                f(prev_net, torch.float64), f(last_cmd, torch.float32)]
        _lib.check(self.L.odg_s2r_set_bookkeeping(self._h, *[_ptr(a) for a in args], self._stream()), "odg_s2r_set_bookkeeping")
        torch.cuda.current_stream(self.device).synchronize()


class _DataView:
    """`env.data.qpos / qvel / ctrl` as numpy (fetched from the device on access)."""

    def __init__(self, env):
        self._e = env

    @property
    def qpos(self):
        return self._e._b.sim.get_state()[0][0].double().cpu().numpy()

    @property
    def qvel(self):
        return self._e._b.sim.get_state()[1][0].double().cpu().numpy()

    @property
    def ctrl(self):
        return self._e._b.sim_target_rad[0].double().cpu().numpy()


class QuadrupedEnv:
    """Single-environment numpy façade with the reference's constructor and 4-tuple API (sim2real/train.py:151)."""

    def __init__(self, model_path: str = "our_robot/walking_scene.xml"):
        if "walking_scene" not in str(model_path) and "our_robot" not in str(model_path):
            raise ValueError("QuadrupedEnv supports the OpenDOG walking scene")
        self._b = BatchedQuadrupedEnv(1, auto_reset=False)
        self.state_dim, self.action_dim = self._b.state_dim, self._b.action_dim
        self.viewer = None
        self.model = self._b.sim.desc
        self.data = _DataView(self)
        self.sim_steps_per_policy_step = 50
        self.episode_policy_step_counter = 0

    def reset(self):
        self.episode_policy_step_counter = 0
        return self._b.reset()[0].cpu().numpy().astype(np.float32)

    reset_to_home_keyframe = reset

    def step(self, policy_actions_scaled_neg1_to_1):
        a = torch.as_tensor(np.asarray(policy_actions_scaled_neg1_to_1, dtype=np.float32).reshape(1, 4))
        obs, rew, done, info = self._b.step(a)
        self.episode_policy_step_counter += 1
        reason = int(self._b.reason[0])
        out = {"sim_target_rad": self._b.sim_target_rad[0].double().cpu().numpy(), "termination_reason": REASONS[reason]}
        if reason == 1:
            out["mj_error"] = True
        return obs[0].cpu().numpy().astype(np.float32), float(rew[0]), bool(done[0]), out

    # viewer hooks of the reference are no-ops here
    def launch_viewer_internal(self): return False
    def close_viewer_internal(self): pass
    def sync_viewer_if_active(self): pass
    def close(self): self._b.close()


class WalkVecEnv:
    """stable-baselines3 `VecEnv` protocol over `BatchedWalkEnv` (replaces `SubprocVecEnv`, train/train.py:81-86)."""

    metadata = {"render_modes": ["human", "rgb_array", "depth_array"], "render_fps": 50}     # WalkEnvironment.py:28-31

    def __init__(self, num_envs: int, device=None, seed: int = 0, **config):
        self.env = BatchedWalkEnv(num_envs, device=device, seed=seed, auto_reset=1, **config)
        self.num_envs = num_envs
        self.observation_shape, self.action_shape = (self.env.obs_dim,), (self.env.act_dim,)
        try:                                                        # real gymnasium spaces when importable
            from gymnasium.spaces import Box
            self.observation_space = Box(-np.inf, np.inf, self.observation_shape, np.float64)   # WalkEnvironment.py:46-48
            self.action_space = Box(-1.0, 1.0, self.action_shape, np.float32)                    # ScaleActionEnvironment.py:19
        except Exception:
            self.observation_space = self.action_space = None
        self._actions = None
        self._h_act = torch.empty(num_envs, self.env.act_dim, pin_memory=True)

    def reset(self):
        return self.env.reset().double().cpu().numpy()

    def step_async(self, actions):
        self._actions = np.asarray(actions, dtype=np.float32)

    def step_wait(self):
        self._h_act.copy_(torch.from_numpy(self._actions))
        obs, rew, done, info = self.env.step(self._h_act.to(self.env.device, non_blocking=True))
        obs = obs.double().cpu().numpy(); rew = rew.double().cpu().numpy(); done = done.cpu().numpy()
        trunc = self.env.truncated.cpu().numpy().astype(bool); term = self.env.terminated.cpu().numpy().astype(bool)
        host = {k: v.cpu().numpy() for k, v in info.items()}
        infos = []
        for i in range(self.num_envs):
            d = {"x_position": float(host["x_position"][i]), "y_position": float(host["y_position"][i]),
                 "distance_from_origin": float(host["distance_from_origin"][i]),
                 "patterns_matches": float(host["patterns_matches"][i]),
                 "paw_contact_forces": {b: host["paw_contact_forces"][i, k].astype(np.float64) for k, b in enumerate((4, 7, 10, 13))},
                 "linear_vel_tracking_reward": float(host["linear_vel_tracking_reward"][i]),
                 "reward_ctrl": float(host["reward_ctrl"][i])}
            if done[i]:
                d["terminal_observation"] = host["terminal_obs"][i].astype(np.float64)
                d["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
            infos.append(d)
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_attr(self, name, indices=None):
        return [getattr(self.env, name)] * self.num_envs

    def seed(self, seed=None):
        return [seed] * self.num_envs
