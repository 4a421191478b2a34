"""ctypes driver of the TEST-ONLY lane emulator (tests/emu/odg_emu.cpp).

Runs the exact source of the CUDA step kernel on the host so kernel numerics can be compared with
the oracle without a GPU. Never imported by the package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from opendog_b200.model.compile import OdgModel, load_compiled, to_struct  # noqa: E402


class OdgEnvConfig(C.Structure):
    _fields_ = [("task", C.c_int), ("frame_skip", C.c_int), ("max_episode_steps", C.c_int),
                ("auto_reset", C.c_int), ("solver_iterations", C.c_int), ("ls_iterations", C.c_int),
                ("solver_tolerance", C.c_float), ("ls_tolerance", C.c_float), ("reset_noise_scale", C.c_float),
                ("scale_actions", C.c_int), ("launch_lanes", C.c_int), ("first_env_id", C.c_int), ("obs_layout", C.c_int),
                ("launch_block", C.c_int), ("launch_lockstep", C.c_int), ("launch_fat", C.c_int)]


_p = C.c_void_p


class StepArgs(C.Structure):
    _fields_ = [("action", _p), ("obs", _p), ("reward", _p), ("terminated", _p), ("truncated", _p),
                ("mode", C.c_int), ("want_info", C.c_int),
                ("x_position", _p), ("y_position", _p), ("distance", _p), ("paw_forces", _p),
                ("patterns_matches", _p), ("lin_vel_reward", _p), ("reward_ctrl", _p), ("terminal_obs", _p),
                ("paws_in_ground", _p), ("gait_reward", _p), ("qacc", _p), ("ncon", _p), ("fn_sum", _p),
                ("solver_iters", _p), ("ls_evals", _p), ("reward_raw", _p), ("cfrc_ext", _p), ("task_terms", _p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libodg_emu.so")
        srcs = [os.path.join(_HERE, "odg_emu.cpp"), os.path.join(_ROOT, "opendog_b200/csrc/odg_core.cuh"),
                os.path.join(_ROOT, "opendog_b200/csrc/odg_prep.h"), os.path.join(_ROOT, "include/odg.h"),
                os.path.join(_ROOT, "include/odg_model.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread",
                                   "-ffp-contract=off", "-o", so, srcs[0]])
        L = C.CDLL(so)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.POINTER(OdgModel), C.POINTER(OdgEnvConfig), C.c_int, C.c_uint64]
        for name in ("emu_destroy", "emu_reset", "emu_step", "emu_get_state", "emu_set_state",
                     "emu_get_env_state", "emu_set_env_state"):
            getattr(L, name).restype = None
        assert L.emu_sizeof_stepargs() == C.sizeof(StepArgs)
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuEnv:
    INFO = dict(x_position=("f", 1), y_position=("f", 1), distance=("f", 1), paw_forces=("f", 24),
                patterns_matches=("f", 1), lin_vel_reward=("f", 1), reward_ctrl=("f", 1),
                terminal_obs=("f", 33), paws_in_ground=("B", 4), gait_reward=("i", 1), qacc=("f", 14),
                ncon=("i", 1), fn_sum=("f", 1), solver_iters=("i", 1), ls_evals=("i", 1), reward_raw=("f", 1),
                cfrc_ext=("f", 78), task_terms=("f", 10))

    def __init__(self, num_envs, model="our_robot", seed=0, **cfg):
        self.desc = load_compiled(model) if isinstance(model, str) else model      # (a descriptor dict: model variants)
        self.m = to_struct(self.desc)
        self.cfg = OdgEnvConfig()
        lib().emu_default_config(C.byref(self.cfg))
        for k, v in cfg.items():
            setattr(self.cfg, k, v)
        self.N = num_envs
        self.h = lib().emu_create(C.byref(self.m), C.byref(self.cfg), num_envs, seed)
        assert self.h
        self.nq, self.nv, self.nu = self.desc["nq"], self.desc["nv"], self.desc["nu"]
        self.obs_dim = 9 + self.nu if self.cfg.task == 1 else (12 if self.cfg.obs_layout else 9) + 3 * self.nu

    def __del__(self):
        if getattr(self, "h", None):
            lib().emu_destroy(C.c_void_p(self.h))
            self.h = None

    def reset(self, mask=None):
        obs = np.zeros((self.N, self.obs_dim), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().emu_reset(C.c_void_p(self.h), _ptr(m), _ptr(obs))
        return obs

    def _run(self, action, mode, info):
        a = np.ascontiguousarray(action, np.float32).reshape(self.N, self.nu)
        obs = np.zeros((self.N, self.obs_dim), np.float32)
        rew = np.zeros(self.N, np.float32)
        term = np.zeros(self.N, np.uint8)
        trunc = np.zeros(self.N, np.uint8)
        A = StepArgs(action=_ptr(a), obs=_ptr(obs), reward=_ptr(rew), terminated=_ptr(term), truncated=_ptr(trunc),
                     mode=mode, want_info=1 if info else 0)
        out = {}
        if info:
            for k, (t, n) in self.INFO.items():
                n = {"terminal_obs": self.obs_dim, "qacc": self.nv}.get(k, n)
                if k in ("cfrc_ext", "task_terms") and self.nu != 12:
                    continue
                dt = {"f": np.float32, "i": np.int32, "B": np.uint8}[t]
                out[k] = np.zeros((self.N, n) if n > 1 else self.N, dt)
                setattr(A, k, _ptr(out[k]))
        lib().emu_step(C.c_void_p(self.h), C.byref(A))
        return obs, rew, term.astype(bool), trunc.astype(bool), out

    def step(self, action, info=True):
        return self._run(action, 0, info)

    def evaluate(self, ctrl, info=True):
        return self._run(ctrl, 1, info)

    def get_state(self):
        qpos = np.zeros((self.N, self.nq), np.float32)
        qvel = np.zeros((self.N, self.nv), np.float32)
        warm = np.zeros((self.N, self.nv), np.float32)
        lib().emu_get_state(C.c_void_p(self.h), _ptr(qpos), _ptr(qvel), _ptr(warm))
        return qpos, qvel, warm

    def set_state(self, qpos, qvel, warm=None):
        qpos = np.ascontiguousarray(qpos, np.float32)
        qvel = np.ascontiguousarray(qvel, np.float32)
        w = None if warm is None else np.ascontiguousarray(warm, np.float32)
        lib().emu_set_state(C.c_void_p(self.h), _ptr(qpos), _ptr(qvel), _ptr(w))

    def get_env_state(self):
        N = self.N
        o = dict(step=np.zeros(N, np.int32), gait_index=np.zeros(N, np.int32), gait_matches=np.zeros(N, np.int32),
                 last_action=np.zeros((N, self.nu), np.float32), desired_velocity=np.zeros((N, 3), np.float32),
                 fresh=np.zeros(N, np.uint8))
        lib().emu_get_env_state(C.c_void_p(self.h), *[_ptr(o[k]) for k in
                                ("step", "gait_index", "gait_matches", "last_action", "desired_velocity", "fresh")])
        return o

    def set_env_state(self, **kw):
        args = []
        for k, dt in (("step", np.int32), ("gait_index", np.int32), ("gait_matches", np.int32),
                      ("last_action", np.float32), ("desired_velocity", np.float32), ("fresh", np.uint8)):
            v = kw.get(k)
            args.append(None if v is None else np.ascontiguousarray(v, dt))
        self._keep = args
        lib().emu_set_env_state(C.c_void_p(self.h), *[_ptr(a) for a in args])
