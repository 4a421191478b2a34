"""ctypes front-end of the CPU oracle (oracle/odg_oracle.c). TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package `opendog_b200` never imports this module. Physics parity: unpinned at the single-step
level, loosely pinned against the reference's shipped MuJoCo walk files (see odg_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from opendog_b200.model.compile import (MAX_GEOM, MAX_JL, MAX_LEG, MAX_NQ, MAX_NU, MAX_NV,  # noqa: E402
                                        OdgModel, load_compiled, to_struct)

MAX_CON = MAX_GEOM * 4
MAX_EFC = MAX_NV + MAX_NV + 6 * MAX_CON
NB = 1 + MAX_LEG * MAX_JL
_d, _i = C.c_double, C.c_int


class OdgoContact(C.Structure):
    _fields_ = [("geom", _i), ("vert", _i), ("efc", _i), ("dim", _i), ("dist", _d), ("pos", _d * 3),
                ("frame", _d * 9), ("force", _d * 6)]


class OdgoData(C.Structure):
    _fields_ = [
        ("qpos", _d * MAX_NQ), ("qvel", _d * MAX_NV), ("time", _d), ("qacc_warmstart", _d * MAX_NV),
        ("ctrl", _d * MAX_NU),
        ("xpos", _d * 3 * NB), ("xmat", _d * 9 * NB), ("xquat", _d * 4 * NB), ("xipos", _d * 3 * NB),
        ("anchor", _d * 3 * MAX_JL * MAX_LEG), ("axis", _d * 3 * MAX_JL * MAX_LEG),
        ("M", _d * (MAX_NV * MAX_NV)),
        ("qfrc_bias", _d * MAX_NV), ("qfrc_passive", _d * MAX_NV), ("qfrc_actuator", _d * MAX_NV),
        ("qfrc_smooth", _d * MAX_NV), ("qacc_smooth", _d * MAX_NV), ("qacc", _d * MAX_NV),
        ("qfrc_constraint", _d * MAX_NV), ("actuator_force", _d * MAX_NU),
        ("ncon", _i), ("nefc", _i), ("solver_iter", _i),
        ("contact", OdgoContact * MAX_CON),
        ("efc_type", _i * MAX_EFC), ("efc_id", _i * MAX_EFC),
        ("efc_J", _d * (MAX_EFC * MAX_NV)),
        ("efc_pos", _d * MAX_EFC), ("efc_margin", _d * MAX_EFC), ("efc_vel", _d * MAX_EFC),
        ("efc_aref", _d * MAX_EFC), ("efc_R", _d * MAX_EFC), ("efc_D", _d * MAX_EFC),
        ("efc_frictionloss", _d * MAX_EFC), ("efc_force", _d * MAX_EFC),
        ("solver_cost", _d), ("solver_gradnorm", _d),
        ("gap_contact", _d), ("gap_support", _d), ("gap_limit", _d),
        ("bs_center", _d * 3 * MAX_GEOM), ("bs_radius", _d * MAX_GEOM), ("bs_ready", _i),
        ("cfrc_ext", _d * 6 * NB),
    ]


class OdgoWalkEnv(C.Structure):
    _fields_ = [
        ("m", C.POINTER(OdgModel)), ("d", OdgoData), ("seed", C.c_uint64), ("env_id", C.c_uint32),
        ("episode", C.c_uint32), ("step", _i), ("frame_skip", _i), ("max_steps", _i),
        ("last_action", C.c_float * 8), ("last_action_is_reset", _i),
        ("gait_index", _i), ("gait_matches", _i), ("desired_velocity", _d * 3),
    ]


class OdgoWalkInfo(C.Structure):
    _fields_ = [
        ("x_position", _d), ("y_position", _d), ("distance_from_origin", _d),
        ("paw_contact_forces", _d * 6 * 4), ("patterns_matches", _d),
        ("linear_vel_tracking_reward", _d), ("reward_ctrl", _d),
        ("paws_in_ground", _i * 4), ("gait_first_call", _i), ("reward_terms", _d * 6), ("min_gap", _d * 3),
    ]


_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libodg_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("odg_oracle.c", "odg_oracle.h")] + \
          [os.path.join(os.path.dirname(_HERE), "include", "odg_model.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        assert L.odgo_sizeof_model() == C.sizeof(OdgModel), (L.odgo_sizeof_model(), C.sizeof(OdgModel))
        assert L.odgo_sizeof_data() == C.sizeof(OdgoData), (L.odgo_sizeof_data(), C.sizeof(OdgoData))
        assert L.odgo_sizeof_walkenv() == C.sizeof(OdgoWalkEnv), (L.odgo_sizeof_walkenv(), C.sizeof(OdgoWalkEnv))
        L.odgo_constraint_cost.restype = C.c_double
        L.odgo_u01.restype = C.c_float
        L.odgo_u01.argtypes = [C.c_uint32]
        L.odgo_philox4x32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_uint32)]
        L.odgo_walk_init.argtypes = [C.POINTER(OdgoWalkEnv), C.POINTER(OdgModel), C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


def _np(ctarr, shape=None):
    a = np.ctypeslib.as_array(ctarr)
    return a if shape is None else a.reshape(shape)


class Sim:
    """One MuJoCo-like (model, data) pair: the physics half of the oracle."""

    def __init__(self, model="our_robot"):
        self.desc = load_compiled(model) if isinstance(model, str) else model
        self.m = to_struct(self.desc)
        self.d = OdgoData()
        self.nq, self.nv, self.nu = self.desc["nq"], self.desc["nv"], self.desc["nu"]
        lib().odgo_reset_data(C.byref(self.m), C.byref(self.d))

    # -- state
    @property
    def qpos(self): return _np(self.d.qpos)[:self.nq]
    @property
    def qvel(self): return _np(self.d.qvel)[:self.nv]
    @property
    def ctrl(self): return _np(self.d.ctrl)[:self.nu]
    @property
    def qacc(self): return _np(self.d.qacc)[:self.nv]
    @property
    def qacc_smooth(self): return _np(self.d.qacc_smooth)[:self.nv]
    @property
    def qacc_warmstart(self): return _np(self.d.qacc_warmstart)[:self.nv]
    @property
    def M(self): return _np(self.d.M)[:self.nv * self.nv].reshape(self.nv, self.nv)
    @property
    def qfrc_bias(self): return _np(self.d.qfrc_bias)[:self.nv]
    @property
    def qfrc_actuator(self): return _np(self.d.qfrc_actuator)[:self.nv]
    @property
    def qfrc_constraint(self): return _np(self.d.qfrc_constraint)[:self.nv]
    @property
    def xpos(self): return _np(self.d.xpos, (NB, 3))
    @property
    def xmat(self): return _np(self.d.xmat, (NB, 3, 3))
    @property
    def xipos(self): return _np(self.d.xipos, (NB, 3))
    @property
    def decision_gaps(self):
        """(contact inclusion [m], support-vertex lead [m], joint-limit distance [rad]) of the last forward pass."""
        return np.array([self.d.gap_contact, self.d.gap_support, self.d.gap_limit])
    @property
    def ncon(self): return self.d.ncon
    @property
    def nefc(self): return self.d.nefc
    @property
    def efc_J(self): return _np(self.d.efc_J)[:self.d.nefc * self.nv].reshape(self.d.nefc, self.nv)
    def efc(self, name): return _np(getattr(self.d, "efc_" + name))[:self.d.nefc]

    def contacts(self):
        out = []
        for c in range(self.d.ncon):
            k = self.d.contact[c]
            out.append(dict(geom=k.geom, vert=k.vert, dim=k.dim, dist=k.dist, pos=np.array(k.pos[:]),
                            force=np.array(k.force[:]), efc=k.efc))
        return out

    def reset_keyframe(self):
        lib().odgo_reset_data(C.byref(self.m), C.byref(self.d))
        self.qpos[:] = self.desc["key_qpos"]
        self.ctrl[:] = self.desc["key_ctrl"]

    def forward(self): lib().odgo_forward(C.byref(self.m), C.byref(self.d))
    def step(self): lib().odgo_step(C.byref(self.m), C.byref(self.d))
    def kinematics(self): lib().odgo_kinematics(C.byref(self.m), C.byref(self.d))
    def mass_matrix(self): lib().odgo_mass_matrix(C.byref(self.m), C.byref(self.d))
    def bias(self): lib().odgo_bias(C.byref(self.m), C.byref(self.d))
    def collision(self): lib().odgo_collision(C.byref(self.m), C.byref(self.d))

    def cfrc_ext(self):
        """data.cfrc_ext as gymnasium leaves it after do_simulation (mj_rnePostConstraint): [nbody, 6], row 0 = world."""
        lib().odgo_cfrc_ext(C.byref(self.m), C.byref(self.d))
        nb = 1 + self.desc["nleg"] * self.desc["njl"]
        return np.vstack([np.zeros((1, 6)), np.array(_np(self.d.cfrc_ext, (NB, 6))[:nb])])

    def cost(self, qacc):
        a = np.zeros(MAX_NV); a[:self.nv] = qacc
        return lib().odgo_constraint_cost(C.byref(self.m), C.byref(self.d), a.ctypes.data_as(C.POINTER(_d)))


class WalkEnv:
    """`ScaleActionWrapper(WalkEnvironmentV0)` restated (reference: WalkEnvironment.py:26-158)."""

    def __init__(self, model="our_robot", seed=0, env_id=0):
        self.desc = load_compiled(model) if isinstance(model, str) else model
        self.m = to_struct(self.desc)
        self.e = OdgoWalkEnv()
        lib().odgo_walk_init(C.byref(self.e), C.byref(self.m), seed, env_id)
        self.nq, self.nv = self.desc["nq"], self.desc["nv"]

    @property
    def qpos(self): return _np(self.e.d.qpos)[:self.nq]
    @property
    def qvel(self): return _np(self.e.d.qvel)[:self.nv]
    @property
    def qacc_warmstart(self): return _np(self.e.d.qacc_warmstart)[:self.nv]
    @property
    def desired_velocity(self): return _np(self.e.desired_velocity)
    @property
    def last_action(self): return _np(self.e.last_action)

    def reset(self):
        obs = np.zeros(33)
        lib().odgo_walk_reset(C.byref(self.e), obs.ctypes.data_as(C.POINTER(_d)))
        return obs

    def _call(self, fn, action, extra_obs=False):
        a = np.ascontiguousarray(action, dtype=np.float32)
        obs = np.zeros(33); rew = _d(); t0 = _i(); t1 = _i(); info = OdgoWalkInfo()
        args = [C.byref(self.e), a.ctypes.data_as(C.POINTER(C.c_float)), obs.ctypes.data_as(C.POINTER(_d)),
                C.byref(rew), C.byref(t0), C.byref(t1)]
        term_obs = None
        if extra_obs:
            term_obs = np.zeros(33)
            args.append(term_obs.ctypes.data_as(C.POINTER(_d)))
        args.append(C.byref(info))
        fn(*args)
        inf = dict(
            x_position=info.x_position, y_position=info.y_position,
            distance_from_origin=info.distance_from_origin,
            paw_contact_forces=np.array(_np(info.paw_contact_forces, (4, 6))),
            patterns_matches=info.patterns_matches,
            linear_vel_tracking_reward=info.linear_vel_tracking_reward, reward_ctrl=info.reward_ctrl,
            paws_in_ground=np.array(info.paws_in_ground[:]), gait_first_call=info.gait_first_call,
            reward_terms=np.array(info.reward_terms[:]), min_gap=np.array(info.min_gap[:]),
        )
        return obs, rew.value, t0.value, t1.value, term_obs, inf

    def step(self, action):
        obs, r, term, trunc, _, info = self._call(lib().odgo_walk_step, action)
        return obs, r, bool(term), bool(trunc), info

    def evaluate(self, scaled_action):
        obs, r, term, trunc, _, info = self._call(lib().odgo_walk_evaluate, scaled_action)
        return obs, r, bool(term), bool(trunc), info

    def step_autoreset(self, action):
        obs, r, done, trunc, tobs, info = self._call(lib().odgo_walk_step_autoreset, action, extra_obs=True)
        return obs, r, bool(done), bool(trunc), tobs, info


def scale_action(action):
    a = np.ascontiguousarray(action, dtype=np.float32)
    out = np.zeros(8, dtype=np.float32)
    lib().odgo_walk_scale_action(a.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def philox(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().odgo_philox4x32(seed, c0, c1, c2, c3, out)
    return list(out)
