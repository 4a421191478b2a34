"""MjcfModel -> `OdgModel` (include/odg_model.h): the constants the step kernel and the oracle use.

Checks that the model has the shape the hot path supports (one free trunk, `nleg` identical-depth
hinge chains), fuses welded leaf bodies (the OpenDOG paws, our_robot.xml:54-57) into the last
jointed link, mixes each collision geom's contact parameters with the floor plane's the way
MuJoCo's `mj_contactParam` does [3P-recalled: condim = max, friction = element-wise max, margin =
max, solref/solimp = solmix-weighted average at equal priority], and evaluates the `mj_setConst`
constants at qpos0.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import rbd_numpy as rbd
from .mjcf import MjcfModel, load_mjcf, quat_mul, quat_to_mat

MAX_LEG, MAX_JL, MAX_NU, MAX_GEOM, MAX_VERT = 4, 3, 12, 48, 2048
MAX_NQ, MAX_NV = 7 + MAX_LEG * MAX_JL, 6 + MAX_LEG * MAX_JL
GEOM_HULL, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX = 0, 1, 2, 3, 4
_PRIMITIVES = {"capsule": GEOM_CAPSULE, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX}

_d = C.c_double
_i = C.c_int


class OdgGeom(C.Structure):
    _fields_ = [
        ("leg", _i), ("link", _i), ("type", _i), ("vert_start", _i), ("vert_count", _i),
        ("mj_geom_id", _i), ("mj_body_id", _i), ("condim", _i),
        ("center", _d * 3), ("radius", _d), ("rot", _d * 9), ("size", _d * 3), ("friction", _d),
        ("friction_torsion", _d), ("friction_roll", _d), ("margin", _d),
        ("solref", _d * 2), ("solimp", _d * 5), ("invweight0", _d),
    ]


class OdgModel(C.Structure):
    _fields_ = [
        ("nleg", _i), ("njl", _i), ("nq", _i), ("nv", _i), ("nu", _i), ("ngeom", _i), ("nvert", _i),
        ("cone", _i),
        ("timestep", _d), ("gravity", _d * 3), ("impratio", _d),
        ("dof_solref", _d * 2), ("dof_solimp", _d * 5), ("lim_solref", _d * 2), ("lim_solimp", _d * 5),
        ("base_mass", _d), ("base_ipos", _d * 3), ("base_inertia", _d * 9),
        ("base_armature", _d * 6), ("base_frictionloss", _d * 6), ("base_damping", _d * 6),
        ("base_invweight0", _d * 6),
        ("body_pos", _d * 3 * MAX_JL * MAX_LEG), ("body_quat", _d * 4 * MAX_JL * MAX_LEG),
        ("jnt_pos", _d * 3 * MAX_JL * MAX_LEG), ("jnt_axis", _d * 3 * MAX_JL * MAX_LEG),
        ("jnt_range", _d * 2 * MAX_JL * MAX_LEG), ("jnt_limited", _i * MAX_JL * MAX_LEG),
        ("armature", _d * MAX_JL * MAX_LEG), ("frictionloss", _d * MAX_JL * MAX_LEG),
        ("damping", _d * MAX_JL * MAX_LEG), ("dof_invweight0", _d * MAX_JL * MAX_LEG),
        ("mass", _d * MAX_JL * MAX_LEG), ("ipos", _d * 3 * MAX_JL * MAX_LEG),
        ("inertia", _d * 9 * MAX_JL * MAX_LEG),
        ("act_leg", _i * MAX_NU), ("act_joint", _i * MAX_NU), ("act_kp", _d * MAX_NU), ("act_kv", _d * MAX_NU),
        ("act_ctrllimited", _i * MAX_NU), ("act_forcelimited", _i * MAX_NU),
        ("act_ctrlrange", _d * 2 * MAX_NU), ("act_forcerange", _d * 2 * MAX_NU),
        ("key_qpos", _d * MAX_NQ), ("key_ctrl", _d * MAX_NU),
        ("geom", OdgGeom * MAX_GEOM), ("vert", _d * 3 * MAX_VERT),
        ("multicontact_tilt", _d),
    ]


def compile_model(m: MjcfModel, key: str = "home", multicontact_tilt: float = 0.1,
                  trunk_collision: bool | None = None, max_condim: int | None = None) -> dict:
    """Return a JSON-able dict mirroring OdgModel. Raises if the tree is not trunk + leg chains."""
    free = [j for j in m.joints if j.type == "free"]
    if len(free) != 1 or m.bodies[free[0].body].parent != 0:
        raise ValueError("expected exactly one free joint on a child of the world")
    trunk = free[0].body
    floor = [g for g in m.geoms if g.body == 0 and g.type == "plane"]
    if len(floor) != 1:
        raise ValueError("expected exactly one floor plane on the world body")
    floor = floor[0]
    floor_id = m.geoms.index(floor)

    children = {i: [] for i in range(len(m.bodies))}
    for i, b in enumerate(m.bodies):
        if b.parent >= 0:
            children[b.parent].append(i)
    legs = []
    for root in children[trunk]:
        chain, welded, b = [], [], root
        while True:
            body = m.bodies[b]
            if len(body.joints) == 1 and m.joints[body.joints[0]].type == "hinge" and not welded:
                chain.append(b)
            elif not body.joints:
                welded.append(b)
            else:
                raise ValueError(f"body {body.name}: unsupported joint layout")
            if not children[b]:
                break
            if len(children[b]) != 1:
                raise ValueError("legs must be serial chains")
            b = children[b][0]
        legs.append((chain, welded))
    njl = len(legs[0][0])
    if any(len(c) != njl for c, _ in legs) or not (1 <= njl <= MAX_JL) or len(legs) > MAX_LEG:
        raise ValueError("legs must have the same number of hinge joints")
    # dof layout must be trunk(6) then legs in order, root to tip
    expect = 6
    for chain, _ in legs:
        for b in chain:
            if m.joints[m.bodies[b].joints[0]].dofadr != expect:
                raise ValueError("unexpected dof order")
            expect += 1

    dof_w, body_w = rbd.invweight0(m)
    tb = m.bodies[trunk]
    fj = free[0]
    out = dict(
        source=os.path.basename(m.source), nleg=len(legs), njl=njl, nq=m.nq, nv=m.nv, nu=len(m.actuators),
        cone=1 if m.option["cone"] == "elliptic" else 0,
        timestep=float(m.option["timestep"]), gravity=m.option["gravity"].tolist(),
        impratio=float(m.option["impratio"]),
        dof_solref=m.option["o_solref"].tolist(), dof_solimp=m.option["o_solimp"].tolist(),
        lim_solref=m.option["o_solref"].tolist(), lim_solimp=m.option["o_solimp"].tolist(),
        base_mass=tb.mass, base_ipos=tb.ipos.tolist(), base_inertia=tb.inertia.reshape(-1).tolist(),
        base_armature=[fj.armature] * 6, base_frictionloss=[fj.frictionloss] * 6,
        base_damping=[fj.damping] * 6, base_invweight0=dof_w[:6].tolist(),
        multicontact_tilt=multicontact_tilt,
    )
    if m.option["cone"] != "elliptic":
        raise NotImplementedError("the reference models use cone=elliptic")

    z = lambda *s: np.zeros(s)
    body_pos, body_quat = z(MAX_LEG, MAX_JL, 3), z(MAX_LEG, MAX_JL, 4)
    body_quat[..., 0] = 1
    jnt_pos, jnt_axis, jnt_range = z(MAX_LEG, MAX_JL, 3), z(MAX_LEG, MAX_JL, 3), z(MAX_LEG, MAX_JL, 2)
    jnt_axis[..., 2] = 1
    jnt_limited = np.zeros((MAX_LEG, MAX_JL), dtype=int)
    arm, fl, damp, dw = z(MAX_LEG, MAX_JL), z(MAX_LEG, MAX_JL), z(MAX_LEG, MAX_JL), z(MAX_LEG, MAX_JL)
    mass, ipos, inertia = z(MAX_LEG, MAX_JL), z(MAX_LEG, MAX_JL, 3), z(MAX_LEG, MAX_JL, 9)
    geoms, verts = [], []
    link_of_body = {}                       # mj body id -> (leg, link, pos_in_link, R_in_link)

    def add_geoms(bid, leg, link, p_in, R_in):
        for gid in m.bodies[bid].geoms:
            g = m.geoms[gid]
            if not ((g.contype & floor.conaffinity) or (floor.contype & g.conaffinity)):
                continue                                   # e.g. chassis contype=conaffinity=0
            if g.priority != floor.priority:
                src = g if g.priority > floor.priority else floor
                condim, fr, sref, simp = src.condim, src.friction, src.solref, src.solimp
            else:
                condim = max(g.condim, floor.condim)
                fr = np.maximum(g.friction, floor.friction)
                mix = g.solmix / (g.solmix + floor.solmix)
                sref = mix * g.solref + (1 - mix) * floor.solref
                simp = mix * g.solimp + (1 - mix) * floor.solimp
            condim = int(condim)
            if condim not in (1, 3, 6):                    # (condim 4 = 6 without the rolling rows: no reference model uses it)
                raise NotImplementedError(f"condim {condim}")
            if max_condim is not None:
                condim = min(condim, max_condim)
            e = dict(leg=leg, link=link, mj_geom_id=gid, mj_body_id=bid, condim=int(condim),
                     friction=float(fr[0]), friction_torsion=float(fr[1]), friction_roll=float(fr[2]), margin=float(max(g.margin, floor.margin) - max(g.gap, floor.gap)),
                     solref=np.asarray(sref).tolist(), solimp=np.asarray(simp).tolist(),
                     invweight0=float(body_w[bid, 0]), center=[0.0, 0.0, 0.0], radius=0.0,
                     rot=np.eye(3).reshape(-1).tolist(), size=[0.0, 0.0, 0.0], vert_start=0, vert_count=0)
            if g.type == "mesh":
                v = g.verts @ R_in.T + p_in
                e.update(type=GEOM_HULL, vert_start=len(verts), vert_count=len(v))
                verts.extend(v.tolist())
            elif g.type == "sphere":
                e.update(type=GEOM_SPHERE, center=(R_in @ g.pos + p_in).tolist(), radius=float(g.size[0]))
            elif g.type in _PRIMITIVES:                    # go1.xml:26-60: capsule / cylinder / box colliders
                size = np.zeros(3); size[:len(g.size)] = g.size
                e.update(type=_PRIMITIVES[g.type], center=(R_in @ g.pos + p_in).tolist(),
                         rot=(R_in @ quat_to_mat(g.quat)).reshape(-1).tolist(), size=size.tolist())
            else:
                continue
            geoms.append(e)

    for li, (chain, welded) in enumerate(legs):
        for ji, b in enumerate(chain):
            body = m.bodies[b]
            j = m.joints[body.joints[0]]
            body_pos[li, ji], body_quat[li, ji] = body.pos, body.quat
            jnt_pos[li, ji], jnt_axis[li, ji], jnt_range[li, ji] = j.pos, j.axis, j.range
            jnt_limited[li, ji] = int(j.limited)
            arm[li, ji], fl[li, ji], damp[li, ji], dw[li, ji] = j.armature, j.frictionloss, j.damping, dof_w[j.dofadr]
            if j.ref != 0.0:
                raise NotImplementedError("joint ref != 0")
            parts = [(body.mass, body.ipos, body.inertia)]
            add_geoms(b, li, ji, np.zeros(3), np.eye(3))
            if ji == njl - 1:                              # fuse welded descendants into the last link
                p, q = np.zeros(3), np.array([1.0, 0, 0, 0])
                for wb in welded:
                    wbody = m.bodies[wb]
                    p = p + quat_to_mat(q) @ wbody.pos
                    q = quat_mul(q, wbody.quat)
                    R = quat_to_mat(q)
                    parts.append((wbody.mass, p + R @ wbody.ipos, R @ wbody.inertia @ R.T))
                    add_geoms(wb, li, ji, p, R)
            mt = sum(pp[0] for pp in parts)
            c = sum(pp[0] * np.asarray(pp[1]) for pp in parts) / mt
            I = np.zeros((3, 3))
            for pm, pc, pI in parts:
                d = np.asarray(pc) - c
                I += pI + pm * (d @ d * np.eye(3) - np.outer(d, d))
            mass[li, ji], ipos[li, ji], inertia[li, ji] = mt, c, I.reshape(-1)
    if trunk_collision is None:
        trunk_collision = True
    if trunk_collision:
        add_geoms(trunk, -1, -1, np.zeros(3), np.eye(3))
    if len(geoms) > MAX_GEOM or len(verts) > MAX_VERT:
        raise ValueError("too many collision geoms / hull vertices")

    dof_to_leg = {}
    for li, (chain, _) in enumerate(legs):
        for ji, b in enumerate(chain):
            dof_to_leg[m.bodies[b].joints[0]] = (li, ji)
    acts = dict(act_leg=[], act_joint=[], act_kp=[], act_kv=[], act_ctrllimited=[], act_forcelimited=[],
                act_ctrlrange=[], act_forcerange=[], act_names=[])
    for a in m.actuators:
        li, ji = dof_to_leg[a.joint]
        acts["act_leg"].append(li); acts["act_joint"].append(ji)
        acts["act_kp"].append(a.kp); acts["act_kv"].append(a.kv)
        acts["act_ctrllimited"].append(int(a.ctrllimited)); acts["act_forcelimited"].append(int(a.forcelimited))
        acts["act_ctrlrange"].append(a.ctrlrange.tolist()); acts["act_forcerange"].append(a.forcerange.tolist())
        acts["act_names"].append(a.name)
    out.update(acts)
    k = m.keys[key]
    out.update(
        key_qpos=k["qpos"].tolist(), key_ctrl=k["ctrl"].tolist(),
        body_pos=body_pos.tolist(), body_quat=body_quat.tolist(), jnt_pos=jnt_pos.tolist(),
        jnt_axis=jnt_axis.tolist(), jnt_range=jnt_range.tolist(), jnt_limited=jnt_limited.tolist(),
        armature=arm.tolist(), frictionloss=fl.tolist(), damping=damp.tolist(), dof_invweight0=dw.tolist(),
        mass=mass.tolist(), ipos=ipos.tolist(), inertia=inertia.tolist(),
        geoms=geoms, verts=verts, ngeom=len(geoms), nvert=len(verts),
        joint_names=[[m.joints[m.bodies[b].joints[0]].name for b in chain] for chain, _ in legs],
        body_names=[b.name for b in m.bodies], floor_geom_id=floor_id,
    )
    return out


def to_struct(d: dict) -> OdgModel:
    """Pack the dict produced by `compile_model` (or loaded from JSON) into the C struct."""
    s = OdgModel()
    for name in ("nleg", "njl", "nq", "nv", "nu", "ngeom", "nvert", "cone"):
        setattr(s, name, int(d[name]))
    for name in ("timestep", "impratio", "base_mass", "multicontact_tilt"):
        setattr(s, name, float(d[name]))

    def fill(field, value):
        arr = np.ascontiguousarray(value)
        dst = getattr(s, field)
        ct = C.c_int if arr.dtype.kind in "iub" else C.c_double
        flat = arr.astype(np.int32 if ct is C.c_int else np.float64).reshape(-1)
        C.memmove(C.addressof(dst), flat.ctypes.data, flat.nbytes)

    for name in ("gravity", "dof_solref", "dof_solimp", "lim_solref", "lim_solimp", "base_ipos",
                 "base_inertia", "base_armature", "base_frictionloss", "base_damping", "base_invweight0",
                 "body_pos", "body_quat", "jnt_pos", "jnt_axis", "jnt_range", "armature", "frictionloss",
                 "damping", "dof_invweight0", "mass", "ipos", "inertia"):
        fill(name, np.asarray(d[name], dtype=np.float64))
    fill("jnt_limited", np.asarray(d["jnt_limited"], dtype=np.int32))
    nu = d["nu"]
    for name, dt in (("act_leg", np.int32), ("act_joint", np.int32), ("act_ctrllimited", np.int32),
                     ("act_forcelimited", np.int32), ("act_kp", np.float64), ("act_kv", np.float64)):
        a = np.zeros(MAX_NU, dtype=dt); a[:nu] = d[name]; fill(name, a)
    for name in ("act_ctrlrange", "act_forcerange"):
        a = np.zeros((MAX_NU, 2)); a[:nu] = d[name]; fill(name, a)
    a = np.zeros(MAX_NQ); a[:d["nq"]] = d["key_qpos"]; fill("key_qpos", a)
    a = np.zeros(MAX_NU); a[:nu] = d["key_ctrl"]; fill("key_ctrl", a)
    v = np.zeros((MAX_VERT, 3)); v[:d["nvert"]] = np.asarray(d["verts"]).reshape(-1, 3); fill("vert", v)
    for gi, g in enumerate(d["geoms"]):
        cg = s.geom[gi]
        for name in ("leg", "link", "type", "vert_start", "vert_count", "mj_geom_id", "mj_body_id", "condim"):
            setattr(cg, name, int(g[name]))
        for name in ("radius", "friction", "margin", "invweight0"):
            setattr(cg, name, float(g[name]))
        cg.friction_torsion = float(g.get("friction_torsion", 0.0)); cg.friction_roll = float(g.get("friction_roll", 0.0))
        cg.center[:] = g["center"]; cg.solref[:] = g["solref"]; cg.solimp[:] = g["solimp"]
        cg.rot[:] = g.get("rot", [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0]); cg.size[:] = g.get("size", [0.0, 0.0, 0.0])
    return s


_ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
BUILTIN = {"our_robot": "our_robot.model.json", "go1": "go1.model.json"}


def load_compiled(name_or_path: str) -> dict:
    """Load a compiled model: a builtin name, a `.json` produced by tools/compile_model.py, or an MJCF path."""
    if name_or_path in BUILTIN:
        name_or_path = os.path.join(_ASSET_DIR, BUILTIN[name_or_path])
    if name_or_path.endswith(".json"):
        with open(name_or_path) as fh:
            return json.load(fh)
    return compile_model(load_mjcf(name_or_path))


def save_compiled(d: dict, path: str):
    def rnd(x):
        if isinstance(x, float):
            return float(repr(x))
        if isinstance(x, list):
            return [rnd(v) for v in x]
        if isinstance(x, dict):
            return {k: rnd(v) for k, v in x.items()}
        return x
    with open(path, "w") as fh:
        json.dump(rnd(d), fh, separators=(",", ":"))
        fh.write("\n")
