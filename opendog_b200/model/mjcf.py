"""MJCF reader: the subset of MuJoCo's XML compiler the OpenDOG hot path needs.

Restates (from the MuJoCo 3.2.3 XML reference, [3P-recalled]) how
`mujoco.MjModel.from_xml_path` resolves the reference's model files
  Code/mujoco/our_robot/our_robot.xml:1-118, walking_scene.xml:1-28,
  Code/mujoco/unitree_go1/go1.xml:1-229, walk_scene.xml
into bodies / joints / geoms / actuators / keyframes: `<include>`, nested
`<default class>` trees with `childclass`, `autolimits`, mesh assets with
`scale`, `inertiafromgeom` (mass from `mass=` or density), `<freejoint/>`
(ignores joint defaults) vs `<joint type="free"/>` (inherits them, which is why
the OpenDOG trunk has armature 0.02 / frictionloss 0.1 on its six base DoFs).

This is host-side model preparation; no dynamics here.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from . import mesh as _mesh


# ----------------------------------------------------------------------------- math
def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n < 1e-15:                      # mju_normalize4: zero quaternion -> identity
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def axis_angle_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w,x,y,z)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0, 0.0]
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return quat_normalize(q)


def _floats(s, n=None):
    v = np.array([float(t) for t in s.split()], dtype=np.float64)
    if n is not None and len(v) != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


# ----------------------------------------------------------------------------- data
@dataclass
class Body:
    name: str
    parent: int
    pos: np.ndarray
    quat: np.ndarray
    mass: float = 0.0
    ipos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))  # about COM, body frame
    joints: list = field(default_factory=list)
    geoms: list = field(default_factory=list)


@dataclass
class Joint:
    name: str
    body: int
    type: str                  # "free" | "hinge"
    pos: np.ndarray
    axis: np.ndarray
    limited: bool
    range: np.ndarray
    armature: float
    damping: float
    frictionloss: float
    ref: float
    qposadr: int = 0
    dofadr: int = 0


@dataclass
class Geom:
    name: str
    body: int
    type: str                  # plane|mesh|sphere|capsule|cylinder|box
    pos: np.ndarray
    quat: np.ndarray
    size: np.ndarray
    contype: int
    conaffinity: int
    condim: int
    priority: int
    friction: np.ndarray
    margin: float
    gap: float
    solref: np.ndarray
    solimp: np.ndarray
    solmix: float
    mass: float | None
    density: float
    mesh: str | None = None
    verts: np.ndarray | None = None      # hull vertices in the BODY frame (mesh geoms)


@dataclass
class Actuator:
    name: str
    joint: int
    kp: float
    kv: float
    ctrllimited: bool
    ctrlrange: np.ndarray
    forcelimited: bool
    forcerange: np.ndarray
    gear: float = 1.0


@dataclass
class MjcfModel:
    bodies: list
    joints: list
    geoms: list
    actuators: list
    keys: dict                 # name -> dict(qpos, ctrl)
    option: dict
    nq: int
    nv: int
    source: str = ""

    def body_id(self, name):
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        raise KeyError(name)

    def joint_id(self, name):
        for i, j in enumerate(self.joints):
            if j.name == name:
                return i
        raise KeyError(name)


# ----------------------------------------------------------------------------- parse
_DEFAULT_OPTION = dict(
    timestep=0.002, gravity=np.array([0.0, 0.0, -9.81]), cone="pyramidal", impratio=1.0,
    integrator="Euler", solver="Newton", iterations=100, tolerance=1e-8,
    o_solref=np.array([0.02, 1.0]), o_solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]),
)
_GEOM_DEFAULT = dict(type="sphere", contype="1", conaffinity="1", condim="3", priority="0",
                     friction="1 0.005 0.0001", margin="0", gap="0", solref="0.02 1",
                     solimp="0.9 0.95 0.001 0.5 2", solmix="1", density="1000")
_JOINT_DEFAULT = dict(type="hinge", pos="0 0 0", axis="0 0 1", armature="0", damping="0",
                      frictionloss="0", ref="0")


def _expand_includes(root, base_dir):
    """Inline `<include file=...>` children (MJCF include = textual splice of the file's sections)."""
    for parent in list(root.iter()):
        for child in list(parent):
            if child.tag == "include":
                inc = ET.parse(os.path.join(base_dir, child.get("file"))).getroot()
                _expand_includes(inc, base_dir)
                idx = list(parent).index(child)
                parent.remove(child)
                for k, sub in enumerate(list(inc)):
                    parent.insert(idx + k, sub)
    return root


class _Defaults:
    """`<default>` class tree: class -> tag -> attribute dict.

    Like MuJoCo's reader, a `<default>` element first applies its own settings
    (on top of a copy of its parent's), then its nested `<default>` children are
    processed, so declaration order inside one element does not matter.
    """

    def __init__(self, root):
        self.table = {"main": {}}
        for d in root.findall("default"):
            self._walk(d, None)

    def _walk(self, elem, parent_cls):
        cls = elem.get("class") or "main"
        if parent_cls is not None and cls not in self.table:
            self.table[cls] = {t: dict(a) for t, a in self.table[parent_cls].items()}
        self.table.setdefault(cls, {})
        for child in elem:
            if child.tag != "default":
                self.table[cls].setdefault(child.tag, {}).update(child.attrib)
        for child in elem:
            if child.tag == "default":
                self._walk(child, cls)

    def resolve(self, tag, elem, childclass):
        cls = elem.get("class") or childclass or "main"
        if cls not in self.table:
            raise KeyError(f"unknown default class {cls!r}")
        attrs = dict(self.table[cls].get(tag, {}))
        attrs.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return attrs


def _orientation(attrs):
    if "quat" in attrs:
        return quat_normalize(_floats(attrs["quat"], 4))
    if "xyaxes" in attrs:
        v = _floats(attrs["xyaxes"], 6)
        x = v[:3] / np.linalg.norm(v[:3])
        y = v[3:] - x * np.dot(x, v[3:])
        y /= np.linalg.norm(y)
        return mat_to_quat(np.stack([x, y, np.cross(x, y)], axis=1))
    if "euler" in attrs or "axisangle" in attrs or "zaxis" in attrs:
        raise NotImplementedError("orientation spec not used by the reference models")
    return np.array([1.0, 0.0, 0.0, 0.0])


def load_mjcf(path: str, inertia_mode: str = "legacy") -> MjcfModel:
    base_dir = os.path.dirname(os.path.abspath(path))
    root = _expand_includes(ET.parse(path).getroot(), base_dir)

    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    if comp.get("angle", "degree") != "radian":
        raise NotImplementedError("reference models use angle=radian")
    autolimits = comp.get("autolimits", "true") == "true"
    meshdir = os.path.join(base_dir, comp.get("meshdir", ""))

    option = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in _DEFAULT_OPTION.items()}
    for o in root.findall("option"):
        for k, v in o.attrib.items():
            if k == "gravity":
                option[k] = _floats(v, 3)
            elif k in ("timestep", "impratio", "tolerance"):
                option[k] = float(v)
            elif k in ("iterations",):
                option[k] = int(v)
            else:
                option[k] = v

    defaults = _Defaults(root)

    # ---- mesh assets: hull vertices + mass properties in mesh-file coordinates (scaled)
    meshes = {}
    for asset in root.findall("asset"):
        for m in asset.findall("mesh"):
            attrs = defaults.resolve("mesh", m, None)
            fname = attrs["file"]
            name = attrs.get("name", os.path.splitext(os.path.basename(fname))[0])
            fpath = os.path.join(meshdir, fname)
            scale = _floats(attrs.get("scale", "1 1 1"), 3)
            meshes[name] = dict(path=fpath, scale=scale)

    def mesh_data(name):
        m = meshes[name]
        if "tris" not in m:
            tris = _mesh.read_stl(m["path"]) * m["scale"]
            m["tris"] = tris
            m["hull"] = _mesh.hull_vertices(tris.reshape(-1, 3))
            m["props"] = _mesh.mesh_mass_properties(tris, inertia_mode)
        return m

    bodies = [Body("world", -1, np.zeros(3), np.array([1.0, 0, 0, 0]))]
    joints, geoms = [], []

    def add_geom(elem, body_id, childclass):
        a = dict(_GEOM_DEFAULT)
        a.update(defaults.resolve("geom", elem, childclass))
        gtype = a["type"]
        if "mesh" in a and "type" not in elem.attrib and "type" not in defaults.resolve("geom", elem, childclass):
            gtype = "mesh"
        pos = _floats(a.get("pos", "0 0 0"), 3)
        quat = _orientation(a)
        size = _floats(a["size"]) if "size" in a else np.zeros(0)
        if "fromto" in a:                      # capsule/cylinder given by end points
            ft = _floats(a["fromto"], 6)
            p0, p1 = ft[:3], ft[3:]
            pos = 0.5 * (p0 + p1)
            z = (p1 - p0)
            length = np.linalg.norm(z)
            z = z / length
            ref = np.array([0.0, 0, 1])
            v = np.cross(ref, z)
            s, c = np.linalg.norm(v), np.dot(ref, z)
            quat = (np.array([1.0, 0, 0, 0]) if s < 1e-12 and c > 0 else
                    np.array([0.0, 1, 0, 0]) if s < 1e-12 else
                    axis_angle_quat(v / s, np.arctan2(s, c)))
            size = np.array([size[0], 0.5 * length])
        g = Geom(
            name=a.get("name", ""), body=body_id, type=gtype, pos=pos, quat=quat, size=size,
            contype=int(a["contype"]), conaffinity=int(a["conaffinity"]), condim=int(a["condim"]),
            priority=int(a["priority"]), friction=_floats(a["friction"]),
            margin=float(a["margin"]), gap=float(a["gap"]), solref=_floats(a["solref"], 2),
            solimp=_floats(a["solimp"]), solmix=float(a["solmix"]),
            mass=float(a["mass"]) if "mass" in a else None, density=float(a["density"]),
            mesh=a.get("mesh"),
        )
        fr = np.array([1.0, 0.005, 0.0001]); fr[:len(g.friction)] = g.friction; g.friction = fr
        si = _DEFAULT_OPTION["o_solimp"].copy(); si[:len(g.solimp)] = g.solimp; g.solimp = si
        # the hull is only needed for geoms that can collide (visual meshes, e.g. Go1's trunk.stl which is not even
        # shipped in the reference, are never read unless a body's inertia must be inferred from them)
        if gtype == "mesh" and (g.contype or g.conaffinity):
            md = mesh_data(g.mesh)
            R = quat_to_mat(quat)
            g.verts = md["hull"] @ R.T + pos
        geoms.append(g)
        bodies[body_id].geoms.append(len(geoms) - 1)

    def add_body(elem, parent_id, childclass):
        cc = elem.get("childclass", childclass)
        b = Body(elem.get("name", f"body{len(bodies)}"), parent_id,
                 _floats(elem.get("pos", "0 0 0"), 3), _orientation(elem.attrib))
        bodies.append(b)
        bid = len(bodies) - 1
        explicit_inertial = None
        for child in elem:
            if child.tag == "inertial":
                explicit_inertial = child
            elif child.tag == "freejoint":
                joints.append(Joint(child.get("name", ""), bid, "free", np.zeros(3), np.array([0.0, 0, 1]),
                                    False, np.zeros(2), 0.0, 0.0, 0.0, 0.0))
                b.joints.append(len(joints) - 1)
            elif child.tag == "joint":
                a = dict(_JOINT_DEFAULT)
                a.update(defaults.resolve("joint", child, cc))
                jtype = a["type"]
                if jtype not in ("free", "hinge"):
                    raise NotImplementedError(f"joint type {jtype}")
                rng = _floats(a["range"], 2) if "range" in a and jtype == "hinge" else np.zeros(2)
                limited = a.get("limited", "auto")
                lim = (limited == "true") or (limited == "auto" and autolimits and "range" in a and jtype == "hinge")
                axis = _floats(a["axis"], 3)
                joints.append(Joint(a.get("name", ""), bid, jtype, _floats(a["pos"], 3),
                                    axis / np.linalg.norm(axis), lim, rng, float(a["armature"]),
                                    float(a["damping"]), float(a["frictionloss"]), float(a["ref"])))
                b.joints.append(len(joints) - 1)
            elif child.tag == "geom":
                add_geom(child, bid, cc)
        # inertia
        if explicit_inertial is not None:
            a = explicit_inertial.attrib
            b.mass = float(a["mass"])
            b.ipos = _floats(a["pos"], 3)
            Rq = quat_to_mat(_orientation(a))
            if "diaginertia" in a:
                b.inertia = Rq @ np.diag(_floats(a["diaginertia"], 3)) @ Rq.T
            else:
                f = _floats(a["fullinertia"], 6)
                b.inertia = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
        else:
            _inertia_from_geoms(b, [geoms[g] for g in b.geoms], mesh_data)
        for child in elem:
            if child.tag == "body":
                add_body(child, bid, cc)

    for wb in root.findall("worldbody"):
        for child in wb:
            if child.tag == "geom":
                add_geom(child, 0, None)
            elif child.tag == "body":
                add_body(child, 0, None)

    # mjModel stores geoms grouped by body id (world geoms first): the reference relies on the
    # floor being geom 0 (rewards/walk_environment_reward_calc.py:319,331)
    order = sorted(range(len(geoms)), key=lambda gi: geoms[gi].body)
    geoms = [geoms[gi] for gi in order]
    for b in bodies:
        b.geoms = []
    for gi, g in enumerate(geoms):
        bodies[g.body].geoms.append(gi)

    # address assignment
    nq = nv = 0
    for j in joints:
        j.qposadr, j.dofadr = nq, nv
        nq += 7 if j.type == "free" else 1
        nv += 6 if j.type == "free" else 1

    # actuators (<position>): gainprm[0]=kp, biasprm=(0,-kp,-kv)
    actuators = []
    jname = {j.name: i for i, j in enumerate(joints)}
    for sec in root.findall("actuator"):
        for act in sec:
            if act.tag != "position":
                raise NotImplementedError(f"actuator <{act.tag}>")
            a = defaults.resolve("position", act, None)
            cr = _floats(a["ctrlrange"], 2) if "ctrlrange" in a else np.zeros(2)
            fr = _floats(a["forcerange"], 2) if "forcerange" in a else np.zeros(2)
            actuators.append(Actuator(
                a.get("name", ""), jname[a["joint"]], float(a.get("kp", 1.0)), float(a.get("kv", 0.0)),
                autolimits and "ctrlrange" in a, cr, autolimits and "forcerange" in a, fr,
                float(a.get("gear", "1").split()[0])))

    keys = {}
    for sec in root.findall("keyframe"):
        for k in sec.findall("key"):
            qpos = _floats(k.get("qpos")) if k.get("qpos") else None
            if qpos is not None:
                for j in joints:                       # compiler normalises keyframe quaternions
                    if j.type == "free":
                        qpos[j.qposadr + 3:j.qposadr + 7] = quat_normalize(qpos[j.qposadr + 3:j.qposadr + 7])
            keys[k.get("name")] = dict(qpos=qpos, ctrl=_floats(k.get("ctrl")) if k.get("ctrl") else None)

    return MjcfModel(bodies, joints, geoms, actuators, keys, option, nq, nv, source=os.path.abspath(path))


def _inertia_from_geoms(body: Body, glist, mesh_data):
    """`inertiafromgeom`: mass/COM/inertia of a body from its geoms (mesh geoms only are needed here)."""
    parts = []
    for g in glist:
        if g.type == "plane":
            continue
        if g.type != "mesh":
            raise NotImplementedError("inertia from primitive geoms is not needed by the reference models")
        vol, com, I_unit = mesh_data(g.mesh)["props"]
        mass = g.mass if g.mass is not None else g.density * vol
        R = quat_to_mat(g.quat)
        parts.append((mass, g.pos + R @ com, R @ (I_unit * (mass / vol)) @ R.T))
    if not parts:
        return
    m = sum(p[0] for p in parts)
    c = sum(p[0] * p[1] for p in parts) / m
    I = np.zeros((3, 3))
    for pm, pc, pI in parts:
        d = pc - c
        I += pI + pm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    body.mass, body.ipos, body.inertia = float(m), c, I
