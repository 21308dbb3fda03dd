"""Minimal MJCF -> rigid-body model compiler for the rollout engine.

The reference loads its robots with ``mujoco.MjModel.from_xml_path`` (examples/walking.py:389-391,
examples/kbot/train.py:1076) and moves them to the device with ``mjx.put_model``
(ksim/utils/mujoco.py:388-391).  Neither ``mujoco`` nor ``mujoco-mjx`` exists in this image, so the
feature subset those two robots exercise is compiled here: default classes,
``childclass`` inheritance, capsule ``fromto``, inertia inferred from geoms or given by ``<inertial>``,
free + hinge joints, motors, sites, sensors, explicit ``<contact><pair>`` / ``<exclude>``.

Everything is float64 numpy; the result is a plain ``Model`` of arrays named after MuJoCo's ``mjModel``
fields so that the plugin layer can read the same names the reference reads.  Anything outside the subset
raises ``MjcfError`` at compile time instead of being silently mis-simulated.
"""

from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np

__all__ = ["Model", "MjcfError", "load_mjcf", "load_mjcf_string", "save_model", "load_model", "GEOM", "JNT", "SENSOR", "OBJ"]


class MjcfError(ValueError):
    """Raised when the MJCF uses a feature outside the supported subset."""


class GEOM:
    PLANE, HFIELD, SPHERE, CAPSULE, ELLIPSOID, CYLINDER, BOX, MESH = range(8)
    BY_NAME = {
        "plane": 0, "hfield": 1, "sphere": 2, "capsule": 3,
        "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7,
    }


class JNT:
    FREE, BALL, SLIDE, HINGE = range(4)


class OBJ:
    BODY, XBODY, GEOM, SITE = 1, 2, 5, 6


class SENSOR:
    """Sensor type codes (own numbering; names follow the MJCF element names)."""

    TOUCH, ACCELEROMETER, VELOCIMETER, GYRO, FORCE, TORQUE, MAGNETOMETER = 0, 1, 2, 3, 4, 5, 6
    FRAMEPOS, FRAMEQUAT, FRAMEXAXIS, FRAMEYAXIS, FRAMEZAXIS = 10, 11, 12, 13, 14
    FRAMELINVEL, FRAMEANGVEL, FRAMELINACC, FRAMEANGACC = 15, 16, 17, 18
    BY_NAME = {
        "touch": (0, 1), "accelerometer": (1, 3), "velocimeter": (2, 3), "gyro": (3, 3),
        "force": (4, 3), "torque": (5, 3), "magnetometer": (6, 3),
        "framepos": (10, 3), "framequat": (11, 4), "framexaxis": (12, 3), "frameyaxis": (13, 3),
        "framezaxis": (14, 3), "framelinvel": (15, 3), "frameangvel": (16, 3),
        "framelinacc": (17, 3), "frameangacc": (18, 3),
    }


MJ_MINVAL = 1e-15

# --------------------------------------------------------------------------------------------------
# small quaternion helpers (wxyz)
# --------------------------------------------------------------------------------------------------


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


def mat_to_quat(m: np.ndarray) -> np.ndarray:
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q / np.linalg.norm(q)


def z_to_quat(vec: np.ndarray) -> np.ndarray:
    """Quaternion rotating the z axis onto ``vec`` (shortest arc)."""
    v = np.asarray(vec, dtype=np.float64)
    n = np.linalg.norm(v)
    if n < MJ_MINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])
    v = v / n
    axis = np.cross([0.0, 0.0, 1.0], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        return np.array([1.0, 0.0, 0.0, 0.0]) if v[2] > 0 else np.array([0.0, 1.0, 0.0, 0.0])
    axis = axis / s
    ang = math.atan2(s, v[2])
    return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


def axis_angle_quat(axis: np.ndarray, angle: float) -> np.ndarray:
    return np.concatenate([[math.cos(angle / 2)], np.asarray(axis) * math.sin(angle / 2)])


def rot_vec(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    return quat_to_mat(q) @ v


# --------------------------------------------------------------------------------------------------
# Model container
# --------------------------------------------------------------------------------------------------


@dataclass
class Options:
    timestep: float = 0.002
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    iterations: int = 100
    ls_iterations: int = 50
    tolerance: float = 1e-8
    ls_tolerance: float = 0.01
    impratio: float = 1.0
    integrator: str = "euler"
    solver: str = "newton"
    cone: str = "pyramidal"
    disable_eulerdamp: bool = False
    disable_refsafe: bool = False
    disable_warmstart: bool = False


@dataclass
class Model:
    """Arrays named after ``mjModel``; see module docstring."""

    name: str = ""
    opt: Options = field(default_factory=Options)
    meaninertia: float = 1.0
    arrays: dict[str, np.ndarray] = field(default_factory=dict)
    names: dict[str, list[str]] = field(default_factory=dict)
    custom_numeric: dict[str, float] = field(default_factory=dict)

    def __getattr__(self, key: str) -> Any:  # noqa: ANN401
        arrays = self.__dict__.get("arrays", {})
        if key in arrays:
            return arrays[key]
        raise AttributeError(key)

    # sizes -----------------------------------------------------------------------------------
    @property
    def nq(self) -> int:
        return int(self.arrays["qpos0"].shape[0])

    @property
    def nv(self) -> int:
        return int(self.arrays["dof_bodyid"].shape[0])

    @property
    def nu(self) -> int:
        return int(self.arrays["actuator_trnid"].shape[0])

    @property
    def nbody(self) -> int:
        return int(self.arrays["body_parentid"].shape[0])

    @property
    def njnt(self) -> int:
        return int(self.arrays["jnt_type"].shape[0])

    @property
    def ngeom(self) -> int:
        return int(self.arrays["geom_type"].shape[0])

    @property
    def nsite(self) -> int:
        return int(self.arrays["site_bodyid"].shape[0])

    @property
    def nsensor(self) -> int:
        return int(self.arrays["sensor_type"].shape[0])

    @property
    def nsensordata(self) -> int:
        return int(self.arrays["sensor_dim"].sum()) if self.nsensor else 0

    @property
    def npair(self) -> int:
        return int(self.arrays["pair_geom1"].shape[0])

    # name lookups (mirrors ksim/utils/mujoco.py:29-168 name->index helpers) --------------------------
    def id(self, kind: str, name: str) -> int:
        try:
            return self.names[kind].index(name)
        except ValueError:
            raise KeyError(f"{kind} '{name}' not found; choices: {self.names[kind]}") from None

    def sensor_range(self, name: str) -> tuple[int, int]:
        i = self.id("sensor", name)
        adr = int(self.arrays["sensor_adr"][i])
        return adr, adr + int(self.arrays["sensor_dim"][i])

    def copy(self) -> "Model":
        return Model(
            name=self.name,
            opt=Options(**{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in vars(self.opt).items()}),
            meaninertia=self.meaninertia,
            arrays={k: v.copy() for k, v in self.arrays.items()},
            names={k: list(v) for k, v in self.names.items()},
            custom_numeric=dict(self.custom_numeric),
        )


# --------------------------------------------------------------------------------------------------
# defaults handling
# --------------------------------------------------------------------------------------------------

_DEFAULT_TAGS = ("geom", "joint", "motor", "site", "general", "position", "velocity", "mesh", "pair")


class _Defaults:
    def __init__(self) -> None:
        self.classes: dict[str, dict[str, dict[str, str]]] = {"main": {t: {} for t in _DEFAULT_TAGS}}

    def parse(self, elem: ET.Element, parent: str | None) -> None:
        name = elem.get("class", "main" if parent is None else None)
        if name is None:
            raise MjcfError("nested <default> requires a class name")
        if parent is None:
            base = {t: {} for t in _DEFAULT_TAGS}
            if name in self.classes:
                base = {t: dict(v) for t, v in self.classes[name].items()}
        else:
            base = {t: dict(v) for t, v in self.classes[parent].items()}
        for child in elem:
            if child.tag in _DEFAULT_TAGS:
                base[child.tag].update(child.attrib)
        self.classes[name] = base
        for child in elem:
            if child.tag == "default":
                self.parse(child, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: str | None) -> dict[str, str]:
        cls = elem.get("class") or childclass or "main"
        if cls not in self.classes:
            raise MjcfError(f"unknown default class '{cls}'")
        out = dict(self.classes[cls].get(tag, {}))
        out.update(elem.attrib)
        return out


def _floats(s: str | None, n: int | None = None, default: list[float] | None = None) -> np.ndarray:
    if s is None:
        if default is None:
            raise MjcfError("missing required attribute")
        return np.array(default, dtype=np.float64)
    vals = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and vals.shape[0] != n:
        if default is not None and vals.shape[0] < n:
            full = np.array(default, dtype=np.float64)
            full[: vals.shape[0]] = vals
            return full
        raise MjcfError(f"expected {n} values, got '{s}'")
    return vals


def _bool(s: str | None, default: bool) -> bool:
    if s is None:
        return default
    if s in ("true", "1"):
        return True
    if s in ("false", "0"):
        return False
    if s == "auto":
        return default
    raise MjcfError(f"bad bool '{s}'")


# --------------------------------------------------------------------------------------------------
# geom mass properties (MuJoCo computation docs, 'inertiafromgeom'; default density 1000)
# --------------------------------------------------------------------------------------------------


def _geom_mass_inertia(gtype: int, size: np.ndarray, density: float) -> tuple[float, np.ndarray]:
    if gtype == GEOM.SPHERE:
        r = size[0]
        m = density * 4.0 / 3.0 * math.pi * r**3
        return m, np.full(3, 0.4 * m * r * r)
    if gtype == GEOM.CAPSULE:
        r, half = size[0], size[1]
        h = 2 * half
        v_cyl = math.pi * r * r * h
        v_sph = 4.0 / 3.0 * math.pi * r**3
        m = density * (v_cyl + v_sph)
        m_sph = density * v_sph
        m_cyl = density * v_cyl
        ixx = m_cyl * (3 * r * r + h * h) / 12.0
        izz = m_cyl * r * r / 2.0
        sph_i = 2.0 * m_sph * r * r / 5.0
        ixx += sph_i + m_sph * h * (3 * r + 2 * h) / 8.0
        izz += sph_i
        return m, np.array([ixx, ixx, izz])
    if gtype == GEOM.BOX:
        x, y, z = size
        m = density * 8 * x * y * z
        return m, m / 3.0 * np.array([y * y + z * z, x * x + z * z, x * x + y * y])
    if gtype == GEOM.CYLINDER:
        r, half = size[0], size[1]
        h = 2 * half
        m = density * math.pi * r * r * h
        ixx = m * (3 * r * r + h * h) / 12.0
        return m, np.array([ixx, ixx, m * r * r / 2.0])
    if gtype == GEOM.ELLIPSOID:
        a, b, c = size
        m = density * 4.0 / 3.0 * math.pi * a * b * c
        return m, m / 5.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    if gtype in (GEOM.PLANE, GEOM.HFIELD, GEOM.MESH):
        return 0.0, np.zeros(3)
    raise MjcfError(f"unsupported geom type {gtype}")


# --------------------------------------------------------------------------------------------------
# compiler
# --------------------------------------------------------------------------------------------------


class _Compiler:
    def __init__(self, root: ET.Element) -> None:
        self.root = root
        self.defaults = _Defaults()
        self.angle_scale = math.pi / 180.0  # MuJoCo default: degrees
        self.bodies: list[dict[str, Any]] = []
        self.joints: list[dict[str, Any]] = []
        self.geoms: list[dict[str, Any]] = []
        self.sites: list[dict[str, Any]] = []

    # .............................................................................................
    def compile(self) -> Model:
        root = self.root
        comp = root.find("compiler")
        if comp is not None:
            ang = comp.get("angle", "degree")
            self.angle_scale = 1.0 if ang == "radian" else math.pi / 180.0
            if comp.get("eulerseq", "xyz") != "xyz" or comp.get("coordinate", "local") != "local":
                raise MjcfError("only local coordinates / xyz euler supported")
        for d in root.findall("default"):
            self.defaults.parse(d, None)

        model = Model(name=root.get("model", ""))
        self._parse_option(root.find("option"), model.opt)
        cust = root.find("custom")
        if cust is not None:
            for num in cust.findall("numeric"):
                model.custom_numeric[num.get("name", "")] = float(num.get("data", "0").split()[0])

        world = root.find("worldbody")
        if world is None:
            raise MjcfError("no <worldbody>")
        self.bodies.append(
            dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), inertial=None, jnts=[], geoms=[])
        )
        self._parse_body_children(world, 0, None)
        self._finalize(model)
        return model

    def _parse_option(self, opt: ET.Element | None, o: Options) -> None:
        if opt is None:
            return
        if opt.get("timestep"):
            o.timestep = float(opt.get("timestep"))  # type: ignore[arg-type]
        if opt.get("gravity"):
            o.gravity = _floats(opt.get("gravity"), 3)
        for key in ("iterations", "ls_iterations"):
            if opt.get(key):
                setattr(o, key, int(opt.get(key)))  # type: ignore[arg-type]
        for key in ("tolerance", "ls_tolerance", "impratio"):
            if opt.get(key):
                setattr(o, key, float(opt.get(key)))  # type: ignore[arg-type]
        for key in ("integrator", "solver", "cone"):
            if opt.get(key):
                setattr(o, key, opt.get(key).lower())  # type: ignore[union-attr]

    # .............................................................................................
    def _orientation(self, a: dict[str, str]) -> np.ndarray:
        if "quat" in a:
            q = _floats(a["quat"], 4)
            return q / np.linalg.norm(q)
        if "euler" in a:
            e = _floats(a["euler"], 3) * self.angle_scale
            q = np.array([1.0, 0, 0, 0])
            for i, ang in enumerate(e):  # intrinsic xyz
                ax = np.zeros(3)
                ax[i] = 1.0
                q = quat_mul(q, axis_angle_quat(ax, ang))
            return q
        if "axisangle" in a:
            aa = _floats(a["axisangle"], 4)
            ax = aa[:3] / np.linalg.norm(aa[:3])
            return axis_angle_quat(ax, aa[3] * self.angle_scale)
        if "zaxis" in a:
            return z_to_quat(_floats(a["zaxis"], 3))
        if "xyaxes" in a:
            xy = _floats(a["xyaxes"], 6)
            x = xy[:3] / np.linalg.norm(xy[:3])
            y = xy[3:] - x * np.dot(x, xy[3:])
            y = y / np.linalg.norm(y)
            return mat_to_quat(np.stack([x, y, np.cross(x, y)], axis=1))
        return np.array([1.0, 0.0, 0.0, 0.0])

    def _parse_body_children(self, elem: ET.Element, body_id: int, childclass: str | None) -> None:
        body = self.bodies[body_id]
        for child in elem:
            tag = child.tag
            if tag == "inertial":
                a = child.attrib
                if "fullinertia" in a:
                    raise MjcfError("fullinertia not supported")
                body["inertial"] = dict(
                    pos=_floats(a.get("pos"), 3),
                    quat=self._orientation(a),
                    mass=float(a["mass"]),
                    inertia=_floats(a.get("diaginertia"), 3),
                )
            elif tag in ("joint", "freejoint"):
                self._parse_joint(child, body_id, childclass, free=(tag == "freejoint"))
            elif tag == "geom":
                self._parse_geom(child, body_id, childclass)
            elif tag == "site":
                a = self.defaults.resolve("site", child, childclass)
                self.sites.append(
                    dict(
                        name=a.get("name", f"site{len(self.sites)}"),
                        body=body_id,
                        pos=_floats(a.get("pos"), 3, [0, 0, 0]),
                        quat=self._orientation(a),
                        size=_floats(a.get("size"), 3, [0.005, 0.005, 0.005]),
                        type=GEOM.BY_NAME[a.get("type", "sphere")],
                    )
                )
            elif tag == "body":
                cc = child.get("childclass", childclass)
                new_id = len(self.bodies)
                self.bodies.append(
                    dict(
                        name=child.get("name", f"body{new_id}"),
                        parent=body_id,
                        pos=_floats(child.get("pos"), 3, [0, 0, 0]),
                        quat=self._orientation(child.attrib),
                        inertial=None,
                        jnts=[],
                        geoms=[],
                    )
                )
                self._parse_body_children(child, new_id, cc)
            elif tag in ("light", "camera"):
                continue
            else:
                raise MjcfError(f"unsupported element <{tag}> in body")

    def _parse_joint(self, elem: ET.Element, body_id: int, childclass: str | None, free: bool) -> None:
        if free:
            a: dict[str, str] = dict(elem.attrib)
            jtype = JNT.FREE
        else:
            a = self.defaults.resolve("joint", elem, childclass)
            tname = a.get("type", "hinge")
            if tname not in ("hinge", "free"):
                raise MjcfError(f"joint type '{tname}' not supported (free + hinge only)")
            jtype = JNT.FREE if tname == "free" else JNT.HINGE
        rng = _floats(a.get("range"), 2, [0.0, 0.0])
        limited_attr = a.get("limited", "auto")
        if limited_attr == "auto":
            limited = bool("range" in a)  # compiler autolimits default true
        else:
            limited = _bool(limited_attr, False)
        if jtype == JNT.HINGE:
            rng = rng * self.angle_scale
        axis = _floats(a.get("axis"), 3, [0, 0, 1])
        axis = axis / np.linalg.norm(axis)
        ref = float(a.get("ref", 0.0)) * (self.angle_scale if jtype == JNT.HINGE else 1.0)
        springref = float(a.get("springref", 0.0)) * (self.angle_scale if jtype == JNT.HINGE else 1.0)
        j = dict(
            name=a.get("name", f"joint{len(self.joints)}"),
            type=jtype,
            body=body_id,
            pos=_floats(a.get("pos"), 3, [0, 0, 0]),
            axis=axis,
            range=rng,
            limited=limited and jtype == JNT.HINGE,
            stiffness=float(a.get("stiffness", 0.0)),
            damping=float(a.get("damping", 0.0)),
            armature=float(a.get("armature", 0.0)),
            frictionloss=float(a.get("frictionloss", 0.0)),
            ref=ref,
            springref=springref,
            margin=float(a.get("margin", 0.0)),
            solreflimit=_floats(a.get("solreflimit"), 2, [0.02, 1.0]),
            solimplimit=_floats(a.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
            solreffriction=_floats(a.get("solreffriction"), 2, [0.02, 1.0]),
            solimpfriction=_floats(a.get("solimpfriction"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
        )
        self.bodies[body_id]["jnts"].append(len(self.joints))
        self.joints.append(j)

    def _parse_geom(self, elem: ET.Element, body_id: int, childclass: str | None) -> None:
        a = self.defaults.resolve("geom", elem, childclass)
        gtype = GEOM.BY_NAME[a.get("type", "sphere")]
        size = _floats(a.get("size"), None, [0.0, 0.0, 0.0])
        size = np.concatenate([size, np.zeros(3 - size.shape[0])])[:3]
        pos = _floats(a.get("pos"), 3, [0, 0, 0])
        quat = self._orientation(a)
        if "fromto" in a:
            if gtype not in (GEOM.CAPSULE, GEOM.CYLINDER, GEOM.BOX, GEOM.ELLIPSOID):
                raise MjcfError("fromto only valid for capsule/cylinder/box/ellipsoid")
            ft = _floats(a["fromto"], 6)
            vec = ft[0:3] - ft[3:6]
            size = np.array([size[0], np.linalg.norm(vec) / 2.0, 0.0])
            pos = 0.5 * (ft[0:3] + ft[3:6])
            quat = z_to_quat(vec)
        self.bodies[body_id]["geoms"].append(len(self.geoms))
        self.geoms.append(
            dict(
                name=a.get("name", f"geom{len(self.geoms)}"),
                type=gtype,
                body=body_id,
                pos=pos,
                quat=quat,
                size=size,
                density=float(a.get("density", 1000.0)),
                mass=float(a["mass"]) if "mass" in a else None,
                contype=int(a.get("contype", 1)),
                conaffinity=int(a.get("conaffinity", 1)),
                condim=int(a.get("condim", 3)),
                priority=int(a.get("priority", 0)),
                friction=_floats(a.get("friction"), 3, [1.0, 0.005, 0.0001]),
                solmix=float(a.get("solmix", 1.0)),
                solref=_floats(a.get("solref"), 2, [0.02, 1.0]),
                solimp=_floats(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                margin=float(a.get("margin", 0.0)),
                gap=float(a.get("gap", 0.0)),
                group=int(a.get("group", 0)),
            )
        )

    # .............................................................................................
    def _body_inertia(self, b: dict[str, Any]) -> tuple[float, np.ndarray, np.ndarray, np.ndarray]:
        """Returns (mass, ipos, iquat, diag inertia) of one body."""
        if b["inertial"] is not None:
            i = b["inertial"]
            return i["mass"], i["pos"], i["quat"], i["inertia"]
        parts = []
        for gid in b["geoms"]:
            g = self.geoms[gid]
            if not 0 <= g["group"] <= 5:
                continue
            m, inertia = _geom_mass_inertia(g["type"], g["size"], g["density"])
            if g["mass"] is not None and m > 0:
                inertia = inertia * (g["mass"] / m)
                m = g["mass"]
            if m > 0:
                parts.append((m, g["pos"], g["quat"], inertia))
        if not parts:
            return 0.0, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3)
        if len(parts) == 1:
            m, pos, quat, inertia = parts[0]
            return m, pos.copy(), quat.copy(), inertia.copy()
        mass = sum(p[0] for p in parts)
        com = sum(p[0] * p[1] for p in parts) / mass
        full = np.zeros((3, 3))
        for m, pos, quat, inertia in parts:
            r = quat_to_mat(quat)
            d = pos - com
            full += r @ np.diag(inertia) @ r.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        evals, evecs = np.linalg.eigh(full)
        order = np.argsort(-evals)
        evals, evecs = evals[order], evecs[:, order]
        if np.linalg.det(evecs) < 0:
            evecs[:, 2] = -evecs[:, 2]
        return mass, com, mat_to_quat(evecs), evals

    def _finalize(self, model: Model) -> None:
        nb = len(self.bodies)
        arr: dict[str, np.ndarray] = {}
        f8, i4 = np.float64, np.int32
        # bodies
        arr["body_parentid"] = np.array([b["parent"] for b in self.bodies], dtype=i4)
        arr["body_pos"] = np.stack([b["pos"] for b in self.bodies]).astype(f8)
        arr["body_quat"] = np.stack([b["quat"] for b in self.bodies]).astype(f8)
        mass, ipos, iquat, inertia = [], [], [], []
        for i, b in enumerate(self.bodies):
            if i == 0:
                m, p, q, d = 0.0, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3)
            else:
                m, p, q, d = self._body_inertia(b)
            mass.append(m)
            ipos.append(p)
            iquat.append(q)
            inertia.append(d)
        arr["body_mass"] = np.array(mass, dtype=f8)
        arr["body_ipos"] = np.stack(ipos).astype(f8)
        arr["body_iquat"] = np.stack(iquat).astype(f8)
        arr["body_inertia"] = np.stack(inertia).astype(f8)
        rootid = np.zeros(nb, dtype=i4)
        for i in range(1, nb):
            p = int(arr["body_parentid"][i])
            rootid[i] = i if p == 0 else rootid[p]
        arr["body_rootid"] = rootid
        sub = arr["body_mass"].copy()
        for i in range(nb - 1, 0, -1):
            sub[arr["body_parentid"][i]] += sub[i]
        arr["body_subtreemass"] = sub

        # joints / dofs
        nj = len(self.joints)
        jtype = np.array([j["type"] for j in self.joints], dtype=i4)
        qposadr, dofadr = np.zeros(nj, dtype=i4), np.zeros(nj, dtype=i4)
        nq = nv = 0
        for k, j in enumerate(self.joints):
            qposadr[k], dofadr[k] = nq, nv
            nq += 7 if j["type"] == JNT.FREE else 1
            nv += 6 if j["type"] == JNT.FREE else 1
        arr["jnt_type"] = jtype
        arr["jnt_qposadr"] = qposadr
        arr["jnt_dofadr"] = dofadr
        arr["jnt_bodyid"] = np.array([j["body"] for j in self.joints], dtype=i4)
        arr["jnt_pos"] = np.stack([j["pos"] for j in self.joints]).astype(f8)
        arr["jnt_axis"] = np.stack([j["axis"] for j in self.joints]).astype(f8)
        arr["jnt_stiffness"] = np.array([j["stiffness"] for j in self.joints], dtype=f8)
        arr["jnt_range"] = np.stack([j["range"] for j in self.joints]).astype(f8)
        arr["jnt_limited"] = np.array([j["limited"] for j in self.joints], dtype=i4)
        arr["jnt_margin"] = np.array([j["margin"] for j in self.joints], dtype=f8)
        arr["jnt_solref"] = np.stack([j["solreflimit"] for j in self.joints]).astype(f8)
        arr["jnt_solimp"] = np.stack([j["solimplimit"] for j in self.joints]).astype(f8)

        qpos0, qpos_spring = np.zeros(nq), np.zeros(nq)
        dof_body, dof_jnt, dof_parent = np.zeros(nv, i4), np.zeros(nv, i4), np.full(nv, -1, i4)
        dof_arm, dof_damp, dof_fl = np.zeros(nv), np.zeros(nv), np.zeros(nv)
        dof_solref, dof_solimp = np.zeros((nv, 2)), np.zeros((nv, 5))
        body_jntnum, body_jntadr = np.zeros(nb, i4), np.full(nb, -1, i4)
        body_dofnum, body_dofadr = np.zeros(nb, i4), np.full(nb, -1, i4)
        last_dof_of_body = np.full(nb, -1, i4)
        for bi, b in enumerate(self.bodies):
            if bi == 0:
                continue
            # last dof of the nearest ancestor that has dofs
            p = int(arr["body_parentid"][bi])
            prev = -1
            while p != 0 and last_dof_of_body[p] < 0:
                p = int(arr["body_parentid"][p])
            if p != 0:
                prev = int(last_dof_of_body[p])
            if b["jnts"]:
                body_jntnum[bi] = len(b["jnts"])
                body_jntadr[bi] = b["jnts"][0]
                body_dofadr[bi] = dofadr[b["jnts"][0]]
                if any(self.joints[k]["type"] == JNT.FREE for k in b["jnts"]) and len(b["jnts"]) != 1:
                    raise MjcfError("free joint must be the only joint of its body")
            for k in b["jnts"]:
                j = self.joints[k]
                n = 6 if j["type"] == JNT.FREE else 1
                if j["type"] == JNT.FREE:
                    if int(arr["body_parentid"][bi]) != 0:
                        raise MjcfError("free joint only allowed on a child of the world")
                    qa = qposadr[k]
                    qpos0[qa : qa + 3] = b["pos"]
                    qpos0[qa + 3 : qa + 7] = b["quat"]
                    qpos_spring[qa : qa + 7] = qpos0[qa : qa + 7]
                else:
                    qpos0[qposadr[k]] = j["ref"]
                    qpos_spring[qposadr[k]] = j["springref"]
                for d in range(n):
                    di = dofadr[k] + d
                    dof_body[di], dof_jnt[di], dof_parent[di] = bi, k, prev
                    dof_arm[di], dof_damp[di], dof_fl[di] = j["armature"], j["damping"], j["frictionloss"]
                    dof_solref[di], dof_solimp[di] = j["solreffriction"], j["solimpfriction"]
                    prev = di
                body_dofnum[bi] += n
            if body_dofnum[bi] > 0:
                last_dof_of_body[bi] = prev
            else:
                last_dof_of_body[bi] = -1
        arr["qpos0"], arr["qpos_spring"] = qpos0, qpos_spring
        arr["dof_bodyid"], arr["dof_jntid"], arr["dof_parentid"] = dof_body, dof_jnt, dof_parent
        arr["dof_armature"], arr["dof_damping"], arr["dof_frictionloss"] = dof_arm, dof_damp, dof_fl
        arr["dof_solref"], arr["dof_solimp"] = dof_solref, dof_solimp
        arr["body_jntnum"], arr["body_jntadr"] = body_jntnum, body_jntadr
        arr["body_dofnum"], arr["body_dofadr"] = body_dofnum, body_dofadr

        # geoms
        ng = len(self.geoms)
        def gstack(key: str, dtype: Any = f8, shape: tuple[int, ...] = ()) -> np.ndarray:  # noqa: ANN401
            if ng == 0:
                return np.zeros((0, *shape), dtype=dtype)
            return np.array([g[key] for g in self.geoms], dtype=dtype)

        arr["geom_type"] = gstack("type", i4)
        arr["geom_bodyid"] = gstack("body", i4)
        arr["geom_pos"] = gstack("pos", f8, (3,))
        arr["geom_quat"] = gstack("quat", f8, (4,))
        arr["geom_size"] = gstack("size", f8, (3,))
        arr["geom_friction"] = gstack("friction", f8, (3,))
        arr["geom_solref"] = gstack("solref", f8, (2,))
        arr["geom_solimp"] = gstack("solimp", f8, (5,))
        arr["geom_solmix"] = gstack("solmix", f8)
        arr["geom_margin"] = gstack("margin", f8)
        arr["geom_gap"] = gstack("gap", f8)
        arr["geom_condim"] = gstack("condim", i4)
        arr["geom_contype"] = gstack("contype", i4)
        arr["geom_conaffinity"] = gstack("conaffinity", i4)
        arr["geom_priority"] = gstack("priority", i4)

        # sites
        ns = len(self.sites)
        arr["site_bodyid"] = np.array([s["body"] for s in self.sites], dtype=i4).reshape(ns)
        arr["site_pos"] = np.array([s["pos"] for s in self.sites], dtype=f8).reshape(ns, 3)
        arr["site_quat"] = np.array([s["quat"] for s in self.sites], dtype=f8).reshape(ns, 4)
        arr["site_size"] = np.array([s["size"] for s in self.sites], dtype=f8).reshape(ns, 3)
        arr["site_type"] = np.array([s["type"] for s in self.sites], dtype=i4).reshape(ns)

        model.names = dict(
            body=[b["name"] for b in self.bodies],
            joint=[j["name"] for j in self.joints],
            geom=[g["name"] for g in self.geoms],
            site=[s["name"] for s in self.sites],
        )
        model.arrays = arr
        self._parse_actuators(model)
        self._parse_sensors(model)
        self._parse_contacts(model)


    # .............................................................................................
    def _parse_actuators(self, model: Model) -> None:
        act = self.root.find("actuator")
        names, trnid, gear, ctrlrange, ctrllimited, forcerange, forcelimited = [], [], [], [], [], [], []
        if act is not None:
            for a_el in act:
                if a_el.tag != "motor":
                    raise MjcfError(f"actuator <{a_el.tag}> not supported (motor only)")
                a = self.defaults.resolve("motor", a_el, None)
                jname = a.get("joint")
                if jname is None:
                    raise MjcfError("motor requires a joint transmission")
                jid = model.id("joint", jname)
                if model.arrays["jnt_type"][jid] != JNT.HINGE:
                    raise MjcfError("motor on non-hinge joint not supported")
                names.append(a.get("name", f"actuator{len(names)}"))
                trnid.append(jid)
                gear.append(_floats(a.get("gear"), None, [1.0])[0])
                cr = _floats(a.get("ctrlrange"), 2, [0.0, 0.0])
                cl = a.get("ctrllimited", "auto")
                ctrllimited.append(("ctrlrange" in a) if cl == "auto" else _bool(cl, False))
                ctrlrange.append(cr)
                fr = _floats(a.get("forcerange"), 2, [0.0, 0.0])
                fl = a.get("forcelimited", "auto")
                forcelimited.append(("forcerange" in a) if fl == "auto" else _bool(fl, False))
                forcerange.append(fr)
        nu = len(names)
        arr = model.arrays
        arr["actuator_trnid"] = np.array(trnid, dtype=np.int32).reshape(nu)
        arr["actuator_gear"] = np.array(gear, dtype=np.float64).reshape(nu)
        arr["actuator_ctrlrange"] = np.array(ctrlrange, dtype=np.float64).reshape(nu, 2)
        arr["actuator_ctrllimited"] = np.array(ctrllimited, dtype=np.int32).reshape(nu)
        arr["actuator_forcerange"] = np.array(forcerange, dtype=np.float64).reshape(nu, 2)
        arr["actuator_forcelimited"] = np.array(forcelimited, dtype=np.int32).reshape(nu)
        model.names["actuator"] = names

    def _parse_sensors(self, model: Model) -> None:
        sen = self.root.find("sensor")
        names, stype, objtype, objid, dim, adr = [], [], [], [], [], []
        n = 0
        if sen is not None:
            for s in sen:
                if s.tag not in SENSOR.BY_NAME:
                    raise MjcfError(f"sensor <{s.tag}> not supported")
                code, d = SENSOR.BY_NAME[s.tag]
                if s.tag.startswith("frame"):
                    ot = s.get("objtype")
                    on = s.get("objname")
                    if ot == "site":
                        objtype.append(OBJ.SITE)
                        objid.append(model.id("site", on))  # type: ignore[arg-type]
                    elif ot in ("body", "xbody"):
                        objtype.append(OBJ.BODY if ot == "body" else OBJ.XBODY)
                        objid.append(model.id("body", on))  # type: ignore[arg-type]
                    else:
                        raise MjcfError(f"frame sensor objtype '{ot}' not supported")
                    if s.get("reftype") is not None:
                        raise MjcfError("frame sensor reftype not supported")
                else:
                    objtype.append(OBJ.SITE)
                    objid.append(model.id("site", s.get("site")))  # type: ignore[arg-type]
                names.append(s.get("name", f"sensor{len(names)}"))
                stype.append(code)
                dim.append(d)
                adr.append(n)
                n += d
        ns = len(names)
        arr = model.arrays
        arr["sensor_type"] = np.array(stype, dtype=np.int32).reshape(ns)
        arr["sensor_objtype"] = np.array(objtype, dtype=np.int32).reshape(ns)
        arr["sensor_objid"] = np.array(objid, dtype=np.int32).reshape(ns)
        arr["sensor_dim"] = np.array(dim, dtype=np.int32).reshape(ns)
        arr["sensor_adr"] = np.array(adr, dtype=np.int32).reshape(ns)
        model.names["sensor"] = names

    # .............................................................................................
    def _pair_params(self, model: Model, g1: int, g2: int, a: dict[str, str]) -> dict[str, Any]:
        """Contact parameters of a geom pair.

        Attributes left undefined on an explicit ``<pair>`` fall back to the geom-derived mixing rule used
        for dynamic pairs (priority, then max condim / max friction / solmix-weighted solref+solimp),
        as MuJoCo's XML reference states for ``contact/pair`` attributes.
        """
        arr = model.arrays
        if arr["geom_type"][g1] > arr["geom_type"][g2]:
            g1, g2 = g2, g1
        p1, p2 = int(arr["geom_priority"][g1]), int(arr["geom_priority"][g2])
        out: dict[str, Any] = dict(geom1=g1, geom2=g2)
        out["margin"] = float(a["margin"]) if "margin" in a else max(arr["geom_margin"][g1], arr["geom_margin"][g2])
        out["gap"] = float(a["gap"]) if "gap" in a else max(arr["geom_gap"][g1], arr["geom_gap"][g2])
        if p1 != p2:
            gh = g1 if p1 > p2 else g2
            condim = int(arr["geom_condim"][gh])
            fr = arr["geom_friction"][gh]
            solref = arr["geom_solref"][gh].copy()
            solimp = arr["geom_solimp"][gh].copy()
        else:
            condim = int(max(arr["geom_condim"][g1], arr["geom_condim"][g2]))
            fr = np.maximum(arr["geom_friction"][g1], arr["geom_friction"][g2])
            s1, s2 = arr["geom_solmix"][g1], arr["geom_solmix"][g2]
            if s1 >= MJ_MINVAL and s2 >= MJ_MINVAL:
                mix = s1 / (s1 + s2)
            elif s1 < MJ_MINVAL and s2 < MJ_MINVAL:
                mix = 0.5
            elif s1 < MJ_MINVAL:
                mix = 0.0
            else:
                mix = 1.0
            r1, r2 = arr["geom_solref"][g1], arr["geom_solref"][g2]
            if r1[0] > 0 and r2[0] > 0:
                solref = mix * r1 + (1 - mix) * r2
            else:
                solref = np.minimum(r1, r2)
            solimp = mix * arr["geom_solimp"][g1] + (1 - mix) * arr["geom_solimp"][g2]
        out["condim"] = int(a["condim"]) if "condim" in a else condim
        if "friction" in a:
            out["friction"] = _floats(a["friction"], 5, [1, 1, 0.005, 0.0001, 0.0001])
        else:
            out["friction"] = np.array([fr[0], fr[0], fr[1], fr[2], fr[2]])
        out["solref"] = _floats(a["solref"], 2) if "solref" in a else solref
        out["solimp"] = _floats(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2.0]) if "solimp" in a else solimp
        return out

    def _parse_contacts(self, model: Model) -> None:
        arr = model.arrays
        contact = self.root.find("contact")
        excludes: set[tuple[int, int]] = set()
        pairs: list[dict[str, Any]] = []
        seen: set[tuple[int, int]] = set()
        if contact is not None:
            for e in contact.findall("exclude"):
                b1, b2 = model.id("body", e.get("body1")), model.id("body", e.get("body2"))  # type: ignore[arg-type]
                excludes.add((min(b1, b2), max(b1, b2)))
            for p in contact.findall("pair"):
                a = self.defaults.resolve("pair", p, None)
                g1, g2 = model.id("geom", a["geom1"]), model.id("geom", a["geom2"])
                pr = self._pair_params(model, g1, g2, a)
                pairs.append(pr)
                seen.add((min(g1, g2), max(g1, g2)))
        # dynamic pairs from contype/conaffinity (same-body, parent-child and excluded pairs filtered)
        ng = model.ngeom
        parent = arr["body_parentid"]
        for g1 in range(ng):
            for g2 in range(g1 + 1, ng):
                if (g1, g2) in seen:
                    continue
                ct1, ca1 = int(arr["geom_contype"][g1]), int(arr["geom_conaffinity"][g1])
                ct2, ca2 = int(arr["geom_contype"][g2]), int(arr["geom_conaffinity"][g2])
                if not ((ct1 & ca2) or (ct2 & ca1)):
                    continue
                b1, b2 = int(arr["geom_bodyid"][g1]), int(arr["geom_bodyid"][g2])
                if b1 == b2:
                    continue
                if (min(b1, b2), max(b1, b2)) in excludes:
                    continue
                # parent-child filter (bodies welded to the world count as the world body)
                if (parent[b1] == b2 and b2 != 0) or (parent[b2] == b1 and b1 != 0):
                    continue
                t1, t2 = int(arr["geom_type"][g1]), int(arr["geom_type"][g2])
                if t1 == GEOM.MESH or t2 == GEOM.MESH:
                    raise MjcfError("collidable mesh geoms are not supported")
                pairs.append(self._pair_params(model, g1, g2, {}))
        supported = {(GEOM.PLANE, GEOM.SPHERE), (GEOM.PLANE, GEOM.CAPSULE), (GEOM.SPHERE, GEOM.SPHERE)}
        for pr in pairs:
            key = (int(arr["geom_type"][pr["geom1"]]), int(arr["geom_type"][pr["geom2"]]))
            if key not in supported:
                raise MjcfError(f"collision between geom types {key} not supported")
            if pr["condim"] not in (1, 3):
                raise MjcfError(f"condim {pr['condim']} not supported (1 or 3)")
        n = len(pairs)
        arr["pair_geom1"] = np.array([p["geom1"] for p in pairs], dtype=np.int32).reshape(n)
        arr["pair_geom2"] = np.array([p["geom2"] for p in pairs], dtype=np.int32).reshape(n)
        arr["pair_dim"] = np.array([p["condim"] for p in pairs], dtype=np.int32).reshape(n)
        arr["pair_friction"] = np.array([p["friction"] for p in pairs], dtype=np.float64).reshape(n, 5)
        arr["pair_solref"] = np.array([p["solref"] for p in pairs], dtype=np.float64).reshape(n, 2)
        arr["pair_solimp"] = np.array([p["solimp"] for p in pairs], dtype=np.float64).reshape(n, 5)
        arr["pair_margin"] = np.array([p["margin"] for p in pairs], dtype=np.float64).reshape(n)
        arr["pair_gap"] = np.array([p["gap"] for p in pairs], dtype=np.float64).reshape(n)


# --------------------------------------------------------------------------------------------------
# qpos0-dependent constants (mj_setConst): dof_invweight0, body_invweight0, meaninertia
# --------------------------------------------------------------------------------------------------


def reference_kinematics(model: Model, qpos: np.ndarray) -> dict[str, np.ndarray]:
    """float64 numpy forward kinematics + composite-rigid-body mass matrix at ``qpos``.

    Used at compile time for the qpos0 constants and by tests as an independent check of the C oracle.
    """
    a = model.arrays
    nb, nv = model.nbody, model.nv
    xpos, xquat = np.zeros((nb, 3)), np.zeros((nb, 4))
    xquat[0, 0] = 1.0
    xanchor, xaxis = np.zeros((model.njnt, 3)), np.zeros((model.njnt, 3))
    for b in range(1, nb):
        p = a["body_parentid"][b]
        jn, ja = a["body_jntnum"][b], a["body_jntadr"][b]
        if jn == 1 and a["jnt_type"][ja] == JNT.FREE:
            qa = a["jnt_qposadr"][ja]
            pos = qpos[qa : qa + 3].copy()
            quat = qpos[qa + 3 : qa + 7] / np.linalg.norm(qpos[qa + 3 : qa + 7])
            xanchor[ja], xaxis[ja] = pos, a["jnt_axis"][ja]
        else:
            pos = xpos[p] + rot_vec(xquat[p], a["body_pos"][b])
            quat = quat_mul(xquat[p], a["body_quat"][b])
            for j in range(ja, ja + jn):
                anchor = rot_vec(quat, a["jnt_pos"][j]) + pos
                axis = rot_vec(quat, a["jnt_axis"][j])
                xanchor[j], xaxis[j] = anchor, axis
                ang = qpos[a["jnt_qposadr"][j]] - a["qpos0"][a["jnt_qposadr"][j]]
                quat = quat_mul(quat, axis_angle_quat(a["jnt_axis"][j], ang))
                pos = anchor - rot_vec(quat, a["jnt_pos"][j])
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    xmat = np.stack([quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, a["body_ipos"])
    ximat = np.stack([quat_to_mat(quat_mul(xquat[b], a["body_iquat"][b])) for b in range(nb)])
    # subtree com
    sub = a["body_mass"][:, None] * xipos
    for b in range(nb - 1, 0, -1):
        sub[a["body_parentid"][b]] += sub[b]
    subtree_com = sub / np.maximum(a["body_subtreemass"], MJ_MINVAL)[:, None]
    # cinert, cdof
    cinert = np.zeros((nb, 10))
    for b in range(1, nb):
        d = xipos[b] - subtree_com[a["body_rootid"][b]]
        m = a["body_mass"][b]
        full = ximat[b] @ np.diag(a["body_inertia"][b]) @ ximat[b].T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        cinert[b] = [full[0, 0], full[1, 1], full[2, 2], full[0, 1], full[0, 2], full[1, 2], m * d[0], m * d[1], m * d[2], m]
    cdof = np.zeros((nv, 6))
    for j in range(model.njnt):
        b = a["jnt_bodyid"][j]
        off = subtree_com[a["body_rootid"][b]] - xanchor[j]
        da = a["jnt_dofadr"][j]
        if a["jnt_type"][j] == JNT.FREE:
            for k in range(3):
                cdof[da + k, 3 + k] = 1.0
                ax = xmat[b][:, k]
                cdof[da + 3 + k, :3] = ax
                cdof[da + 3 + k, 3:] = np.cross(ax, off)
        else:
            cdof[da, :3] = xaxis[j]
            cdof[da, 3:] = np.cross(xaxis[j], off)
    # crb
    crb = cinert.copy()
    for b in range(nb - 1, 0, -1):
        if a["body_parentid"][b] > 0:
            crb[a["body_parentid"][b]] += crb[b]
    def mul_inert(i: np.ndarray, v: np.ndarray) -> np.ndarray:
        return np.array(
            [
                i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5],
                i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5],
                i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4],
                i[8] * v[1] - i[7] * v[2] + i[9] * v[3],
                i[6] * v[2] - i[8] * v[0] + i[9] * v[4],
                i[7] * v[0] - i[6] * v[1] + i[9] * v[5],
            ]
        )
    qm = np.zeros((nv, nv))
    for i in range(nv):
        buf = mul_inert(crb[a["dof_bodyid"][i]], cdof[i])
        j = i
        while j >= 0:
            qm[i, j] = qm[j, i] = np.dot(cdof[j], buf)
            j = a["dof_parentid"][j]
        qm[i, i] += a["dof_armature"][i]
    return dict(
        xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, ximat=ximat, subtree_com=subtree_com,
        cinert=cinert, cdof=cdof, qM=qm, xanchor=xanchor, xaxis=xaxis,
    )


def body_jacobian(model: Model, kin: dict[str, np.ndarray], body: int, point: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    a = model.arrays
    nv = model.nv
    jacp, jacr = np.zeros((3, nv)), np.zeros((3, nv))
    if body == 0:
        return jacp, jacr
    off = point - kin["subtree_com"][a["body_rootid"][body]]
    b = body
    while b != 0 and a["body_dofnum"][b] == 0:
        b = a["body_parentid"][b]
    if b == 0:
        return jacp, jacr
    i = a["body_dofadr"][b] + a["body_dofnum"][b] - 1
    while i >= 0:
        cd = kin["cdof"][i]
        jacr[:, i] = cd[:3]
        jacp[:, i] = cd[3:] + np.cross(cd[:3], off)
        i = a["dof_parentid"][i]
    return jacp, jacr


def set_const(model: Model) -> None:
    """qpos0-dependent constants, following MuJoCo's ``mj_setConst`` documentation."""
    a = model.arrays
    nv, nb = model.nv, model.nbody
    kin = reference_kinematics(model, a["qpos0"])
    qm = kin["qM"]
    model.meaninertia = float(np.trace(qm) / max(nv, 1)) if nv else 1.0
    minv = np.linalg.inv(qm) if nv else np.zeros((0, 0))
    dof_inv = np.diag(minv).copy()
    for j in range(model.njnt):
        da = a["jnt_dofadr"][j]
        if a["jnt_type"][j] == JNT.FREE:
            dof_inv[da : da + 3] = dof_inv[da : da + 3].mean()
            dof_inv[da + 3 : da + 6] = dof_inv[da + 3 : da + 6].mean()
    a["dof_invweight0"] = dof_inv
    binv = np.zeros((nb, 2))
    for b in range(1, nb):
        jp, jr = body_jacobian(model, kin, b, kin["xipos"][b])
        binv[b, 0] = np.trace(jp @ minv @ jp.T) / 3.0
        binv[b, 1] = np.trace(jr @ minv @ jr.T) / 3.0
    a["body_invweight0"] = binv


def _compile(root: ET.Element) -> Model:
    if root.tag != "mujoco":
        raise MjcfError("root element must be <mujoco>")
    for tag in ("equality", "tendon", "keyframe"):
        el = root.find(tag)
        if el is not None and len(el):
            if tag == "keyframe":
                continue
            raise MjcfError(f"<{tag}> not supported")
    model = _Compiler(root).compile()
    set_const(model)
    return model


def load_mjcf(path: str | Path) -> Model:
    return _compile(ET.parse(str(path)).getroot())


def load_mjcf_string(xml: str) -> Model:
    return _compile(ET.fromstring(xml))


# --------------------------------------------------------------------------------------------------
# compiled-model (de)serialisation: one .npz with every array plus a JSON side-car entry
# --------------------------------------------------------------------------------------------------


def save_model(model: Model, path: str | Path) -> None:
    """Writes a compiled model as ``.npz`` (arrays) with names / options / constants as a JSON string entry."""
    import json

    meta = dict(
        name=model.name, meaninertia=model.meaninertia, names=model.names, custom_numeric=model.custom_numeric,
        opt={k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in vars(model.opt).items()},
    )
    np.savez_compressed(str(path), __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **model.arrays)


def load_model(path: str | Path) -> Model:
    import json

    with np.load(str(path)) as z:
        meta = json.loads(bytes(z["__meta__"]).decode())
        arrays = {k: z[k] for k in z.files if k != "__meta__"}
    opt = Options(**{k: (np.array(v, dtype=np.float64) if k == "gravity" else v) for k, v in meta["opt"].items()})
    return Model(name=meta["name"], opt=opt, meaninertia=float(meta["meaninertia"]), arrays=arrays,
                 names={k: list(v) for k, v in meta["names"].items()}, custom_numeric=dict(meta["custom_numeric"]))
