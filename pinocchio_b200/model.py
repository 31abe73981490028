"""Host-side Model: the source of the flattened constants staged on the GPU.

Restates the parts of the reference ``ModelTpl`` the batched-dynamics path reads
(reference: include/pinocchio/multibody/model.hpp:97-205):

* ``Model.addJoint`` / ``appendBodyToJoint`` / ``addFrame`` bookkeeping
  (include/pinocchio/multibody/model.hxx:61-170, 406-413, 491-513),
* the sample models named by BASELINE.json
  (include/pinocchio/multibody/sample-models.hxx:58-139, 235-308, 310-403),
* the URDF -> Model rules for the two URDF configs
  (src/parsers/urdf/model.cpp:30-42, 67-294; include/pinocchio/parsers/urdf/model.hxx:205-372,
  437-480, 564-616),
* ``Model.flat()``: the POD arrays behind ``brbd_flat_model`` (include/pinocchio_b200.h).

Only numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# joint type tags == brbd_joint_type (include/pinocchio_b200.h)
JOINT_RX, JOINT_RY, JOINT_RZ = 0, 1, 2
JOINT_PX, JOINT_PY, JOINT_PZ = 3, 4, 5
JOINT_FREEFLYER, JOINT_SPHERICAL, JOINT_PLANAR = 6, 7, 8
JOINT_REVOLUTE_UNALIGNED, JOINT_PRISMATIC_UNALIGNED = 9, 10  # axis given per joint (joint-{revolute,prismatic}-unaligned.hpp)
# unbounded revolute joints, q = (cos, sin): joint-revolute-unbounded.hpp:121-235 (URDF "continuous", parsers/urdf/model.hxx:269-273)
JOINT_RUBX, JOINT_RUBY, JOINT_RUBZ, JOINT_REVOLUTE_UNBOUNDED_UNALIGNED = 11, 12, 13, 14
JOINT_UNIVERSE = -1

_JOINT_NAMES = {
    JOINT_RX: "JointModelRX", JOINT_RY: "JointModelRY", JOINT_RZ: "JointModelRZ",
    JOINT_PX: "JointModelPX", JOINT_PY: "JointModelPY", JOINT_PZ: "JointModelPZ",
    JOINT_FREEFLYER: "JointModelFreeFlyer", JOINT_SPHERICAL: "JointModelSpherical",
    JOINT_PLANAR: "JointModelPlanar", JOINT_UNIVERSE: "universe",
    JOINT_REVOLUTE_UNALIGNED: "JointModelRevoluteUnaligned", JOINT_PRISMATIC_UNALIGNED: "JointModelPrismaticUnaligned",
    JOINT_RUBX: "JointModelRUBX", JOINT_RUBY: "JointModelRUBY", JOINT_RUBZ: "JointModelRUBZ",
    JOINT_REVOLUTE_UNBOUNDED_UNALIGNED: "JointModelRevoluteUnboundedUnaligned",
}


def joint_is_unbounded(t: int) -> bool:
    return JOINT_RUBX <= t <= JOINT_REVOLUTE_UNBOUNDED_UNALIGNED


def joint_has_axis(t: int) -> bool:
    return t in (JOINT_REVOLUTE_UNALIGNED, JOINT_PRISMATIC_UNALIGNED, JOINT_REVOLUTE_UNBOUNDED_UNALIGNED)


def joint_nq(t: int) -> int:
    if joint_is_unbounded(t):
        return 2
    return 1 if (0 <= t <= JOINT_PZ or t >= JOINT_REVOLUTE_UNALIGNED) else (7 if t == JOINT_FREEFLYER else 4)


def joint_nv(t: int) -> int:
    return 1 if (0 <= t <= JOINT_PZ or t >= JOINT_REVOLUTE_UNALIGNED) else (6 if t == JOINT_FREEFLYER else 3)


# --------------------------------------------------------------------------------------------
# Spatial helpers (float64, same operation order as the reference where it matters)
# --------------------------------------------------------------------------------------------
class SE3:
    """(R, p); ``a * b`` = (Ra Rb, pa + Ra pb) — spatial/se3-tpl.hpp:314-317."""

    __slots__ = ("R", "p")

    def __init__(self, R=None, p=None):
        self.R = np.eye(3) if R is None else np.array(R, dtype=np.float64).reshape(3, 3)
        self.p = np.zeros(3) if p is None else np.array(p, dtype=np.float64).reshape(3)

    @staticmethod
    def Identity() -> "SE3":
        return SE3()

    def __mul__(self, o: "SE3") -> "SE3":
        return SE3(self.R @ o.R, self.p + self.R @ o.p)

    def copy(self) -> "SE3":
        return SE3(self.R.copy(), self.p.copy())


def quat_to_matrix(x, y, z, w) -> np.ndarray:
    """Eigen ``Quaternion::toRotationMatrix``."""
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1.0 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1.0 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
    ])


def angle_axis_matrix(angle: float, axis) -> np.ndarray:
    """Eigen ``AngleAxis::toRotationMatrix`` (used by sample-models.hxx:215)."""
    ax = np.asarray(axis, dtype=np.float64)
    s, c = math.sin(angle), math.cos(angle)
    cos1_axis = (1.0 - c) * ax
    R = np.empty((3, 3))
    tmp = cos1_axis[0] * ax[1]
    R[0, 1] = tmp - s * ax[2]
    R[1, 0] = tmp + s * ax[2]
    tmp = cos1_axis[0] * ax[2]
    R[0, 2] = tmp + s * ax[1]
    R[2, 0] = tmp - s * ax[1]
    tmp = cos1_axis[1] * ax[2]
    R[1, 2] = tmp - s * ax[0]
    R[2, 1] = tmp + s * ax[0]
    R[0, 0] = cos1_axis[0] * ax[0] + c
    R[1, 1] = cos1_axis[1] * ax[1] + c
    R[2, 2] = cos1_axis[2] * ax[2] + c
    return R


def sym3_rotate(d: np.ndarray, R: np.ndarray) -> np.ndarray:
    """R S R^T for a packed Symmetric3 — spatial/symmetric3.hpp:561-601."""
    L = np.array([[d[0] - d[5], d[1]], [d[1], d[2] - d[5]], [2 * d[3], d[4] + d[4]]])
    Y = R[1:3, :] @ L
    r = np.zeros(6)
    r[1] = Y[0, 0] * R[0, 0] + Y[0, 1] * R[0, 1]
    r[2] = Y[0, 0] * R[1, 0] + Y[0, 1] * R[1, 1]
    r[3] = Y[1, 0] * R[0, 0] + Y[1, 1] * R[0, 1]
    r[4] = Y[1, 0] * R[1, 0] + Y[1, 1] * R[1, 1]
    r[5] = Y[1, 0] * R[2, 0] + Y[1, 1] * R[2, 1]
    rr = np.array([-R[0, 0] * d[4] + R[0, 1] * d[3], -R[1, 0] * d[4] + R[1, 1] * d[3],
                   -R[2, 0] * d[4] + R[2, 1] * d[3]])
    r[0] = L[0, 0] + L[1, 1] - r[2] - r[5]
    r[0] += d[5]
    r[1] += rr[2]
    r[2] += d[5]
    r[3] -= rr[1]
    r[4] += rr[0]
    r[5] += d[5]
    return r


class Inertia:
    """(mass, lever, Symmetric3 packed xx,xy,yy,xz,yz,zz) — spatial/inertia.hpp:286-291."""

    __slots__ = ("mass", "lever", "sym")

    def __init__(self, mass=0.0, lever=None, inertia=None):
        self.mass = float(mass)
        self.lever = np.zeros(3) if lever is None else np.array(lever, dtype=np.float64).reshape(3)
        if inertia is None:
            self.sym = np.zeros(6)
        else:
            I = np.array(inertia, dtype=np.float64)
            if I.shape == (3, 3):  # Symmetric3(Matrix3): lower-triangle entries
                self.sym = np.array([I[0, 0], I[1, 0], I[1, 1], I[2, 0], I[2, 1], I[2, 2]])
            else:
                self.sym = I.reshape(6).copy()

    @staticmethod
    def Zero() -> "Inertia":
        return Inertia()

    def copy(self) -> "Inertia":
        return Inertia(self.mass, self.lever.copy(), self.sym.copy())

    def isZero(self, prec: float = 0.0) -> bool:
        return (abs(self.mass) <= prec and bool(np.all(np.abs(self.lever) <= prec))
                and bool(np.all(np.abs(self.sym) <= prec)))

    def se3Action(self, M: SE3) -> "Inertia":
        """aI = aXb.act(bI) — spatial/inertia.hpp:872-880."""
        return Inertia(self.mass, M.p + M.R @ self.lever, sym3_rotate(self.sym, M.R))

    def __iadd__(self, Yb: "Inertia") -> "Inertia":
        """spatial/inertia.hpp:659-673 (``__pequ__``)."""
        eps = np.finfo(np.float64).eps
        mab = self.mass + Yb.mass
        mab_inv = 1.0 / max(mab, eps)
        AB = self.lever - Yb.lever
        self.lever = self.lever * (self.mass * mab_inv)
        self.lever = self.lever + (Yb.mass * mab_inv) * Yb.lever
        self.sym = self.sym + Yb.sym
        m = self.mass * Yb.mass * mab_inv
        x, y, z = AB
        self.sym = self.sym + np.array([m * (y * y + z * z), -(m * x * y), m * (x * x + z * z),
                                        -(m * x * z), -(m * y * z), m * (x * x + y * y)])
        self.mass = mab
        return self

    def matrix(self) -> np.ndarray:
        """6x6 spatial inertia (linear first) — spatial/inertia.hpp:480-491."""
        m, c = self.mass, self.lever
        cx = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])
        d = self.sym
        I = np.array([[d[0], d[1], d[3]], [d[1], d[2], d[4]], [d[3], d[4], d[5]]])
        M = np.zeros((6, 6))
        M[:3, :3] = m * np.eye(3)
        M[3:, :3] = m * cx
        M[:3, 3:] = -m * cx
        M[3:, 3:] = I - m * cx @ cx
        return M


@dataclass
class Frame:
    name: str
    parentJoint: int
    parentFrame: int
    placement: SE3
    type: str
    inertia: Inertia = field(default_factory=Inertia.Zero)


class Model:
    """Kinematic tree + constants; joint 0 is the universe (model.hxx:35-58)."""

    def __init__(self):
        self.name = ""
        self.njoints = 1
        self.nq = 0
        self.nv = 0
        self.joint_types: List[int] = [JOINT_UNIVERSE]
        self.parents: List[int] = [0]
        self.names: List[str] = ["universe"]
        self.idx_qs: List[int] = [0]
        self.idx_vs: List[int] = [0]
        self.nqs: List[int] = [0]
        self.nvs: List[int] = [0]
        self.jointPlacements: List[SE3] = [SE3.Identity()]
        self.axes: List[np.ndarray] = [np.zeros(3)]  # unit axis of the unaligned joints, zeros elsewhere
        self.inertias: List[Inertia] = [Inertia.Zero()]
        self.armature = np.zeros(0)
        self.lowerPositionLimit = np.zeros(0)
        self.upperPositionLimit = np.zeros(0)
        self.gravity = np.array([0.0, 0.0, -9.81])  # model.hxx:40 (linear part)
        self.frames: List[Frame] = [Frame("universe", 0, 0, SE3.Identity(), "FIXED_JOINT")]

    # -- model.hxx:61-170 ---------------------------------------------------------------
    def addJoint(self, parent: int, joint_type: int, placement: SE3, name: str,
                 min_config=None, max_config=None, axis=None) -> int:
        if not (0 <= parent < self.njoints):
            raise ValueError("The index of the parent joint is not valid.")
        if joint_type not in _JOINT_NAMES or joint_type == JOINT_UNIVERSE:
            raise ValueError(f"unsupported joint type tag {joint_type}")
        jid = self.njoints
        self.njoints += 1
        nqj, nvj = joint_nq(joint_type), joint_nv(joint_type)
        self.joint_types.append(joint_type)
        self.parents.append(parent)
        self.names.append(name)
        self.idx_qs.append(self.nq)
        self.idx_vs.append(self.nv)
        self.nqs.append(nqj)
        self.nvs.append(nvj)
        self.jointPlacements.append(placement.copy())
        if joint_has_axis(joint_type):
            ax = np.asarray(axis, dtype=np.float64)
            if ax.shape != (3,) or not np.linalg.norm(ax) > 0:
                raise ValueError("an unaligned joint needs a non-zero axis")
            self.axes.append(ax / np.linalg.norm(ax))  # the reference normalises too (joint-revolute-unaligned.hpp:611-615)
        else:
            self.axes.append(np.zeros(3))
        self.inertias.append(Inertia.Zero())
        self.nq += nqj
        self.nv += nvj
        self.armature = np.concatenate([self.armature, np.zeros(nvj)])
        lo = np.full(nqj, -np.inf) if min_config is None else np.asarray(min_config, dtype=np.float64)
        hi = np.full(nqj, np.inf) if max_config is None else np.asarray(max_config, dtype=np.float64)
        self.lowerPositionLimit = np.concatenate([self.lowerPositionLimit, lo])
        self.upperPositionLimit = np.concatenate([self.upperPositionLimit, hi])
        return jid

    # -- model.hxx:406-413 ----------------------------------------------------------------
    def appendBodyToJoint(self, joint_index: int, Y: Inertia, body_placement: Optional[SE3] = None):
        M = SE3.Identity() if body_placement is None else body_placement
        self.inertias[joint_index] += Y.se3Action(M)

    # -- model.hxx:491-513 ----------------------------------------------------------------
    def addFrame(self, frame: Frame, append_inertia: bool = True) -> int:
        for k, f in enumerate(self.frames):
            if f.name == frame.name and f.type == frame.type:
                return k
        self.frames.append(frame)
        if append_inertia:
            self.inertias[frame.parentJoint] += frame.inertia.se3Action(frame.placement)
        return len(self.frames) - 1

    def getFrameId(self, name: str, types=("JOINT", "FIXED_JOINT", "BODY")) -> int:
        for k, f in enumerate(self.frames):
            if f.name == name and f.type in types:
                return k
        return len(self.frames)

    def addJointFrame(self, joint_index: int, previous_frame: int = -1) -> int:  # model.hxx:213-232
        if previous_frame < 0:
            previous_frame = self.getFrameId(self.names[self.parents[joint_index]], ("JOINT", "FIXED_JOINT"))
        return self.addFrame(Frame(self.names[joint_index], joint_index, previous_frame, SE3.Identity(), "JOINT"))

    def addBodyFrame(self, body_name: str, parentJoint: int, placement: Optional[SE3] = None,
                     parentFrame: int = -1) -> int:  # model.hxx:416-430
        if parentFrame < 0:
            parentFrame = self.getFrameId(self.names[parentJoint], ("JOINT", "FIXED_JOINT"))
        M = SE3.Identity() if placement is None else placement
        return self.addFrame(Frame(body_name, parentJoint, parentFrame, M, "BODY"))

    def getJointId(self, name: str) -> int:
        return self.names.index(name) if name in self.names else self.njoints

    # -- topology tables (multibody/data.hxx:197-315) ---------------------------------------
    def nvSubtree(self) -> List[int]:
        last = [-1] * self.njoints
        nvs = [0] * self.njoints
        for i in range(self.njoints - 1, -1, -1):
            if last[i] == -1:
                last[i] = i
            p = self.parents[i]
            last[p] = max(last[i], last[p])
            lc = last[i]
            nvs[i] = 0 if lc == 0 else self.idx_vs[lc] + self.nvs[lc] - (self.idx_vs[i] if i else 0)
        return nvs

    def depth(self) -> List[int]:
        d = [0] * self.njoints
        for i in range(1, self.njoints):
            d[i] = d[self.parents[i]] + 1
        return d

    def is_compact(self) -> bool:
        """CRBAChecker: descendants of i are stored contiguously after i (crba.hxx:573-595)."""
        def is_desc(j, root):
            while j > 0:
                if j == root:
                    return True
                j = self.parents[j]
            return root == 0
        for i in range(1, self.njoints - 1):
            k = i + 1
            while k < self.njoints and is_desc(k, i):
                k += 1
            for kk in range(k, self.njoints):
                if is_desc(kk, i):
                    return False
        return True

    # -- flattening ---------------------------------------------------------------------------
    def flat(self) -> Dict[str, np.ndarray]:
        n = self.njoints
        placement = np.zeros((n, 12))
        inertia = np.zeros((n, 10))
        for i in range(n):
            placement[i, :9] = self.jointPlacements[i].R.reshape(9)
            placement[i, 9:] = self.jointPlacements[i].p
            inertia[i, 0] = self.inertias[i].mass
            inertia[i, 1:4] = self.inertias[i].lever
            inertia[i, 4:] = self.inertias[i].sym
        return {
            "njoints": n, "nq": self.nq, "nv": self.nv,
            "parents": np.array(self.parents, dtype=np.int32),
            "joint_type": np.array(self.joint_types, dtype=np.int32),
            "idx_q": np.array(self.idx_qs, dtype=np.int32),
            "idx_v": np.array(self.idx_vs, dtype=np.int32),
            "placement": np.ascontiguousarray(placement),
            "inertia": np.ascontiguousarray(inertia),
            "armature": np.ascontiguousarray(self.armature, dtype=np.float64),
            "gravity": np.array(self.gravity, dtype=np.float64),
            "axis": np.ascontiguousarray(np.array(self.axes, dtype=np.float64).reshape(-1)),
        }

    # -- (de)serialisation of the flattened model: fixtures that travel to the GPU box ---------
    def to_json(self) -> str:
        f = self.flat()
        d = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in f.items()}
        # repr() round-trips float64 exactly
        d["names"] = self.names
        d["name"] = self.name
        d["lowerPositionLimit"] = [repr(float(x)) for x in self.lowerPositionLimit]
        d["upperPositionLimit"] = [repr(float(x)) for x in self.upperPositionLimit]
        d["placement"] = [[repr(float(x)) for x in row] for row in f["placement"]]
        d["inertia"] = [[repr(float(x)) for x in row] for row in f["inertia"]]
        d["armature"] = [repr(float(x)) for x in f["armature"]]
        d["gravity"] = [repr(float(x)) for x in f["gravity"]]
        d["axis"] = [repr(float(x)) for x in f["axis"]]
        return json.dumps(d, indent=1)

    @staticmethod
    def from_json(text: str) -> "Model":
        d = json.loads(text)
        m = Model()
        m.name = d.get("name", "")
        n = d["njoints"]
        placement = np.array([[float(x) for x in row] for row in d["placement"]])
        inertia = np.array([[float(x) for x in row] for row in d["inertia"]])
        lo = np.array([float(x) for x in d["lowerPositionLimit"]])
        hi = np.array([float(x) for x in d["upperPositionLimit"]])
        axes = np.array([float(x) for x in d["axis"]]).reshape(-1, 3) if "axis" in d else np.zeros((n, 3))
        for i in range(1, n):
            t = d["joint_type"][i]
            iq = d["idx_q"][i]
            jid = m.addJoint(d["parents"][i], t, SE3(placement[i, :9].reshape(3, 3), placement[i, 9:]),
                             d["names"][i], lo[iq:iq + joint_nq(t)], hi[iq:iq + joint_nq(t)], axis=axes[i])
            m.inertias[jid] = Inertia(inertia[i, 0], inertia[i, 1:4], inertia[i, 4:])
        m.inertias[0] = Inertia(inertia[0, 0], inertia[0, 1:4], inertia[0, 4:])
        m.armature = np.array([float(x) for x in d["armature"]])
        m.gravity = np.array([float(x) for x in d["gravity"]])
        assert m.nq == d["nq"] and m.nv == d["nv"]
        return m

    def __repr__(self):
        return f"Model(name={self.name!r}, njoints={self.njoints}, nq={self.nq}, nv={self.nv})"


# --------------------------------------------------------------------------------------------
# Sample models — include/pinocchio/multibody/sample-models.hxx
# --------------------------------------------------------------------------------------------
def _add_joint_and_body(model: Model, jtype: int, parent_name: str, name: str,
                        placement: SE3, inertia: Inertia, lo=-3.14, hi=3.14) -> int:
    """details::addJointAndBody (sample-models.hxx:20-55) with deterministic limits."""
    nqj = joint_nq(jtype)
    idx = model.addJoint(model.getJointId(parent_name), jtype, placement, name + "_joint",
                         np.full(nqj, lo), np.full(nqj, hi))
    model.addJointFrame(idx)
    model.appendBodyToJoint(idx, inertia, SE3.Identity())
    model.addBodyFrame(name + "_body", idx)
    return idx


def _add_manipulator(model: Model, root_joint_idx: int = 0, Mroot: Optional[SE3] = None, pre: str = ""):
    """details::addManipulator (sample-models.hxx:58-139): RX, RY, RZ, RY, RX, RY."""
    Mroot = SE3.Identity() if Mroot is None else Mroot
    Marm = SE3(np.eye(3), [0.0, 0.0, 1.0])
    Id4 = SE3.Identity()
    Ijoint = Inertia(0.1, np.zeros(3), np.eye(3) * 0.01)
    Iarm = Inertia(1.0, [0.0, 0.0, 0.5], np.eye(3))
    spec = [
        (JOINT_RX, "shoulder1", Mroot, Ijoint), (JOINT_RY, "shoulder2", Id4, Ijoint),
        (JOINT_RZ, "shoulder3", Id4, Iarm), (JOINT_RY, "elbow", Marm, Iarm),
        (JOINT_RX, "wrist1", Marm, Ijoint), (JOINT_RY, "wrist2", Id4, Iarm),
    ]
    parent_name = model.names[root_joint_idx]
    for jt, nm, M, Y in spec:
        jid = _add_joint_and_body(model, jt, parent_name, pre + nm, M, Y)
        model.inertias[jid] = Y.copy()  # sample-models.hxx:86,91,95,100,106,127
        parent_name = model.names[jid]


def buildSampleModelManipulator() -> Model:
    """buildModels::manipulator — nq = nv = 6 (unittest/sample-models.cpp:70-71)."""
    m = Model()
    m.name = "manipulator"
    _add_manipulator(m)
    return m


def buildSampleModelHumanoid(usingFF: bool = True) -> Model:
    """buildModels::humanoid — nq = 35, nv = 34 with a free-flyer (unittest/sample-models.cpp:92-93)."""
    if not usingFF:
        raise ValueError("composite (translation + sphericalZYX) root joint is out of scope")
    m = Model()
    m.name = "humanoid"
    Ijoint = Inertia(0.1, np.zeros(3), np.eye(3) * 0.01)
    Iarm = Inertia(1.0, [0.0, 0.0, 0.5], np.eye(3))
    lo7 = np.array([-np.inf] * 3 + [-1.0] * 4)
    hi7 = np.array([np.inf] * 3 + [1.0] * 4)
    ffidx = m.addJoint(0, JOINT_FREEFLYER, SE3.Identity(), "root_joint", lo7, hi7)
    m.appendBodyToJoint(ffidx, Ijoint)
    m.addJointFrame(ffidx)
    pi = math.pi
    Rx = angle_axis_matrix(pi, [1.0, 0.0, 0.0])
    _add_manipulator(m, ffidx, SE3(Rx, [0.0, -0.2, -0.1]), "rleg_")
    _add_manipulator(m, ffidx, SE3(Rx, [0.0, 0.2, -0.1]), "lleg_")
    m.jointPlacements[7].R = angle_axis_matrix(pi / 2, [0.0, 1.0, 0.0])
    m.jointPlacements[13].R = angle_axis_matrix(pi / 2, [0.0, 1.0, 0.0])
    lim = (np.array([-3.14]), np.array([3.14]))

    def add(parent, jt, M, name, Y):
        idx = m.addJoint(parent, jt, M, name + "_joint", *lim)
        m.appendBodyToJoint(idx, Y)
        m.addJointFrame(idx)
        m.addBodyFrame(name + "_body", idx)
        return idx

    idx = add(ffidx, JOINT_RX, SE3.Identity(), "chest1", Ijoint)
    idx = add(idx, JOINT_RY, SE3.Identity(), "chest2", Iarm)
    chest = idx
    idx = add(idx, JOINT_RX, SE3(np.eye(3), [0.0, 0.0, 1.0]), "head1", Ijoint)
    idx = add(idx, JOINT_RY, SE3.Identity(), "head2", Iarm)
    _add_manipulator(m, chest, SE3(Rx, [0.0, -0.3, 1.0]), "rarm_")
    _add_manipulator(m, chest, SE3(Rx, [0.0, 0.3, 1.0]), "larm_")
    return m


class _Rng:
    """Seeded stand-in for the libc ``rand()`` / Eigen ``::Random`` draws of the reference.

    ``humanoidRandom`` is not reproducible bit-for-bit across Eigen/libc versions (SURVEY §8c
    caveat); we keep its topology and distributions and fix the stream with our own seed.
    """

    def __init__(self, seed: int):
        self.rs = np.random.RandomState(seed)

    def unit(self) -> float:  # rand()/RAND_MAX
        return float(self.rs.random_sample())

    def sym(self, n: int) -> np.ndarray:  # Eigen Random(): uniform [-1, 1]
        return self.rs.uniform(-1.0, 1.0, size=n)

    def se3(self) -> SE3:  # SE3::Random, se3-tpl.hpp:159-167 + math/quaternion.hpp:115-137
        u1, u2, u3 = self.unit(), self.unit(), self.unit()
        m1, m2 = math.sqrt(1.0 - u1), math.sqrt(u1)
        s2, c2 = math.sin(2 * math.pi * u2), math.cos(2 * math.pi * u2)
        s3, c3 = math.sin(2 * math.pi * u3), math.cos(2 * math.pi * u3)
        w, x, y, z = m1 * s2, m1 * c2, m2 * s3, m2 * c3
        return SE3(quat_to_matrix(x, y, z, w), self.sym(3))

    def inertia(self) -> Inertia:  # Inertia::Random, inertia.hpp:362-367; symmetric3.hpp:291-302
        mass = float(self.sym(1)[0]) + 1.0
        lever = self.sym(3)
        a, b, c, d, e, f = self.sym(6)
        sym = np.array([a * a + b * b + d * d, a * b + b * c + d * e, b * b + c * c + e * e,
                        a * d + b * e + d * f, b * d + c * e + e * f, d * d + e * e + f * f])
        return Inertia(mass, lever, sym)


def buildSampleModelHumanoidRandom(usingFF: bool = True, seed: int = 0) -> Model:
    """buildModels::humanoidRandom (sample-models.hxx:235-308): FF + 26 revolute, nq 33 / nv 32."""
    if not usingFF:
        raise ValueError("composite root joint is out of scope")
    rng = _Rng(seed)
    m = Model()
    m.name = "humanoidRandom"

    def add(jt, parent_name, name, placement=None, lo=None, hi=None):
        M = rng.se3() if placement is None else placement
        nqj = joint_nq(jt)
        # setRandomLimits: qmin = Random - 1, qmax = Random + 1 (sample-models.hxx:40-43)
        lo_ = rng.sym(nqj) - 1.0 if lo is None else lo
        hi_ = rng.sym(nqj) + 1.0 if hi is None else hi
        idx = m.addJoint(m.getJointId(parent_name), jt, M, name + "_joint", lo_, hi_)
        m.addJointFrame(idx)
        m.appendBodyToJoint(idx, rng.inertia(), SE3.Identity())
        m.addBodyFrame(name + "_body", idx)
        return idx

    add(JOINT_FREEFLYER, "universe", "root", SE3.Identity())
    m.lowerPositionLimit[3:7] = -1.0
    m.upperPositionLimit[3:7] = 1.0
    chain = [JOINT_RX, JOINT_RY, JOINT_RZ, JOINT_RY, JOINT_RY, JOINT_RX]
    for limb in ("lleg", "rleg"):
        parent = "root_joint"
        for k, jt in enumerate(chain):
            add(jt, parent, f"{limb}{k + 1}")
            parent = f"{limb}{k + 1}_joint"
    add(JOINT_RY, "root_joint", "torso1")
    add(JOINT_RZ, "torso1_joint", "chest")
    for limb in ("rarm", "larm"):
        parent = "chest_joint"
        for k, jt in enumerate(chain):
            add(jt, parent, f"{limb}{k + 1}")
            parent = f"{limb}{k + 1}_joint"
    return m


# --------------------------------------------------------------------------------------------
# URDF -> Model (minimal loader restating the reference's rules)
# --------------------------------------------------------------------------------------------
def _rpy_to_quat(r: float, p: float, y: float) -> Tuple[float, float, float, float]:
    """urdfdom_headers ``Rotation::setFromRPY`` (urdf_model/pose.h, urdfdom 4.0.1) + normalize()."""
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    yy = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(x * x + yy * yy + z * z + w * w)
    if s == 0.0:
        return 0.0, 0.0, 0.0, 1.0
    return x / s, yy / s, z / s, w / s


def _parse_origin(elem) -> SE3:
    if elem is None:
        return SE3.Identity()
    xyz = [float(t) for t in elem.get("xyz", "0 0 0").split()]
    rpy = [float(t) for t in elem.get("rpy", "0 0 0").split()]
    x, y, z, w = _rpy_to_quat(*rpy)
    return SE3(quat_to_matrix(x, y, z, w), xyz)  # src/parsers/urdf/utils.cpp:10-15


def _parse_inertial(link) -> Inertia:
    """convertFromUrdf(urdf::Inertial) — src/parsers/urdf/model.cpp:30-42."""
    ine = link.find("inertial")
    if ine is None:
        return Inertia.Zero()
    M = _parse_origin(ine.find("origin"))
    mass = float(ine.find("mass").get("value"))
    I = ine.find("inertia")
    g = lambda k: float(I.get(k, "0"))
    Im = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
    return Inertia(mass, M.p, M.R @ Im @ M.R.T)


def _axis_tag(axis: np.ndarray) -> Optional[int]:
    """extractCartesianAxis (parsers/urdf/model.hxx:564-574): Eigen isApprox at 1e-12."""
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1.0
        if np.linalg.norm(axis - e) <= 1e-12 * min(np.linalg.norm(axis), 1.0):
            return k
    return None


def buildModelFromUrdf(path_or_xml: str, root_joint: Optional[int] = None,
                       root_joint_name: str = "root_joint") -> Model:
    """pinocchio::urdf::buildModel(filename[, JointModelFreeFlyer()], model).

    Child links are visited in urdfdom's order (children attached while iterating the
    name-sorted joint map) and depth-first (src/parsers/urdf/model.cpp:67-294); fixed joints
    merge their body into the parent joint; axis-aligned revolute / continuous / prismatic axes
    map to RX/RY/RZ / PX/PY/PZ, any other axis to RevoluteUnaligned / PrismaticUnaligned (which the engine re-frames
    into RZ / PZ at brbd_model_create, model_build.hpp).
    ``continuous`` joints become RUBX / RUBY / RUBZ / RevoluteUnboundedUnaligned (nq = 2, q = (cos, sin)) as in
    parsers/urdf/model.hxx:269-273; mimic joints are rejected with the reference's wording.
    """
    text = path_or_xml
    if not path_or_xml.lstrip().startswith("<"):
        with open(path_or_xml, "r") as fh:
            text = fh.read()
    robot = ET.fromstring(text)
    links = {l.get("name"): l for l in robot.findall("link")}
    joints = {j.get("name"): j for j in robot.findall("joint")}
    children: Dict[str, List[Tuple[str, str]]] = {n: [] for n in links}
    has_parent = set()
    for jname in sorted(joints):  # std::map<std::string, JointSharedPtr> iteration order
        j = joints[jname]
        p, c = j.find("parent").get("link"), j.find("child").get("link")
        children[p].append((jname, c))
        has_parent.add(c)
    roots = [n for n in links if n not in has_parent]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link")
    root = roots[0]

    model = Model()
    model.name = robot.get("name", "")
    body_frame: Dict[str, int] = {}

    # addRootJoint — parsers/urdf/model.hxx:205-211 / 601-616
    Yroot = _parse_inertial(links[root])
    if root_joint is None:
        body_frame[root] = model.addFrame(Frame(root, 0, 0, SE3.Identity(), "BODY", Yroot))
    else:
        lo = hi = None
        idx = model.addJoint(0, root_joint, SE3.Identity(), root_joint_name, lo, hi)
        jf = model.addJointFrame(idx, 0)
        body_frame[root] = _urdf_append_body(model, jf, Yroot, SE3.Identity(), root)

    def visit(link_name: str):
        for jname, child in children[link_name]:
            j = joints[jname]
            jtype = j.get("type")
            parent_fid = body_frame[link_name]
            frame = model.frames[parent_fid]
            placement = _parse_origin(j.find("origin"))
            Y = _parse_inertial(links[child])
            ax_el = j.find("axis")
            axis = np.array([float(t) for t in ax_el.get("xyz").split()]) if ax_el is not None else np.array([1.0, 0, 0])
            lim = j.find("limit")
            if jtype != "fixed" and j.find("mimic") is not None:
                # mimic joints (parsers/urdf/model.hxx:298-319) are not supported by the batched engine
                raise ValueError(f"Cannot mimic this type. Only revolute, prismatic and helicoidal can be mimicked (joint {jname})")
            if jtype == "fixed":
                # addFixedJointAndBody — parsers/urdf/model.hxx:347-361
                M = frame.placement * placement
                fid = model.addFrame(Frame(jname, frame.parentJoint, parent_fid, M, "FIXED_JOINT", Y))
                body_frame[child] = model.addBodyFrame(child, frame.parentJoint, M, fid)
            elif jtype in ("revolute", "prismatic"):
                k = _axis_tag(axis)
                if k is None:  # CartesianAxis AXIS_UNALIGNED: Joint*Unaligned(axis.normalized()) — parsers/urdf/model.hxx:437-480
                    tag = JOINT_REVOLUTE_UNALIGNED if jtype == "revolute" else JOINT_PRISMATIC_UNALIGNED
                else:
                    tag = (JOINT_RX if jtype == "revolute" else JOINT_PX) + k
                lo = [float(lim.get("lower", "0"))] if lim is not None else None
                hi = [float(lim.get("upper", "0"))] if lim is not None else None
                jid = model.addJoint(frame.parentJoint, tag, frame.placement * placement, jname, lo, hi,
                                     axis=axis if k is None else None)
                jf = model.addJointFrame(jid, parent_fid)
                body_frame[child] = _urdf_append_body(model, jf, Y, SE3.Identity(), child)
            elif jtype == "continuous":
                k = _axis_tag(axis)
                tag = JOINT_REVOLUTE_UNBOUNDED_UNALIGNED if k is None else JOINT_RUBX + k
                # an unbounded joint has no position limits; (cos, sin) is sampled on the circle whatever the bounds
                jid = model.addJoint(frame.parentJoint, tag, frame.placement * placement, jname, [-1.01, -1.01], [1.01, 1.01],
                                     axis=axis if k is None else None)
                jf = model.addJointFrame(jid, parent_fid)
                body_frame[child] = _urdf_append_body(model, jf, Y, SE3.Identity(), child)
            elif jtype in ("floating", "planar"):
                tag = JOINT_FREEFLYER if jtype == "floating" else JOINT_PLANAR
                jid = model.addJoint(frame.parentJoint, tag, frame.placement * placement, jname)
                jf = model.addJointFrame(jid, parent_fid)
                body_frame[child] = _urdf_append_body(model, jf, Y, SE3.Identity(), child)
            else:
                raise ValueError(f"The type of joint {jname} ({jtype}) is not supported.")
            visit(child)

    visit(root)
    if root_joint == JOINT_FREEFLYER:
        # randomConfiguration needs finite bounds on the translation; the quaternion part is
        # sampled on S^3 regardless (special-orthogonal.hpp:683-688)
        pass
    return model


def _urdf_append_body(model: Model, fid: int, Y: Inertia, placement: SE3, body_name: str) -> int:
    """UrdfVisitor::appendBodyToJoint — parsers/urdf/model.hxx:363-384."""
    frame = model.frames[fid]
    p = frame.placement * placement
    if not Y.isZero(0.0):
        model.appendBodyToJoint(frame.parentJoint, Y, p)
    return model.addBodyFrame(body_name, frame.parentJoint, p, fid)
