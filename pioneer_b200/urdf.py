"""Host-side URDF flattener: robot XML -> struct-of-arrays kinematic-chain tables.

This replaces what the reference delegates to Bullet's URDF importer at every reset
(reference: pioneer/envs/bullet/bullet_env.py:101-148 ``load_scene`` -> ``loadURDF`` /
``getJointInfo``).  It is run ONCE; the tables are uploaded to ``__constant__`` memory by
``pnr_create`` (include/pioneer_b200.h).

Conventions restated from the URDF specification (Bullet itself is not available here):
  * child_link_frame = parent_link_frame * T(origin xyz, rpy) * Rot(axis, q)
  * rpy is fixed-axis roll/pitch/yaw: R = Rz(yaw) * Ry(pitch) * Rx(roll)
  * joint/link index i = i-th joint in depth-first order from the root link; link i is the
    child link of joint i (reference: bullet_env.py:116-122 iterates ``getNumJoints``)
  * only ``revolute`` joints become degrees of freedom, ``fixed`` joints are ignored, anything
    else raises AssertionError (reference: bullet_env.py:141-146)
  * the position reported for a link is its centre of mass, i.e. the link frame origin moved by
    the ``<inertial><origin>`` (reference: bullet_scene.py:58 reads ``link_world_position``)
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

DEFAULT_URDF = os.path.join(os.path.dirname(__file__), "assets", "pioneer_reach_6dof.urdf")

JOINT_REVOLUTE = 0  # pybullet.JOINT_REVOLUTE
JOINT_FIXED = 4     # pybullet.JOINT_FIXED


def find_unique(parent: ET.Element, tag_or_path: str,
                attr_name: Optional[str] = None, attr_value: Optional[str] = None) -> ET.Element:
    """Unique-child lookup (same contract as reference pioneer/xml_util.py:5-12)."""
    found = parent.findall(tag_or_path)
    if attr_name is not None:
        found = [e for e in found if e.attrib.get(attr_name) == attr_value]
    assert len(found) == 1, f"expected exactly one <{tag_or_path}>, got {len(found)}"
    return found[0]


def _vec(text: Optional[str], n: int, default: float = 0.0) -> np.ndarray:
    if text is None:
        return np.full(n, default, dtype=np.float64)
    vals = [float(x) for x in text.split()]
    assert len(vals) == n, f"expected {n} numbers, got {text!r}"
    return np.array(vals, dtype=np.float64)


def rpy_to_matrix(rpy) -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


def axis_angle_matrix(axis, q: float) -> np.ndarray:
    k = np.asarray(axis, dtype=np.float64)
    c, s = math.cos(q), math.sin(q)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=np.float64)
    return np.eye(3) * c + s * K + (1.0 - c) * np.outer(k, k)


@dataclass
class UrdfLink:
    name: str
    mass: float = 0.0
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    com_xyz: np.ndarray = field(default_factory=lambda: np.zeros(3))
    com_rot: np.ndarray = field(default_factory=lambda: np.eye(3))
    capsules: List[Tuple[float, np.ndarray, np.ndarray]] = field(default_factory=list)


@dataclass
class UrdfJoint:
    """One joint as Bullet's ``getJointInfo`` would describe it (bullet_bindings.py:11-27)."""
    index: int
    name: str
    joint_type: int
    parent_link: str
    child_link: str
    origin_xyz: np.ndarray
    origin_rot: np.ndarray
    axis: np.ndarray
    lower: float
    upper: float
    effort: float
    max_velocity: float
    damping: float
    friction: float
    parent_index: int


@dataclass
class ChainModel:
    """Serial chain with fixed joints folded into their parents (all float64)."""
    dof: int
    joint_names: List[str]
    body_names: List[str]            # child link of each revolute joint
    axis: np.ndarray                 # [dof,3] unit joint axis in the joint frame
    origin_xyz: np.ndarray           # [dof,3] previous moving frame -> joint frame translation
    origin_rot: np.ndarray           # [dof,3,3] previous moving frame -> joint frame rotation
    base_xyz: np.ndarray             # [3] world -> first joint's parent frame (static part)
    base_rot: np.ndarray             # [3,3]
    tip_name: str
    tip_xyz: np.ndarray              # [3] tracked point in the last moving frame
    lower: np.ndarray                # [dof]
    upper: np.ndarray
    effort: np.ndarray
    max_velocity: np.ndarray
    damping: np.ndarray
    friction: np.ndarray
    body_mass: np.ndarray            # [dof]   composite rigid body per moving frame
    body_com: np.ndarray             # [dof,3] in the moving frame
    body_inertia: np.ndarray         # [dof,3,3] about the COM, moving-frame axes
    capsules: List[Tuple[int, float, np.ndarray, np.ndarray]]  # (body, radius, p0, p1) moving frame
    joints: List[UrdfJoint]          # every joint in Bullet index order (for the Scene mirror)
    links: Dict[str, UrdfLink]
    root_link: str
    robot_name: str

    # ---- float64 forward kinematics on the flattened tables (host utility, not the hot path)
    def tip_position(self, q) -> np.ndarray:
        p = self.tip_xyz.copy()
        for j in range(self.dof - 1, -1, -1):
            p = self.origin_xyz[j] + self.origin_rot[j] @ (axis_angle_matrix(self.axis[j], float(q[j])) @ p)
        return self.base_xyz + self.base_rot @ p


def parse_urdf(path: str = DEFAULT_URDF):
    root = ET.parse(path).getroot()
    assert root.tag == "robot", f"not a URDF: root tag <{root.tag}>"
    links: Dict[str, UrdfLink] = {}
    for le in root.findall("link"):
        link = UrdfLink(name=le.attrib["name"])
        ine = le.find("inertial")
        if ine is not None:
            link.mass = float(find_unique(ine, "mass").attrib["value"])
            ie = find_unique(ine, "inertia").attrib
            ixx, ixy, ixz = float(ie["ixx"]), float(ie.get("ixy", 0)), float(ie.get("ixz", 0))
            iyy, iyz, izz = float(ie["iyy"]), float(ie.get("iyz", 0)), float(ie["izz"])
            link.inertia = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=np.float64)
            oe = ine.find("origin")
            if oe is not None:
                link.com_xyz = _vec(oe.attrib.get("xyz"), 3)
                link.com_rot = rpy_to_matrix(_vec(oe.attrib.get("rpy"), 3))
        for ce in le.findall("collision"):
            cap = ce.find("geometry/capsule")
            if cap is not None:  # repo-specific extension, see assets/pioneer_reach_6dof.urdf
                link.capsules.append((float(cap.attrib["radius"]),
                                      _vec(cap.attrib["from"], 3), _vec(cap.attrib["to"], 3)))
        assert link.name not in links, f"duplicate link {link.name}"
        links[link.name] = link

    raw = []
    for je in root.findall("joint"):
        jt = je.attrib["type"]
        oe, ae, lim, dyn = je.find("origin"), je.find("axis"), je.find("limit"), je.find("dynamics")
        axis = _vec(ae.attrib.get("xyz"), 3) if ae is not None else np.array([1.0, 0.0, 0.0])
        raw.append(dict(
            name=je.attrib["name"], type=jt,
            parent=find_unique(je, "parent").attrib["link"], child=find_unique(je, "child").attrib["link"],
            xyz=_vec(oe.attrib.get("xyz"), 3) if oe is not None else np.zeros(3),
            rot=rpy_to_matrix(_vec(oe.attrib.get("rpy"), 3)) if oe is not None else np.eye(3),
            axis=axis,
            lower=float(lim.attrib.get("lower", 0.0)) if lim is not None else 0.0,
            upper=float(lim.attrib.get("upper", -1.0)) if lim is not None else -1.0,
            effort=float(lim.attrib.get("effort", 0.0)) if lim is not None else 0.0,
            velocity=float(lim.attrib.get("velocity", 0.0)) if lim is not None else 0.0,
            damping=float(dyn.attrib.get("damping", 0.0)) if dyn is not None else 0.0,
            friction=float(dyn.attrib.get("friction", 0.0)) if dyn is not None else 0.0))
    return root.attrib.get("name", "robot"), links, raw


def flatten_urdf(path: str = DEFAULT_URDF, tip_link: str = "robot:pointer") -> ChainModel:
    robot_name, links, raw = parse_urdf(path)
    children = {j["child"] for j in raw}
    roots = [n for n in links if n not in children]
    assert len(roots) == 1, f"URDF must have exactly one root link, got {roots}"
    root_link = roots[0]

    # depth-first order from the root = Bullet's link/joint indexing
    by_parent: Dict[str, list] = {}
    for j in raw:
        by_parent.setdefault(j["parent"], []).append(j)
    ordered: List[dict] = []
    link_index = {root_link: -1}

    def walk(link_name: str):
        for j in by_parent.get(link_name, []):
            j["index"] = len(ordered)
            j["parent_index"] = link_index[link_name]
            link_index[j["child"]] = j["index"]
            ordered.append(j)
            walk(j["child"])

    walk(root_link)
    assert len(ordered) == len(raw), "URDF joints do not form a single tree"

    joints: List[UrdfJoint] = []
    for j in ordered:
        if j["type"] == "revolute":
            jt = JOINT_REVOLUTE
        elif j["type"] == "fixed":
            jt = JOINT_FIXED
        else:
            # same failure as the reference scene loader (bullet_env.py:146)
            raise AssertionError(f"Only revolute and fixed joints are supported atm, got: {j['name']} ({j['type']})")
        n = float(np.linalg.norm(j["axis"]))
        joints.append(UrdfJoint(index=j["index"], name=j["name"], joint_type=jt, parent_link=j["parent"],
                                child_link=j["child"], origin_xyz=j["xyz"], origin_rot=j["rot"],
                                axis=j["axis"] / n if n > 0 else j["axis"], lower=j["lower"], upper=j["upper"],
                                effort=j["effort"], max_velocity=j["velocity"], damping=j["damping"],
                                friction=j["friction"], parent_index=j["parent_index"]))

    # the chain from the root to the tip link must be serial
    assert tip_link in links, f"tip link {tip_link!r} not in URDF"
    path_joints: List[UrdfJoint] = []
    cur = tip_link
    by_child = {j.child_link: j for j in joints}
    while cur != root_link:
        path_joints.append(by_child[cur])
        cur = by_child[cur].parent_link
    path_joints.reverse()
    revolute = [j for j in joints if j.joint_type == JOINT_REVOLUTE]
    assert all(j in path_joints for j in revolute), "every revolute joint must lie on the root->tip chain"
    dof = len(revolute)

    # fold fixed joints: walk the chain accumulating the static transform since the last moving frame
    axis = np.zeros((dof, 3)); oxyz = np.zeros((dof, 3)); orot = np.zeros((dof, 3, 3))
    acc_R, acc_p = np.eye(3), np.zeros(3)
    base_R, base_p = np.eye(3), np.zeros(3)
    # rigid bodies: list of (mass, com, inertia_about_com) in the current moving frame
    body_parts: List[List[Tuple[float, np.ndarray, np.ndarray]]] = [[] for _ in range(dof)]
    capsules: List[Tuple[int, float, np.ndarray, np.ndarray]] = []
    k = -1  # index of the current moving frame (-1 = static base)

    def attach(link: UrdfLink, R: np.ndarray, p: np.ndarray):
        if k < 0:
            return
        if link.mass > 0.0:
            Rc = R @ link.com_rot
            body_parts[k].append((link.mass, p + R @ link.com_xyz, Rc @ link.inertia @ Rc.T))
        for rad, a, b in link.capsules:
            capsules.append((k, rad, p + R @ a, p + R @ b))

    for j in path_joints:
        # transform of the joint frame in the current moving frame
        jp = acc_p + acc_R @ j.origin_xyz
        jR = acc_R @ j.origin_rot
        if j.joint_type == JOINT_REVOLUTE:
            k += 1
            if k == 0:
                # static part world -> first joint frame is kept separately so origin[0] stays "as written"
                base_R, base_p = np.eye(3), np.zeros(3)
            axis[k], oxyz[k], orot[k] = j.axis, jp, jR
            acc_R, acc_p = np.eye(3), np.zeros(3)
        else:
            acc_R, acc_p = jR, jp
        attach(links[j.child_link], acc_R, acc_p)
    tip = links[tip_link]
    tip_xyz = acc_p + acc_R @ tip.com_xyz

    mass = np.zeros(dof); com = np.zeros((dof, 3)); inertia = np.zeros((dof, 3, 3))
    for b in range(dof):
        m = sum(part[0] for part in body_parts[b])
        mass[b] = m
        if m > 0:
            com[b] = sum(part[0] * part[1] for part in body_parts[b]) / m
            for pm, pc, pI in body_parts[b]:
                d = pc - com[b]
                inertia[b] += pI + pm * (float(d @ d) * np.eye(3) - np.outer(d, d))

    return ChainModel(
        dof=dof, joint_names=[j.name for j in revolute], body_names=[j.child_link for j in revolute],
        axis=axis, origin_xyz=oxyz, origin_rot=orot, base_xyz=base_p, base_rot=base_R,
        tip_name=tip_link, tip_xyz=tip_xyz,
        lower=np.array([j.lower for j in revolute]), upper=np.array([j.upper for j in revolute]),
        effort=np.array([j.effort for j in revolute]), max_velocity=np.array([j.max_velocity for j in revolute]),
        damping=np.array([j.damping for j in revolute]), friction=np.array([j.friction for j in revolute]),
        body_mass=mass, body_com=com, body_inertia=inertia, capsules=capsules,
        joints=joints, links=links, root_link=root_link, robot_name=robot_name)
