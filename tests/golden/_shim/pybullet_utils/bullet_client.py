"""Stand-in for pybullet_utils.bullet_client.BulletClient (see ../README.md).

Implements only the calls the reference env makes on top of a float64
rigid-tree forward kinematics with 4x4 homogeneous transforms.  URDF conventions:
child = parent * T(origin xyz, rpy) * Rot(axis, q); rpy = fixed-axis roll, pitch, yaw.
"""
import math
import xml.etree.ElementTree as ET

import numpy as np
import pybullet as _pb


def _floats(text, n, default=0.0):
    if text is None:
        return [default] * n
    vals = [float(x) for x in text.split()]
    assert len(vals) == n
    return vals


def _T(xyz=(0, 0, 0), rpy=(0, 0, 0)):
    r, p, y = rpy
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = xyz
    return T


def _rot_axis(axis, q):
    x, y, z = axis
    c, s, C = math.cos(q), math.sin(q), 1.0 - math.cos(q)
    T = np.eye(4)
    T[:3, :3] = [[x * x * C + c, x * y * C - z * s, x * z * C + y * s],
                 [y * x * C + z * s, y * y * C + c, y * z * C - x * s],
                 [z * x * C - y * s, z * y * C + x * s, z * z * C + c]]
    return T


def _quat_from_matrix(R):
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return ((R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s)
    i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    q = [0.0, 0.0, 0.0, 0.0]
    q[i] = 0.25 * s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    q[3] = (R[k, j] - R[j, k]) / s
    return tuple(q)


class _MultiBody:
    def __init__(self, path):
        root = ET.parse(path).getroot()
        self.robot_name = root.attrib["name"]
        self.link_inertial_xyz = {}
        for le in root.findall("link"):
            oe = le.find("inertial/origin")
            self.link_inertial_xyz[le.attrib["name"]] = _floats(oe.attrib.get("xyz"), 3) if oe is not None else [0.0] * 3
        joints = []
        for je in root.findall("joint"):
            oe, ae, lim, dyn = je.find("origin"), je.find("axis"), je.find("limit"), je.find("dynamics")
            joints.append(dict(
                name=je.attrib["name"], type=je.attrib["type"],
                parent=je.find("parent").attrib["link"], child=je.find("child").attrib["link"],
                xyz=_floats(oe.attrib.get("xyz") if oe is not None else None, 3),
                rpy=_floats(oe.attrib.get("rpy") if oe is not None else None, 3),
                axis=_floats(ae.attrib.get("xyz"), 3) if ae is not None else [1.0, 0.0, 0.0],
                lower=float(lim.attrib.get("lower", 0)) if lim is not None else 0.0,
                upper=float(lim.attrib.get("upper", -1)) if lim is not None else -1.0,
                effort=float(lim.attrib.get("effort", 0)) if lim is not None else 0.0,
                velocity=float(lim.attrib.get("velocity", 0)) if lim is not None else 0.0,
                damping=float(dyn.attrib.get("damping", 0)) if dyn is not None else 0.0,
                friction=float(dyn.attrib.get("friction", 0)) if dyn is not None else 0.0))
        children = {j["child"] for j in joints}
        self.base_link = [le.attrib["name"] for le in root.findall("link") if le.attrib["name"] not in children][0]
        self.joints = []            # depth-first from the base: Bullet's joint == link index
        index = {self.base_link: -1}

        def walk(link):
            for j in joints:
                if j["parent"] == link:
                    j["parent_index"] = index[link]
                    index[j["child"]] = len(self.joints)
                    self.joints.append(j)
                    walk(j["child"])

        walk(self.base_link)
        assert len(self.joints) == len(joints)
        self.q = [0.0] * len(self.joints)
        self.qd = [0.0] * len(self.joints)

    def link_frames(self):
        frames = []
        for i, j in enumerate(self.joints):
            parent = np.eye(4) if j["parent_index"] < 0 else frames[j["parent_index"]]
            T = parent @ _T(j["xyz"], j["rpy"])
            if j["type"] == "revolute":
                T = T @ _rot_axis(j["axis"], self.q[i])
            elif j["type"] != "fixed":
                raise NotImplementedError(j["type"])
            frames.append(T)
        return frames


class _StaticBody:
    def __init__(self, position, orientation):
        self.position = tuple(float(x) for x in position)
        self.orientation = tuple(float(x) for x in orientation)


class BulletClient:
    def __init__(self, connection_mode=None):
        assert connection_mode == _pb.DIRECT, "the stand-in only runs headless"
        self._bodies = []
        self._gravity = (0.0, 0.0, 0.0)
        self._timestep = 1 / 240
        self._shapes = 0

    def __getattr__(self, name):
        # the real client forwards unknown attributes to the pybullet module (constants)
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(_pb, name)

    # ---- world
    def setGravity(self, x, y, z):
        self._gravity = (float(x), float(y), float(z))

    def setTimeStep(self, dt):
        self._timestep = float(dt)

    def stepSimulation(self):
        if any(g != 0 for g in self._gravity):
            raise NotImplementedError("stand-in has no dynamics: gravity must be 0")
        for b in self._bodies:
            if isinstance(b, _MultiBody) and any(v != 0 for v in b.qd):
                raise NotImplementedError("stand-in has no dynamics: joint velocities must be 0")
        # zero gravity, zero joint velocity, no motor target ever set by the env: nothing moves

    # ---- model
    def loadURDF(self, path, flags=0):
        self._bodies.append(_MultiBody(path))
        return len(self._bodies) - 1

    def getBodyInfo(self, body_id):
        b = self._bodies[body_id]
        return (b.base_link.encode("utf8"), b.robot_name.encode("utf8"))

    def getNumJoints(self, body_id):
        return len(self._bodies[body_id].joints)

    def getJointInfo(self, body_id, joint_index):
        b = self._bodies[body_id]
        j = b.joints[joint_index]
        jt = {"revolute": _pb.JOINT_REVOLUTE, "fixed": _pb.JOINT_FIXED, "prismatic": _pb.JOINT_PRISMATIC}[j["type"]]
        moving = jt != _pb.JOINT_FIXED
        n_before = sum(1 for k in b.joints[:joint_index] if k["type"] != "fixed")
        return (joint_index, j["name"].encode("utf8"), jt,
                7 + n_before if moving else -1, 6 + n_before if moving else -1, int(moving),
                j["damping"], j["friction"], j["lower"], j["upper"], j["effort"], j["velocity"],
                j["child"].encode("utf8"), tuple(j["axis"]) if moving else (0.0, 0.0, 0.0),
                tuple(j["xyz"]), _quat_from_matrix(_T(j["xyz"], j["rpy"])[:3, :3]), j["parent_index"])

    # ---- joints / links
    def resetJointState(self, bodyUniqueId, jointIndex, targetValue, targetVelocity=0.0):
        b = self._bodies[bodyUniqueId]
        b.q[jointIndex] = float(targetValue)
        b.qd[jointIndex] = float(targetVelocity)

    def getJointState(self, body_id, joint_index):
        b = self._bodies[body_id]
        return (b.q[joint_index], b.qd[joint_index], (0.0,) * 6, 0.0)

    def getLinkState(self, body_id, link_index, computeLinkVelocity=0, computeForwardKinematics=0):
        b = self._bodies[body_id]
        T = b.link_frames()[link_index]
        local_com = np.array(b.link_inertial_xyz[b.joints[link_index]["child"]])
        com_world = T[:3, 3] + T[:3, :3] @ local_com
        quat = _quat_from_matrix(T[:3, :3])
        return (tuple(float(x) for x in com_world), quat, tuple(float(x) for x in local_com), (0.0, 0.0, 0.0, 1.0),
                tuple(float(x) for x in T[:3, 3]), quat, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0))

    def getBasePositionAndOrientation(self, body_id):
        b = self._bodies[body_id]
        if isinstance(b, _StaticBody):
            return (b.position, b.orientation)
        return ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0))

    def getBaseVelocity(self, body_id):
        return ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))

    def resetBasePositionAndOrientation(self, bodyUniqueId, posObj, ornObj):
        b = self._bodies[bodyUniqueId]
        assert isinstance(b, _StaticBody)
        b.position, b.orientation = tuple(posObj), tuple(ornObj)

    # ---- primitive bodies
    def createVisualShape(self, *args, **kwargs):
        self._shapes += 1
        return self._shapes

    def createCollisionShape(self, *args, **kwargs):
        self._shapes += 1
        return self._shapes

    def createMultiBody(self, baseMass=0.0, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), **kwargs):
        self._bodies.append(_StaticBody(basePosition, baseOrientation))
        return len(self._bodies) - 1

    # ---- helpers
    def getQuaternionFromEuler(self, rpy):
        return _quat_from_matrix(_T((0, 0, 0), rpy)[:3, :3])

    def getEulerFromQuaternion(self, quat):
        x, y, z, w = quat
        roll = math.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
        pitch = math.asin(max(-1.0, min(1.0, 2 * (w * y - z * x))))
        yaw = math.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
        return (roll, pitch, yaw)
