"""Host-side URDF flattener (pioneer_b200/urdf.py): what replaces loadURDF / getJointInfo of the reference scene
loader (pioneer/envs/bullet/bullet_env.py:101-148).  CPU only."""
import numpy as np
import pytest

from pioneer_b200.urdf import DEFAULT_URDF, find_unique, flatten_urdf, parse_urdf, rpy_to_matrix, axis_angle_matrix
from tests._util import load_golden

G = load_golden()


def test_tables_of_the_shipped_robot():
    m = flatten_urdf()
    assert m.dof == 6 and m.joint_names == list(G["const_joint_names"])          # revolute joints in Bullet order
    assert [j.name for j in m.joints if j.joint_type == 4] == ["world_to_base", "robot:rotator1_to_hinge1",
                                                               "robot:rotator2_to_hinge2", "robot:rotator3_to_effector",
                                                               "robot:effector_to_pointer"]
    assert np.array_equal(m.axis, [[0, 0, 1], [0, 1, 0], [0, 1, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0]])
    assert np.array_equal(m.origin_xyz, [[0, 0, 0], [0, 0, 3], [0, 0, 11], [0, 1, 0], [11, 0, 0], [0, 0, 0]])
    assert all(np.array_equal(r, np.eye(3)) for r in m.origin_rot)
    assert np.array_equal(m.tip_xyz, [3.6, 0, 1.9]) and m.tip_name == "robot:pointer"
    assert np.array_equal(m.lower.astype(np.float32), G["const_r_lo"]) and np.array_equal(m.upper.astype(np.float32), G["const_r_hi"])
    assert set(G["const_item_names"]) - {"target"} <= set(m.links)
    assert m.root_link == "world" and len(m.joints) == 11


def test_tip_position_known_answers():
    m = flatten_urdf()
    np.testing.assert_allclose(m.tip_position(np.zeros(6)), (14.6, 1.0, 15.9), atol=1e-12)
    np.testing.assert_allclose(m.tip_position([0.3, -0.4, 0.9, 1.1, -0.7, 2.0]),
                               (7.904033967723, 1.072102752440, 5.624962379045), atol=2e-12)


def _tree_fk(path, q, tip):
    """Independent forward kinematics straight from the parsed XML with 4x4 homogeneous matrices (no folding)."""
    _, links, raw = parse_urdf(path)
    by_child = {j["child"]: j for j in raw}
    rev = [j["name"] for j in raw if j["type"] == "revolute"]
    chain, cur = [], tip
    while cur in by_child:
        chain.append(by_child[cur])
        cur = by_child[cur]["parent"]
    T = np.eye(4)
    for j in reversed(chain):
        A = np.eye(4)
        A[:3, :3], A[:3, 3] = j["rot"], j["xyz"]
        T = T @ A
        if j["type"] == "revolute":
            R = np.eye(4)
            R[:3, :3] = axis_angle_matrix(j["axis"] / np.linalg.norm(j["axis"]), q[rev.index(j["name"])])
            T = T @ R
    return (T @ np.append(links[tip].com_xyz, 1.0))[:3]


def test_fixed_joint_folding_with_rotations(tmp_path):
    text = open(DEFAULT_URDF).read()
    text = text.replace('<joint name="robot:rotator1_to_hinge1" type="fixed"><parent link="robot:rotator1"/><child link="robot:hinge1"/>',
                        '<joint name="robot:rotator1_to_hinge1" type="fixed"><parent link="robot:rotator1"/><child link="robot:hinge1"/>'
                        '<origin xyz="0.2 -0.1 0.4" rpy="0.1 0.2 -0.3"/>')
    text = text.replace('<origin xyz="0 0 11"/><axis xyz="0 1 0"/>', '<origin xyz="0 0 11" rpy="0.3 -0.2 0.5"/><axis xyz="0 3 4"/>')
    text = text.replace('<origin xyz="3.6 0 1.9"/>', '<origin xyz="3.6 0 1.9" rpy="0.4 0 0"/>')
    path = tmp_path / "rot.urdf"
    path.write_text(text)
    m = flatten_urdf(str(path))
    assert m.dof == 6 and np.allclose(m.axis[2], (0, 0.6, 0.8))                    # axis normalised
    assert not np.allclose(m.origin_rot[1], np.eye(3))                              # fixed joint folded into the next frame
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.uniform(m.lower, m.upper)
        np.testing.assert_allclose(m.tip_position(q), _tree_fk(str(path), q, "robot:pointer"), atol=1e-10)
    # composite bodies: total mass is conserved by the folding (world/base are static and excluded)
    assert np.isclose(m.body_mass.sum(), 10.0)


def test_unsupported_joint_type_raises_like_the_reference(tmp_path):
    text = open(DEFAULT_URDF).read().replace('name="robot:hinge1_to_arm1" type="revolute"', 'name="robot:hinge1_to_arm1" type="prismatic"')
    path = tmp_path / "bad.urdf"
    path.write_text(text)
    with pytest.raises(AssertionError, match="Only revolute and fixed joints are supported"):    # bullet_env.py:146
        flatten_urdf(str(path))


def test_find_unique_contract():
    import xml.etree.ElementTree as ET
    root = ET.fromstring('<r><a k="1"/><a k="2"/><b/></r>')
    assert find_unique(root, "b").tag == "b"
    assert find_unique(root, "a", "k", "2").attrib["k"] == "2"
    with pytest.raises(AssertionError):
        find_unique(root, "a")
    with pytest.raises(AssertionError):
        find_unique(root, "c")


def test_rpy_convention():
    # fixed-axis roll/pitch/yaw: R = Rz(yaw) Ry(pitch) Rx(roll)
    R = rpy_to_matrix((0.1, 0.2, 0.3))
    want = axis_angle_matrix((0, 0, 1), 0.3) @ axis_angle_matrix((0, 1, 0), 0.2) @ axis_angle_matrix((1, 0, 0), 0.1)
    np.testing.assert_allclose(R, want, atol=1e-15)


def _random_urdf(rng, path):
    """Six revolute joints with random axes, fixed joints (with rotations) in between, every link with a random mass,
    inertial origin (xyz + rpy) and full inertia tensor; the tracked tip hangs off the last link by a fixed joint."""
    def f(v):
        return " ".join(f"{x:.9f}" for x in v)

    def link(name):
        a = rng.normal(size=(3, 3))
        I = a @ a.T + np.eye(3)
        return (f'<link name="{name}"><inertial><mass value="{rng.uniform(0.5, 3.0):.9f}"/>'
                f'<origin xyz="{f(rng.uniform(-0.5, 0.5, 3))}" rpy="{f(rng.uniform(-1, 1, 3))}"/>'
                f'<inertia ixx="{I[0, 0]:.9f}" ixy="{I[0, 1]:.9f}" ixz="{I[0, 2]:.9f}" iyy="{I[1, 1]:.9f}" iyz="{I[1, 2]:.9f}" '
                f'izz="{I[2, 2]:.9f}"/></inertial></link>')

    def joint(name, kind, parent, child):
        body = (f'<joint name="{name}" type="{kind}"><parent link="{parent}"/><child link="{child}"/>'
                f'<origin xyz="{f(rng.uniform(-2, 2, 3))}" rpy="{f(rng.uniform(-1, 1, 3))}"/>')
        if kind == "revolute":
            body += f'<axis xyz="{f(rng.normal(size=3))}"/><limit lower="-2.5" upper="2.5" effort="10" velocity="5"/>'
        return body + "</joint>"

    parts, prev = ['<link name="world"/>'], "world"
    n = 0
    for r in range(6):
        for _ in range(int(rng.integers(0, 3))):              # 0..2 fixed joints before every revolute one
            n += 1
            parts += [link(f"f{n}"), joint(f"jf{n}", "fixed", prev, f"f{n}")]
            prev = f"f{n}"
        parts += [link(f"m{r}"), joint(f"jr{r}", "revolute", prev, f"m{r}")]
        prev = f"m{r}"
    parts += [link("tip"), joint("jtip", "fixed", prev, "tip")]
    path.write_text('<robot name="random">' + "".join(parts) + "</robot>")


def _link_frames(path, q):
    """World transform of every link, straight from the parsed XML (no folding)."""
    _, links, raw = parse_urdf(path)
    rev = [j["name"] for j in raw if j["type"] == "revolute"]
    T = {"world": np.eye(4)}
    todo = list(raw)
    while todo:
        for j in list(todo):
            if j["parent"] in T:
                A = np.eye(4)
                A[:3, :3], A[:3, 3] = j["rot"], j["xyz"]
                if j["type"] == "revolute":
                    R = np.eye(4)
                    R[:3, :3] = axis_angle_matrix(j["axis"] / np.linalg.norm(j["axis"]), q[rev.index(j["name"])])
                    A = A @ R
                T[j["child"]] = T[j["parent"]] @ A
                todo.remove(j)
    return links, T


def test_folded_composite_bodies_carry_the_kinetic_energy_of_the_unfolded_links(tmp_path):
    """Fixed joints (with rotations) are folded into composite bodies: mass, centre of mass and inertia about it.  The
    kinetic energy of the UNFOLDED links -- from finite differences of every link's world pose, nothing else -- must equal
    1/2 qd^T M(q) qd with the joint-space inertia M of the FOLDED chain (oracle CRBA).  Also: the tracked point."""
    from oracle.dynamics_oracle import DynChain, crba
    rng = np.random.default_rng(17)
    for case in range(3):
        path = tmp_path / f"random{case}.urdf"
        _random_urdf(rng, path)
        m = flatten_urdf(str(path), tip_link="tip")
        assert m.dof == 6
        ch = DynChain.from_model(m)
        for _ in range(3):
            q, qd = rng.uniform(-2.0, 2.0, 6), rng.normal(size=6)
            h = 1e-6
            links, Tp = _link_frames(str(path), q + h * qd)
            _, Tm = _link_frames(str(path), q - h * qd)
            _, T0 = _link_frames(str(path), q)
            ke = 0.0
            for name, ln in links.items():
                if ln.mass <= 0.0:
                    continue
                cp = (Tp[name] @ np.append(ln.com_xyz, 1.0))[:3]
                cm = (Tm[name] @ np.append(ln.com_xyz, 1.0))[:3]
                v = (cp - cm) / (2 * h)
                W = (Tp[name][:3, :3] @ Tm[name][:3, :3].T - np.eye(3)) / (2 * h)          # ~ [omega]x
                om = 0.5 * np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]])
                Rw = T0[name][:3, :3] @ ln.com_rot
                ke += 0.5 * ln.mass * v @ v + 0.5 * om @ (Rw @ ln.inertia @ Rw.T) @ om
            want = 0.5 * qd @ crba(ch, q) @ qd
            assert abs(ke - want) <= 1e-6 * max(1.0, abs(want)), (case, ke, want)
            # getLinkState(...)[0] is the world position of the link's CENTRE OF MASS (bullet_bindings.py:38-45)
            np.testing.assert_allclose(m.tip_position(q), (T0["tip"] @ np.append(links["tip"].com_xyz, 1.0))[:3], atol=1e-9)
