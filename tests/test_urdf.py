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
