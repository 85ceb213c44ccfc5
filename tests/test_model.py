"""MJCF loader + closed-form constants against the numbers derived in SURVEY.md 8a-11."""
import json
import os

import numpy as np
import pytest

from uav_reinforcement_learning_control_b200 import model as M

REF_XML = "/root/reference/model/drone/drone.xml"


def test_tree_shape():
    t = M.load_mjcf(M.default_model_path())
    assert (t.nq, t.nv, t.nu, t.nbody) == (11, 10, 4, 6)
    assert t.timestep == 0.01 and t.density == 1.225 and t.viscosity == 1.8e-5
    assert t.integrator == "Euler" and list(t.gravity) == [0, 0, -9.81]
    assert all(cr is not None and list(cr) == [0.0, 13.0] for cr in t.act_ctrlrange)   # default class + autolimits
    np.testing.assert_allclose(t.qpos0(), [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0])


def test_derived_constants_match_survey():
    _, c = M.load_default()
    assert c.mass == pytest.approx(0.22274432, abs=1e-9)
    np.testing.assert_allclose(c.com, [0, 0, 0.0040605594], atol=1e-9)
    I_O = c.I_C + c.mass * (np.dot(c.com, c.com) * np.eye(3) - np.outer(c.com, c.com))
    np.testing.assert_allclose(np.diag(I_O), [4.9694749e-4, 5.0394749e-4, 6.3990521e-4], rtol=1e-7)
    assert np.abs(I_O - np.diag(np.diag(I_O))).max() < 1.8e-10
    np.testing.assert_allclose(c.rotor_J, 3.75335e-6)
    assert (c.I_C[2, 2] - 4 * c.rotor_J[0]) / c.I_C[2, 2] == pytest.approx(0.97654, abs=1e-5)
    np.testing.assert_allclose(c.base_box, [0.1293772, 0.1277016, 0.0963966], atol=1e-7)
    np.testing.assert_allclose(c.rot_box[0], [0.0019970, 0.0569807, 0.0569807], atol=1e-7)
    assert c.hover_thrust_per_motor == pytest.approx(0.5462804, abs=1e-7)
    # the site wrench reproduces the reference mixer rows (hover_env.py:94-99)
    l, k = 0.039799, 0.0201
    np.testing.assert_allclose(c.wrench[2], 1.0)
    np.testing.assert_allclose(c.wrench[3:], [[-l, -l, l, l], [-l, l, l, -l], [k, -k, k, -k]], atol=1e-15)
    # known answers (from rest, identity attitude, no fluid): yaw 1e-3 N m, roll 1e-3 N m
    assert 1e-3 * c.Ieff_inv[2, 2] == pytest.approx(1.600277018, rel=1e-9)
    assert 1e-3 * c.Ieff_inv[0, 0] == pytest.approx(2.027267355, rel=1e-9)


@pytest.mark.skipif(not os.path.exists(REF_XML), reason="reference tree not present (GPU box)")
def test_shipped_asset_equals_reference_xml():
    """The re-authored dynamics-only MJCF carries exactly the reference's physical parameters."""
    a = M.load_mjcf(M.default_model_path())
    b = M.load_mjcf(REF_XML)            # exercises: nested defaults, mesh geoms, sensors, bad keyframe ignored
    sa, sb = a.summary(), b.summary()
    # the reference has one extra, sensor-only site ("imu"); everything dynamic must be identical
    for key in sa:
        if key.startswith("site_") or key == "act_site":
            continue
        assert json.dumps(sa[key]) == json.dumps(sb[key]), key
    for k in range(4):
        np.testing.assert_array_equal(a.site_pos[a.act_site[k]], b.site_pos[b.act_site[k]])
    ca, cb = M.derive_constants(a), M.derive_constants(b)
    for f in ca.__dataclass_fields__:
        np.testing.assert_array_equal(getattr(ca, f), getattr(cb, f), err_msg=f)


def test_bad_models_raise(tmp_path):
    src = open(M.default_model_path()).read()
    cases = {
        "unbalanced": src.replace('inertial pos="0 0 -0.001" quat', 'inertial pos="0.001 0 -0.001" quat', 1),
        "slide": src.replace('<joint name="prop1" pos="0 0 0" axis="0 0 1"/>', '<joint name="prop1" type="slide"/>'),
        "axis": src.replace('<joint name="prop2" pos="0 0 0" axis="0 0 1"/>', '<joint name="prop2" axis="1 0 0"/>'),
        "rk4": src.replace('timestep="0.01"', 'timestep="0.01" integrator="RK4"'),
        "garbage": "<notmujoco/>",
    }
    for name, text in cases.items():
        p = tmp_path / f"{name}.xml"
        p.write_text(text)
        with pytest.raises(M.ModelError):
            M.derive_constants(M.load_mjcf(str(p)))
    with pytest.raises(M.ModelError):
        M.load_mjcf(str(tmp_path / "missing.xml"))


def test_joint_damping_and_armature_are_parsed():
    import tempfile
    src = open(M.default_model_path()).read().replace(
        '<joint name="prop1" pos="0 0 0" axis="0 0 1"/>', '<joint name="prop1" pos="0 0 0" axis="0 0 1" damping="1e-6" armature="2e-6"/>')
    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
        f.write(src)
    t = M.load_mjcf(f.name)
    c = M.derive_constants(t)
    assert c.rotor_damping[0] == 1e-6 and c.rotor_Js[0] == pytest.approx(3.75335e-6 + 2e-6 + 0.01 * 1e-6)
    os.unlink(f.name)


# ---------------------------------------------------------------------------------------------------------------
# contact guard (drone.xml:51,56,61,66,71 geoms with default contype / conaffinity, scene.xml:22 floor)
# ---------------------------------------------------------------------------------------------------------------
from uav_reinforcement_learning_control_b200 import config as Q  # noqa: E402

ALL_CFGS = [Q.EnvConfig.north_star(), Q.EnvConfig.hover_gym(), Q.EnvConfig.traj_gym(), Q.EnvConfig.mjx_brax(),
            Q.EnvConfig.hover_brax(), Q.EnvConfig.mjx_playground()]


def _report(tree, cfg):
    return M.contact_report(tree, *cfg.position_bounds(), overshoot=cfg.position_overshoot(tree.timestep))


def test_contact_guard_passes_the_shipped_model():
    t = M.load_mjcf(M.default_model_path())
    assert len(t.geom_name) == 5 and all(c == 1 for c in t.geom_contype + t.geom_conaffinity)
    for cfg in ALL_CFGS:
        assert _report(t, cfg) == []
        M.check_contacts(t, *cfg.position_bounds())
    # the tightest pair: neighbouring propeller discs, 0.0796 m apart, radius 0.0329 m each -> 13.8 mm clearance
    # (the conservative bounding spheres add the half thickness, so a little less here)
    c1 = np.asarray(t.body_pos[2]) + t.geom_center[1]; c2 = np.asarray(t.body_pos[3]) + t.geom_center[2]
    gap = np.linalg.norm(c1 - c2) - t.geom_rbound[1] - t.geom_rbound[2]
    assert 0.0130 < gap < 0.0139


@pytest.mark.skipif(not os.path.exists(REF_XML), reason="reference tree not present (GPU box)")
def test_contact_guard_passes_reference_xml_and_shipped_bounds_enclose_its_meshes():
    ref = M.load_mjcf(REF_XML)                     # bounding radii measured from assets/drone/*.STL
    ours = M.load_mjcf(M.default_model_path())
    assert ref.geom_type == ["mesh"] * 5 and all(np.isfinite(ref.geom_rbound))
    for cfg in ALL_CFGS:
        assert _report(ref, cfg) == []
    for g in range(5):
        # every mesh's bounding sphere lies inside the shipped primitive's bounding sphere
        d = np.linalg.norm(np.asarray(ref.geom_center[g]) - np.asarray(ours.geom_center[g]))
        assert d + ref.geom_rbound[g] <= ours.geom_rbound[g] + 1e-4, (g, d, ref.geom_rbound[g], ours.geom_rbound[g])


def test_contact_guard_rejects_a_floor_and_touching_geoms(tmp_path):
    src = open(M.default_model_path()).read()
    floor = src.replace("<worldbody>", '<worldbody>\n    <geom name="floor" size="0 0 0.05" type="plane"/>')   # scene.xml:22
    p = tmp_path / "floor.xml"; p.write_text(floor)
    t = M.load_mjcf(str(p))
    for cfg in ALL_CFGS:
        rep = _report(t, cfg)
        assert rep and all("floor" in r for r in rep)
        with pytest.raises(M.ModelError, match="reachable contacts"):
            M.check_contacts(t, *cfg.position_bounds())
        M.check_contacts(t, *cfg.position_bounds(), assume_no_contact=True)        # explicit override only
    # a floor far below the termination box is unreachable for the bounded envs, still reachable for the unbounded one
    p.write_text(floor.replace('type="plane"', 'type="plane" pos="0 0 -1"'))
    t = M.load_mjcf(str(p))
    assert _report(t, Q.EnvConfig.north_star()) == [] and _report(t, Q.EnvConfig.mjx_brax()) == []
    assert _report(t, Q.EnvConfig.mjx_playground())
    # a static obstacle inside the box
    p.write_text(src.replace("<worldbody>", '<worldbody>\n    <geom name="pole" type="capsule" fromto="1 1 0 1 1 2" size="0.05"/>'))
    assert any("pole" in r for r in _report(M.load_mjcf(str(p)), Q.EnvConfig.north_star()))
    # propeller discs grown until neighbours overlap
    p.write_text(src.replace('size="0.0329 0.001"', 'size="0.0400 0.001"'))
    rep = _report(M.load_mjcf(str(p)), Q.EnvConfig.north_star())
    assert len(rep) == 4 and all("overlap" in r for r in rep)                      # the 4 neighbouring pairs, not the diagonals
    # the same discs with collisions switched off are fine; parent-child (base <-> prop) is always filtered
    p.write_text(src.replace('size="0.0329 0.001"', 'size="0.0400 0.001" contype="0" conaffinity="0"'))
    assert _report(M.load_mjcf(str(p)), Q.EnvConfig.north_star()) == []
    # a mesh whose file is missing has an unknown extent
    p.write_text(src.replace("<worldbody>", '<asset><mesh name="m" file="nope.stl"/></asset>\n  <worldbody>').replace(
        '<geom name="prop1_bound" type="cylinder" pos="0 0 -0.001" size="0.0329 0.001" mass="0"/>', '<geom name="prop1_bound" type="mesh" mesh="m"/>'))
    assert any("extent unknown" in r for r in _report(M.load_mjcf(str(p)), Q.EnvConfig.north_star()))
    # <contact> sections are not interpreted -> reported
    p.write_text(src.replace("<actuator>", "<contact><exclude body1='prop1' body2='prop2'/></contact>\n  <actuator>"))
    assert any("<contact>" in r for r in _report(M.load_mjcf(str(p)), Q.EnvConfig.north_star()))
