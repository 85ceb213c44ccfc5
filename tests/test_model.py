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
