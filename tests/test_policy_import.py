"""SB3 .zip policy import: a zip laid out like ``PPO.save`` (train.py:141) -> the kernels' packed vector,
checked through the policy oracle against a plain torch forward of the same state_dict."""
import io
import zipfile

import numpy as np
import pytest
import torch

from oracle import ppo_ref
from uav_reinforcement_learning_control_b200 import ppo


def _fake_sb3_zip(tmp_path, obs_dim=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g) * 0.3
    sd = {
        "log_std": r(4),
        "mlp_extractor.policy_net.0.weight": r(128, obs_dim), "mlp_extractor.policy_net.0.bias": r(128),
        "mlp_extractor.policy_net.2.weight": r(128, 128), "mlp_extractor.policy_net.2.bias": r(128),
        "mlp_extractor.value_net.0.weight": r(128, obs_dim), "mlp_extractor.value_net.0.bias": r(128),
        "mlp_extractor.value_net.2.weight": r(128, 128), "mlp_extractor.value_net.2.bias": r(128),
        "action_net.weight": r(4, 128), "action_net.bias": r(4),
        "value_net.weight": r(1, 128), "value_net.bias": r(1),
    }
    buf = io.BytesIO(); torch.save(sd, buf)
    path = tmp_path / "final_model.zip"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", "{}"); z.writestr("policy.pth", buf.getvalue()); z.writestr("_stable_baselines3_version", "2.3.0")
    return str(path), sd


def test_sb3_zip_import_matches_torch_forward(tmp_path):
    path, sd = _fake_sb3_zip(tmp_path)
    packed = ppo.load_sb3_policy_zip(path)
    assert packed.numel() == ppo_ref.param_count(12, 0)
    pp = ppo_ref.unpack(packed.numpy(), 12, 0)
    obs = torch.randn(64, 12, generator=torch.Generator().manual_seed(1))
    head, value = ppo_ref.forward(pp, obs.numpy())
    lin = torch.nn.functional.linear
    h = torch.relu(lin(torch.relu(lin(obs, sd["mlp_extractor.policy_net.0.weight"], sd["mlp_extractor.policy_net.0.bias"])),
                       sd["mlp_extractor.policy_net.2.weight"], sd["mlp_extractor.policy_net.2.bias"]))
    mean = lin(h, sd["action_net.weight"], sd["action_net.bias"])
    c = torch.relu(lin(torch.relu(lin(obs, sd["mlp_extractor.value_net.0.weight"], sd["mlp_extractor.value_net.0.bias"])),
                       sd["mlp_extractor.value_net.2.weight"], sd["mlp_extractor.value_net.2.bias"]))
    v = lin(c, sd["value_net.weight"], sd["value_net.bias"])[:, 0]
    np.testing.assert_allclose(head, mean.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(value, v.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(pp["log_std"], sd["log_std"].numpy(), rtol=0, atol=0)


def test_sb3_import_rejects_other_files(tmp_path):
    p = tmp_path / "x.zip"
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("data", "{}")
    with pytest.raises(ValueError):
        ppo.load_sb3_policy_zip(str(p))
    with pytest.raises(KeyError):
        ppo.sb3_state_dict_to_packed({"log_std": torch.zeros(4)})


def test_sb3_export_round_trip(tmp_path):
    """packed -> SB3 state_dict -> zip -> packed is the identity, and the exported tensors have torch.nn.Linear's layout."""
    path, sd = _fake_sb3_zip(tmp_path, seed=3)
    packed = ppo.load_sb3_policy_zip(path)
    sd2 = ppo.packed_to_sb3_state_dict(packed)
    assert set(sd2) == set(sd)
    for k in sd:
        assert sd2[k].shape == sd[k].shape, k
        np.testing.assert_array_equal(sd2[k].numpy(), sd[k].numpy())
    out = str(tmp_path / "exported.zip")
    ppo.save_sb3_policy_zip(out, packed)
    np.testing.assert_array_equal(ppo.load_sb3_policy_zip(out).numpy(), packed.numpy())
    bad = packed.clone(); bad[-1] = 2.0            # a non-identity observation normaliser cannot be exported
    with pytest.raises(ValueError):
        ppo.packed_to_sb3_state_dict(bad)
