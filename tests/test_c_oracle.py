"""The C restatement (oracle/quad_step_c.c) against the torch oracle: two independent CPU implementations of the same
float32 step must agree bit for bit (and both are what the CUDA kernel is held to).  CPU only."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("kw", [dict(), dict(fault_mode=1, dr_enable=1, pomdp_mode=3, pomdp_prob=0.2, noise_sigma=0.1,
                                             max_episode_length=40, lin_drag=0.05, yaw_km=0.016, env_id_base=777)])
def test_c_oracle_bit_exact_vs_torch_oracle(kw):
    from ouzelum_b200 import _lib
    from oracle.c_oracle import COracle
    from oracle.quad_step import QuadStepOracle
    n = 333
    cfg = _lib.default_cfg(n, seed=99, **kw)
    co, to = COracle(cfg), QuadStepOracle(cfg.to_dict())
    g = torch.Generator().manual_seed(8)
    for t in range(120):
        a = torch.rand(n, 4, generator=g) * 2 - 1
        if t % 5 == 0:
            a = a * 2.5
        co.step(a.numpy())
        to.step(a)
        assert np.array_equal(co.reset_buf, to.reset_buf.numpy()), t
        assert np.array_equal(co.progress_buf, to.progress_buf.numpy()), t
        assert np.array_equal(co.root, to.root.numpy()), f"root t={t} max {np.abs(co.root - to.root.numpy()).max()}"
        assert np.array_equal(co.obs_buf, to.obs_buf.numpy()), t
        assert np.array_equal(co.rew_buf, to.rew_buf.numpy()), t
        assert np.array_equal(co.timeout_buf.astype(bool), to.timeout_buf.numpy()), t
        assert np.array_equal(co.target, to.target.numpy()) and np.array_equal(co.thrust, to.thrust.numpy()), t
    assert to.mcnt[1] > 0
