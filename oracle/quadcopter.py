"""CPU oracle of the stock Quadcopter hover task's observation / reward / reset (kernel K1q) -- TEST INFRASTRUCTURE.

Restates isaacgymenvs/tasks/quadcopter.py:359-370 (compute_observations) and :386-418 (compute_quadcopter_reward).
The articulated 9-body PhysX dynamics of that task are out of parity scope.
"""
import math

import torch

from .quad_step import div_by_scalar, ieee_sqrt, quat_axis


def compute_quadcopter_reward(root_positions, root_quats, root_linvels, root_angvels, reset_buf, progress_buf, max_episode_length):
    """quadcopter.py:386-418: target (0,0,1), up_reward 1/(1+tilt^2), die dist>3 or z<0.3."""
    target_dist = ieee_sqrt(root_positions[..., 0] * root_positions[..., 0] +
                            root_positions[..., 1] * root_positions[..., 1] +
                            (1 - root_positions[..., 2]) * (1 - root_positions[..., 2]))
    pos_reward = 1.0 / (1.0 + target_dist * target_dist)
    ups = quat_axis(root_quats, 2)
    tiltage = torch.abs(1 - ups[..., 2])
    up_reward = 1.0 / (1.0 + tiltage * tiltage)
    spinnage = torch.abs(root_angvels[..., 2])
    spinnage_reward = 1.0 / (1.0 + spinnage * spinnage)
    reward = pos_reward + pos_reward * (up_reward + spinnage_reward)
    ones = torch.ones_like(reset_buf)
    die = torch.zeros_like(reset_buf)
    die = torch.where(target_dist > 3.0, ones, die)
    die = torch.where(root_positions[..., 2] < 0.3, ones, die)
    reset = torch.where(progress_buf >= max_episode_length - 1, ones, die)
    return reward, reset


def compute_observations(root_states, dof_positions):
    """quadcopter.py:359-370: 21 observations."""
    n = root_states.shape[0]
    obs = torch.empty(n, 21, dtype=root_states.dtype)
    obs[..., 0] = div_by_scalar(0.0 - root_states[..., 0], 3)
    obs[..., 1] = div_by_scalar(0.0 - root_states[..., 1], 3)
    obs[..., 2] = div_by_scalar(1.0 - root_states[..., 2], 3)
    obs[..., 3:7] = root_states[..., 3:7]
    obs[..., 7:10] = div_by_scalar(root_states[..., 7:10], 2)
    obs[..., 10:13] = div_by_scalar(root_states[..., 10:13], math.pi)
    obs[..., 13:21] = dof_positions
    return obs
