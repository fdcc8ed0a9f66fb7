"""Trainer-side env wrappers (mirror of isaacgymenvs/<trainer>/utils.py:4-39, identical in all eight trainers)."""
import torch

from ._lib import check, lib


class RecordEpisodeStatisticsTorch:
    """isaacgymenvs/RPO-LSTM/utils.py:4-35 -- six element-wise launches per step fused into one kernel."""

    def __init__(self, env, device):
        self.env = env
        self.num_envs = getattr(env, "num_envs", 1)
        self.device = device
        self.episode_returns = None
        self.episode_lengths = None

    def __getattr__(self, name):           # gym.Wrapper forwards unknown attributes to the wrapped env
        return getattr(self.env, name)

    def reset(self, **kwargs):
        observations = self.env.reset(**kwargs)
        z = lambda dt: torch.zeros(self.num_envs, dtype=dt, device=self.device)
        self.episode_returns, self.episode_lengths = z(torch.float32), z(torch.int32)
        self.returned_episode_returns, self.returned_episode_lengths = z(torch.float32), z(torch.int32)
        return observations

    def step(self, action):
        observations, rewards, dones, infos = self.env.step(action)
        check(lib.ozl_episode_stats(self.num_envs, rewards.data_ptr(), dones.data_ptr(), self.episode_returns.data_ptr(),
                                    self.episode_lengths.data_ptr(), self.returned_episode_returns.data_ptr(),
                                    self.returned_episode_lengths.data_ptr(), torch.cuda.current_stream().cuda_stream))
        infos["r"] = self.returned_episode_returns
        infos["l"] = self.returned_episode_lengths
        return observations, rewards, dones, infos


class ExtractObsWrapper:
    """isaacgymenvs/RPO-LSTM/utils.py:37-39."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def observation(self, obs):
        return obs["obs"]

    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        return self.observation(observation), reward, done, info
