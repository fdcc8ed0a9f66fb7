"""Trainer-side env wrappers (mirror of isaacgymenvs/<trainer>/utils.py:4-39, identical in all eight trainers)."""
import torch

from ._lib import check, lib


class RecordEpisodeStatisticsTorch:
    """isaacgymenvs/RPO-LSTM/utils.py:4-35 -- six element-wise launches per step fused into one kernel."""

    def __init__(self, env, device):
        self.env = env
        self.num_envs = getattr(env, "num_envs", 1)
        self.device = device                                   # where the trainer wants `infos["r"]` / `infos["l"]` (reference: any device)
        # the statistics kernel runs on the env's sim device; results are moved to `device` when the two differ
        self._sim_device = torch.device(getattr(env, "device", device))
        if self._sim_device.type != "cuda":
            raise RuntimeError("ouzelum_b200.RecordEpisodeStatisticsTorch needs an env on a CUDA device (no CPU fallback)")
        self._same_device = torch.device(device) == self._sim_device
        self.episode_returns = None
        self.episode_lengths = None

    def __getattr__(self, name):           # gym.Wrapper forwards unknown attributes to the wrapped env
        return getattr(self.env, name)

    def reset(self, **kwargs):
        observations = self.env.reset(**kwargs)
        z = lambda dt: torch.zeros(self.num_envs, dtype=dt, device=self._sim_device)
        self.episode_returns, self.episode_lengths = z(torch.float32), z(torch.int32)
        self.returned_episode_returns, self.returned_episode_lengths = z(torch.float32), z(torch.int32)
        return observations

    def step(self, action):
        observations, rewards, dones, infos = self.env.step(action)
        if self.episode_returns is None:
            raise RuntimeError("RecordEpisodeStatisticsTorch.step called before reset()")
        # the kernel dereferences raw device pointers: bring the inputs to the sim device / dtype / layout it expects (rl_device
        # may be the CPU -- vec_task.py:353-359 -- and a caller may hand back e.g. a bool `dones`)
        r = rewards if (rewards.device == self._sim_device and rewards.dtype == torch.float32 and rewards.is_contiguous()) \
            else rewards.to(device=self._sim_device, dtype=torch.float32).contiguous()
        d = dones if (dones.device == self._sim_device and dones.dtype == torch.int64 and dones.is_contiguous()) \
            else dones.to(device=self._sim_device, dtype=torch.int64).contiguous()
        if r.shape != (self.num_envs,) or d.shape != (self.num_envs,):
            raise ValueError(f"expected rewards / dones of shape ({self.num_envs},), got {tuple(r.shape)} / {tuple(d.shape)}")
        with torch.cuda.device(self._sim_device):
            check(lib.ozl_episode_stats(self.num_envs, r.data_ptr(), d.data_ptr(), self.episode_returns.data_ptr(),
                                        self.episode_lengths.data_ptr(), self.returned_episode_returns.data_ptr(),
                                        self.returned_episode_lengths.data_ptr(), torch.cuda.current_stream().cuda_stream))
        infos["r"] = self.returned_episode_returns if self._same_device else self.returned_episode_returns.to(self.device)
        infos["l"] = self.returned_episode_lengths if self._same_device else self.returned_episode_lengths.to(self.device)
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
