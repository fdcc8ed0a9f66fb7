"""`Env` / `VecTask`: the reference's vectorised-task surface on top of the fused CUDA step.

Mirrors isaacgymenvs/tasks/base/vec_task.py:60-406 -- same constructor signature, attributes
(`num_envs num_obs num_acts num_states observation_space action_space state_space device rl_device
obs_buf rew_buf reset_buf progress_buf timeout_buf randomize_buf states_buf extras control_freq_inv
max_episode_length`), same `step / reset / reset_done / reset_idx / get_state / zero_actions`
contracts -- so the reference's trainers (e.g. isaacgymenvs/RPO-LSTM/main.py:41-58,81,102) run
unchanged.  What differs: `step` is ONE kernel launch (no pre/post hooks, no host syncs, no
`gym.simulate`), and there is no CPU pipeline.
"""
from typing import Any, Dict, Tuple

import numpy as np
import torch

from .spaces import Box

import os as _os
_NVTX = _os.environ.get("OUZELUM_B200_NVTX", "0") == "1"


class Env:
    def __init__(self, config: Dict[str, Any], rl_device: str, sim_device: str, graphics_device_id: int, headless: bool):
        split_device = str(sim_device).split(":")
        self.device_type = split_device[0]
        self.device_id = int(split_device[1]) if len(split_device) > 1 else 0
        if self.device_type.lower() not in ("cuda", "gpu"):
            # vec_task.py:74-81 falls back to the CPU pipeline here; this framework has none.
            raise RuntimeError(f"sim_device={sim_device!r}: ouzelum_b200 has no CPU pipeline (sm_100a CUDA only)")
        self.device = "cuda:" + str(self.device_id)
        self.rl_device = rl_device
        self._tdev = torch.device(self.device)
        self._rl_on_sim_device = torch.device(rl_device) == self._tdev
        self.headless = headless
        self.graphics_device_id = -1 if headless else graphics_device_id

        self.num_environments = config["env"]["numEnvs"]
        self.num_agents = config["env"].get("numAgents", 1)
        self.num_observations = config["env"]["numObservations"]
        self.num_states = config["env"].get("numStates", 0)
        self.num_actions = config["env"]["numActions"]
        self.control_freq_inv = config["env"].get("controlFrequencyInv", 1)

        self.obs_space = Box(np.ones(self.num_obs) * -np.inf, np.ones(self.num_obs) * np.inf)
        self.state_space = Box(np.ones(self.num_states) * -np.inf, np.ones(self.num_states) * np.inf)
        self.act_space = Box(np.ones(self.num_actions) * -1., np.ones(self.num_actions) * 1.)

        self.clip_obs = config["env"].get("clipObservations", np.inf)
        self.clip_actions = config["env"].get("clipActions", np.inf)

    @property
    def observation_space(self):
        return self.obs_space

    @property
    def action_space(self):
        return self.act_space

    @property
    def num_envs(self) -> int:
        return self.num_environments

    @property
    def num_acts(self) -> int:
        return self.num_actions

    @property
    def num_obs(self) -> int:
        return self.num_observations


class VecTask(Env):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 24}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, config, rl_device, sim_device, graphics_device_id, headless,
                 virtual_screen_capture: bool = False, force_render: bool = False):
        super().__init__(config, rl_device, sim_device, graphics_device_id, headless)
        self.cfg = config
        engine = config.get("physics_engine", "physx")
        if engine not in ("physx", "b200"):
            raise ValueError(f"Invalid physics engine backend: {engine}")          # vec_task.py:194-196
        up = config.get("sim", {}).get("up_axis", "z")
        if up != "z":
            raise ValueError(f"Invalid physics up-axis: {up}")                     # vec_task.py:454-457
        self.virtual_screen_capture = virtual_screen_capture
        self.force_render = force_render
        self.viewer = None
        self.dr_randomizations = {}
        self.sim_initialized = False
        self.create_sim()
        self.sim_initialized = True
        self.allocate_buffers()
        self.obs_dict = {}

    # ---- buffers (vec_task.py:254-277) --------------------------------------------------------------
    def allocate_buffers(self):
        dev = self.device
        self.obs_buf = torch.zeros((self.num_envs, self.num_obs), device=dev, dtype=torch.float)
        self.states_buf = torch.zeros((self.num_envs, self.num_states), device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(self.num_envs, device=dev, dtype=torch.long)
        self.timeout_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.bool)
        self.progress_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.long)
        self.randomize_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.long)
        self.extras = {}

    # ---- hooks a task implements --------------------------------------------------------------------
    def create_sim(self):
        raise NotImplementedError

    def _fused_step(self, actions: torch.Tensor):
        """Launch the task's fused step kernel: consumes `actions`, updates every *_buf in place."""
        raise NotImplementedError

    # ---- surface ------------------------------------------------------------------------------------
    def get_state(self):
        return torch.clamp(self.states_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)

    def step(self, actions: torch.Tensor) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor, Dict[str, Any]]:
        """vec_task.py:313-359.  Action clamp, pre/post physics, reward, reset flags, time-outs and the
        observation clamp all happen inside the one kernel `_fused_step` launches."""
        if actions.dtype != torch.float32 or actions.device != self._tdev or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if _NVTX:                                   # OUZELUM_B200_NVTX=1: one NVTX range per env step (nsys / ncu --nvtx)
            torch.cuda.nvtx.range_push(f"{type(self).__name__}.step")
            self._fused_step(actions)
            torch.cuda.nvtx.range_pop()
        else:
            self._fused_step(actions)
        if self._rl_on_sim_device:                  # the usual case: no `.to()` round trips through the dispatcher
            self.extras["time_outs"] = self.timeout_buf
            self.obs_dict["obs"] = self.obs_buf
            if self.num_states > 0:
                self.obs_dict["states"] = self.get_state()
            return self.obs_dict, self.rew_buf, self.reset_buf, self.extras
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = self.obs_buf.to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def zero_actions(self) -> torch.Tensor:
        return torch.zeros([self.num_envs, self.num_actions], dtype=torch.float32, device=self.rl_device)

    def reset_idx(self, env_idx):
        pass

    def reset(self):
        """vec_task.py:377-389: returns the (initially zero) observation buffer; no simulation."""
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict

    def reset_done(self):
        """vec_task.py:391-406."""
        done_env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        if len(done_env_ids) > 0:
            self.reset_idx(done_env_ids)
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, done_env_ids

    def render(self, mode="rgb_array"):
        return None                      # headless only: no viewer in this framework

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self
