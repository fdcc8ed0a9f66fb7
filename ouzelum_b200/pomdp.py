"""`POMDPWrapper`: the reference's sensor-fault model (isaacgymenvs/utils/POMDP.py:4-42) as one kernel launch.

Same constructor and `observation(obs)` contract.  Differences, both deliberate: draws come from the counter RNG
(seed, env, call index) instead of the CPU generator + H2D copy of a noise tensor every call, and the output stays on
the input's CUDA device instead of the hard-coded "cuda:0".
"""
import torch

from ._lib import check, lib

_MODES = {"flicker": 1, "random_noise": 2, "flickering_and_random_noise": 3}


class POMDPWrapper:
    def __init__(self, pomdp="flicker", pomdp_prob=0.1, seed=0, env_id_base=0, stream_id=1):
        self.pomdp = pomdp
        self.flicker_prob = pomdp_prob
        self.random_noise_sigma = pomdp_prob
        self.range = (1 - self.random_noise_sigma, 1 + self.random_noise_sigma)
        if pomdp not in _MODES:
            raise ValueError("pomdp was not in ['remove_velocity', 'flickering', 'random_noise', 'random_sensor_missing']!")
        self.prob = pomdp_prob
        if pomdp == "flickering_and_random_noise":
            self.flicker_prob = 0.1                                           # POMDP.py:16-18
        self.mode = _MODES[pomdp]
        self.seed, self.env_id_base, self.stream_id = int(seed), int(env_id_base), int(stream_id)
        self.calls = 0
        self._step_ptr = None

    def follow_step_counter(self, env):
        """Take the call index from `env`'s device step counter instead of the host-side call count: `observation` then has
        no host-changing argument and can be captured in a CUDA graph together with the env step."""
        import ctypes as C
        p = C.c_void_p()
        check(lib.ozl_step_counter_ptr(env.sim._h, C.byref(p)))
        self._step_ptr = p.value

    def observation(self, obs):
        x = obs.to(torch.float32).contiguous()
        if x.device.type != "cuda":
            raise RuntimeError("ouzelum_b200.POMDPWrapper needs a CUDA tensor (no CPU fallback)")
        flat = x.reshape(-1, x.shape[-1]) if x.dim() > 1 else x.reshape(1, -1)
        out = torch.empty_like(flat)
        if self._step_ptr is not None:
            check(lib.ozl_pomdp_observation_dev(flat.shape[0], flat.shape[1], self.mode, float(self.prob), self.seed,
                                                self._step_ptr, self.env_id_base, self.stream_id, flat.data_ptr(),
                                                out.data_ptr(), torch.cuda.current_stream().cuda_stream), ValueError)
        else:
            check(lib.ozl_pomdp_observation(flat.shape[0], flat.shape[1], self.mode, float(self.prob), self.seed, self.calls,
                                            self.env_id_base, self.stream_id, flat.data_ptr(), out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), ValueError)
        self.calls += 1
        return out.reshape(x.shape)
