"""ouzelum_b200 -- B200-native vectorised quadcopter environments behind the reference's VecTask surface.

`make()` has the signature of `isaacgymenvs.make` (isaacgymenvs/__init__.py:14-55) and returns a
`VecTask`-compatible env whose `step` is one fused sm_100a kernel.  Importing this package loads
libouzelum_b200.so; there is no CPU or PyTorch fallback.
"""
from . import _lib                      # noqa: F401  (fails loudly if the native library is missing)
from .cfg import task_config, task_names
from .sim import QuadSim
from .tasks import task_map

__all__ = ["make", "task_map", "task_config", "task_names", "QuadSim"]


def make(seed: int, task: str, num_envs: int, sim_device: str = "cuda:0", rl_device: str = "cuda:0",
         graphics_device_id: int = -1, headless: bool = True, multi_gpu: bool = False,
         virtual_screen_capture: bool = False, force_render: bool = False, cfg=None):
    """Drop-in for `isaacgymenvs.make`.  `cfg` may be a plain nested dict (or an object with `.task`) in the
    layout of isaacgymenvs/cfg/task/<Task>.yaml; when omitted the built-in mirror of that YAML is used.
    The reference accepts `seed` and ignores it (utils/rlgames_utils.py:42,78-86 -- env randomness comes
    from the global torch generator); here it seeds the counter-based RNG unless cfg["env"]["seed"] is set."""
    if cfg is None:
        cfg_dict = task_config(task, num_envs)
        cfg_dict["env"]["seed"] = int(seed)
    else:
        cfg_dict = getattr(cfg, "task", cfg)
        cfg_dict["env"].setdefault("seed", int(seed))
    if multi_gpu:
        import os
        rank = int(os.getenv("LOCAL_RANK", "0"))      # isaacgymenvs/train.py:74-82
        world = int(os.getenv("WORLD_SIZE", "1"))
        cfg_dict["env"].setdefault("envIdBase", int(os.getenv("RANK", str(rank))) * int(cfg_dict["env"]["numEnvs"]))
        del world
    name = cfg_dict["name"]
    if name not in task_map:
        raise KeyError(f"unknown task {name!r}; known: {sorted(task_map)}")
    return task_map[name](cfg=cfg_dict, rl_device=rl_device, sim_device=sim_device,
                          graphics_device_id=graphics_device_id, headless=headless,
                          virtual_screen_capture=virtual_screen_capture, force_render=force_render)
