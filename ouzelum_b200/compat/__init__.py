"""Run the reference's trainers UNMODIFIED against ouzelum_b200 ("compat launcher").

The reference scripts (e.g. isaacgymenvs/RPO-LSTM/main.py:6-14) do

    import gym, isaacgym, isaacgymenvs
    from isaacgymenvs.utils.POMDP import POMDPWrapper
    from torch.utils.tensorboard import SummaryWriter
    envs = isaacgymenvs.make(seed=0, task="Landing", num_envs=4096, sim_device="cuda:0", rl_device="cuda:0", ...)

`install()` registers, for every one of those module names that is NOT importable in the running interpreter, a small
stand-in in `sys.modules`:

    isaacgym                       empty shell (the trainers import it only for its side effects)
    isaacgymenvs                   `make` = ouzelum_b200.make, `isaacgymenvs.utils.POMDP.POMDPWrapper` = ouzelum_b200.pomdp's
    gym                            `spaces.Box` (ouzelum_b200.spaces.Box), `Wrapper`, `ObservationWrapper` with gym 0.24 semantics
                                   (what isaacgymenvs/RPO-LSTM/utils.py:4-39 subclasses)
    torch.utils.tensorboard        `SummaryWriter` that drops everything (only if tensorboard is absent)

Real packages always win: nothing that imports is replaced.  Launcher:

    python -m ouzelum_b200.compat path/to/isaacgymenvs/RPO-LSTM/main.py --env Landing --num_envs 4096 --total_steps 200000

runs the script as `__main__` with its own directory first on `sys.path` (so `from agent import PPO` resolves to the
reference's file next to it) and the working directory unchanged.  `OUZELUM_COMPAT_DEFAULT_DEVICE=cuda:0` additionally calls
`torch.set_default_device`, which the torch-1.x-era update step of the reference's PPO needs under current torch (see `run`).
"""
import importlib
import importlib.util
import sys
import types


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _make_gym():
    from ..spaces import Box

    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box

    class Wrapper:
        """gym.Wrapper (gym 0.24): holds `env`, forwards unknown attributes, `step` / `reset` / `close` pass through."""

        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_") or name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return getattr(self.env, "unwrapped", self.env)

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def close(self):
            return self.env.close()

    class ObservationWrapper(Wrapper):
        def reset(self, **kwargs):
            return self.observation(self.env.reset(**kwargs))

        def step(self, action):
            observation, reward, done, info = self.env.step(action)
            return self.observation(observation), reward, done, info

        def observation(self, observation):
            raise NotImplementedError

    class Env:
        pass

    gym.spaces, gym.Wrapper, gym.ObservationWrapper, gym.Env = spaces, Wrapper, ObservationWrapper, Env
    gym.__version__ = "0.24.1-ouzelum_b200-shim"
    return gym, spaces


def _make_isaacgymenvs():
    import ouzelum_b200
    from ..pomdp import POMDPWrapper

    pkg = types.ModuleType("isaacgymenvs")
    pkg.__path__ = []                                    # a package, so `isaacgymenvs.utils.POMDP` can be imported
    pkg.make = ouzelum_b200.make
    utils = types.ModuleType("isaacgymenvs.utils")
    utils.__path__ = []
    pomdp = types.ModuleType("isaacgymenvs.utils.POMDP")
    pomdp.POMDPWrapper = POMDPWrapper
    tasks = types.ModuleType("isaacgymenvs.tasks")
    tasks.isaacgym_task_map = ouzelum_b200.task_map
    pkg.utils, utils.POMDP, pkg.tasks = utils, pomdp, tasks
    return {"isaacgymenvs": pkg, "isaacgymenvs.utils": utils, "isaacgymenvs.utils.POMDP": pomdp, "isaacgymenvs.tasks": tasks}


def _make_tensorboard():
    mod = types.ModuleType("torch.utils.tensorboard")

    class SummaryWriter:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):                     # add_scalar, add_text, flush, close, ...
            return lambda *a, **k: None

    mod.SummaryWriter = SummaryWriter
    return mod


def install():
    """Register the stand-ins described in the module docstring; returns the list of names that were installed."""
    done = []
    if _missing("isaacgym"):
        sys.modules["isaacgym"] = types.ModuleType("isaacgym")
        done.append("isaacgym")
    if _missing("isaacgymenvs"):
        sys.modules.update(_make_isaacgymenvs())
        done.append("isaacgymenvs")
    if _missing("gym"):
        gym, spaces = _make_gym()
        sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
        done.append("gym")
    if _missing("tensorboard"):
        import torch.utils
        mod = _make_tensorboard()
        sys.modules["torch.utils.tensorboard"] = mod
        torch.utils.tensorboard = mod
        done.append("torch.utils.tensorboard")
    return done


def run(script, argv=()):
    """Run `script` as __main__ with the stand-ins installed (what `python -m ouzelum_b200.compat script args...` does)."""
    import os
    import runpy
    install()
    dd = os.environ.get("OUZELUM_COMPAT_DEFAULT_DEVICE")
    if dd:
        # The reference's trainers were written against torch 1.x, which let a CPU index tensor be indexed with CUDA indices
        # (RPO-LSTM/agent.py:74,81: `flatinds = torch.arange(...)` on the CPU, `flatinds[:, mbenvinds]` with CUDA indices);
        # current torch refuses.  Making factory functions default to the GPU keeps those scripts unmodified.
        import torch
        torch.set_default_device(dd)
    script = os.path.abspath(script)
    old_argv, old_path = sys.argv, list(sys.path)
    sys.argv = [script] + list(argv)
    sys.path.insert(0, os.path.dirname(script))
    try:
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
