"""Minimal `Box` space so the env surface works without `gym` (absent from this image).

If `gym` is importable its `spaces.Box` is used, so `gym.Wrapper`s (isaacgymenvs/RPO-LSTM/utils.py:4-39)
see the real thing; otherwise this duck-typed stand-in provides `.shape/.low/.high/.dtype/.sample()`.
Mirrors isaacgymenvs/tasks/base/vec_task.py:102-105.
"""
import numpy as np

try:                                  # pragma: no cover - gym is not in the build image
    from gym.spaces import Box        # type: ignore
except Exception:                     # noqa: BLE001
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            if shape is not None:
                low = np.broadcast_to(low, shape).copy()
                high = np.broadcast_to(high, shape).copy()
            self.low, self.high = low, high
            self.shape = low.shape
            self.dtype = np.dtype(dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
