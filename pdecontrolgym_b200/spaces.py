"""`Box` space: gymnasium's when it is installed, else a minimal stand-in with the same attributes
(`low`, `high`, `shape`, `dtype`, `sample`) -- the reference only ever builds Box spaces
(base_env_1d.py:33-35, base_env_2d.py:37-44)."""
import numpy as np

try:  # pragma: no cover - gymnasium is absent from the build image
    from gymnasium.spaces import Box  # type: ignore
except Exception:  # noqa: BLE001
    class Box:  # type: ignore
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low) if np.ndim(low) else np.shape(high)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def __repr__(self):
            return f"Box(shape={self.shape}, dtype={self.dtype})"
