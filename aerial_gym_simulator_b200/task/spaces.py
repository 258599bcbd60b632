"""gym / gymnasium spaces when installed, else minimal stand-ins with the attributes the RL
wrappers read (rl_training/rl_games/runner.py:68-79, sample_factory/.../train_aerialgym.py:38-46)."""
import numpy as np

try:  # pragma: no cover - neither is installed in the build image
    from gymnasium.spaces import Box, Dict
except Exception:  # noqa: BLE001
    try:
        from gym.spaces import Box, Dict
    except Exception:  # noqa: BLE001

        class Box:
            def __init__(self, low, high, shape, dtype=np.float32):
                self.low = np.full(shape, low, dtype=dtype)
                self.high = np.full(shape, high, dtype=dtype)
                self.shape, self.dtype = tuple(shape), dtype

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

        class Dict(dict):
            def __init__(self, spaces):
                super().__init__(spaces)
                self.spaces = spaces
