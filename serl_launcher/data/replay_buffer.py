"""reference data/replay_buffer.py:40-90 -> serl_b200."""
from serl_b200.data.replay_buffer import ReplayBuffer  # noqa: F401
