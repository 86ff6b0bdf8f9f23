"""reference data/memory_efficient_replay_buffer.py:12-164 -> serl_b200."""
from serl_b200.data.memory_efficient_replay_buffer import MemoryEfficientReplayBuffer  # noqa: F401
