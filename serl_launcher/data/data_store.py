"""reference data/data_store.py:26-162 -> serl_b200."""
from serl_b200.data.data_store import (MemoryEfficientReplayBufferDataStore, ReplayBufferDataStore,  # noqa: F401
                                       populate_data_store)
