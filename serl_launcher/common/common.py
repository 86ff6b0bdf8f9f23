"""reference common/common.py:81-245 (JaxRLTrainState) -> serl_b200 TrainState."""
from serl_b200.common.common import TrainState  # noqa: F401
JaxRLTrainState = TrainState
