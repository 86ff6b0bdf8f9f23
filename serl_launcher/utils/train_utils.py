"""reference utils/train_utils.py:16-130 -> serl_b200."""
from serl_b200.utils.train_utils import _unpack, concat_batches, load_resnet10_params  # noqa: F401
