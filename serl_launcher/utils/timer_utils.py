"""reference utils/timer_utils.py -> serl_b200."""
from serl_b200.utils.timer_utils import Timer  # noqa: F401
