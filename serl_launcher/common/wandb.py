"""reference common/wandb.py -> serl_b200."""
from serl_b200.common.wandb import WandBLogger  # noqa: F401
