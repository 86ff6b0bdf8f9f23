"""reference utils/launcher.py:50-272 -> serl_b200."""
from serl_b200.utils.launcher import (make_bc_agent, make_drq_agent, make_replay_buffer, make_sac_agent, make_trainer_config,  # noqa: F401
                                      make_wandb_logger)
