"""reference agents/continuous/sac.py -> serl_b200."""
from serl_b200.agents.continuous.sac import SACAgent  # noqa: F401
