"""reference agents/continuous/drq.py -> serl_b200."""
from serl_b200.agents.continuous.drq import DrQAgent  # noqa: F401
