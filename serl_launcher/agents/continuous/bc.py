"""reference agents/continuous/bc.py -> serl_b200."""
from serl_b200.agents.continuous.bc import BCAgent  # noqa: F401
