from serl_b200.agents.continuous.drq import DrQAgent  # noqa: F401
from serl_b200.agents.continuous.sac import SACAgent  # noqa: F401
