"""Drop-in import surface: the reference's package name and module paths (`serl_launcher.utils.launcher`,
`serl_launcher.utils.train_utils`, `serl_launcher.data.data_store`, `serl_launcher.agents.continuous.{sac,drq}` ...),
served by the B200 learner in `serl_b200`.  With this directory ahead of the reference checkout on `sys.path` the learner
side of `examples/async_*_sim/*.py` imports resolve here; the actor side and everything outside the
learner hot path (wrappers, envs, vision definitions) stay with the reference package."""
from serl_b200 import __doc__ as _doc  # noqa: F401
