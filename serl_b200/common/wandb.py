"""Logger with the surface the learner scripts use (reference common/wandb.py: `WandBLogger.get_default_config()`,
`WandBLogger(wandb_config, variant, debug)`, `.log(data, step)`): nested info dicts are flattened to "a/b" keys and
device scalars converted on the way.  wandb itself is optional: without it (or with debug=True) records are kept in
`self.history` so that a learner loop runs unchanged offline."""
from __future__ import annotations


def _flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        key = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(_flatten(v, key))
        else:
            out[key] = float(v) if hasattr(v, "__float__") else v
    return out


class WandBLogger:
    @staticmethod
    def get_default_config():
        return {"project": "serl_launcher", "entity": None, "exp_descriptor": "", "unique_identifier": "", "tag": None, "group": None}

    def __init__(self, wandb_config, variant, wandb_output_dir=None, debug=False):
        self.config, self.variant, self.history = dict(wandb_config), variant, []
        self.run = None
        if not debug:
            try:
                import wandb
                self.run = wandb.init(project=self.config.get("project"), entity=self.config.get("entity"), config=variant,
                                      tags=[self.config["tag"]] if self.config.get("tag") else None, group=self.config.get("group"),
                                      dir=wandb_output_dir, name=self.config.get("exp_descriptor") or None)
            except Exception:                          # noqa: BLE001  (wandb absent / offline): keep records locally
                self.run = None

    def log(self, data: dict, step: int = None):
        flat = _flatten(data)
        if self.run is not None:
            self.run.log(flat, step=step)
        else:
            self.history.append((step, flat))
