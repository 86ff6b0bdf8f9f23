"""Launcher factories with the reference's names and hyper-parameters (utils/launcher.py:50-116,201-272):
`make_bc_agent`, `make_sac_agent`, `make_drq_agent`, `make_replay_buffer`, `make_trainer_config`, `make_wandb_logger`.
"""
from __future__ import annotations

from typing import Optional

from ..agents.continuous.bc import BCAgent
from ..agents.continuous.drq import DrQAgent
from ..agents.continuous.sac import SACAgent
from ..data.data_store import MemoryEfficientReplayBufferDataStore, ReplayBufferDataStore


def make_bc_agent(seed, sample_obs, sample_action, image_keys=("image",), encoder_type="small", precision="fp32", device=None):
    """utils/launcher.py:26-47."""
    return BCAgent.create(
        seed, sample_obs, sample_action,
        network_kwargs={"activations": "tanh", "use_layer_norm": False, "hidden_dims": [256, 256]},
        policy_kwargs={"tanh_squash_distribution": False, "std_parameterization": "exp", "std_min": 1e-5, "std_max": 5},
        use_proprio=True, encoder_type=encoder_type, image_keys=image_keys, precision=precision, device=device)


def make_sac_agent(seed, sample_obs, sample_action, discount=0.99, device=None):
    """utils/launcher.py:50-76."""
    return SACAgent.create_states(
        seed, sample_obs, sample_action,
        policy_kwargs={"tanh_squash_distribution": True, "std_parameterization": "exp", "std_min": 1e-5, "std_max": 5},
        temperature_init=1e-2, discount=discount, backup_entropy=False, critic_ensemble_size=10, critic_subsample_size=2,
        device=device)


def make_drq_agent(seed, sample_obs, sample_action, image_keys=("image",), encoder_type="small", discount=0.96,
                   precision="fp32", device=None):
    """utils/launcher.py:79-116."""
    return DrQAgent.create_drq(
        seed, sample_obs, sample_action, encoder_type=encoder_type, use_proprio=True, image_keys=image_keys,
        policy_kwargs={"tanh_squash_distribution": True, "std_parameterization": "exp", "std_min": 1e-5, "std_max": 5},
        temperature_init=1e-2, discount=discount, backup_entropy=False, critic_ensemble_size=10, critic_subsample_size=2,
        precision=precision, device=device)


def make_replay_buffer(env, capacity: int = 1000000, rlds_logger_path: Optional[str] = None, type: str = "replay_buffer",
                       image_keys: list = [], preload_rlds_path: Optional[str] = None, preload_data_transform=None,
                       device=None, seed=None):
    """utils/launcher.py:201-272 (RLDS logging / tfds preload are outside the hot path and unsupported)."""
    if rlds_logger_path or preload_rlds_path:
        raise NotImplementedError("RLDS logging / preload need oxe_envlogger + tensorflow_datasets (not on the hot path)")
    if type == "replay_buffer":
        return ReplayBufferDataStore(env.observation_space, env.action_space, capacity=capacity, device=device, seed=seed)
    if type == "memory_efficient_replay_buffer":
        return MemoryEfficientReplayBufferDataStore(env.observation_space, env.action_space, capacity=capacity,
                                                    image_keys=image_keys, device=device, seed=seed)
    raise ValueError(f"Unsupported replay_buffer_type: {type}")


def make_trainer_config(port_number: int = 5488, broadcast_port: int = 5489):
    """utils/launcher.py:171-177 (needs agentlace, which is consumed unchanged)."""
    from agentlace.trainer import TrainerConfig
    return TrainerConfig(port_number=port_number, broadcast_port=broadcast_port, request_types=["send-stats"])


def make_wandb_logger(project: str = "agentlace", description: str = "serl_launcher", debug: bool = False):
    """utils/launcher.py:180-198."""
    from ..common.wandb import WandBLogger
    cfg = WandBLogger.get_default_config()
    cfg.update({"project": project, "exp_descriptor": description, "tag": description})
    return WandBLogger(wandb_config=cfg, variant={}, debug=debug)
