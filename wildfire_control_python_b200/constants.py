"""Environment parameters -- same key names and defaults as the reference's module-global
``METADATA`` dict (Simulation/constants.py:30-47) and ``grass`` cell parameters
(Simulation/utility.py:94-102).  The DQN hyper-parameters of that dict (constants.py:48-56) are
carried along for the learners of ``agents.py``; the environment ignores them.

Unlike the reference (WIDTH/HEIGHT frozen at import, environment.py:22-23) the size is an
ordinary per-instance parameter here.
"""
from __future__ import annotations

SIZE = 10
A_SPEED = 1

METADATA = {
    # reward measure
    "death_penalty": -1000 * A_SPEED,
    "contained_bonus": 1000 * A_SPEED,
    "default_reward": -1,
    # simulation constants
    "width": SIZE,
    "height": SIZE,
    "wind": [0.54, (0, 0)],
    "debug": 1,
    "n_actions": 4,
    "a_speed": A_SPEED,
    "a_speed_iter": A_SPEED,
    "make_rivers": False,
    "containment_wins": False,
    "allow_dig_toggle": False,
}

# DQN parameters (Simulation/constants.py:48-56), read by agents.py only
DQN_DEFAULTS = {
    "memory_size": 20000,
    "max_eps": 1.0,
    "min_eps": 0.01,
    "eps_decay_rate": 0.005,
    "gamma": 0.999,
    "alpha": 0.005,
    "target_update": 20,
    "batch_size": 32,
}

grass = {"heat": 0.3, "fuel": 20, "threshold": 3, "radius": 1}

# Simulation/utility.py:115-140
layer = {"type": 0, "gray": 1, "temp": 2, "heat": 3, "fuel": 4, "threshold": 5, "agent_pos": 6,
         "fire_mobility": 7, "agent_mobility": 8}
types = {0: "grass", 1: "fire", 2: "burnt", 3: "dirt", 4: "water",
         "grass": 0, "fire": 1, "burnt": 2, "dirt": 3, "water": 4}
# ForestFire.render symbols (utility.py:143-149, via the grey level of each type)
ascii_of_type = {0: "+", 1: "@", 2: "#", 3: "0", 4: "x"}

# Extensions of the batched implementation (not in the reference's dict)
EXTRA_DEFAULTS = {"seed": 0, "extra_ignitions": 0, "auto_reset": False, "env_id_base": 0}


def make_metadata(**overrides) -> dict:
    m = dict(METADATA)
    m.update(DQN_DEFAULTS)
    m.update(grass)
    m.update(EXTRA_DEFAULTS)
    if "size" in overrides:
        s = overrides.pop("size")
        m["width"] = m["height"] = s
    unknown = set(overrides) - set(m)
    if unknown:
        raise KeyError(f"unknown METADATA keys: {sorted(unknown)}")
    m.update(overrides)
    m["a_speed_iter"] = m["a_speed"]
    return m
