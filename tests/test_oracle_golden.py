"""The C oracle must reproduce every golden trajectory recorded from the Python reference."""
import numpy as np
import pytest

from oracle import wf_oracle as wo
from tests.golden_util import env_cfg, expected_obs, golden_names, load_golden


def check_frame(tag, g, f, env, obs):
    p = env.planes()
    for k in ("type", "burning", "fm_inf", "apos"):
        assert np.array_equal(p[k], g[k][f]), f"{tag}: plane {k}"
    assert np.array_equal(p["fuel"].astype(np.uint8), g["fuel"][f]), f"{tag}: fuel"
    for k in ("alive", "ax", "ay", "fire_at_border", "running", "wind_x", "wind_y"):
        assert p[k] == int(g[k][f]), f"{tag}: {k}"
    assert p["wind_speed"] == float(g["wind_speed"][f])
    grass = g["type"][f] == 0
    assert np.abs(p["temp"] - g["temp"][f])[grass].max(initial=0.0) <= 1e-12, f"{tag}: temp"
    assert np.array_equal(obs, expected_obs(g, f)), f"{tag}: obs"


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(name):
    g = load_golden(name)
    env = wo.OracleEnv(env_cfg(g))
    F = len(g["kind"])
    assert g["kind"][0] == 0
    for f in range(F):
        tag = f"{name} frame {f}"
        if g["kind"][f] == 0:
            obs = env.reset()
        else:
            obs, r, d, _ = env.step(int(g["action"][f]))
            assert r == float(g["reward"][f]), f"{tag}: reward {r} != {g['reward'][f]}"
            assert d == bool(g["done"][f]), f"{tag}: done"
        check_frame(tag, g, f, env, obs)


def test_free_burn_ignition_map_14():
    """ignition ticks of a 14x14 free burn, agent parked at (7, 10)."""
    want_ul = np.array([
        [165, 155, 147, 141, 137, 135, 134, 133],
        [155, 143, 133, 125, 120, 117, 115, 114],
        [147, 133, 121, 111, 104, 99, 96, 95],
        [141, 125, 111, 99, 89, 82, 78, 76],
        [137, 120, 104, 89, 76, 66, 60, 57],
        [135, 117, 99, 82, 66, 53, 43, 38],
        [134, 115, 96, 78, 60, 43, 29, 19],
        [133, 114, 95, 76, 57, 38, 19, 0]])  # rows = y, cols = x
    env = wo.OracleEnv(dict(width=14, height=14, seed=0))
    env.reset(start=(7, 10))
    ign = np.full((14, 14), -1)
    ign[7, 7] = 0
    t, done = 0, False
    while not done:
        _, r, done, _ = env.step(5)  # no-op action
        t += 1
        typ = env.planes()["type"]
        ign[(typ == 1) & (ign < 0)] = t
    assert t == 185
    assert (env.planes()["type"] == 0).sum() == 0
    assert np.array_equal(ign[:8, :8].T, want_ul)


@pytest.mark.parametrize("wind,ticks,healthy", [([0.85, (1, 0)], 92, 188), ([0.7, (1, 1)], 20, 194)])
def test_free_burn_wind_known_answers(wind, ticks, healthy):
    """wind line fire / non-spreading diagonal wind."""
    env = wo.OracleEnv(dict(width=14, height=14, seed=0, wind=wind))
    env.reset(start=(7, 10))
    t, done = 0, False
    while not done:
        _, r, done, _ = env.step(5)
        t += 1
    assert t == ticks
    assert (env.planes()["type"] == 0).sum() == healthy
    assert r == 1000 * (healthy / 196)
