"""The single-env ForestFire facade must look like Simulation.forest_fire.ForestFire to its callers
(DQN.py / misc.py call sites) and follow the reference's golden trajectory."""
import numpy as np
import pytest

from tests.golden_util import env_cfg, expected_obs, load_golden

pytestmark = pytest.mark.gpu


def test_facade_follows_reference_golden_c1():
    from wildfire_control_python_b200 import ForestFire
    g = load_golden("c1_10x10_seed0")
    sim = ForestFire(**env_cfg(g))
    # construction resets once (environment.py:183): episode 0 is consumed, like World.__init__ does
    # with its own throw-away draw; re-create the episode numbering by resetting through the batch
    sim._batch.set_state(scalars=_episode_minus_one(sim))
    F = len(g["kind"])
    for f in range(F):
        if g["kind"][f] == 0:
            state = sim.reset()
        else:
            state, reward, done, info = sim.step(int(g["action"][f]))
            assert reward == g["reward"][f] and done == bool(g["done"][f]) and info == {}
            assert isinstance(done, bool)
        assert state.dtype == np.float64 and state.shape == (10, 10, 3)
        assert np.array_equal(state, expected_obs(g, f).astype(np.float64))
        assert (len(sim.W.agents) == 1) == bool(g["alive"][f])
        if g["alive"][f]:
            assert (sim.W.agents[0].x, sim.W.agents[0].y) == (g["ax"][f], g["ay"][f])
        assert sim.W.RUNNING == bool(g["running"][f])
        env = sim.W.env
        assert env.shape == (10, 10, 9)
        assert np.array_equal(env[:, :, sim.layer["type"]], g["type"][f])
        assert np.array_equal(np.isinf(env[:, :, sim.layer["fire_mobility"]]), g["fm_inf"][f] != 0)


def _episode_minus_one(sim):
    import torch
    from wildfire_control_python_b200 import _lib
    sc = sim._batch.get_state()["scalars"].clone()
    sc[:, _lib.S_EPISODE] = -1
    return sc


def test_facade_attribute_surface_and_errors():
    from wildfire_control_python_b200 import ForestFire
    sim = ForestFire(width=10, height=10, seed=3)
    assert sim.n_actions == 4 and sim.width == 10 and sim.height == 10 and sim.DEBUG == 1
    assert (sim.W.WIDTH, sim.W.HEIGHT, sim.W.DEPTH) == (10, 10, 3)
    assert sim.W.wind_speed == 0.54 and sim.W.wind_vector == (0, 0)
    assert sim.METADATA["contained_bonus"] == 1000 and len(sim.W.border_points) == 40
    assert sim.W.burning_cells == {(5, 5)}
    text = sim.render(print_map=False)
    rows = text.strip("\n").split("\n")
    assert len(rows) == 10 and all(len(r) == 10 for r in rows)
    assert rows[5][5] == "@" and text.count("A") == 1 and text.count("@") == 1
    assert sim.get_name(10, 1, 0, "x").startswith("x-10s-1k-0m-")
    a = sim.W.agents[0]
    assert isinstance(a.fire_in_direction(0), bool)
    # walk into the fire: the reference returns the death penalty and done on that same step (Q2)
    sim2 = ForestFire(width=10, height=10, seed=3)
    sim2._batch.reset(starts=[[5, 6]])
    sim2._cache = None
    state, reward, done, _ = sim2.step("N")
    assert reward == -1000 and done and sim2.W.agents == []
    with pytest.raises(IndexError):
        sim2.step(0)
    sim2.reset()
    assert len(sim2.W.agents) == 1 and sim2.W.RUNNING
