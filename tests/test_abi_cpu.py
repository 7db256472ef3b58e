"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/wildfire.h declares, validates arguments, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

from wildfire_control_python_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wildfire.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wf_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert declared_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_default_config(lib):
    assert lib.wf_abi_version() == 2
    cfg = _lib.WfConfig()
    lib.wf_default_config(C.byref(cfg), 14)
    # Simulation/constants.py:30-47 + utility.py:94-102
    assert (cfg.width, cfg.height, cfg.n_actions, cfg.a_speed) == (14, 14, 4, 1)
    assert (cfg.wind_speed, cfg.wind_x, cfg.wind_y) == (0.54, 0, 0)
    assert (cfg.death_penalty, cfg.contained_bonus, cfg.default_reward) == (-1000.0, 1000.0, -1.0)
    assert (cfg.heat, cfg.fuel, cfg.threshold, cfg.radius) == (0.3, 20, 3.0, 1)
    assert not (cfg.make_rivers or cfg.allow_dig_toggle or cfg.containment_wins or cfg.wind_random)


def test_config_struct_matches_header_size():
    # 14 int32 + 6 double + uint64 + int64
    assert C.sizeof(_lib.WfConfig) == 14 * 4 + 6 * 8 + 8 + 8


def test_create_validates_arguments_and_never_falls_back(lib):
    cfg = _lib.WfConfig()
    lib.wf_default_config(C.byref(cfg), 10)
    h = C.c_void_p()
    cfg.radius = 2
    assert lib.wf_create(C.byref(cfg), 4, 0, C.byref(h)) == _lib.WF_ERR_INVALID
    assert b"radius" in lib.wf_last_error()
    cfg.radius = 1
    cfg.width = cfg.height = 8
    assert lib.wf_create(C.byref(cfg), 4, 0, C.byref(h)) == _lib.WF_ERR_INVALID
    cfg.width = cfg.height = 10
    import torch
    if not torch.cuda.is_available():
        rc = lib.wf_create(C.byref(cfg), 4, 0, C.byref(h))
        assert rc == _lib.WF_ERR_CUDA and b"no CPU fallback" in lib.wf_last_error()
        assert not h.value


def test_host_class_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from wildfire_control_python_b200 import BatchedForestFire
    with pytest.raises(_lib.WildfireError):
        BatchedForestFire(4)


def test_metadata_defaults_follow_the_reference():
    from wildfire_control_python_b200.constants import make_metadata
    m = make_metadata()
    assert m["wind"] == [0.54, (0, 0)] and m["n_actions"] == 4 and m["a_speed"] == 1
    assert m["death_penalty"] == -1000 and m["contained_bonus"] == 1000 and m["default_reward"] == -1
    with pytest.raises(KeyError):
        make_metadata(not_a_key=1)
