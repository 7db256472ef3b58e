"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the oracle-side generators."""
import pytest

from oracle import philox as pyphilox
from oracle import wf_oracle as wo

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", KATS)
def test_python_philox_kat(ctr, key, want):
    assert tuple(pyphilox.philox4x32_10(ctr, key)) == want


@pytest.mark.parametrize("ctr,key,want", KATS)
def test_c_oracle_philox_kat(ctr, key, want):
    assert tuple(wo.philox(ctr, key)) == want


def test_streams_agree_between_python_and_c():
    e = wo.OracleEnv(dict(width=10, height=10, seed=0x1234_5678_9ABC))
    e.reset()
    for t in range(5):
        want = pyphilox.action_draw(0x1234_5678_9ABC, 0, 0, t) % 4
        assert e.random_action() == want
        e.step(5)  # no-op action advances t
