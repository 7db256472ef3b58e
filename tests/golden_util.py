"""Helpers to replay tests/golden/*.npz (fixtures written by oracle/gen_golden.py from the
unmodified Python reference)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["cfg"] = json.loads(str(g["cfg"]))
    return g


def env_cfg(g):
    """The METADATA-style config of a fixture (without the recorder's own keys)."""
    return {k: v for k, v in g["cfg"].items() if k not in ("policy", "steps")}


def expected_obs(g, f):
    """World.get_state() of frame f (environment.py:399-402) from the recorded planes."""
    return np.stack([g["apos"][f], (g["type"][f] == 1).astype(np.uint8), 1 - g["fm_inf"][f]], axis=-1)
