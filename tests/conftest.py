import glob
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def golden_id(path):
    return os.path.splitext(os.path.basename(path))[0]


def load_golden(path):
    """-> (dict of arrays, VineConfig, plain task cfg dict) for a fixture made by generate_golden.py."""
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    z = np.load(path)
    data = {k: z[k] for k in z.files}
    overrides = json.loads(str(data.pop("overrides")))
    task_cfg = vcfg.task_config(overrides)
    return data, vcfg.task_cfg_to_vine_config(task_cfg), task_cfg


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.build()
    return O
