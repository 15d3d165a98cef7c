"""CPU: the C-ABI library loads and exports every symbol include/vine_b200.h declares; the ctypes
mirror matches the header; config errors are reported at vine_create without touching a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import REPO
from vine_robot_isaacgymenvs_b200 import abi, build as vbuild


@pytest.fixture(scope="module")
def lib():
    vbuild.build()
    return abi.load_library()


def header_functions():
    src = open(os.path.join(REPO, "include", "vine_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|void|const char\*)\s+(vine_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 18
    assert sorted(names) == sorted(abi.EXPORTED_SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layout_matches_library(lib):
    cfg = abi.VineConfig()
    assert lib.vine_config_defaults(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(abi.VineConfig)
    # spot-check fields at both ends and in the middle of the struct (YT defaults)
    assert (cfg.substeps, cfg.dt, cfg.control_freq_inv, cfg.max_episode_length) == (10, 0.00833, 4, 500)
    assert cfg.observation_type == abi.OBSERVATION_TYPES["POS_AND_FD_VEL_AND_OBJ_INFO"]
    assert list(cfg.reward_weights) == [0, 0, 1, 0, 0.1, 0, 0, 0, 0, 1, 0, 0, 0.10]
    assert (cfg.contact_stiffness, cfg.contact_damping, cfg.contact_rest_offset) == (2000.0, 2.0, 0.001)
    assert lib.vine_abi_version() == abi.ABI_VERSION


def test_ppo_argument_structs_match_the_library(lib):
    for i, st in enumerate(abi.PPO_STRUCTS):
        assert lib.vine_abi_struct_size(i) == C.sizeof(st), (st.__name__, lib.vine_abi_struct_size(i), C.sizeof(st))
    assert lib.vine_abi_struct_size(len(abi.PPO_STRUCTS)) < 0


def test_observation_widths(lib):
    for name, t in abi.OBSERVATION_TYPES.items():
        assert lib.vine_num_observations(t) == abi.NUM_OBSERVATIONS[t]
    assert lib.vine_num_observations(17) < 0


def test_config_errors_are_reported_at_create(lib):
    h = C.c_void_p()
    cfg = abi.default_config()
    cfg.observation_type = abi.OBSERVATION_TYPES["POS_AND_FD_VEL"]   # scaled + this type: V5:267-268 raises
    assert lib.vine_create(C.byref(cfg), 16, 0, 0, 42, C.byref(h)) == abi.ERR_UNSUPPORTED
    assert b"not implemented" in lib.vine_last_error(None)
    cfg = abi.default_config()
    cfg.struct_size = 12
    assert lib.vine_create(C.byref(cfg), 16, 0, 0, 42, C.byref(h)) == abi.ERR_ABI_MISMATCH
    cfg = abi.default_config()
    cfg.action_delay = 99
    assert lib.vine_create(C.byref(cfg), 16, 0, 0, 42, C.byref(h)) == abi.ERR_INVALID_ARG
    assert lib.vine_create(C.byref(abi.default_config()), 0, 0, 0, 42, C.byref(h)) == abi.ERR_INVALID_ARG
    assert lib.vine_step(None, None) == abi.ERR_INVALID_ARG


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(REPO, "vine_robot_isaacgymenvs_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "vine_oracle" not in text, f


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback|not found"):
        abi.load_library(str(tmp_path / "nope.so"))


def test_the_library_reads_no_environment_variable():
    """Launch tuning goes through VineConfig / the argument structs, never through a hidden getenv channel behind the C ABI."""
    import glob
    import os
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vine_robot_isaacgymenvs_b200", "csrc")
    for path in glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.h")):
        assert "getenv" not in open(path).read(), path
