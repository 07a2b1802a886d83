"""The C-ABI library loads and exports every symbol include/mcsim.h declares; ctypes mirrors match the header."""
import ctypes as C
import os
import re

import pytest

from interactive_highway_simulation_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mcsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcs_[a-z0-9_]+)\s*\(", text)))


def declared_env_functions():
    text = open(os.path.join(ROOT, "include", "mcsim_env.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mce_[a-z0-9_]+)\s*\(", text)))


def test_env_header_entry_points_are_exported(backend_lib):
    names = declared_env_functions()
    for n in ("mce_map_create", "mce_map_destroy", "mce_env_match", "mce_env_neighbors", "mce_frenet", "mce_project_center",
              "mce_step_along_path", "mce_last_error", "mce_last_kernel_ms", "mce_center_distance", "mce_border_distance"):
        assert n in names
    L = C.CDLL(backend_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libmcsim_b200.so does not export %s" % n
    from interactive_highway_simulation_b200 import outer_env
    assert C.sizeof(outer_env.CorridorPOD) == 24 + 3 * 16 * 16          # mce_corridor


def test_env_no_cpu_fallback_without_a_device(backend_lib):
    if backend_lib.lib().mcs_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from interactive_highway_simulation_b200 import outer_env
    c = {"id": 0, "left": [[0, 3.5], [100, 3.5]], "right": [[0, 0], [100, 0]], "center": [[0, 1.75], [100, 1.75]],
         "left_id": None, "right_id": None}
    with pytest.raises(outer_env.EnvError) as e:
        outer_env.DeviceMap([c])
    assert e.value.code == -2                                            # MCE_ERR_CUDA


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for n in ("mcs_scene_create", "mcs_rollouts", "mcs_rollouts_range", "mcs_reduce_features", "mcs_trajectories",
              "mcs_decide_mobil", "mcs_decide_fixed_direction", "mcs_decide_fixed_gap", "mcs_decide_closest_gap"):
        assert n in names


def test_library_exports_every_declared_symbol(backend_lib):
    L = C.CDLL(backend_lib.LIB_PATH)
    for n in declared_functions():
        assert hasattr(L, n), "libmcsim_b200.so does not export %s" % n
    assert set(declared_functions()) == set(backend_lib.EXPORTS)
    assert b"sm_100a" in backend_lib.lib().mcs_version()


def test_struct_layouts_match_header():
    # sizes the header promises (mcs_status / mcs_feature are 64-byte records)
    assert C.sizeof(abi.Status) == 64 and C.sizeof(abi.Feature) == 64
    assert C.sizeof(abi.Agent) == 240 and C.sizeof(abi.Corridor) == 112
    assert C.sizeof(abi.SimParam) == 16 and C.sizeof(abi.Candidate) == 16
    assert abi.Feature.fallBackRate.offset == 0 and abi.Feature.successRate.offset == 8      # BehaviorFeature order
    assert abi.Feature.fromNSims.offset == 0x38


def test_no_cpu_fallback_without_a_device(backend_lib):
    """Without a CUDA device every compute entry point must fail loudly (MCS_ERR_CUDA), never compute on the host."""
    if backend_lib.lib().mcs_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from scenes import probe_scene
    s = probe_scene()
    with pytest.raises(backend_lib.McsimError) as e:
        backend_lib.DeviceScene(s.c_corridors(), s.c_agents())
    assert e.value.code == abi.MCS_ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "interactive_highway_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text.replace("under oracle/", "") and "pyoracle" not in text and "mcsim_oracle" not in text, f
