"""CPU tests: the C-ABI library builds, loads and exports every declared symbol; host-side logic."""
import importlib
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_loads_and_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    header = open(os.path.join(ROOT, "include", "cstr_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cstr_[a-z0-9_]+)\s*\(", header))
    declared -= {"cstr_env_params", "cstr_actor_f32"}
    assert declared == set(pkg._lib.EXPORTED_SYMBOLS), declared ^ set(pkg._lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cstr_b200_abi_version() == pkg._lib.ABI_VERSION
    m = re.search(r"#define CSTR_B200_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == pkg._lib.ABI_VERSION


def test_no_cpu_fallback(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CstrLibraryError):
        pkg.GpuCSTRVecEnv(8)
    with pytest.raises(pkg.CstrLibraryError):
        pkg.GpuReplayBuffer(100)
    with pytest.raises(pkg.CstrLibraryError):
        pkg.TwoSeriesCSTREnv()


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "pytorch-rl-enhancedstablebaselines_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cstr_oracle" not in src.replace("oracle/cstr_oracle.c:cstr_expf_shared", "") and "build_oracle" not in src, f
                assert "refload" not in src and "/root/reference" not in src, f


def test_alias_module():
    import b200rl

    assert b200rl is importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")


def test_lazy_infos(pkg):
    env_mod = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200.env")
    t = np.zeros(6, bool)
    t[4] = True
    infos = env_mod.LazyInfos(6, {4: {"TimeLimit.truncated": True, "terminal_observation": np.ones(4), "episode": {"r": 1.0, "l": 400, "t": 0.1}}}, t)
    assert len(infos) == 6 and infos[-2]["TimeLimit.truncated"] is True and infos[0].get("TimeLimit.truncated", False) is False
    assert [i.get("TimeLimit.truncated", False) for i in infos] == [False] * 4 + [True, False]  # buffers.py:278
    assert [idx for idx, i in enumerate(infos) if i.get("episode")] == [4]  # base_class.py:476
    assert "terminal_observation" not in infos[1] and infos[1].get("is_success") is None
    with pytest.raises(IndexError):
        infos[6]
    from copy import deepcopy

    c = deepcopy(infos)
    assert c[4]["episode"]["l"] == 400 and c[4] is not infos[4]
    infos[0]["x"] = 1  # writes to the shared blank row are dropped
    assert "x" not in infos[1]
    rep = env_mod._Repeat(10**6, dict)
    assert len(rep) == 10**6 and rep[5] == {} and rep[-1] == {}


def test_host_pcg64_reset_matches_reference_fixture(pkg, golden):
    """reset_rng='pcg64' host draws (product code, not the oracle) == the reference's reset, bit for bit."""
    env_mod = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200.env")
    g = golden("reset.npz")
    for k, seed in enumerate(g["seeds"]):
        gen = env_mod._pcg64(int(seed))
        for e in range(g["random_obs"].shape[1]):
            raw = env_mod._initial_raw_state(gen, "random", None)
            assert np.array_equal(raw, g["random_raw"][k, e])
            assert np.array_equal(env_mod._normalize_f64(raw).astype(np.float32), g["random_obs"][k, e])
        gen = env_mod._pcg64(int(seed))
        base = np.array(env_mod.STATIC_INIT_STATE)
        for e in range(g["static_obs"].shape[1]):
            raw = env_mod._initial_raw_state(gen, "static", base)
            assert np.array_equal(env_mod._normalize_f64(raw).astype(np.float32), g["static_obs"][k, e])


def test_affine_helpers_match_oracle(pkg):
    import cstr_oracle as O

    env_mod = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200.env")
    x = np.random.default_rng(0).uniform(-1, 1, (100, 4)).astype(np.float32)
    assert np.array_equal(env_mod._denormalize_state(x), O.denormalize_state(x))
    assert np.array_equal(env_mod._normalize_state(env_mod._denormalize_state(x)), O.normalize_state(O.denormalize_state(x)))


def test_header_is_plain_c_and_struct_sizes_match_the_binding(pkg, tmp_path):
    """include/cstr_b200.h must be consumable from C (the ABI is `extern "C"`, plain pointers and sizes): compile it as strict C99
    and compare sizeof() of every struct with the ctypes mirror in _lib.py."""
    import ctypes
    import shutil
    import subprocess

    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {"cstr_env_params": pkg._lib.EnvParams, "cstr_actor_f32": pkg._lib.ActorF32, "cstr_episode_stats": pkg._lib.EpisodeStatsStruct,
               "cstr_norm_params": pkg._lib.NormParams, "cstr_td3_config": pkg._lib.Td3Config, "cstr_td3_state": pkg._lib.Td3State,
               "cstr_sac_config": pkg._lib.SacConfig}
    src = tmp_path / "abi.c"
    body = "".join(f'    printf("{name} %zu\\n", sizeof({name}));\n' for name in structs)
    src.write_text('#include <stdio.h>\n#include "cstr_b200.h"\nint main(void) {\n' + body + "    return 0;\n}\n")
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.splitlines())
    for name, cls in structs.items():
        assert int(sizes[name]) == ctypes.sizeof(cls), name
