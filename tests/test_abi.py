"""CPU-side checks of the drop-in boundary: libdqlb200.so loads, exports every symbol include/dqlb200.h declares,
struct layouts agree with the ctypes mirrors, and the product refuses to run without its CUDA pieces (no fallback)."""
import ctypes as C
import pathlib
import re

import pytest

from dql_multirotor_landing_b200 import _ffi
from dql_multirotor_landing_b200 import constants as K

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "dqlb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dqlb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.load()
    syms = declared_symbols()
    assert len(syms) >= 17
    assert sorted(_ffi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_struct_layouts_match_header():
    lib = _ffi.load()
    assert lib.dqlb200_abi_version() == K.ABI_VERSION
    assert lib.dqlb200_config_bytes() == C.sizeof(K.Config)
    assert lib.dqlb200_population_state_bytes() == C.sizeof(K.PopulationState) == 320
    assert lib.dqlb200_eval2d_params_bytes() == C.sizeof(K.Eval2DParams) == 72
    assert lib.dqlb200_termination_string(2) == b"SUCCESS: Goal state reached"
    assert lib.dqlb200_termination_string(0) is None
    for code, text in K.TERMINATION_STRINGS.items():
        assert lib.dqlb200_termination_string(code).decode() == text


def test_argument_validation_without_gpu():
    lib = _ffi.load()
    cfg = K.build_config(1, 1)
    cfg.struct_bytes = 12
    h = C.c_void_p()
    lut = K.alpha_lut()
    pp = (K.PopulationParams * 1)()
    rc = lib.dqlb200_create(C.byref(cfg), lut.ctypes.data_as(C.POINTER(C.c_float)), pp, 0, C.byref(h))
    assert rc == -1 and b"mismatch" in lib.dqlb200_last_error()
    assert lib.dqlb200_create(None, None, None, 0, C.byref(h)) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dql_multirotor_landing_b200.engine import Engine
    from dql_multirotor_landing_b200.mdp import TrainingMdp
    with pytest.raises(_ffi.Dqlb200Error):
        Engine(1, 8)
    with pytest.raises(_ffi.Dqlb200Error):
        TrainingMdp(0, 22.92, 20, 4.5)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    for p in (ROOT / "dql_multirotor_landing_b200").rglob("*.py"):
        assert "oracle" not in re.sub(r"#.*", "", p.read_text()), p


def test_npy_layout_of_committed_assets():
    import numpy as np
    for name in ("Q_table_a.npy", "Q_table_b.npy", "state_action_count.npy"):
        a = np.load(ROOT / "assets" / name)
        assert a.shape == (5, 3, 3, 3, 7, 3) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]


def test_compiled_in_defaults_match_constants_py():
    """The production kernel instances carry the reference-default constants as literals (csrc/dqlb200_device.cuh: KDef).  They
    must equal what constants.py derives for the default parameters bit for bit -- otherwise every default run would silently
    fall back to the slower generic instance.  Host-only check, no GPU needed."""
    import ctypes as C
    lib = _ffi.load()
    assert lib.dqlb200_config_is_default(C.byref(K.build_config(4, 100, 128))) == 1
    assert lib.dqlb200_config_is_default(C.byref(K.build_config(4, 100, 128, dp=K.DynamicsParameters(c_d=0.21)))) == 0
    assert lib.dqlb200_config_is_default(C.byref(K.build_config(4, 100, 128, mp=K.MdpParameters(p_max=4.0)))) == 0
    assert lib.dqlb200_config_is_default(C.byref(K.build_config(4, 100, 128, dp=K.DynamicsParameters(noise_vel_sd=0.1)))) == 0
    assert lib.dqlb200_config_is_default(C.byref(K.build_config(4, 100, 128, dp=K.DynamicsParameters(accel_mode="kalman", n_sub=4)))) == 0
    assert lib.dqlb200_config_is_default(C.byref(K.build_config(4, 100, 128, dp=K.DynamicsParameters(v_mp=0.8, r_mp=3.0)))) == 1
