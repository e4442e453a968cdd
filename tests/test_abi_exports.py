"""The C-ABI library builds, loads without a GPU and exports every symbol include/le_b200.h declares."""
import ctypes as C
import os
import re

from learning_environments_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "le_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(le_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_all_declared_symbols():
    from learning_environments_b200.csrc import build as le_build
    le_build.build()
    lib = _abi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in include/le_b200.h is not exported" % name
    assert sorted(_abi.EXPORTED_SYMBOLS) == declared
    assert lib.le_version() == 100
    assert lib.le_sizeof_lane_cfg() == C.sizeof(_abi.LaneCfg)


def test_bad_arguments_report_errors_without_gpu():
    lib = _abi.load_library()
    cfg = _abi.LaneCfg()
    cfg.sd, cfg.ad, cfg.real_env, cfg.env_kind = 5, 2, 0, 0      # no kernel set for state_dim 5
    rc = lib.le_inner_loop_workspace_bytes(C.byref(cfg), C.c_int(4), C.c_int(1))
    assert rc < 0 and b"real_env" in lib.le_last_error()
    cfg.sd, cfg.env_kind, cfg.rn_type = 4, 1, 3                   # info-vector reward type
    rc = lib.le_rn_reward(C.byref(cfg), None, C.c_int(1), C.c_int(1), None, None, None, None, None)
    assert rc == -3 and b"info vector" in lib.le_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "learning_environments_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/philox.py", "").replace("oracle/le_oracle.c", ""), \
                    "%s mentions the oracle" % os.path.join(dirpath, f)
