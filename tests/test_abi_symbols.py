"""The C-ABI library loads (no GPU needed) and exports every symbol the header declares."""
import ctypes
import re

from roskfpos_b200 import lib as L


def declared_symbols():
    src = open(L.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kfpos_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = declared_symbols()
    assert len(names) >= 20
    l = ctypes.CDLL(L.SO_PATH)
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/kfpos_b200.h but not exported"
    assert set(names) == set(L.SIGNATURES), set(names) ^ set(L.SIGNATURES)


def test_abi_version_and_strerror(kflib):
    l = kflib.lib()
    assert l.kfpos_abi_version() == 1
    assert b"no CPU fallback" in l.kfpos_strerror(-2)


def test_no_cpu_fallback_without_gpu(kflib):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    from roskfpos_b200.batch import Batch
    try:
        Batch(kflib.MODEL_T6, 8, accel_noise=0.5)
    except kflib.KfposError as e:
        assert e.code == -2
    else:
        raise AssertionError("Batch() succeeded without a GPU")
    # the handle-free entry points as well
    import ctypes as C
    import numpy as np
    x = np.ones(4)
    assert kflib.lib().kfpos_selftest_math(0, 4, C.c_void_p(x.ctypes.data), None, None, None, None) == -2
    a = np.zeros((2, 1), dtype=np.uint8); r = np.zeros((2, 1), dtype=np.int32); t = np.zeros((2, 1))
    o = np.zeros((1, 4, 1), dtype=np.int32); dt = np.zeros((1, 1))
    rc = kflib.lib().kfpos_assemble_epochs(0, 1, 2, 4, C.c_void_p(a.ctypes.data), C.c_void_p(a.ctypes.data),
                                           C.c_void_p(r.ctypes.data), None, C.c_void_p(t.ctypes.data), 1, 0, 0.1,
                                           C.c_void_p(o.ctypes.data), None, C.c_void_p(dt.ctypes.data), None, None)
    assert rc == -2
    v = C.c_double(0.0)
    assert kflib.lib().kfpos_measure_fp64_peak(0, C.byref(v)) == -2


def test_config_struct_layout(kflib):
    """ctypes mirror and the C struct agree on size (checked through load_xml writes)."""
    from roskfpos_b200.batch import make_config
    cfg = make_config(xml=['<config><mag angleOffset="0.25" covarianceMag="0.0001"/></config>',
                           '<config><imu useFixedCovarianceAcceleration="1" covarianceAcceleration="0.003" '
                           'useFixedCovarianceAngularVelocityZ="1" covarianceAngularVelocityZ="0.089"/></config>'])
    assert cfg.mag_angle_offset == 0.25 and cfg.mag_cov == 0.0001
    assert cfg.imu_use_fixed_cov_acc == 1 and cfg.imu_cov_gyro_z == 0.089
    assert list(cfg.ml_start) == [1.0, 1.0, 4.0]


def test_header_is_plain_c(tmp_path):
    """include/kfpos_b200.h must be consumable from C (cgo / JNI / ctypes style bindings), not only C++;
    a C program links against the library with nothing but the header."""
    import os
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "kfpos_b200.h"\n#include <stdio.h>\n'
                   "int main(void) { kfpos_config c; kfpos_config_default(&c);\n"
                   '  printf("%d %s\\n", kfpos_abi_version(), kfpos_strerror(KFPOS_ERR_CUDA));\n'
                   "  return kfpos_abi_version() == KFPOS_ABI_VERSION ? 0 : 1; }\n")
    exe = tmp_path / "abi"
    inc = os.path.dirname(L.HEADER_PATH)
    libdir = os.path.dirname(L.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{inc}", str(src), "-o", str(exe),
                           f"-L{libdir}", "-lkfpos_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("1 "), (out.stdout, out.stderr)
