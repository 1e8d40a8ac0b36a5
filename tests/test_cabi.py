"""CPU-only checks of the C-ABI boundary: the library loads without a GPU and exports every
symbol include/gpbt.h declares; argument validation errors come back as codes, not crashes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from gpbt_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gpbt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gpbt_[a-z_]+)\s*\(", hdr))
    assert declared == set(lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib.lib, name), name
    assert lib.lib.gpbt_version() >= 100


def test_bad_arguments_return_codes(lib):
    h = C.c_void_p()
    rc = lib.lib.gpbt_emulator_create(C.byref(h), 0, 0, 0, 0, 0, 0, *([None] * 10))
    assert rc == -1 and b"gpbt_emulator_create" in lib.lib.gpbt_last_error()
    assert lib.lib.gpbt_log_posterior(None, None, 0.0, None, None, 1, 0, None) == -1
    assert lib.lib.gpbt_mvn_loglike(None, None, None, None, None, None, 0.0, 1, 3, None) == -1
    with pytest.raises(lib.GpbtError):
        lib.check(rc)


def test_parameter_file_and_pickles(tmp_path):
    import gpbt_b200
    from gpbt_b200 import synthetic
    from gpbt_b200.emulator import read_training_pickle
    from gpbt_b200.mcmc import read_experiment_pickle
    paths = synthetic.write_fixture(str(tmp_path), p=3, n=12, m=4)
    par = gpbt_b200.parse_model_parameter_file(paths["par"])
    assert list(par) == ["par0", "par1", "par2"] and par["par1"][1:3] == [0.0, 1.25]
    design, data, err, dropped = read_training_pickle(paths["train"])
    assert design.shape == (12, 3) and data.shape == (12, 4) and dropped == 0
    assert np.allclose(err, 0.01)
    y, cov = read_experiment_pickle(paths["exp"])
    assert y.shape == (1, 4) and np.allclose(np.diag(cov), (0.03 * np.abs(y[0])) ** 2)
    logd = read_training_pickle(paths["train"], log_trafo=True)[1]
    assert np.allclose(logd, np.log(np.abs(data) + 1e-30))


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tests import goldens
    from tests.helpers import product_states
    from gpbt_b200.device import DeviceChain
    g = goldens.load("odd_shape")
    states, _ = product_states(g)
    ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
    with pytest.raises(RuntimeError):
        ch.log_target(g["X"], -np.inf)


def test_header_is_plain_c(tmp_path):
    """include/gpbt.h is a C header (no C++-isms): a C translation unit that includes it compiles."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "use_gpbt.c"
    src.write_text('#include "gpbt.h"\nint main(void) { gpbt_chain_t c = 0; gpbt_emulator_t e = 0; (void)c; (void)e; return GPBT_PATH_AUTO; }\n')
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
