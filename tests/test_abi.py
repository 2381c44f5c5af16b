"""The C-ABI boundary without a GPU: libmfc5.so loads, exports every symbol include/mfc5.h declares,
the ctypes mirror of mfc5_params matches the C struct, and the entry points fail loudly (no CPU
fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from mfc_b200 import lib
from mfc_b200.params import Mfc5Params, UnsupportedCase, params_from_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mfc5.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfc5_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_match_binding_list():
    assert _declared() == sorted(lib.EXPORTED_SYMBOLS)


@pytest.mark.parametrize("name", ["libmfc5.so", "libmfc5_strict.so"])
def test_library_exports_every_declared_symbol(name):
    path = os.path.join(ROOT, "mfc_b200", name)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-j", "8", "-C", os.path.join(ROOT, "mfc_b200", "csrc")])
    L = C.CDLL(path)
    for sym in _declared():
        assert hasattr(L, sym), f"{name} does not export {sym}"
    L.mfc5_version.restype = C.c_char_p
    assert b"sm_100a" in L.mfc5_version()


def test_params_struct_matches_c_layout(tmp_path):
    """sizeof / offsetof of the C struct, compiled with gcc, against the ctypes mirror."""
    fields = [f[0] for f in Mfc5Params._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "mfc5_params.h"\nint main(){printf("%zu\\n", sizeof(mfc5_params));\n'
    for f in fields:
        prog += f'printf("%zu\\n", offsetof(mfc5_params, {f}));\n'
    prog += "return 0;}\n"
    src = tmp_path / "t.c"
    src.write_text(prog)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(Mfc5Params)
    for f, off in zip(fields, out[1:]):
        assert getattr(Mfc5Params, f).offset == off, f


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from golden_util import setup_case
    case, P, cbs, cons, dbuf = setup_case("2AB32975")
    with pytest.raises(lib.Mfc5Error, match="no CUDA device"):
        lib.Mfc5(P, cbs)


def test_out_of_envelope_cases_are_rejected():
    from golden_util import CASES
    base = dict(CASES["2AB32975"]["params"])
    for key, val in (("viscous", "T"), ("bubbles_euler", "T"), ("model_eqns", 3), ("riemann_solver", 4), ("weno_order", 9),
                     ("igr", "T"), ("alt_soundspeed", "T"), ("low_Mach", 3), ("cont_damage", "T")):
        with pytest.raises(UnsupportedCase):  # (alt_soundspeed: the base case has one fluid)
            params_from_case(dict(base, **{key: val}))
    two = dict(CASES["A0B82851"]["params"])
    assert params_from_case(dict(two, alt_soundspeed="T")).alt_soundspeed == 1
    assert params_from_case(dict(base, low_Mach=1)).low_Mach == 1
    assert params_from_case(dict(base, mp_weno="T")).mp_weno == 1
    with pytest.raises(UnsupportedCase):  # case_validator.py:967
        params_from_case(dict(base, mp_weno="T", weno_order=3))
    with pytest.raises(UnsupportedCase):  # HLL with interface alphas + alt_soundspeed: no nc_iface_vel on this path
        params_from_case(dict(two, alt_soundspeed="T", riemann_solver=1))
    with pytest.raises(UnsupportedCase):  # low_Mach = 2 is HLLC only
        params_from_case(dict(base, low_Mach=2, riemann_solver=1))
    with pytest.raises(UnsupportedCase):  # m_checker.fpp:52-69
        params_from_case(dict(base, m=10))


def test_c_host_program_compiles_and_links_against_the_header(tmp_path):
    """tests/c_driver/abi_driver.c (the shim's call sequence in plain C) builds with gcc -Wall -Werror against include/mfc5.h
    and links with libmfc5.so; its compute path runs in tests/test_gpu_abi.py."""
    exe = str(tmp_path / "abi_driver")
    libdir = os.path.join(ROOT, "mfc_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_driver", "abi_driver.c"), "-L", libdir, "-lmfc5", f"-Wl,-rpath,{libdir}", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
