"""The C ABI: libb200zkp.so loads without a GPU, exports every symbol include/b200zkp.h declares, fails loudly
instead of falling back to the CPU, and agrees with the oracle on the (static) seal layout."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "b200zkp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(b200lib):
    from boundless_b200 import lib
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(b200lib, n), "missing export " + n
    # and the Python binding table covers exactly the header
    assert sorted(lib.SYMBOLS) == names


def test_no_cpu_fallback_without_gpu(b200lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from boundless_b200 import B200Error, ProverOpts, get_prover_server
    assert b200lib.b200_device_count() <= 0
    err = b200lib.b200_init(0)
    assert err is not None and b"no CUDA device" in err
    assert b200lib.b200_batch_intt(None, 10, 1, None) is not None
    with pytest.raises(B200Error):
        get_prover_server(ProverOpts(segment_po2=10))


@pytest.mark.parametrize("circ", [(9, 16, 32, 8, 0), (10, 16, 208, 32, 0), (12, 16, 208, 32, 0), (18, 16, 128, 16, 1), (20, 16, 208, 32, 0),
                                  (22, 4, 4, 4, 2)])
def test_seal_layout_matches_oracle(b200lib, oracle, circ):
    from boundless_b200 import Circuit
    c = Circuit(*circ)
    assert b200lib.b200_seal_words(C.byref(c)) == oracle.seal_words(*circ)


def test_circuit_validation(b200lib):
    from boundless_b200 import Circuit
    for bad in [(8, 16, 32, 8, 0), (25, 16, 32, 8, 0), (10, 0, 32, 8, 0), (10, 16, 30, 8, 0), (10, 16, 8, 16, 0), (10, 16, 500, 8, 0)]:
        assert b200lib.b200_seal_words(C.byref(Circuit(*bad))) == 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under boundless_b200/ may reference it."""
    hits = subprocess.run(["grep", "-rIl", "-E", r"(from|import)\s+oracle|liboracle|pyoracle|oracle\.h", os.path.join(ROOT, "boundless_b200")],
                          capture_output=True, text=True).stdout.split()
    hits = [h for h in hits if not h.endswith(".so") and "/_obj/" not in h]
    assert hits == [], hits


def test_constants_match_independent_generators(b200lib, oracle):
    """csrc/constants.inc (Python Grain generator) == oracle's C Grain generator, converted out of Montgomery form."""
    import numpy as np
    txt = open(os.path.join(ROOT, "boundless_b200", "csrc", "constants.inc")).read()
    def arr(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", txt, re.S).group(1)
        return np.array([int(x, 16) for x in re.findall(r"0x([0-9a-f]+)u", body)], dtype=np.uint32)
    L = oracle.lib()
    rc = np.ctypeslib.as_array(L.oracle_p2_rc_canon(), shape=(213,))
    dg = np.ctypeslib.as_array(L.oracle_p2_diag_canon(), shape=(24,))
    assert np.array_equal(oracle.from_mont(arr("B200_P2_RC_MONT")), rc)
    assert np.array_equal(oracle.from_mont(arr("B200_P2_DIAG_MONT")), dg)
    fwd = oracle.from_mont(arr("B200_ROU_FWD_MONT"))
    assert fwd[27] == 137 and fwd[1] == oracle.P - 1 and fwd[0] == 1
    for k in range(28):
        assert int(fwd[k]) == oracle.from_mont(np.array([L.oracle_rou_fwd(k)], dtype=np.uint32))[0]


def test_rust_sys_binding_declares_the_same_symbols():
    """bindings/rust/b200zkp-sys (source only; no cargo here) must list exactly the functions of include/b200zkp.h."""
    src = open(os.path.join(ROOT, "bindings", "rust", "b200zkp-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (b200_[a-z0-9_]+)\s*\(", src)))
    assert rust == header_functions()


def compat_header_functions():
    txt = open(os.path.join(ROOT, "include", "b200_risc0_sys_compat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:sppark|supra)_[A-Za-z0-9_]+)\s*\(", txt)))


def test_risc0_sys_compat_symbols_are_exported(b200lib):
    """include/b200_risc0_sys_compat.h: the original risc0-sys / sppark names, so the Rust side links unchanged."""
    from boundless_b200 import lib
    names = compat_header_functions()
    assert names == sorted(lib.COMPAT_SYMBOLS) and len(names) == 8
    for n in names:
        assert hasattr(b200lib, n), "missing export " + n


def test_risc0_sys_compat_fails_loudly_without_gpu(b200lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from boundless_b200 import lib
    for call in (lambda: b200lib.sppark_init(), lambda: b200lib.sppark_batch_iNTT(None, 10, 1),
                 lambda: b200lib.sppark_batch_expand(None, None, 10, 2, 1), lambda: b200lib.sppark_poseidon2_rows(None, None, 4, 4),
                 lambda: b200lib.supra_poly_divide(None, 4, None, None)):
        err = call()
        assert err.code != 0 and err.message
        with pytest.raises(lib.B200Error, match="no CUDA device"):
            lib.check_sppark(err)        # also frees the malloc()ed message, like sppark::Error's Drop
