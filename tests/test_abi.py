"""CPU-side checks of the C ABI: the library builds/loads here (nvcc cross-compiles, no GPU needed)
and exports every symbol include/pd_b200.h declares.  No compute calls."""
import ctypes
import subprocess

from pydreamer_b200 import _native


def test_header_parses_all_entry_points():
    protos = _native.parse_header()
    assert len(protos) >= 39
    for must in ("pd_create", "pd_gemm", "pd_ln_elu_fwd", "pd_gru_bwd", "pd_cat_sample", "pd_kl", "pd_im2col",
                 "pd_col2im_imgloss", "pd_gae_critic", "pd_adamw"):
        assert must in protos
    assert len(protos["pd_gemm"][1]) == 21


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in _native.parse_header():
        assert name in exported, f"{name} declared in pd_b200.h but not exported"
        assert hasattr(lib, name)


def test_version_and_error_paths_without_gpu():
    lib = _native.load()
    assert b"sm_100a" in lib.pd_version()
    h = ctypes.c_void_p()
    rc = lib.pd_create(0, ctypes.byref(h))
    # no GPU in the authoring container: must fail cleanly, never fall back
    if rc != 0:
        assert not h.value
    else:
        lib.pd_destroy(h)


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass   # tcgen05.mma
    assert "UTMALDG" in sass                         # TMA tensor load
    assert "LDTM" in sass                            # tcgen05.ld
