"""ctypes binding of libpd_b200.so.

The prototypes are parsed from include/pd_b200.h, so the header stays the single source of
truth for the C ABI (tests/test_abi.py checks every declared symbol is exported).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "pd_b200.h")
LIB_PATH = os.path.join(HERE, "libpd_b200.so")

_SCALARS = {
    "int": ctypes.c_int,
    "long": ctypes.c_long,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
}


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every function declared in the header."""
    with open(path) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"#[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pd_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret:
            continue
        if "char" in ret and "*" in ret:
            restype = ctypes.c_char_p
        elif ret == "void":
            restype = None
        else:
            restype = _SCALARS.get(ret.replace("const", "").strip(), ctypes.c_int)
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(a.split("*")[-1].strip())
                else:
                    toks = a.replace("const", "").split()
                    argtypes.append(_SCALARS[toks[0]])
                    argnames.append(toks[-1])
        protos[name] = (restype, argtypes, argnames)
    return protos


_lib = None
_protos = None


def load(build_if_missing=True):
    """Load (building in-tree if needed) the native library. Raises if that is impossible:
    there is no Python / CPU fallback for the product path."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build

        try:
            _build.build()
        except Exception as e:  # nvcc missing on the box: fall through to the prebuilt file
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"libpd_b200.so is missing and could not be built: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libpd_b200.so not found; run `python -m pydreamer_b200.build`")
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argtypes, _) in _protos.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def prototypes():
    if _protos is None:
        return parse_header()
    return _protos
