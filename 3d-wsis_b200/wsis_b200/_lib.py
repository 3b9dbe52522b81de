"""ctypes loader for the C-ABI library declared in include/wsis_b200.h.

The prototypes are parsed from the header itself, so the Python side cannot drift from the ABI.  There is
no CPU fallback: if libwsis_b200.so is missing the import of any op raises (build it with
`python __graft_entry__.py build` or `make -C 3d-wsis_b200/csrc`).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.environ.get("WSIS_B200_LIB") or os.path.join(_HERE, "libwsis_b200.so")  # override: experiment builds
HEADER_PATH = os.path.join(_ROOT, "include", "wsis_b200.h")

_SCALARS = {
    "int": ctypes.c_int,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "wsis_stream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [(ctype, param_name), ...])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char \*|int64_t|int)\s*(wsis_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "const char *": ctypes.c_char_p}[ret]
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                pname = re.findall(r"(\w+)(?:\[\d*\])?$", a)[0]
                if "*" in a or "[" in a:
                    params.append((ctypes.c_void_p, pname))
                else:
                    base = a.replace("const ", "").split()[0]
                    params.append((_SCALARS[base], pname))
        protos[name] = (restype, params)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "wsis_b200: %s not found -- the CUDA library is mandatory (no CPU fallback). "
                "Build it: python __graft_entry__.py build" % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, params) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = [t for t, _ in params]

    def last_error(self):
        msg = self.cdll.wsis_last_error()
        return msg.decode() if msg else ""

    def call(self, name, *args):
        """Calls a status-returning entry point; non-zero status -> RuntimeError (TV_ASSERT_RT_ERR analogue)."""
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise RuntimeError("%s: %s" % (name, self.last_error()))

    def value(self, name, *args):
        """Calls a value-returning helper (sizes, counts)."""
        return getattr(self.cdll, name)(*args)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
