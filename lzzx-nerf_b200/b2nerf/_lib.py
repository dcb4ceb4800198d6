"""ctypes binding of libb2nerf.so (the C ABI declared in include/b2nerf.h).

The header is the single source of truth: prototypes are parsed from it, so a symbol that is declared
but not exported (or the other way round) fails at import, not at first use.  There is NO fallback:
if the shared library is missing or a call fails, a RuntimeError is raised — the reference's
TORCH_CHECK behaviour (gridencoder.cu:425-441) — and nothing is ever routed to a CPU path.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
HEADER_PATHS = [os.path.join(_REPO, "include", "b2nerf.h"), os.path.join(_REPO, "include", "b2nerf_fused.h")]
LIB_PATH = os.path.join(_HERE, "libb2nerf.so")

_CTYPES = {
    "int": ctypes.c_int, "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float, "b2n_dtype": ctypes.c_int, "void": None,
}


def parse_header(path):
    """Return {name: (restype, [argtypes], [argnames])} for every `b2n_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"\b(const\s+char\s*\*\s*|int\s+|void\s+|uint64_t\s+)(b2n_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret.strip()]
        argtypes, argnames = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                nm = re.search(r"(\w+)\s*(\[\d*\])?$", a).group(1)
                if "*" in a or "[" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPES[base])
                argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


def declared_symbols():
    out = {}
    for h in HEADER_PATHS:
        if os.path.exists(h):
            out.update(parse_header(h))
    return out


class B2NError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise B2NError(
                f"libb2nerf.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C lzzx-nerf_b200/csrc`). There is no CPU fallback.")
        self._dll = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        self._dll.b2n_last_error.restype = ctypes.c_char_p
        self.protos = declared_symbols()
        for name, (restype, argtypes, _) in self.protos.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError as e:
                raise B2NError(f"libb2nerf.so does not export {name} declared in include/*.h") from e
            fn.restype = restype
            fn.argtypes = argtypes
        if self._dll.b2n_version() < 100:
            raise B2NError("libb2nerf.so is older than the headers")

    def raw(self, name):
        return getattr(self._dll, name)

    def call(self, name, *args):
        """Call an `int b2n_*` entry point; raise RuntimeError with the library's message on failure."""
        rc = getattr(self._dll, name)(*args)
        if rc != 0:
            msg = self._dll.b2n_last_error()
            raise B2NError(f"{name}: {msg.decode() if msg else 'error %d' % rc}")

    def launch_count(self):
        return int(self._dll.b2n_launch_count())


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
