"""ctypes binding of libfsb200.so — the only way Python reaches the kernels.

The prototypes are parsed from include/fsb200.h so the binding cannot drift from the header.
There is NO fallback: if the library is missing or a symbol is absent, import of the product path
fails with a RuntimeError.
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

from . import _build

HEADER = Path(__file__).resolve().parent.parent / "include" / "fsb200.h"

_CTYPES = {
    "int": ctypes.c_int,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
}

_DECL_RE = re.compile(r"^\s*(int|size_t|int64_t|uint64_t)\s+(fsb_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)


def parse_header(path: Path = HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every fsb_* prototype in the header."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for ret, name, args in _DECL_RE.findall(text):
        argtypes, argnames = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(a.split("*")[-1].strip())
                else:
                    toks = [t for t in a.split() if t != "const"]
                    argtypes.append(_CTYPES[toks[0]])
                    argnames.append(toks[-1])
        protos[name] = (_CTYPES[ret], argtypes, argnames)
    return protos


class FsbError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._cdll = None
        self._protos = None

    def load(self):
        if self._cdll is not None:
            return self._cdll
        path = _build.LIB_PATH
        if not path.exists() or _build.is_stale():
            # a library built from other sources than the ones on disk would be bound to prototypes parsed from the
            # current header: silent ABI drift.  Rebuild (no-op when the stamp matches) or refuse.
            if _build.find_nvcc() is None:
                raise FsbError(
                    f"{path} is missing or was built from different sources (stamp mismatch) and nvcc is not "
                    "available to build it; fusionsense_b200 has no CPU fallback. Run `python -m fusionsense_b200._build`."
                )
            _build.build()
        try:
            cdll = ctypes.CDLL(str(path))
        except OSError as e:  # pragma: no cover
            raise FsbError(f"cannot load {path}: {e}") from e
        protos = parse_header()
        for name, (ret, argtypes, _) in protos.items():
            try:
                fn = getattr(cdll, name)
            except AttributeError as e:
                raise FsbError(f"libfsb200.so does not export {name} (declared in include/fsb200.h)") from e
            fn.restype = ret
            fn.argtypes = argtypes
        self._cdll, self._protos = cdll, protos
        return cdll

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.load(), name)


lib = _Lib()


def check(code: int, what: str):
    """Raise on a non-zero status from an fsb_* call."""
    if code != 0:
        if code == 10001:
            raise FsbError(f"{what}: argument refused by libfsb200 (FSB_E_ARG)")
        raise FsbError(f"{what}: CUDA error {code}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else t.data_ptr()
