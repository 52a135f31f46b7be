"""ctypes binding of libpacingpseudo_b200.so (the C ABI declared in include/pacingpseudo_b200.h).

The signatures are parsed from the public header, so the binding cannot drift from the declared
ABI. There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised — the product path never routes through PyTorch ops or the CPU oracle.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER_PATH = os.path.join(_HERE, "..", "include", "pacingpseudo_b200.h")
LIB_PATH = os.path.join(_HERE, "libpacingpseudo_b200.so")

F32, BF16 = 0, 1
CR_VARIANTS = {None: 0, "none": 0, "ce_loss": 1, "l1_loss": 2, "l2_loss": 3, "kl_loss": 4}

_PROTO_RE = re.compile(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(pp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", re.M | re.S)


def _ctype_of(decl, is_return=False):
    d = " ".join(decl.replace("*", " * ").split())
    if "*" in d:
        if is_return and d.startswith("const char"):
            return ctypes.c_char_p
        return ctypes.c_void_p
    base = d.replace("const ", "").strip()
    # strip the parameter name, if any
    toks = base.split()
    if toks[:2] == ["long", "long"]:
        return ctypes.c_longlong
    t = toks[0]
    if t == "void":
        return None
    return {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
            "pp_unet_t": ctypes.c_void_p}[t]


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function prototype in the public header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in _PROTO_RE.finditer(text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = args.strip()
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        protos[name] = (_ctype_of(ret, is_return=True), argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "pacingpseudo_b200: %s not found. Build it with `python -m pacingpseudo_b200.build` "
                "(nvcc, sm_100a). There is no fallback path." % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._init_devices = set()
        self._lock = threading.Lock()

    def last_error(self):
        msg = self.cdll.pp_last_error()
        return msg.decode() if msg else ""

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, rc, self.last_error()))

    def ensure_init(self, device_index):
        if device_index in self._init_devices:
            return
        with self._lock:
            if device_index not in self._init_devices:
                self.call("pp_init", int(device_index))
                self._init_devices.add(device_index)


_lib = None
_lib_lock = threading.Lock()


def get_lib():
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                _lib = _Lib()
    return _lib


def ptr(t):
    """Device pointer of a (contiguous) torch tensor, or NULL for None."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def current_stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("pacingpseudo_b200: %s must be a CUDA tensor (got %s); there is no CPU path" % (what, t.device))
    get_lib().ensure_init(t.device.index if t.device.index is not None else 0)
