"""SpeckvLib: the ctypes front end of the eight frozen speckv_* functions.

Drop-in for the class of the same name in the reference (host/python/speckv_ctypes.py:7-98): same constructor
(`SpeckvLib(path, dev_path="/dev/speckv0")`, which initialises the library), same methods and argument meaning
(`alloc`, `free`, `access`, `prefetch`, `set_prefetch_depth`, `set_compression_scheme`), the same attributes callers
reach into (`.lib`, `.handle_t`, `.AllocHint`) and the same failure convention: a non-zero status raises
RuntimeError("<function> failed: <status>").  The binding itself is table-driven (one signature table applied in a
loop) and the calls go through one checked-call helper.  Additions: `finalize()`, use as a context manager, and
`dev_path` may be "cuda" / "cuda:<n>" (see include/speckv.h).
"""
import ctypes as C

# name -> (restype, argtypes); pointer-to-struct arguments are filled in per instance below
_U64 = C.c_uint64
_SIGNATURES = {
    "speckv_init": (C.c_int, [C.c_char_p]),
    "speckv_finalize": (None, []),
    "speckv_alloc": (C.c_int, [C.c_size_t, "hint*", "handle*"]),
    "speckv_free": (C.c_int, [_U64]),
    "speckv_access": (C.c_int, [_U64, C.c_uint64, C.c_size_t, C.POINTER(C.c_void_p)]),
    "speckv_prefetch": (C.c_int, [C.c_uint32, C.c_uint16, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32]),
    "speckv_set_prefetch_depth": (C.c_int, [C.c_uint32]),
    "speckv_set_compression_scheme": (C.c_int, [C.c_int]),
}


class _AllocHint(C.Structure):          # speckv_alloc_hint_t
    _fields_ = [("preferred_node", C.c_uint32), ("reserved", C.c_uint32)]


class SpeckvLib:
    handle_t = _U64
    AllocHint = _AllocHint

    def __init__(self, path: str, dev_path: str = "/dev/speckv0"):
        self.lib = C.CDLL(path)
        placeholders = {"hint*": C.POINTER(_AllocHint), "handle*": C.POINTER(_U64)}
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(self.lib, name)
            fn.restype = restype
            fn.argtypes = [placeholders.get(a, a) for a in argtypes]
        self._call("speckv_init", dev_path.encode("ascii"))

    def _call(self, name: str, *args) -> None:
        status = getattr(self.lib, name)(*args)
        if status != 0:
            raise RuntimeError(f"{name} failed: {status}")

    # -- the reference's methods ------------------------------------------------------------------
    def alloc(self, bytes_needed, preferred_node=0):
        out = _U64()
        self._call("speckv_alloc", bytes_needed, C.byref(_AllocHint(preferred_node, 0)), C.byref(out))
        return out.value

    def free(self, handle):
        self._call("speckv_free", handle)

    def access(self, handle, offset, length):
        address = C.c_void_p()
        self._call("speckv_access", handle, offset, length, C.byref(address))
        return address.value

    def prefetch(self, req_id, layer, cur_pos, depth_k, tokens):
        window = (C.c_int32 * len(tokens))(*tokens)
        self._call("speckv_prefetch", req_id, layer, cur_pos, depth_k, window, len(tokens))

    def set_prefetch_depth(self, depth_k):
        self._call("speckv_set_prefetch_depth", depth_k)

    def set_compression_scheme(self, scheme):
        self._call("speckv_set_compression_scheme", scheme)

    # -- additions ------------------------------------------------------------------------------------
    def finalize(self):
        self.lib.speckv_finalize()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.finalize()
        return False
