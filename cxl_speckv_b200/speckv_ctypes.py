"""SpeckvLib -- same class, methods, argument meaning and error behaviour as the
reference's host/python/speckv_ctypes.py:7-98, bound to the B200 libcxlspeckv.so.

    SpeckvLib(path, dev_path="/dev/speckv0")   calls speckv_init in the constructor
    .alloc(bytes_needed, preferred_node=0) -> handle
    .free(handle) / .access(handle, offset, length) -> address
    .prefetch(req_id, layer, cur_pos, depth_k, tokens)
    .set_prefetch_depth(depth_k) / .set_compression_scheme(scheme)
Every failure raises RuntimeError("<function> failed: <status>") like the reference.
dev_path may also be "cuda" / "cuda:<n>" (see include/speckv.h).
"""
import ctypes
from ctypes import c_uint32, c_uint16, c_uint64, c_size_t, c_int32, c_void_p, c_char_p, c_int


class SpeckvLib:
    def __init__(self, path: str, dev_path: str = "/dev/speckv0"):
        self.lib = ctypes.CDLL(path)
        self.handle_t = c_uint64

        class AllocHint(ctypes.Structure):
            _fields_ = [("preferred_node", c_uint32), ("reserved", c_uint32)]

        self.AllocHint = AllocHint
        L = self.lib
        L.speckv_init.argtypes = [c_char_p]
        L.speckv_init.restype = c_int
        L.speckv_finalize.argtypes = []
        L.speckv_finalize.restype = None
        L.speckv_alloc.argtypes = [c_size_t, ctypes.POINTER(AllocHint), ctypes.POINTER(self.handle_t)]
        L.speckv_alloc.restype = c_int
        L.speckv_free.argtypes = [self.handle_t]
        L.speckv_free.restype = c_int
        L.speckv_access.argtypes = [self.handle_t, c_uint64, c_size_t, ctypes.POINTER(c_void_p)]
        L.speckv_access.restype = c_int
        L.speckv_prefetch.argtypes = [c_uint32, c_uint16, c_uint32, c_uint32, ctypes.POINTER(c_int32), c_uint32]
        L.speckv_prefetch.restype = c_int
        L.speckv_set_prefetch_depth.argtypes = [c_uint32]
        L.speckv_set_prefetch_depth.restype = c_int
        L.speckv_set_compression_scheme.argtypes = [c_int]
        L.speckv_set_compression_scheme.restype = c_int

        ret = L.speckv_init(dev_path.encode("ascii"))
        if ret != 0:
            raise RuntimeError(f"speckv_init failed: {ret}")

    def finalize(self):
        self.lib.speckv_finalize()

    def alloc(self, bytes_needed, preferred_node=0):
        hint = self.AllocHint(preferred_node, 0)
        handle = self.handle_t()
        ret = self.lib.speckv_alloc(bytes_needed, ctypes.byref(hint), ctypes.byref(handle))
        if ret != 0:
            raise RuntimeError(f"speckv_alloc failed: {ret}")
        return handle.value

    def free(self, handle):
        ret = self.lib.speckv_free(handle)
        if ret != 0:
            raise RuntimeError(f"speckv_free failed: {ret}")

    def access(self, handle, offset, length):
        gpu_ptr = c_void_p()
        ret = self.lib.speckv_access(handle, offset, length, ctypes.byref(gpu_ptr))
        if ret != 0:
            raise RuntimeError(f"speckv_access failed: {ret}")
        return gpu_ptr.value

    def prefetch(self, req_id, layer, cur_pos, depth_k, tokens):
        arr = (c_int32 * len(tokens))(*tokens)
        ret = self.lib.speckv_prefetch(req_id, layer, cur_pos, depth_k, arr, len(tokens))
        if ret != 0:
            raise RuntimeError(f"speckv_prefetch failed: {ret}")

    def set_prefetch_depth(self, depth_k):
        ret = self.lib.speckv_set_prefetch_depth(depth_k)
        if ret != 0:
            raise RuntimeError(f"speckv_set_prefetch_depth failed: {ret}")

    def set_compression_scheme(self, scheme):
        ret = self.lib.speckv_set_compression_scheme(scheme)
        if ret != 0:
            raise RuntimeError(f"speckv_set_compression_scheme failed: {ret}")
