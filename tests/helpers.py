"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

F16, BF16, F32 = 0, 1, 2


def f32_bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def bf16_to_f32(b):
    return (np.ascontiguousarray(b).astype(np.uint32) << 16).view(np.float32)


def bf16_from_f32(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 0x40).astype(np.uint16)
    return r


def case_input(codec, meta, name):
    """Returns (raw array as stored, fp32 widened values, dtype code)."""
    dtype = meta["codec"][name]["dtype"]
    raw = codec[name + ".in"]
    if dtype == F16:
        raw = raw.view(np.float16)
        return raw, raw.astype(np.float32), dtype
    if dtype == BF16:
        return raw, bf16_to_f32(raw), dtype
    return raw, raw.astype(np.float32), dtype


def narrow(y_f32, dtype):
    """fp32 -> boundary dtype bit patterns (uint16) with RN-even, or fp32 bits."""
    if dtype == F16:
        with np.errstate(over="ignore"):
            return np.ascontiguousarray(y_f32, np.float32).astype(np.float16).view(np.uint16)
    if dtype == BF16:
        return bf16_from_f32(y_f32)
    return f32_bits(y_f32)
