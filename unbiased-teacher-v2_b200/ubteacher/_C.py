"""ctypes binding of the C-ABI kernel library ``libut2_sm100.so`` (declared in ``include/ut2.h``).

There is deliberately no CPU fallback: if the shared library is missing, or a call returns a
non-zero status, a RuntimeError is raised (the product path must fail loudly).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UT2_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libut2_sm100.so")

_lib = None


class f32(float):
    """Marks a Python number that must cross the ABI as a C ``float``."""


class f64(float):
    """Marks a Python number that must cross the ABI as a C ``double``."""


class i64(int):
    """Marks a Python int that must cross the ABI as ``long long``."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " (there is no CPU fallback for the UT2 hot path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ut2_last_error_string.restype = ctypes.c_char_p
    return _lib


def _conv(a):
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, (ctypes.Array, ctypes._SimpleCData)):
        return a
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, f32):
        return ctypes.c_float(float(a))
    if isinstance(a, f64):
        return ctypes.c_double(float(a))
    if isinstance(a, i64):
        return ctypes.c_longlong(int(a))
    if isinstance(a, int):
        return ctypes.c_int(a)
    if isinstance(a, float):
        return ctypes.c_float(a)
    raise TypeError(f"cannot pass {type(a)} across the C ABI")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Invoke ``int name(args..., void* stream)`` on torch's current CUDA stream."""
    fn = getattr(lib(), name)
    rc = fn(*[_conv(a) for a in args], stream_ptr())
    if rc != 0:
        msg = lib().ut2_last_error_string().decode()
        raise RuntimeError(f"{name} failed with status {rc}: {msg}")


launch_count = 0
# set to {} by tools/bench to bracket EVERY C-ABI call with CUDA events: {tag: [(start, end), ...]}
EVENT_PROFILE = None


def counted_call(name, *args, tag=None):
    global launch_count
    launch_count += 1
    if EVENT_PROFILE is None:
        call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    EVENT_PROFILE.setdefault(tag or name, []).append((e0, e1))
