"""ctypes binding of tests/selftest/libdfp_selftest.so: the tcgen05 building-block self-test and the shared-memory
descriptor probes.  Test infrastructure -- the product library exports none of this."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfp_selftest.so")
_lib = None

SYMBOLS = {
    "dfpt_last_error": (C.c_char_p, []),
    "dfpt_selftest_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "dfpt_probe_tc": (C.c_int, [C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p]),
    "dfpt_probe_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p]),
    "dfpt_probe_bf16_k": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
    "dfpt_probe_bf16_kt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built (python -m attentive_dfprior_b200.build)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libdfp_selftest error {rc}: {lib().dfpt_last_error().decode()}")
