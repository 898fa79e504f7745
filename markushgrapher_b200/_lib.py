"""ctypes loader for libmg_b200.so (the C-ABI product library).

There is no fallback: if the shared library is missing or a call fails the caller gets an exception.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MG_B200_LIB", os.path.join(_HERE, "lib", "libmg_b200.so"))  # override: instrumented builds

_lib = None


class MgError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MgError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(markushgrapher_b200 has no CPU/PyTorch fallback)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mg_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mg_last_error().decode("utf-8", "replace")
        raise MgError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """device/host pointer of a torch tensor (or None) as c_void_p"""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_contiguous(), "tensor must be contiguous"
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
