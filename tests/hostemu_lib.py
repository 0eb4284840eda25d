"""ctypes wrapper over tests/hostemu/libhostemu.so — TEST INFRASTRUCTURE ONLY (CPU build of the K1/K2 unit logic)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(os.path.join(HERE, "libhostemu.so"))
        vp = C.c_void_p
        L.emu_huffman.argtypes = [vp, C.c_ulonglong, vp, C.c_longlong, vp, vp, vp]
        L.emu_huffman.restype = None
        L.emu_huffman_staged.argtypes = [vp, C.c_ulonglong, vp, C.c_longlong, C.c_int, C.c_int, vp, vp, vp]
        L.emu_huffman_staged.restype = None
        L.emu_requant.argtypes = [vp, C.c_longlong, vp, vp, vp, vp]
        L.emu_requant.restype = None
        L.emu_table.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.emu_table.restype = C.POINTER(C.c_float)
        L.emu_powtab34.argtypes = [C.POINTER(C.c_int)]
        L.emu_powtab34.restype = C.POINTER(C.c_double)
        L.emu_huff_one.argtypes = [C.c_int, vp, C.c_int, C.POINTER(C.c_int * 4)]
        _lib = L
    return _lib


def huffman(main_data: np.ndarray, units: np.ndarray):
    n = len(units)
    is16 = np.zeros((n, 576), np.int16)
    meta = np.zeros(n, np.uint32)
    sf = np.zeros((n, 64), np.uint8)
    main_data = np.ascontiguousarray(main_data)
    units = np.ascontiguousarray(units)
    lib().emu_huffman(main_data.ctypes.data, (len(main_data) - 64) * 8, units.ctypes.data, n, is16.ctypes.data, meta.ctypes.data, sf.ctypes.data)
    return is16, meta, sf


def huffman_staged(main_data: np.ndarray, units: np.ndarray, tile: int = 256, cap16: int = 1 << 20):
    """K1 through the staged cursor, tiled like k_huffman (tile units per stretch, at most cap16 16-byte chunks staged)."""
    n = len(units)
    is16 = np.zeros((n, 576), np.int16)
    meta = np.zeros(n, np.uint32)
    sf = np.zeros((n, 64), np.uint8)
    main_data = np.ascontiguousarray(main_data)
    units = np.ascontiguousarray(units)
    lib().emu_huffman_staged(main_data.ctypes.data, (len(main_data) - 64) * 8, units.ctypes.data, n, tile, cap16, is16.ctypes.data,
                             meta.ctypes.data, sf.ctypes.data)
    return is16, meta, sf


def requant(units: np.ndarray, is16, meta, sf):
    ng = len(units) // 2
    xr = np.zeros((ng, 2, 576), np.float32)
    units = np.ascontiguousarray(units)
    lib().emu_requant(units.ctypes.data, ng, is16.ctypes.data, meta.ctypes.data, sf.ctypes.data, xr.ctypes.data)
    return xr


def table(which: int) -> np.ndarray:
    n = C.c_int()
    p = lib().emu_table(which, C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def powtab34() -> np.ndarray:
    n = C.c_int()
    p = lib().emu_powtab34(C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()
