"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of go-mp3's decode path (oracle/mp3_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

ORC_OK, ORC_EOF = 0, 1


class Taps(C.Structure):
    _fields_ = [("capacity_frames", C.c_int), ("n_frames", C.c_int), ("header", C.POINTER(C.c_uint32)),
                ("main_data_begin", C.POINTER(C.c_int32)), ("position", C.POINTER(C.c_int64)),
                ("is_", C.POINTER(C.c_int16)), ("count1", C.POINTER(C.c_int32)), ("scalefac_l", C.POINTER(C.c_uint8)),
                ("scalefac_s", C.POINTER(C.c_uint8)), ("part2_start", C.POINTER(C.c_int32)),
                ("xr_requant", C.POINTER(C.c_float)), ("xr_reorder", C.POINTER(C.c_float)),
                ("xr_stereo", C.POINTER(C.c_float)), ("xr_alias", C.POINTER(C.c_float)), ("hybrid", C.POINTER(C.c_float))]


class OrcBits(C.Structure):
    _fields_ = [("vec", C.c_void_p), ("len", C.c_int), ("bit_pos", C.c_int), ("byte_pos", C.c_int), ("err", C.c_int)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    vp, i64 = C.c_void_p, C.c_int64
    L.orc_new_decoder.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
    L.orc_new_decoder.restype = vp
    L.orc_free_decoder.argtypes = [vp]
    L.orc_free_decoder.restype = None
    L.orc_set_taps.argtypes = [vp, C.POINTER(Taps)]
    L.orc_set_taps.restype = None
    L.orc_read.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_int)]
    L.orc_read.restype = C.c_long
    L.orc_read_all.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int)]
    L.orc_read_all.restype = C.c_long
    L.orc_free.argtypes = [vp]
    L.orc_free.restype = None
    L.orc_seek.argtypes = [vp, i64, C.c_int, C.POINTER(C.c_int)]
    L.orc_seek.restype = i64
    for n in ("length", "bytes_per_frame", "duration_ns", "position_ns", "remaining_ns", "sample_position", "sample_count"):
        f = getattr(L, "orc_" + n)
        f.argtypes = [vp]
        f.restype = i64
    L.orc_sample_rate.argtypes = [vp]
    L.orc_sample_rate.restype = C.c_int
    L.orc_progress.argtypes = [vp]
    L.orc_progress.restype = C.c_double
    for n in ("seek_to_sample", "skip", "seek_to_time"):
        f = getattr(L, "orc_" + n)
        f.argtypes = [vp, i64]
        f.restype = C.c_int
    L.orc_num_frame_starts.argtypes = [vp]
    L.orc_num_frame_starts.restype = C.c_int
    L.orc_frame_start.argtypes = [vp, C.c_int]
    L.orc_frame_start.restype = i64
    L.orc_error_string.argtypes = [C.c_int]
    L.orc_error_string.restype = C.c_char_p
    L.orc_decode_streams_mt.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int,
                                        C.POINTER(i64), C.POINTER(C.c_uint64)]
    L.orc_decode_streams_mt.restype = C.c_double
    L.orc_bits_init.argtypes = [C.POINTER(OrcBits), vp, C.c_int]
    L.orc_bits_init.restype = None
    L.orc_bits_bit.argtypes = [C.POINTER(OrcBits)]
    L.orc_bits_bits.argtypes = [C.POINTER(OrcBits), C.c_int]
    L.orc_bits_pos.argtypes = [C.POINTER(OrcBits)]
    L.orc_bits_set_pos.argtypes = [C.POINTER(OrcBits), C.c_int]
    L.orc_bits_set_pos.restype = None
    L.orc_huffman_decode.argtypes = [C.POINTER(OrcBits), C.c_int, C.POINTER(C.c_int * 4)]
    L.orc_huffman_table_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_huffman_table_code.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_uint32)]
    for n in ("is_valid", "bitrate", "sampling_frequency_value", "side_info_size", "bytes_per_frame", "samples_per_frame",
              "bytes_per_second"):
        f = getattr(L, "orc_header_" + n)
        f.argtypes = [C.c_uint32]
        f.restype = C.c_int
    L.orc_header_frame_size.argtypes = [C.c_uint32, C.POINTER(C.c_int)]
    L.orc_header_frame_duration_ns.argtypes = [C.c_uint32]
    L.orc_header_frame_duration_ns.restype = i64
    L.orc_frameheader_read_mem.argtypes = [vp, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(i64),
                                           C.POINTER(C.c_size_t), C.POINTER(i64)]
    for n, t in (("imdct_win", C.c_float), ("cos_n12", C.c_float), ("cos_n36", C.c_float), ("synth_nwin", C.c_float),
                 ("synth_dtbl", C.c_float), ("powtab34", C.c_double)):
        f = getattr(L, "orc_table_" + n)
        f.argtypes = []
        f.restype = C.POINTER(t)
    _lib = L
    return L


class TapArrays:
    """numpy-backed orc_taps for up to `cap` frames. Arrays are indexed [frame][gr][ch][...]."""

    def __init__(self, cap: int, stages: bool = True):
        self.cap = cap
        self.header = np.zeros(cap, np.uint32)
        self.main_data_begin = np.zeros(cap, np.int32)
        self.position = np.zeros(cap, np.int64)
        self.is_ = np.zeros((cap, 2, 2, 576), np.int16)
        self.count1 = np.zeros((cap, 2, 2), np.int32)
        self.scalefac_l = np.zeros((cap, 2, 2, 22), np.uint8)
        self.scalefac_s = np.zeros((cap, 2, 2, 39), np.uint8)
        self.part2_start = np.zeros((cap, 2, 2), np.int32)
        self.stage_names = ("xr_requant", "xr_reorder", "xr_stereo", "xr_alias", "hybrid")
        for n in self.stage_names:
            setattr(self, n, np.zeros((cap, 2, 2, 576), np.float32) if stages else None)
        t = Taps()
        t.capacity_frames = cap
        t.n_frames = 0

        def ptr(a, ct):
            return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None

        t.header = ptr(self.header, C.c_uint32)
        t.main_data_begin = ptr(self.main_data_begin, C.c_int32)
        t.position = ptr(self.position, C.c_int64)
        t.is_ = ptr(self.is_, C.c_int16)
        t.count1 = ptr(self.count1, C.c_int32)
        t.scalefac_l = ptr(self.scalefac_l, C.c_uint8)
        t.scalefac_s = ptr(self.scalefac_s, C.c_uint8)
        t.part2_start = ptr(self.part2_start, C.c_int32)
        for n in self.stage_names:
            a = getattr(self, n)
            if a is not None:
                setattr(t, n, ptr(a, C.c_float))
        self.c = t

    @property
    def n_frames(self) -> int:
        return self.c.n_frames


class OracleDecoder:
    def __init__(self, data: bytes, seekable: bool = True, taps: TapArrays | None = None):
        self.L = lib()
        self._data = bytes(data)
        err = C.c_int(0)
        self.taps = taps
        self.h = self.L.orc_new_decoder(self._data, len(self._data), 1 if seekable else 0, C.byref(err))
        self.open_err = err.value
        if self.h and taps is not None:
            self.L.orc_set_taps(self.h, C.byref(taps.c))

    def ok(self) -> bool:
        return bool(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.orc_free_decoder(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def read(self, n: int):
        buf = C.create_string_buffer(n)
        err = C.c_int(0)
        got = self.L.orc_read(self.h, buf, n, C.byref(err))
        return buf.raw[:got], err.value

    def read_all(self):
        out = C.POINTER(C.c_uint8)()
        err = C.c_int(0)
        n = self.L.orc_read_all(self.h, C.byref(out), C.byref(err))
        data = C.string_at(out, n) if n > 0 else b""
        self.L.orc_free(out)
        return data, err.value

    def seek(self, off: int, whence: int = 0):
        err = C.c_int(0)
        r = self.L.orc_seek(self.h, off, whence, C.byref(err))
        return r, err.value

    def __getattr__(self, name):
        # length, sample_rate, duration_ns, ... -> orc_<name>(h)
        f = getattr(lib(), "orc_" + name)
        return lambda *a: f(self.h, *a)


def decode_with_taps(data: bytes, cap_frames: int, stages: bool = True):
    """NewDecoder + io.ReadAll with taps on every frame (including frame 0, which NewDecoder decodes).

    The oracle attaches taps after open, so frame 0 is re-decoded by seeking to 0 first: Seek(0) from a fresh
    decoder re-reads frame 0 from zero state (decode.go:135-142), which is exactly the open state.
    """
    taps = TapArrays(cap_frames, stages)
    d = OracleDecoder(data, True, taps)
    if not d.ok():
        return None, None, d.open_err, taps
    d.seek(0, 0)
    pcm, err = d.read_all()
    return d, pcm, err, taps


def decode_streams_mt(streams, threads: int):
    L = lib()
    n = len(streams)
    arr = (C.c_char_p * n)(*streams)
    lens = (C.c_size_t * n)(*[len(s) for s in streams])
    pcm_bytes = (C.c_int64 * n)()
    cks = (C.c_uint64 * n)()
    secs = L.orc_decode_streams_mt(arr, lens, n, threads, pcm_bytes, cks)
    return secs, list(pcm_bytes), list(cks)
