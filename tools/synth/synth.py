"""ctypes wrapper + workload presets for the stream synthesiser (tools/synth/mp3synth.cc).

Presets mirror BASELINE.json's configs:
  cfg3(i)  MPEG-1 44.1 kHz 128 kbps CBR plain stereo, long blocks only, shallow reservoir, 1,149 frames (30 s)
  cfg4(i)  VBR, long/short/mixed blocks, joint stereo with MS + intensity, deep reservoir; 5 % LSF streams
  cfg5()   one 320 kbps stream, mixed features, 413,438 frames (3 h)
  wild(i)  quirk coverage (Q1/Q4/Q5/Q6/Q8/Q15, region clamp, big linbits); fuzz(i): random side info over random bits
Seeds are 0x6D7033 + stream index.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED0 = 0x6D7033


class Cfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_frames", C.c_int32), ("lsf", C.c_int32), ("sfreq", C.c_int32),
                ("mode", C.c_int32), ("crc", C.c_int32), ("bitrate_lo", C.c_int32), ("bitrate_hi", C.c_int32),
                ("padding", C.c_int32), ("blocks", C.c_int32), ("mode_ext_mask", C.c_int32), ("reservoir", C.c_int32),
                ("gain_lo", C.c_int32), ("gain_hi", C.c_int32), ("wild", C.c_int32), ("id3v2_bytes", C.c_int32),
                ("trailer", C.c_int32), ("fill", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libmp3synth.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(so)
        L.synth_bound.argtypes = [C.POINTER(Cfg)]
        L.synth_bound.restype = C.c_size_t
        L.synth_stream.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_size_t]
        L.synth_stream.restype = C.c_size_t
        L.synth_batch.argtypes = [C.POINTER(Cfg), C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.synth_batch.restype = None
        _lib = L
    return _lib


def make(**kw) -> Cfg:
    d = dict(seed=SEED0, n_frames=32, lsf=0, sfreq=0, mode=0, crc=0, bitrate_lo=9, bitrate_hi=9, padding=3, blocks=0,
             mode_ext_mask=1, reservoir=1, gain_lo=150, gain_hi=185, wild=0, id3v2_bytes=0, trailer=0, fill=92)
    d.update(kw)
    return Cfg(**d)


def cfg3(i: int, n_frames: int = 1149) -> Cfg:
    return make(seed=SEED0 + i, n_frames=n_frames)


def cfg4(i: int, n_frames: int = 1149) -> Cfg:
    seed = SEED0 + i
    if i % 20 == 19:  # 5 % LSF streams: 22.05 / 24 / 16 kHz, stereo + mono (no mixed blocks: the reference panics on them)
        k = i // 20
        return make(seed=seed, n_frames=n_frames, lsf=1, sfreq=k % 3, mode=(1, 3, 0)[k % 3], bitrate_lo=6, bitrate_hi=12,
                    padding=0, blocks=1, mode_ext_mask=0xF, reservoir=2)
    return make(seed=seed, n_frames=n_frames, sfreq=i % 3, mode=1, bitrate_lo=6, bitrate_hi=14, padding=2, blocks=2,
                mode_ext_mask=0xF, reservoir=2, crc=(i % 7 == 3))


def cfg5(n_frames: int = 413438) -> Cfg:
    return make(seed=SEED0 + 5, n_frames=n_frames, mode=1, bitrate_lo=14, bitrate_hi=14, padding=3, blocks=2, mode_ext_mask=0xF,
                reservoir=2)


def wild(i: int, n_frames: int = 48) -> Cfg:
    lsf = i % 3 == 2
    return make(seed=SEED0 + 1000003 * (i + 1), n_frames=n_frames, lsf=int(lsf), sfreq=i % 3, mode=(1, 0, 3, 2)[i % 4],
                crc=i % 2, bitrate_lo=(4 if lsf else 5), bitrate_hi=(14 if not lsf else 13), padding=0 if lsf else 2,
                blocks=1 if lsf else 2, mode_ext_mask=0xF, reservoir=2, gain_lo=120, gain_hi=200, wild=1,
                id3v2_bytes=(37 if i % 5 == 0 else 0), trailer=i % 4)


def fuzz(i: int, n_frames: int = 24) -> Cfg:
    c = wild(i, n_frames)
    c.wild = 2
    c.seed = SEED0 + 7777777 * (i + 1)
    return c


def stream(cfg: Cfg) -> bytes:
    L = lib()
    cap = L.synth_bound(C.byref(cfg))
    buf = C.create_string_buffer(cap)
    n = L.synth_stream(C.byref(cfg), buf, cap)
    assert n <= cap
    return buf.raw[:n]


def batch(cfgs, threads: int = 0):
    """Generate many streams on `threads` threads into one numpy buffer; returns (buffer, offsets, lens)."""
    L = lib()
    n = len(cfgs)
    arr = (Cfg * n)(*cfgs)
    caps = [L.synth_bound(C.byref(arr[i])) for i in range(n)]
    offs = np.zeros(n, dtype=np.uint64)
    if n > 1:
        offs[1:] = np.cumsum(np.array(caps[:-1], dtype=np.uint64))
    total = int(sum(caps))
    buf = np.zeros(total, dtype=np.uint8)
    lens = (C.c_size_t * n)()
    c_offs = (C.c_size_t * n)(*[int(o) for o in offs])
    c_caps = (C.c_size_t * n)(*caps)
    L.synth_batch(arr, n, threads or (os.cpu_count() or 1), buf.ctypes.data, c_offs, c_caps, lens)
    return buf, [int(o) for o in offs], [int(x) for x in lens]
