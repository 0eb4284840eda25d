"""go-mp3_b200 — B200 (sm_100a) MP3 Layer III decode engine behind go-mp3's API.

This Python layer is a thin ctypes binding over the two C-ABI libraries built in-tree:

* ``libmp3gpu.so``  (include/mp3gpu.h)  — the device engine: hand-written CUDA kernels for the
  per-granule hot path (Huffman/scalefactors, requantise/stereo/reorder/alias, IMDCT, synthesis).
* ``libmp3host.so`` (include/mp3host.h) — the host-side mirror of the reference's ``package mp3``
  (``NewDecoder``, ``Decoder.Read/Seek/...``, ``DecodeBatch``): tag skipping, header sync, side
  info, bit-reservoir resolution.  In the product this layer is Go + cgo;
  no Go toolchain exists in the build image, so it is mirrored in C++.

There is no CPU decode path and no fallback: if the libraries are missing, or no CUDA device is
present, creating an :class:`Engine` raises.

The directory name contains a hyphen, so the package is imported through
``importlib`` (see ``__graft_entry__.load_package``) under the module name ``go_mp3_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MP3_OK = 0
MP3_EOF = 1
MP3_ERR_UNEXPECTED_EOF = -1
MP3_ERR_SYNC_LIMIT = -2
MP3_ERR_FREE_FORMAT = -3
MP3_ERR_MPEG25 = -4
MP3_ERR_LAYER = -5
MP3_ERR_FRAMESIZE = -6
MP3_ERR_MAINDATA_SIZE = -7
MP3_ERR_ISPOS = -8
MP3_ERR_SEEK_UNSUPPORTED = -11
MP3_ERR_WHENCE = -12
MP3_ERR_REF_PANIC = -14
MP3_ERR_NO_XING_HEADER = -20
MP3_ERR_DEVICE = -50
MP3_ERR_INVALID = -51

TAP_IS, TAP_COUNT1, TAP_SCALEFAC, TAP_XR, TAP_HYBRID = 0, 1, 2, 3, 4
PCM_BYTES_PER_GRANULE = 2304
W2_VALID = 1 << 25
W2_ZERO_STATE = 1 << 26


class Mp3Error(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        super().__init__(msg or f"mp3 error {code}")


class Unit(C.Structure):
    """mp3gpu_unit (include/mp3gpu.h)."""
    _fields_ = [("bit_start", C.c_uint64), ("buf_end_rel", C.c_int32), ("w0", C.c_uint32), ("w1", C.c_uint32),
                ("w2", C.c_uint32), ("reserved", C.c_uint32 * 2)]


UNIT_DTYPE = np.dtype([("bit_start", "<u8"), ("buf_end_rel", "<i4"), ("w0", "<u4"), ("w1", "<u4"), ("w2", "<u4"),
                       ("r0", "<u4"), ("r1", "<u4")])
assert UNIT_DTYPE.itemsize == 32 and C.sizeof(Unit) == 32


class GpuOpts(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("wave_granules", C.c_uint32), ("keep_intermediates", C.c_uint32),
                ("reserved", C.c_uint32)]


class GpuTimings(C.Structure):
    _fields_ = [("k1_huffman_ms", C.c_float), ("k_hybrid_ms", C.c_float), ("k_synth_ms", C.c_float), ("reserved_ms", C.c_float),
                ("total_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("waves", C.c_uint32), ("launches", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved_ms"}


MAX_DEVICES = 16


class EngineOpts(C.Structure):
    """mp3_engine_opts (include/mp3host.h)."""
    _fields_ = [("device", C.c_int), ("host_threads", C.c_int), ("wave_granules", C.c_uint32),
                ("chunk_frames", C.c_uint32), ("keep_intermediates", C.c_uint32), ("use_exact_library", C.c_uint32),
                ("n_devices", C.c_int), ("devices", C.c_int * MAX_DEVICES), ("trim_gapless", C.c_uint32),
                ("reserved", C.c_uint32)]


class LameInfoStruct(C.Structure):
    """mp3_lame_info (include/mp3host.h) = lameinfo.Info (lameinfo/lameinfo.go:20-49)."""
    _fields_ = [("is_xing", C.c_int32), ("flags", C.c_uint32), ("frame_count", C.c_uint32), ("byte_count", C.c_uint32),
                ("toc", C.c_uint8 * 100), ("vbr_scale", C.c_uint32), ("has_lame_info", C.c_int32),
                ("lame_version", C.c_char * 12), ("encoder_delay", C.c_uint16), ("encoder_padding", C.c_uint16)]


class StreamResult(C.Structure):
    _fields_ = [("pcm_offset", C.c_int64), ("pcm_bytes", C.c_int64), ("sample_rate", C.c_int32),
                ("status", C.c_int32), ("frames", C.c_int64)]


class BatchTimings(C.Structure):
    _fields_ = [("parse_s", C.c_double), ("gather_s", C.c_double), ("device_s", C.c_double), ("total_s", C.c_double),
                ("main_data_bytes", C.c_uint64), ("n_granules", C.c_uint64), ("pcm_bytes", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Parsed(C.Structure):
    _fields_ = [("main_data", C.POINTER(C.c_uint8)), ("main_data_len", C.c_size_t), ("units", C.POINTER(Unit)),
                ("n_granules", C.c_size_t), ("streams", C.POINTER(StreamResult)), ("n_streams", C.c_size_t)]


def _lib_path(name: str) -> str:
    p = os.path.join(_HERE, name)
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with __graft_entry__.build() (make -C go-mp3_b200/csrc). "
                          "There is no fallback path.")
    return p


_host = None
_gpu = {}


def host_lib() -> C.CDLL:
    """libmp3host.so with prototypes set."""
    global _host
    if _host is not None:
        return _host
    L = C.CDLL(_lib_path("libmp3host.so"))
    vp, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
    pp_u8 = C.POINTER(C.c_char_p)
    L.mp3_engine_create.argtypes = [C.POINTER(EngineOpts), C.POINTER(vp)]
    L.mp3_engine_create.restype = C.c_int
    L.mp3_engine_destroy.argtypes = [vp]
    L.mp3_engine_destroy.restype = None
    L.mp3_engine_last_error.argtypes = [vp]
    L.mp3_engine_last_error.restype = C.c_char_p
    L.mp3_engine_gpu.argtypes = [vp]
    L.mp3_engine_gpu.restype = vp
    L.mp3_error_string.argtypes = [C.c_int]
    L.mp3_error_string.restype = C.c_char_p
    L.mp3_new_decoder.argtypes = [vp, C.c_void_p, sz, C.c_int, C.POINTER(C.c_int)]
    L.mp3_new_decoder.restype = vp
    L.mp3_decoder_free.argtypes = [vp]
    L.mp3_decoder_free.restype = None
    L.mp3_decoder_read.argtypes = [vp, C.c_void_p, sz, C.POINTER(C.c_int)]
    L.mp3_decoder_read.restype = C.c_long
    L.mp3_decoder_seek.argtypes = [vp, i64, C.c_int, C.POINTER(C.c_int)]
    L.mp3_decoder_seek.restype = i64
    for name in ("length", "bytes_per_frame", "duration_ns", "position_ns", "remaining_ns", "sample_position",
                 "sample_count"):
        f = getattr(L, "mp3_decoder_" + name)
        f.argtypes = [vp]
        f.restype = i64
    L.mp3_decoder_sample_rate.argtypes = [vp]
    L.mp3_decoder_sample_rate.restype = C.c_int
    L.mp3_decoder_progress.argtypes = [vp]
    L.mp3_decoder_progress.restype = C.c_double
    for name in ("seek_to_sample", "skip", "seek_to_time"):
        f = getattr(L, "mp3_decoder_" + name)
        f.argtypes = [vp, i64]
        f.restype = C.c_int
    L.mp3_decode_batch.argtypes = [vp, pp_u8, C.POINTER(sz), sz, C.POINTER(StreamResult), C.POINTER(C.c_void_p),
                                   C.POINTER(BatchTimings)]
    L.mp3_decode_batch.restype = C.c_int
    L.mp3_engine_device_count.argtypes = [vp]
    L.mp3_engine_device_count.restype = C.c_int
    L.mp3_engine_gpu_at.argtypes = [vp, C.c_int]
    L.mp3_engine_gpu_at.restype = vp
    L.mp3_new_decoder_on.argtypes = [vp, C.c_int, C.c_void_p, sz, C.c_int, C.POINTER(C.c_int)]
    L.mp3_new_decoder_on.restype = vp
    L.mp3_stream_index_create.argtypes = [C.c_void_p, sz, C.POINTER(vp)]
    L.mp3_stream_index_create.restype = C.c_int
    L.mp3_stream_index_free.argtypes = [vp]
    L.mp3_stream_index_free.restype = None
    L.mp3_stream_index_frames.argtypes = [vp]
    L.mp3_stream_index_frames.restype = i64
    L.mp3_stream_index_sample_rate.argtypes = [vp]
    L.mp3_stream_index_sample_rate.restype = C.c_int
    L.mp3_stream_index_frame_pos.argtypes = [vp, i64]
    L.mp3_stream_index_frame_pos.restype = i64
    L.mp3_stream_index_pcm_bytes.argtypes = [vp, i64, i64]
    L.mp3_stream_index_pcm_bytes.restype = i64
    L.mp3_decode_frames.argtypes = [vp, C.c_int, vp, i64, i64, C.c_void_p, C.POINTER(i64)]
    L.mp3_decode_frames.restype = C.c_int
    L.mp3_decode_stream_split.argtypes = [vp, vp, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(BatchTimings)]
    L.mp3_decode_stream_split.restype = C.c_int
    L.mp3_lameinfo_parse.argtypes = [C.c_void_p, sz, C.POINTER(LameInfoStruct)]
    L.mp3_lameinfo_parse.restype = C.c_int
    L.mp3_lameinfo_parse_from_reader.argtypes = [C.c_void_p, sz, C.POINTER(LameInfoStruct)]
    L.mp3_lameinfo_parse_from_reader.restype = C.c_int
    L.mp3_lameinfo_total_delay.argtypes = [C.POINTER(LameInfoStruct)]
    L.mp3_lameinfo_total_delay.restype = C.c_int
    L.mp3_lameinfo_total_padding.argtypes = [C.POINTER(LameInfoStruct)]
    L.mp3_lameinfo_total_padding.restype = C.c_int
    L.mp3_lameinfo_is_lame_version.argtypes = [C.c_void_p, sz]
    L.mp3_lameinfo_is_lame_version.restype = C.c_int
    L.mp3_parse_streams.argtypes = [pp_u8, C.POINTER(sz), sz, C.c_int, C.POINTER(C.POINTER(Parsed))]
    L.mp3_parse_streams.restype = C.c_int
    L.mp3_parsed_free.argtypes = [C.POINTER(Parsed)]
    L.mp3_parsed_free.restype = None
    _host = L
    return L


def gpu_lib(exact=False) -> C.CDLL:
    """libmp3gpu.so (False), the no-contraction build libmp3gpu_exact.so (True) or the bounds-checked build
    libmp3gpu_checked.so ("checked"), with prototypes set."""
    if exact in _gpu:
        return _gpu[exact]
    name = {False: "libmp3gpu.so", True: "libmp3gpu_exact.so", "checked": "libmp3gpu_checked.so"}[exact]
    if exact is False and os.environ.get("MP3GPU_LIB_VARIANT"):  # tools/ only: a diagnostic build of the product library (make probe)
        name = "libmp3gpu_%s.so" % os.environ["MP3GPU_LIB_VARIANT"]
    L = C.CDLL(_lib_path(name))
    vp, sz = C.c_void_p, C.c_size_t
    L.mp3gpu_create.argtypes = [C.c_int, C.POINTER(GpuOpts), C.POINTER(vp)]
    L.mp3gpu_create.restype = C.c_int
    L.mp3gpu_destroy.argtypes = [vp]
    L.mp3gpu_destroy.restype = None
    L.mp3gpu_last_error.argtypes = [vp]
    L.mp3gpu_last_error.restype = C.c_char_p
    L.mp3gpu_decode.argtypes = [vp, vp, sz, vp, sz, vp]
    L.mp3gpu_decode.restype = C.c_int
    L.mp3gpu_decode_range.argtypes = [vp, vp, sz, vp, sz, sz, vp]
    L.mp3gpu_decode_range.restype = C.c_int
    L.mp3gpu_decode_device.argtypes = [vp, vp, sz, vp, sz, vp]
    L.mp3gpu_decode_device.restype = C.c_int
    L.mp3gpu_decode_device_async.argtypes = [vp, vp, sz, vp, sz, vp]
    L.mp3gpu_decode_device_async.restype = C.c_int
    L.mp3gpu_event_record.argtypes = [vp, C.c_int]
    L.mp3gpu_event_record.restype = C.c_int
    L.mp3gpu_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.mp3gpu_event_elapsed_ms.restype = C.c_int
    L.mp3gpu_host_alloc.argtypes = [sz]
    L.mp3gpu_host_alloc.restype = vp
    L.mp3gpu_host_free.argtypes = [vp]
    L.mp3gpu_host_free.restype = None
    L.mp3gpu_device_alloc.argtypes = [vp, sz]
    L.mp3gpu_device_alloc.restype = vp
    L.mp3gpu_device_free.argtypes = [vp, vp]
    L.mp3gpu_device_free.restype = None
    L.mp3gpu_copy_to_device.argtypes = [vp, vp, vp, sz]
    L.mp3gpu_copy_to_device.restype = C.c_int
    L.mp3gpu_copy_to_host.argtypes = [vp, vp, vp, sz]
    L.mp3gpu_copy_to_host.restype = C.c_int
    L.mp3gpu_synchronize.argtypes = [vp]
    L.mp3gpu_synchronize.restype = C.c_int
    L.mp3gpu_last_timings.argtypes = [vp, C.POINTER(GpuTimings)]
    L.mp3gpu_last_timings.restype = C.c_int
    L.mp3gpu_debug_read.argtypes = [vp, C.c_int, sz, sz, vp]
    L.mp3gpu_debug_read.restype = C.c_int
    L.mp3gpu_device_info.argtypes = [vp, C.c_char_p, sz, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mp3gpu_device_info.restype = C.c_int
    L.mp3gpu_device_pci_bus_id.argtypes = [vp, C.c_char_p, sz]
    L.mp3gpu_device_pci_bus_id.restype = C.c_int
    L.mp3gpu_pcm_to_f32_planar.argtypes = [vp, vp, sz, vp, vp]
    L.mp3gpu_pcm_to_f32_planar.restype = C.c_int
    L.mp3gpu_measure_d2h.argtypes = [vp, vp, sz, C.c_int, C.POINTER(C.c_double)]
    L.mp3gpu_measure_d2h.restype = C.c_int
    L.mp3gpu_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.mp3gpu_measure_fp32_peak.restype = C.c_int
    _gpu[exact] = L
    return L


def shard_streams(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """Half-open range [first, last) of the streams rank `rank` of `world` decodes.  Streams are independent
    (SURVEY.md 8e), so multi-GPU decoding shards the stream set; there is no data-path collective."""
    base, extra = divmod(n_streams, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def frame_range_job(pb: "ParsedBatch", stream: int, f0: int, f1: int):
    """Self-contained decode job for frames [f0, f1) of one stream of a parsed batch (BASELINE.json configs[4]: one long
    stream split at frame boundaries, e.g. one range per GPU, or a random-access read).

    The job starts with a halo of whole frames covering the two granules in front of f0: PCM of a granule needs the
    IMDCT overlap of the granule before it and 15 slots of V history, which in turn need the overlap of the one before
    that.  The reservoir needs no halo: bit-slices were resolved on the host over the whole stream.
    Returns (main_data_window, window_len, units, halo_granules); decode it and drop the first halo_granules * 2304 bytes.
    """
    st = pb.streams[stream]
    frames = st["frames"]
    if not (0 <= f0 <= f1 <= frames):
        raise ValueError("frame range outside the stream")
    gpf = (st["pcm_bytes"] // 2304) // frames if frames else 2   # granules per frame: 2 (MPEG-1) or 1 (LSF)
    g_base = st["pcm_offset"] // 2304
    halo_frames = min(f0, 2 // gpf if gpf else 1)
    ga, gb = g_base + (f0 - halo_frames) * gpf, g_base + f1 * gpf
    units = pb.units[2 * ga:2 * gb].copy()
    if len(units) == 0:
        return np.zeros(64, np.uint8), 0, units, 0
    valid = (units["w2"] & W2_VALID) != 0
    lo_bit = int(units["bit_start"][valid].min())
    hi_bit = int((units["bit_start"][valid].astype(np.int64) + units["buf_end_rel"][valid]).max())
    lo = (lo_bit // 8) & ~3
    hi = min(pb.main_data_len, (max(hi_bit, lo_bit) + 7) // 8)
    window = np.zeros(hi - lo + 64, np.uint8)
    window[:hi - lo] = pb.main_data[lo:hi]
    units["bit_start"] -= np.uint64(lo * 8)
    return window, hi - lo, units, halo_frames * gpf


def error_string(code: int) -> str:
    return host_lib().mp3_error_string(code).decode()


# --------------------------------------------------------------------------------------------------
# Host-only stage: parse + reservoir resolution (no GPU needed)
# --------------------------------------------------------------------------------------------------
class ParsedBatch:
    """Result of the host stage for a batch of streams: main data M, unit descriptors, per-stream results."""

    def __init__(self, main_data: np.ndarray, units: np.ndarray, streams: List[dict]):
        self.main_data = main_data      # uint8, padded by 64 zero bytes (main_data_len excludes the pad)
        self.units = units              # UNIT_DTYPE, 2 per granule
        self.streams = streams
        self.main_data_len = len(main_data) - 64
        self.n_granules = len(units) // 2


class StreamBuffer:
    """Many streams stored in one uint8 numpy buffer: stream i = buf[offsets[i] : offsets[i] + lens[i]]."""

    def __init__(self, buf: np.ndarray, offsets: Sequence[int], lens: Sequence[int]):
        self.buf, self.offsets, self.lens = buf, list(offsets), list(lens)

    def __len__(self):
        return len(self.offsets)

    def stream(self, i: int) -> bytes:
        return self.buf[self.offsets[i]:self.offsets[i] + self.lens[i]].tobytes()


def _stream_args(streams):
    """Pointer + length arrays for a list of bytes objects or a StreamBuffer.  Returns (ptrs, lens, n, keepalive)."""
    n = len(streams)
    arr = (C.c_char_p * max(n, 1))()
    lens = (C.c_size_t * max(n, 1))()
    if isinstance(streams, StreamBuffer):
        base = streams.buf.ctypes.data
        ptrs = (C.c_void_p * max(n, 1))()
        for i in range(n):
            ptrs[i] = base + streams.offsets[i]
            lens[i] = streams.lens[i]
        return C.cast(ptrs, C.POINTER(C.c_char_p)), lens, n, (ptrs, streams.buf)
    for i, s in enumerate(streams):
        arr[i] = s
        lens[i] = len(s)
    return arr, lens, n, streams


def _result_dict(r: StreamResult) -> dict:
    return {"pcm_offset": r.pcm_offset, "pcm_bytes": r.pcm_bytes, "sample_rate": r.sample_rate, "status": r.status,
            "frames": r.frames}


def unit_slots_upper_bound(stream: bytes) -> int:
    """The header-only bound DecodeBatch sizes its arenas with (mp3_debug_unit_slots_upper_bound; test hook)."""
    L = host_lib()
    L.mp3_debug_unit_slots_upper_bound.argtypes = [C.c_char_p, C.c_size_t]
    L.mp3_debug_unit_slots_upper_bound.restype = C.c_size_t
    return int(L.mp3_debug_unit_slots_upper_bound(stream, len(stream)))


def main_bytes_upper_bound(stream: bytes) -> int:
    """The same walk's bound of the stream's main-data bytes (mp3_debug_main_bytes_upper_bound; test hook)."""
    L = host_lib()
    L.mp3_debug_main_bytes_upper_bound.argtypes = [C.c_char_p, C.c_size_t]
    L.mp3_debug_main_bytes_upper_bound.restype = C.c_size_t
    return int(L.mp3_debug_main_bytes_upper_bound(stream, len(stream)))


def parse_streams(streams: Sequence[bytes], host_threads: int = 0) -> ParsedBatch:
    """Host half of DecodeBatch: tags, headers, side info, reservoir -> bit-slices (mp3_parse_streams)."""
    L = host_lib()
    arr, lens, n, _keep = _stream_args(streams)
    out = C.POINTER(Parsed)()
    rc = L.mp3_parse_streams(arr, lens, n, host_threads, C.byref(out))
    if rc != MP3_OK:
        raise Mp3Error(rc, "mp3_parse_streams failed")
    try:
        p = out.contents
        md = np.ctypeslib.as_array(p.main_data, shape=(p.main_data_len + 64,)).copy()
        nu = p.n_granules * 2
        if nu:
            raw = np.ctypeslib.as_array(C.cast(p.units, C.POINTER(C.c_uint8)), shape=(nu * 32,))
            units = raw.copy().view(UNIT_DTYPE)
        else:
            units = np.zeros(0, dtype=UNIT_DTYPE)
        res = [_result_dict(p.streams[i]) for i in range(p.n_streams)]
    finally:
        L.mp3_parsed_free(out)
    return ParsedBatch(md, units, res)


# --------------------------------------------------------------------------------------------------
# Device engine (mp3gpu.h) — used directly by bench.py and the stage-level parity tests
# --------------------------------------------------------------------------------------------------
class GpuEngine:
    def __init__(self, device: int = 0, wave_granules: int = 0, keep_intermediates: bool = False, exact: bool = False,
                 checked: bool = False):
        self.lib = gpu_lib("checked" if checked else exact)
        self.ctx = C.c_void_p()
        opts = GpuOpts(1, wave_granules, 1 if keep_intermediates else 0, 0)
        rc = self.lib.mp3gpu_create(device, C.byref(opts), C.byref(self.ctx))
        if rc != 0:
            self.ctx = None
            why = {-1: "no CUDA device (there is no CPU path)", -2: "CUDA error", -3: "invalid tables/arguments (see stderr)",
                   -4: "out of memory"}.get(rc, "unknown")
            raise Mp3Error(rc, f"mp3gpu_create failed ({rc}): {why}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.mp3gpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise Mp3Error(rc, self.lib.mp3gpu_last_error(self.ctx).decode())

    def decode(self, main_data: np.ndarray, main_data_len: int, units: np.ndarray) -> np.ndarray:
        """Host-buffer decode (mp3gpu_decode): returns int16 [n_granules*576, 2]."""
        n_gr = len(units) // 2
        pcm = np.empty((n_gr * 576, 2), dtype=np.int16)
        main_data = np.ascontiguousarray(main_data)
        units = np.ascontiguousarray(units)
        self._check(self.lib.mp3gpu_decode(self.ctx, main_data.ctypes.data, main_data_len, units.ctypes.data, n_gr,
                                           pcm.ctypes.data))
        return pcm

    def decode_device(self, d_main: int, main_len: int, d_units: int, n_granules: int, d_pcm: int, sync: bool = True):
        """Device-resident decode (raw device pointers, e.g. torch tensors' data_ptr())."""
        f = self.lib.mp3gpu_decode_device if sync else self.lib.mp3gpu_decode_device_async
        self._check(f(self.ctx, d_main, main_len, d_units, n_granules, d_pcm))

    def pcm_to_f32_planar(self, d_pcm: int, n_samples: int, d_left: int, d_right: int):
        """Output side: device-resident s16 interleaved PCM -> two float32 planes (x 1/32768), queued behind the decode."""
        self._check(self.lib.mp3gpu_pcm_to_f32_planar(self.ctx, d_pcm, n_samples, d_left, d_right))

    def decode_host(self, p_main: int, main_len: int, p_units: int, n_granules: int, p_pcm: int):
        """Host-buffer decode with raw host pointers (pinned buffers from host_alloc make the copies asynchronous)."""
        self._check(self.lib.mp3gpu_decode(self.ctx, p_main, main_len, p_units, n_granules, p_pcm))

    def event_record(self, which: int):
        self._check(self.lib.mp3gpu_event_record(self.ctx, which))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self.lib.mp3gpu_event_elapsed_ms(self.ctx, a, b, C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._check(self.lib.mp3gpu_synchronize(self.ctx))

    def host_alloc(self, nbytes: int) -> int:
        p = self.lib.mp3gpu_host_alloc(nbytes)
        if not p:
            raise MemoryError(f"pinned host allocation of {nbytes} bytes failed")
        return p

    def host_free(self, p: int):
        self.lib.mp3gpu_host_free(p)

    def decode_frames(self, pb: "ParsedBatch", stream: int, f0: int, f1: int) -> np.ndarray:
        """PCM of frames [f0, f1) of one stream, decoded on their own (halo handled): int16 [samples, 2]."""
        window, n, units, halo = frame_range_job(pb, stream, f0, f1)
        if len(units) == 0:
            return np.zeros((0, 2), np.int16)
        return self.decode(window, n, units)[halo * 576:]

    def timings(self) -> dict:
        t = GpuTimings()
        self._check(self.lib.mp3gpu_last_timings(self.ctx, C.byref(t)))
        return t.as_dict()

    def tap(self, which: int, first: int, n: int) -> np.ndarray:
        shapes = {TAP_IS: ((n, 2, 576), np.int16), TAP_COUNT1: ((n, 2), np.int32), TAP_SCALEFAC: ((n, 2, 64), np.uint8),
                  TAP_XR: ((n, 2, 576), np.float32), TAP_HYBRID: ((n, 2, 576), np.float32)}
        shape, dt = shapes[which]
        out = np.zeros(shape, dtype=dt)
        self._check(self.lib.mp3gpu_debug_read(self.ctx, which, first, n, out.ctypes.data))
        return out

    def device_info(self) -> dict:
        name = C.create_string_buffer(256)
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.mp3gpu_device_info(self.ctx, name, 256, C.byref(sm), C.byref(maj), C.byref(mnr)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": f"{maj.value}.{mnr.value}"}

    def bind_host_to_gpu_numa_node(self):
        """Pin the calling process to the CPUs of the NUMA node this GPU hangs off, so that pinned buffers allocated
        afterwards (first touch) and the copies' host side stay local.  Returns the node, or None if unknown."""
        buf = C.create_string_buffer(64)
        if self.lib.mp3gpu_device_pci_bus_id(self.ctx, buf, 64) != 0:
            return None
        bus = buf.value.decode().lower()
        try:
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            if node < 0:
                return None
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
            return node
        except (OSError, ValueError):
            return None

    def fp32_peak_tflops(self) -> float:
        v = C.c_double()
        self._check(self.lib.mp3gpu_measure_fp32_peak(self.ctx, C.byref(v)))
        return v.value


# --------------------------------------------------------------------------------------------------
# Host mirror of package mp3 (mp3host.h)
# --------------------------------------------------------------------------------------------------
class Engine:
    """One GPU + the host stage.  Mirrors what the Go package holds as package-level state."""

    def __init__(self, device: int = 0, host_threads: int = 0, wave_granules: int = 0, chunk_frames: int = 0,
                 keep_intermediates: bool = False, exact: bool = False, devices: Optional[Sequence[int]] = None,
                 trim_gapless: bool = False, checked: bool = False):
        """`devices`: CUDA ordinals of a multi-GPU engine (DecodeBatch shards its streams over them, decode_stream_split
        cuts one stream into a frame range per device); the same ordinal may appear twice (two device engines on one GPU)."""
        self.lib = host_lib()
        self.h = C.c_void_p()
        opts = EngineOpts()
        opts.device, opts.host_threads, opts.wave_granules, opts.chunk_frames = device, host_threads, wave_granules, chunk_frames
        opts.keep_intermediates, opts.use_exact_library = (1 if keep_intermediates else 0), (2 if checked else (1 if exact else 0))
        if devices:
            if len(devices) > MAX_DEVICES:
                raise ValueError("too many devices")
            opts.n_devices = len(devices)
            for i, d in enumerate(devices):
                opts.devices[i] = int(d)
        opts.trim_gapless = 1 if trim_gapless else 0
        rc = self.lib.mp3_engine_create(C.byref(opts), C.byref(self.h))
        if rc != MP3_OK:
            self.h = None
            raise Mp3Error(rc, "mp3_engine_create failed: a CUDA device and libmp3gpu.so are required; "
                               "there is no CPU decode path")
        self.exact = "checked" if checked else exact

    def close(self):
        if getattr(self, "h", None):
            self.lib.mp3_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def new_decoder(self, data: bytes, seekable: bool = True, slot: int = 0) -> "Decoder":
        return Decoder(self, data, seekable, slot)

    def device_count(self) -> int:
        return self.lib.mp3_engine_device_count(self.h)

    def measure_d2h_ceiling(self, bytes_per_device: int, reps: int = 3) -> dict:
        """Plain cudaMemcpyAsync device -> pinned host on every device of the engine AT THE SAME TIME (one thread per
        device): the copy ceiling the end-to-end path is measured against.  Returns aggregate and per-device GB/s."""
        import threading
        g = gpu_lib(getattr(self, "exact", False))
        n = self.device_count()
        bufs = [g.mp3gpu_host_alloc(bytes_per_device) for _ in range(n)]
        if not all(bufs):
            for b in bufs:
                if b:
                    g.mp3gpu_host_free(b)
            raise MemoryError("pinned allocation for the copy ceiling failed")
        secs = [C.c_double(0) for _ in range(n)]
        rcs = [0] * n
        start = threading.Barrier(n)

        def run(i):
            ctx = self.lib.mp3_engine_gpu_at(self.h, i)
            start.wait()
            rcs[i] = g.mp3gpu_measure_d2h(ctx, bufs[i], bytes_per_device, reps, C.byref(secs[i]))

        th = [threading.Thread(target=run, args=(i,)) for i in range(n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for b in bufs:
            g.mp3gpu_host_free(b)
        if any(rcs):
            raise Mp3Error(-2, "mp3gpu_measure_d2h failed")
        per = [bytes_per_device * reps / s.value / 1e9 for s in secs]
        return {"aggregate_gbs": bytes_per_device * reps * n / max(s.value for s in secs) / 1e9, "per_device_gbs": per}

    def decode_frames(self, index: "StreamIndex", f0: int, f1: int, slot: int = 0) -> Tuple[np.ndarray, int]:
        """PCM bytes of frames [f0, f1) of an indexed stream, identical to that stretch of the linear decode
        (mp3_decode_frames).  Returns (uint8 array, status)."""
        want = index.pcm_bytes(f0, f1)
        out = np.empty(max(want, 0), dtype=np.uint8)
        got = C.c_int64(0)
        rc = self.lib.mp3_decode_frames(self.h, slot, index.h, f0, f1, out.ctypes.data, C.byref(got))
        if rc == MP3_ERR_DEVICE or rc == MP3_ERR_INVALID:
            raise Mp3Error(rc, self.lib.mp3_engine_last_error(self.h).decode() or error_string(rc))
        return out[:got.value], rc

    def decode_stream_split(self, index: "StreamIndex") -> Tuple[np.ndarray, int, dict]:
        """One stream cut into a frame range per device, ranges decoded concurrently (mp3_decode_stream_split).
        Returns (PCM bytes view valid until the next batch/split call, status, timings)."""
        base = C.c_void_p()
        n = C.c_int64(0)
        tm = BatchTimings()
        rc = self.lib.mp3_decode_stream_split(self.h, index.h, C.byref(base), C.byref(n), C.byref(tm))
        if rc == MP3_ERR_DEVICE or rc == MP3_ERR_INVALID:
            raise Mp3Error(rc, self.lib.mp3_engine_last_error(self.h).decode() or error_string(rc))
        pcm = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(n.value,)) if n.value else np.zeros(0, np.uint8)
        return pcm, rc, tm.as_dict()

    def decode_batch(self, streams: Sequence[bytes]) -> Tuple[List[dict], np.ndarray, dict]:
        """DecodeBatch: returns (per-stream results, PCM bytes view (valid until the next call), timings)."""
        arr, lens, n, _keep = _stream_args(streams)
        res = (StreamResult * max(n, 1))()
        base = C.c_void_p()
        tm = BatchTimings()
        rc = self.lib.mp3_decode_batch(self.h, arr, lens, n, res, C.byref(base), C.byref(tm))
        if rc != MP3_OK:
            raise Mp3Error(rc, self.lib.mp3_engine_last_error(self.h).decode())
        total = int(tm.pcm_bytes)
        if total:
            pcm = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(total,))
        else:
            pcm = np.zeros(0, dtype=np.uint8)
        return [_result_dict(res[i]) for i in range(n)], pcm, tm.as_dict()


class Decoder:
    """Drop-in mirror of *mp3.Decoder (decode.go): Read/Seek/SampleRate/Length/... over in-memory data."""

    def __init__(self, engine: Engine, data: bytes, seekable: bool = True, slot: int = 0):
        self.engine = engine
        self.lib = engine.lib
        self._data = bytes(data)  # must outlive the decoder
        err = C.c_int(0)
        self.h = self.lib.mp3_new_decoder_on(engine.h, slot, self._data, len(self._data), 1 if seekable else 0, C.byref(err))
        if not self.h:
            raise Mp3Error(err.value, "NewDecoder: " + error_string(err.value))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mp3_decoder_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def read(self, n: int) -> Tuple[bytes, int]:
        """Decoder.Read: returns (bytes, err) with err MP3_OK, MP3_EOF or a fatal code.  Like the reference it hands out at
        most the rest of one frame per call."""
        if getattr(self, "_rbuf", None) is None or len(self._rbuf) < n:
            self._rbuf = C.create_string_buffer(max(n, 8192))
        err = C.c_int(0)
        got = self.lib.mp3_decoder_read(self.h, self._rbuf, n, C.byref(err))
        return C.string_at(self._rbuf, got), err.value

    def read_all(self) -> Tuple[bytes, int]:
        """io.ReadAll(d): (bytes, err) with err MP3_OK on clean EOF."""
        chunks = []
        while True:
            b, err = self.read(8192)
            if not b:
                return b"".join(chunks), (MP3_OK if err == MP3_EOF else err)
            chunks.append(b)

    def seek(self, offset: int, whence: int = 0) -> int:
        err = C.c_int(0)
        r = self.lib.mp3_decoder_seek(self.h, offset, whence, C.byref(err))
        if err.value != MP3_OK:
            raise Mp3Error(err.value, error_string(err.value))
        return r

    def sample_rate(self) -> int: return self.lib.mp3_decoder_sample_rate(self.h)
    def length(self) -> int: return self.lib.mp3_decoder_length(self.h)
    def bytes_per_frame(self) -> int: return self.lib.mp3_decoder_bytes_per_frame(self.h)
    def duration_ns(self) -> int: return self.lib.mp3_decoder_duration_ns(self.h)
    def position_ns(self) -> int: return self.lib.mp3_decoder_position_ns(self.h)
    def remaining_ns(self) -> int: return self.lib.mp3_decoder_remaining_ns(self.h)
    def progress(self) -> float: return self.lib.mp3_decoder_progress(self.h)
    def sample_position(self) -> int: return self.lib.mp3_decoder_sample_position(self.h)
    def sample_count(self) -> int: return self.lib.mp3_decoder_sample_count(self.h)

    def _chk(self, rc):
        if rc != MP3_OK:
            raise Mp3Error(rc, error_string(rc))

    def seek_to_sample(self, s: int): self._chk(self.lib.mp3_decoder_seek_to_sample(self.h, s))
    def skip(self, delta_ns: int): self._chk(self.lib.mp3_decoder_skip(self.h, delta_ns))
    def seek_to_time(self, t_ns: int): self._chk(self.lib.mp3_decoder_seek_to_time(self.h, t_ns))


class StreamIndex:
    """Frame index of one stream (mp3_stream_index): what frame-range decode and the split over devices work from."""

    def __init__(self, data):
        self.lib = host_lib()
        if isinstance(data, np.ndarray):
            self._data = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, n = self._data.ctypes.data, self._data.size
        else:
            self._data = bytes(data)
            ptr, n = C.cast(C.c_char_p(self._data), C.c_void_p), len(self._data)
        self.h = C.c_void_p()
        rc = self.lib.mp3_stream_index_create(ptr, n, C.byref(self.h))
        if rc != MP3_OK:
            self.h = None
            raise Mp3Error(rc, "mp3_stream_index_create: " + error_string(rc))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mp3_stream_index_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def frames(self) -> int: return self.lib.mp3_stream_index_frames(self.h)
    def sample_rate(self) -> int: return self.lib.mp3_stream_index_sample_rate(self.h)
    def pcm_bytes(self, f0: int, f1: int) -> int: return self.lib.mp3_stream_index_pcm_bytes(self.h, f0, f1)
    def frame_pos(self, f: int) -> int: return self.lib.mp3_stream_index_frame_pos(self.h, f)


# --------------------------------------------------------------------------------------------------
# lameinfo (mirror of package lameinfo, lameinfo/lameinfo.go)
# --------------------------------------------------------------------------------------------------
LAME_FLAG_FRAME_COUNT, LAME_FLAG_BYTE_COUNT, LAME_FLAG_TOC, LAME_FLAG_VBR_SCALE = 1, 2, 4, 8
LAME_DECODER_DELAY = 529


class NoXingHeader(Mp3Error):
    """lameinfo.ErrNoXingHeader."""


class LameInfo:
    """lameinfo.Info with its helper methods (lameinfo.go:20-108)."""

    def __init__(self, raw: Optional[LameInfoStruct] = None, **kw):
        self._raw = raw if raw is not None else LameInfoStruct()
        for k, v in kw.items():  # tests build Info values directly, like the reference's do
            if k == "lame_version":
                self._raw.lame_version = v if isinstance(v, bytes) else v.encode()
                self._raw.has_lame_info = 1 if v else 0
            else:
                setattr(self._raw, k, v)

    is_xing = property(lambda s: bool(s._raw.is_xing))
    flags = property(lambda s: s._raw.flags)
    frame_count = property(lambda s: s._raw.frame_count)
    byte_count = property(lambda s: s._raw.byte_count)
    toc = property(lambda s: bytes(s._raw.toc))
    vbr_scale = property(lambda s: s._raw.vbr_scale)
    encoder_delay = property(lambda s: s._raw.encoder_delay)
    encoder_padding = property(lambda s: s._raw.encoder_padding)

    @property
    def lame_version(self) -> bytes:
        """The 9 bytes of the version field (b"" if there is no LAME tag); may end in NULs (e.g. b"LAME3.99\x00")."""
        return C.string_at(C.addressof(self._raw) + LameInfoStruct.lame_version.offset, 9) if self._raw.has_lame_info else b""

    def toc_offset(self, fraction: float, stream_bytes: int = 0) -> int:
        """Byte offset (from the first audio frame) at `fraction` of the playing time, interpolated in the TOC; -1 without one."""
        L = host_lib()
        L.mp3_lameinfo_toc_offset.argtypes = [C.POINTER(LameInfoStruct), C.c_double, C.c_uint64]
        L.mp3_lameinfo_toc_offset.restype = C.c_int64
        return int(L.mp3_lameinfo_toc_offset(C.byref(self._raw), float(fraction), int(stream_bytes)))

    def has_frame_count(self): return bool(self.flags & LAME_FLAG_FRAME_COUNT)
    def has_byte_count(self): return bool(self.flags & LAME_FLAG_BYTE_COUNT)
    def has_toc(self): return bool(self.flags & LAME_FLAG_TOC)
    def has_vbr_scale(self): return bool(self.flags & LAME_FLAG_VBR_SCALE)
    def has_lame_info(self): return bool(self._raw.has_lame_info)
    def total_delay(self) -> int: return host_lib().mp3_lameinfo_total_delay(C.byref(self._raw))
    def total_padding(self) -> int: return host_lib().mp3_lameinfo_total_padding(C.byref(self._raw))


def _lame_result(rc: int, raw: LameInfoStruct) -> LameInfo:
    if rc == MP3_ERR_NO_XING_HEADER:
        raise NoXingHeader(rc, error_string(rc))
    if rc != MP3_OK:
        raise Mp3Error(rc, "EOF" if rc == MP3_EOF else error_string(rc))
    return LameInfo(raw)


def lameinfo_parse(frame: bytes) -> LameInfo:
    """lameinfo.Parse (lameinfo.go:139-270)."""
    raw = LameInfoStruct()
    return _lame_result(host_lib().mp3_lameinfo_parse(bytes(frame), len(frame), C.byref(raw)), raw)


def lameinfo_parse_from_reader(data: bytes) -> LameInfo:
    """lameinfo.ParseFromReader (lameinfo.go:288-328) over bytes positioned at the first frame."""
    raw = LameInfoStruct()
    return _lame_result(host_lib().mp3_lameinfo_parse_from_reader(bytes(data), len(data), C.byref(raw)), raw)


def is_lame_version(s: bytes) -> bool:
    return bool(host_lib().mp3_lameinfo_is_lame_version(bytes(s), len(s)))
