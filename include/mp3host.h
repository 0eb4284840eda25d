/*
 * mp3host.h — C API of the host-side mirror of go-mp3's public package, above the mp3gpu C ABI.
 *
 * No Go toolchain exists in the build image, so the host layer that the product would write in Go
 * (go-mp3_b200/go/mp3) is mirrored here in C++ with the same names, argument
 * meaning and error behaviour as the reference's `package mp3`:
 *
 *   mp3.NewDecoder(r)            decode.go:361-388   -> mp3_new_decoder
 *   (*Decoder).Read              decode.go:70-80     -> mp3_decoder_read
 *   (*Decoder).Seek              decode.go:89-145    -> mp3_decoder_seek
 *   SampleRate/Length/BytesPerFrame/Duration/Position/Remaining/Progress/
 *   SamplePosition/SampleCount/SeekToSample/Skip/SeekToTime   decode.go:150-341
 *   (new) DecodeBatch            north star          -> mp3_decode_batch
 *
 * The serial stream work stays on the host (tag skipping source.go:42-83, header sync
 * frameheader.go:279-328, side info sideinfo.go:66-156, bit-reservoir resolution
 * maindata.go:290-323); everything from "bit-slice" to PCM runs on the GPU through mp3gpu.h.
 * There is no CPU decode path: without a CUDA device mp3_engine_create fails.
 */
#ifndef MP3HOST_H
#define MP3HOST_H

#include <stddef.h>
#include <stdint.h>

#include "mp3gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Status / error codes.  > 0: end of stream; < 0: the reference's error values. */
enum {
    MP3_OK = 0,
    MP3_EOF = 1,                   /* io.EOF */
    MP3_ERR_UNEXPECTED_EOF = -1,   /* *consts.UnexpectedEOFError (mapped to io.EOF by readFrame, decode.go:52-56) */
    MP3_ERR_SYNC_LIMIT = -2,       /* *frameheader.SyncSearchLimitError (mapped to io.EOF, decode.go:59-62) */
    MP3_ERR_FREE_FORMAT = -3,      /* "mp3: free bitrate format is not supported..." frameheader.go:323-326 */
    MP3_ERR_MPEG25 = -4,           /* "mp3: MPEG version 2.5 is not supported" frame.go:79-81 */
    MP3_ERR_LAYER = -5,            /* "mp3: only layer3 ... is supported" frame.go:82-84 */
    MP3_ERR_FRAMESIZE = -6,        /* "mp3: framesize = %d" sideinfo.go:72-74 */
    MP3_ERR_MAINDATA_SIZE = -7,    /* "mp3: size = %d" maindata.go:291-293 */
    MP3_ERR_ISPOS = -8,            /* "mp3: isPos was too big: %d" maindata/huffman.go:68-70 */
    MP3_ERR_SEEK_UNSUPPORTED = -11,/* "mp3: seek not supported on non-seekable source" decode.go:291,323 */
    MP3_ERR_WHENCE = -12,          /* "mp3: invalid whence" decode.go:104 */
    MP3_ERR_REF_PANIC = -14,       /* input on which the reference panics (LSF mixed blocks, maindata.go:172-178) */
    MP3_ERR_NO_XING_HEADER = -20,  /* lameinfo.ErrNoXingHeader "lameinfo: no Xing/Info header found" lameinfo.go:111 */
    MP3_ERR_DEVICE = -50,          /* CUDA / engine failure (message from mp3_engine_last_error) */
    MP3_ERR_INVALID = -51
};

typedef struct mp3_engine mp3_engine;   /* one or more GPUs + host thread pool + pinned arenas */
typedef struct mp3_decoder mp3_decoder; /* mirrors *mp3.Decoder */

#define MP3_MAX_DEVICES 16
typedef struct mp3_engine_opts {
    int device;              /* CUDA device ordinal (used when n_devices == 0) */
    int host_threads;        /* stream-parsing threads for DecodeBatch (0 = hardware concurrency), shared by the devices */
    uint32_t wave_granules;  /* passed to mp3gpu_opts (0 = default) */
    uint32_t chunk_frames;   /* Decoder decode-ahead per GPU call (0 = default 256) */
    uint32_t keep_intermediates; /* passed to mp3gpu_opts */
    uint32_t use_exact_library;  /* 1: load libmp3gpu_exact.so (no FMA contraction), 2: libmp3gpu_checked.so (every input-dependent
                                    global access guarded; for tests) instead of libmp3gpu.so */
    /* Multi-GPU: work is partitioned by stream (DecodeBatch) or by frame range of one stream
     * (mp3_decode_stream_split); streams share nothing, so there is no collective — each device gets its own device
     * engine, worker thread and region of the pinned arenas.  n_devices == 0 means the single device `device`. */
    int n_devices;
    int devices[MP3_MAX_DEVICES];
    uint32_t trim_gapless;   /* DecodeBatch: report each stream's PCM without the LAME encoder delay / padding
                                (lameinfo TotalDelay / TotalPadding, as the reference's README.md:110-195 example does) */
    uint32_t reserved;
} mp3_engine_opts;

int mp3_engine_create(const mp3_engine_opts *opts, mp3_engine **out);
void mp3_engine_destroy(mp3_engine *e);
const char *mp3_engine_last_error(const mp3_engine *e);
mp3gpu_ctx *mp3_engine_gpu(mp3_engine *e); /* the device engine of slot 0 (for taps / timings) */
int mp3_engine_device_count(const mp3_engine *e);
mp3gpu_ctx *mp3_engine_gpu_at(mp3_engine *e, int slot);
const char *mp3_error_string(int code);    /* the reference's error message for a status code */

/* ---- Decoder: drop-in for *mp3.Decoder over an in-memory source ------------------------- */
/* seekable = 0 models a plain io.Reader: Length() = -1 and the Seek* methods fail. The data
 * must stay valid for the decoder's lifetime. On failure returns NULL and sets *err. */
mp3_decoder *mp3_new_decoder(mp3_engine *e, const uint8_t *data, size_t len, int seekable, int *err);
/* Same, decoding on device slot `slot` of a multi-device engine (decoders on different slots may be used concurrently). */
mp3_decoder *mp3_new_decoder_on(mp3_engine *e, int slot, const uint8_t *data, size_t len, int seekable, int *err);
void mp3_decoder_free(mp3_decoder *d);
/* Read: copies up to n bytes; returns the count (> 0) or 0 with *err = MP3_EOF or a fatal code. */
long mp3_decoder_read(mp3_decoder *d, uint8_t *buf, size_t n, int *err);
int64_t mp3_decoder_seek(mp3_decoder *d, int64_t offset, int whence, int *err);
int mp3_decoder_sample_rate(const mp3_decoder *d);
int64_t mp3_decoder_length(const mp3_decoder *d);
int64_t mp3_decoder_bytes_per_frame(const mp3_decoder *d);
int64_t mp3_decoder_duration_ns(const mp3_decoder *d);
int64_t mp3_decoder_position_ns(const mp3_decoder *d);
int64_t mp3_decoder_remaining_ns(const mp3_decoder *d);
double mp3_decoder_progress(const mp3_decoder *d);
int64_t mp3_decoder_sample_position(const mp3_decoder *d);
int64_t mp3_decoder_sample_count(const mp3_decoder *d);
int mp3_decoder_seek_to_sample(mp3_decoder *d, int64_t sample);
int mp3_decoder_skip(mp3_decoder *d, int64_t delta_ns);
int mp3_decoder_seek_to_time(mp3_decoder *d, int64_t t_ns);

/* ---- DecodeBatch: many independent streams in one call ------------------------------------ */
typedef struct mp3_stream_result {
    int64_t pcm_offset;  /* byte offset of this stream's PCM in the batch PCM buffer */
    int64_t pcm_bytes;   /* what io.ReadAll(NewDecoder(stream)) returns */
    int32_t sample_rate; /* from the first frame (decode.go:377-381); 0 if the stream failed to open */
    int32_t status;      /* MP3_OK: clean end of stream; < 0: NewDecoder/Read error of the reference */
    int64_t frames;      /* frames decoded */
} mp3_stream_result;

typedef struct mp3_batch_timings {
    double parse_s, gather_s, device_s, total_s;  /* gather_s: 0 since the parse writes straight into the arenas */
    uint64_t main_data_bytes, n_granules, pcm_bytes;
} mp3_batch_timings;

/* Decodes n streams on all devices of the engine (chunks of streams are pulled by the devices from a common queue; which
 * device decodes a stream does not change its PCM).  PCM lands in an engine-owned pinned host buffer that stays valid
 * until the next mp3_decode_batch / mp3_decode_stream_split / mp3_engine_destroy on this engine; *pcm_base receives its
 * address.  Stream i's PCM is results[i].pcm_bytes bytes at results[i].pcm_offset, in stream order; for well-formed
 * streams the buffer is dense, a stream that ends early (truncated, garbage) leaves a hole behind its PCM that no
 * result covers.  One call at a time per engine. */
int mp3_decode_batch(mp3_engine *e, const uint8_t *const *data, const size_t *lens, size_t n,
                     mp3_stream_result *results, const uint8_t **pcm_base, mp3_batch_timings *timings);

/* ---- One long stream: frame-range decode and the split over the engine's devices (BASELINE.json configs[4]) ------ */
/* Frame index of a stream (the reference builds the same table for Seek: decode.go:154-216): offsets of all frame
 * headers behind the tags, from a header-only walk.  `data` must outlive the index. */
typedef struct mp3_stream_index mp3_stream_index;
int mp3_stream_index_create(const uint8_t *data, size_t len, mp3_stream_index **out);
void mp3_stream_index_free(mp3_stream_index *idx);
int64_t mp3_stream_index_frames(const mp3_stream_index *idx);
int mp3_stream_index_sample_rate(const mp3_stream_index *idx);
int64_t mp3_stream_index_pcm_bytes(const mp3_stream_index *idx, int64_t f0, int64_t f1); /* PCM bytes of frames [f0, f1) */
int64_t mp3_stream_index_frame_pos(const mp3_stream_index *idx, int64_t f); /* byte offset of frame f's header; -1 if out of range */

/* PCM of frames [f0, f1) of the stream, byte-identical to that stretch of a linear decode of the whole stream, decoded
 * on their own on device slot `slot`.  The library re-creates the state a linear decode has at f0: it parses a lead-in of
 * earlier frames so that the bit reservoir (main_data_begin reaches up to 511 bytes back, maindata.go:290-323) resolves
 * exactly, and decodes a halo of whole frames covering the two granules in front of f0 (IMDCT overlap and 15 slots of
 * synthesis history, frame.go:473-476,637-653) whose PCM is dropped.  pcm_out (host memory; pinned makes the copy
 * asynchronous) receives mp3_stream_index_pcm_bytes(idx, f0, f1) bytes unless a frame fails to parse; *pcm_bytes = bytes
 * written.  Calls on different slots may run concurrently. */
int mp3_decode_frames(mp3_engine *e, int slot, const mp3_stream_index *idx, int64_t f0, int64_t f1, uint8_t *pcm_out,
                      int64_t *pcm_bytes);

/* The whole stream cut into one contiguous frame range per device of the engine, the ranges decoded concurrently (one
 * host thread per device); the PCM lands, in stream order, in the engine's pinned buffer (valid until the next
 * mp3_decode_batch / mp3_decode_stream_split on this engine).  No collective: ranges share nothing once the host has
 * resolved the reservoir. */
int mp3_decode_stream_split(mp3_engine *e, const mp3_stream_index *idx, const uint8_t **pcm_base, int64_t *pcm_bytes,
                            mp3_batch_timings *timings);

/* ---- lameinfo: mirror of package lameinfo (lameinfo/lameinfo.go) ----------------------------------------------- */
#define MP3_LAME_FLAG_FRAME_COUNT 0x0001 /* lameinfo.go:52-57 */
#define MP3_LAME_FLAG_BYTE_COUNT 0x0002
#define MP3_LAME_FLAG_TOC 0x0004
#define MP3_LAME_FLAG_VBR_SCALE 0x0008
#define MP3_LAME_DECODER_DELAY 529       /* lameinfo.DecoderDelay, lameinfo.go:86 */
typedef struct mp3_lame_info {           /* lameinfo.Info, lameinfo.go:20-49 */
    int32_t is_xing;          /* tag was "Xing" (VBR) rather than "Info" (CBR) */
    uint32_t flags;
    uint32_t frame_count;     /* valid if flags & MP3_LAME_FLAG_FRAME_COUNT */
    uint32_t byte_count;      /* valid if flags & MP3_LAME_FLAG_BYTE_COUNT */
    uint8_t toc[100];         /* valid if flags & MP3_LAME_FLAG_TOC */
    uint32_t vbr_scale;       /* valid if flags & MP3_LAME_FLAG_VBR_SCALE */
    int32_t has_lame_info;    /* HasLAMEInfo(): LAMEVersion != "" */
    char lame_version[12];    /* the 9 bytes of the version field as they are (may contain NULs), zero padded */
    uint16_t encoder_delay;   /* valid if has_lame_info */
    uint16_t encoder_padding;
} mp3_lame_info;
/* lameinfo.Parse (lameinfo.go:139-270): `frame` is the complete first MP3 frame.  MP3_OK or MP3_ERR_NO_XING_HEADER. */
int mp3_lameinfo_parse(const uint8_t *frame, size_t len, mp3_lame_info *out);
/* lameinfo.ParseFromReader (lameinfo.go:288-328) over a reader positioned at `data`: reads the header, sizes the frame,
 * reads it, parses it.  Also returns MP3_EOF / MP3_ERR_UNEXPECTED_EOF where io.ReadFull would. */
int mp3_lameinfo_parse_from_reader(const uint8_t *data, size_t len, mp3_lame_info *out);
int mp3_lameinfo_total_delay(const mp3_lame_info *info);   /* TotalDelay(), lameinfo.go:88-93 */
int mp3_lameinfo_total_padding(const mp3_lame_info *info); /* TotalPadding(), lameinfo.go:97-108 */
int mp3_lameinfo_is_lame_version(const uint8_t *s, size_t n); /* isLAMEVersion, lameinfo.go:273-282 (exported for its test) */
/* Coarse seek without a frame index (not in the reference, which parses the TOC and never uses it): byte offset from the
 * first audio frame at which `fraction` (0..1) of the playing time has passed, interpolated in the 100-entry TOC;
 * `stream_bytes` is used when the tag carries no byte count.  -1 if the tag has no TOC. */
int64_t mp3_lameinfo_toc_offset(const mp3_lame_info *info, double fraction, uint64_t stream_bytes);

/* Test hook: upper bound of the unit slots (2 per granule) mp3_parse_streams / mp3_decode_batch produce for one
 * stream, from a header-only frame walk; DecodeBatch sizes its pinned arenas with it. */
size_t mp3_debug_unit_slots_upper_bound(const uint8_t *data, size_t len);
/* The same walk's bound of the stream's main-data bytes (DecodeBatch parses every stream straight into its arena slot). */
size_t mp3_debug_main_bytes_upper_bound(const uint8_t *data, size_t len);

/* Host-only stage of DecodeBatch (no GPU): parse + reservoir resolution into caller-visible arrays.
 * Used by tests (host logic) and by bench.py to stage device-resident inputs.  Buffers are owned by
 * the engine-independent parse result; free with mp3_parsed_free. */
typedef struct mp3_parsed {
    uint8_t *main_data;     /* padded by 64 zero bytes */
    size_t main_data_len;
    mp3gpu_unit *units;     /* 2 * n_granules */
    size_t n_granules;
    mp3_stream_result *streams; /* pcm_offset/pcm_bytes/sample_rate/status/frames per stream */
    size_t n_streams;
} mp3_parsed;
int mp3_parse_streams(const uint8_t *const *data, const size_t *lens, size_t n, int host_threads, mp3_parsed **out);
void mp3_parsed_free(mp3_parsed *p);

#ifdef __cplusplus
}
#endif
#endif
