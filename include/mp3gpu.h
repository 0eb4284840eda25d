/*
 * mp3gpu.h — C ABI of the B200 (sm_100a) MP3 Layer III granule decode engine.
 *
 * This is the drop-in boundary for go-mp3's hot path.  The reference has no FFI today
 * (pure Go); the seam this ABI replaces is internal:
 *
 *     Decoder.readFrame                       decode.go:45-67
 *       frame.Read(source, pos, prev)         internal/frame/frame.go:67-115   [stays on the host]
 *         maindata.Read -> scalefactors + readHuffman
 *                                             internal/maindata/maindata.go:119-288,
 *                                             internal/maindata/huffman.go:27-138   -> K1 (device)
 *       (*Frame).Decode() []byte              internal/frame/frame.go:121-138      -> K2..K4 (device)
 *
 * The host (Go via cgo in the product; the C++ mirror in go-mp3_b200/csrc/host here) keeps the
 * serial stream work: tag skipping, frame-header sync, side-info parsing and bit-reservoir
 * resolution (maindata.go:290-323).  It hands the device
 *   - `main_data`: for every stream, the concatenation of each frame's main-data bytes
 *     (header, CRC and side info stripped; ancillary bytes kept), streams back to back, and
 *   - one `mp3gpu_unit` per (granule, channel slot): the absolute bit position where that
 *     granule-channel's part2 (scalefactor) bits start, the end of the frame's logical
 *     reservoir buffer, and the side-info fields.
 * The device returns interleaved 16-bit stereo PCM, 576 stereo samples (2304 bytes) per
 * granule, in granule order — the byte stream Decoder.Read would have produced.
 *
 * Granule g owns unit slots 2g (channel 0) and 2g+1 (channel 1); a mono granule leaves slot
 * 2g+1 with MP3GPU_W2_VALID clear and its PCM duplicates channel 0 (frame.go:671-678).
 * A granule flagged MP3GPU_W2_ZERO_STATE starts from zeroed overlap/V state
 * (Frame.store / Frame.vVec, frame.go:48-49), i.e. the first granule of a stream or of a
 * Seek (decode.go:106-108).  All other cross-granule state is recomputed on the device
 * from the preceding granules of the same submission.
 *
 * All functions return 0 on success, a negative MP3GPU_E_* code on failure; the message is
 * available from mp3gpu_last_error().  There is no CPU fallback: without a CUDA device
 * mp3gpu_create fails.
 */
#ifndef MP3GPU_H
#define MP3GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP3GPU_ABI_VERSION 1

enum {
    MP3GPU_OK = 0,
    MP3GPU_E_NO_DEVICE = -1,
    MP3GPU_E_CUDA = -2,
    MP3GPU_E_INVALID = -3,
    MP3GPU_E_NOMEM = -4      /* a device or pinned allocation failed (the context stays usable) */
};

/* One (granule, channel) bit-slice + side info.  32 bytes.
 * Replaces the arguments of readHuffman (maindata/huffman.go:27) and the SideInfo fields
 * (sideinfo.go:33-55) the hot path reads. */
typedef struct mp3gpu_unit {
    uint64_t bit_start;   /* absolute bit index into main_data of this unit's first part2 bit
                             (= part2Start, maindata.go:133,202) */
    int32_t buf_end_rel;  /* (bit index one past the frame's logical buffer) - bit_start.
                             Reads at/after it return 0 and do not advance (bits.go:46-49,65-68).
                             May be <= 0: the cursor was parked beyond the buffer (bits.go:83-86). */
    uint32_t w0;          /* part2_3_length:12 | big_values:9 | global_gain:8 | win_switch:1 | block_type:2 */
    uint32_t w1;          /* scalefac_compress:9 | table_select[0]:5 | [1]:5 | [2]:5 |
                             region0_count:4 | region1_count:4 (the implicit value of window-switched
                             granules, 20 - region0 = 12 or 13, needs 4 bits: sideinfo.go:128-136) */
    uint32_t w2;          /* subblock_gain[0]:3 | [1]:3 | [2]:3 | preflag:1 | scalefac_scale:1 |
                             count1table_select:1 | scfsi:4 (band0 = lsb) | lsf:1 | sfreq:2 |
                             mode:2 | mode_ext:2 | gr:1 | ch:1 | valid:1 | zero_state:1 |
                             mixed_block_flag:1 */
    uint32_t reserved[2];
} mp3gpu_unit;

/* w0 */
#define MP3GPU_W0_P23LEN_SHIFT 0
#define MP3GPU_W0_BIGVAL_SHIFT 12
#define MP3GPU_W0_GGAIN_SHIFT 21
#define MP3GPU_W0_WINSW_SHIFT 29
#define MP3GPU_W0_BTYPE_SHIFT 30
/* w1 */
#define MP3GPU_W1_SFCOMP_SHIFT 0
#define MP3GPU_W1_TSEL0_SHIFT 9
#define MP3GPU_W1_TSEL1_SHIFT 14
#define MP3GPU_W1_TSEL2_SHIFT 19
#define MP3GPU_W1_REG0_SHIFT 24
#define MP3GPU_W1_REG1_SHIFT 28
/* w2 */
#define MP3GPU_W2_SBG0_SHIFT 0
#define MP3GPU_W2_SBG1_SHIFT 3
#define MP3GPU_W2_SBG2_SHIFT 6
#define MP3GPU_W2_PREFLAG_SHIFT 9
#define MP3GPU_W2_SFSCALE_SHIFT 10
#define MP3GPU_W2_C1TSEL_SHIFT 11
#define MP3GPU_W2_SCFSI_SHIFT 12
#define MP3GPU_W2_LSF_SHIFT 16
#define MP3GPU_W2_SFREQ_SHIFT 17
#define MP3GPU_W2_MODE_SHIFT 19
#define MP3GPU_W2_MODEEXT_SHIFT 21
#define MP3GPU_W2_GR_SHIFT 23
#define MP3GPU_W2_CH_SHIFT 24
#define MP3GPU_W2_VALID (1u << 25)
#define MP3GPU_W2_ZERO_STATE (1u << 26)
#define MP3GPU_W2_MIXED_SHIFT 27

#define MP3GPU_SAMPLES_PER_GRANULE 576
#define MP3GPU_PCM_BYTES_PER_GRANULE 2304 /* 576 stereo samples x 2 ch x int16 */

typedef struct mp3gpu_opts {
    uint32_t abi_version;       /* MP3GPU_ABI_VERSION */
    uint32_t wave_granules;     /* granules decoded per kernel wave (0 = default 2097152, capped at 16777216); bounds the workspace, which is sized on demand */
    uint32_t keep_intermediates;/* 1: keep per-stage buffers of the last wave readable via mp3gpu_debug_read */
    uint32_t reserved;
} mp3gpu_opts;

/* Per-kernel device time of the last decode call, from CUDA events on the launch stream. */
typedef struct mp3gpu_timings {
    float k1_huffman_ms;   /* k_huffman : scalefactors + Huffman                                   */
    float k_hybrid_ms;     /* k_hybrid  : requantise/reorder/stereo/alias + IMDCT/window/overlap   */
    float k_synth_ms;      /* k_synth   : polyphase synthesis + int16 clamp/interleave             */
    float reserved_ms;
    float total_ms;        /* first kernel start -> last kernel end (includes copies overlapped in between) */
    float h2d_ms, d2h_ms;  /* copy time on the copy streams (host-buffer calls only) */
    uint32_t waves;
    uint32_t launches;     /* kernel launches issued by the call */
} mp3gpu_timings;

typedef struct mp3gpu_ctx mp3gpu_ctx;

/* Create an engine on CUDA device `device` (tables uploaded, streams and workspace created). */
int mp3gpu_create(int device, const mp3gpu_opts *opts, mp3gpu_ctx **out);
void mp3gpu_destroy(mp3gpu_ctx *ctx);
const char *mp3gpu_last_error(const mp3gpu_ctx *ctx);

/* Decode `n_granules` granules (2*n_granules units).  All pointers are HOST memory; host<->device
 * copies happen inside the call, pipelined wave by wave (pinned memory from mp3gpu_host_alloc
 * makes them asynchronous).  pcm_out receives n_granules * 2304 bytes.
 * Replaces: maindata scalefactor/Huffman read + (*Frame).Decode() for every frame of the batch. */
int mp3gpu_decode(mp3gpu_ctx *ctx, const uint8_t *main_data, size_t main_data_len,
                  const mp3gpu_unit *units, size_t n_granules, int16_t *pcm_out);

/* Same, but pcm_out receives only granules [first_out, n_granules): the granules in front are decoded for the state they
 * leave behind (the halo of a frame-range job, mp3host.h: mp3_decode_frames) and their PCM is not copied back. */
int mp3gpu_decode_range(mp3gpu_ctx *ctx, const uint8_t *main_data, size_t main_data_len,
                        const mp3gpu_unit *units, size_t n_granules, size_t first_out, int16_t *pcm_out);

/* Same, with every pointer DEVICE-resident on the context's device (no copies).
 * Requirements on the caller's buffers (checked where they can be; MP3GPU_E_INVALID): d_main_data and d_units aligned to
 * 16 bytes, d_pcm_out to 4 bytes (cudaMalloc and whole torch tensors are; a sliced tensor may not be), and d_main_data
 * readable for 64 bytes past main_data_len (the kernels read whole 16-byte chunks and one word ahead; the bytes
 * themselves are never used). */
int mp3gpu_decode_device(mp3gpu_ctx *ctx, const uint8_t *d_main_data, size_t main_data_len,
                         const mp3gpu_unit *d_units, size_t n_granules, int16_t *d_pcm_out);

/* Same, but returns as soon as the kernels are queued on the context's compute stream (no synchronise).
 * Together with the two event calls below this lets a caller time K back-to-back passes on the device. */
int mp3gpu_decode_device_async(mp3gpu_ctx *ctx, const uint8_t *d_main_data, size_t main_data_len,
                               const mp3gpu_unit *d_units, size_t n_granules, int16_t *d_pcm_out);
/* Record user event `which` (0..7) on the compute stream; elapsed time between two recorded events
 * (synchronises on `to`). */
int mp3gpu_event_record(mp3gpu_ctx *ctx, int which);
int mp3gpu_event_elapsed_ms(mp3gpu_ctx *ctx, int from, int to, float *ms);

/* Pinned host memory (page-locked) for main_data / units / pcm buffers. */
void *mp3gpu_host_alloc(size_t bytes);
void mp3gpu_host_free(void *p);

/* Plain device memory + copies on the context's device, for callers that keep inputs resident. */
void *mp3gpu_device_alloc(mp3gpu_ctx *ctx, size_t bytes);
void mp3gpu_device_free(mp3gpu_ctx *ctx, void *p);
int mp3gpu_copy_to_device(mp3gpu_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);
int mp3gpu_copy_to_host(mp3gpu_ctx *ctx, void *dst_host, const void *src_device, size_t bytes);
int mp3gpu_synchronize(mp3gpu_ctx *ctx);

int mp3gpu_last_timings(mp3gpu_ctx *ctx, mp3gpu_timings *out);

/* Debug taps (opts.keep_intermediates = 1): per-stage outputs of the LAST wave of the last call,
 * for parity tests against the oracle.  `first_granule`/`n_granules` index within that wave. */
enum {
    MP3GPU_TAP_IS = 0,       /* int16  [gr][2][576]  Huffman integers, zero-filled above count1 */
    MP3GPU_TAP_COUNT1 = 1,   /* int32  [gr][2]       */
    MP3GPU_TAP_SCALEFAC = 2, /* uint8  [gr][2][64]   [0..21] scalefac_l, [22..60] scalefac_s[13][3], [61] preflag */
    MP3GPU_TAP_XR = 3,       /* float  [gr][2][576]  after requantise+reorder+stereo+antialias, index sb*18+i */
    MP3GPU_TAP_HYBRID = 4    /* float  [gr][2][576]  after IMDCT/overlap/frequency inversion, index sb*18+i */
};
int mp3gpu_debug_read(mp3gpu_ctx *ctx, int tap, size_t first_granule, size_t n_granules, void *host_out);

/* Device properties, for logs. */
int mp3gpu_device_info(mp3gpu_ctx *ctx, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor);

/* PCI bus id of the context's device ("0000:1b:00.0"), so a host can place its pinned buffers and threads on the
 * GPU's NUMA node (/sys/bus/pci/devices/<id>/numa_node). */
int mp3gpu_device_pci_bus_id(mp3gpu_ctx *ctx, char *out, size_t out_len);

/* Output side: hands decoded PCM to a consumer on the same GPU without crossing PCIe.  Converts n_samples stereo samples of
 * device-resident s16le interleaved PCM (what mp3gpu_decode_device leaves in HBM; the contract of Decoder.Read,
 * decode.go:356-360) into two float32 planes scaled by 1/32768, queued on the context's compute stream behind the decode
 * (no synchronise; use mp3gpu_synchronize or the event calls).  All three pointers 16-byte aligned. */
int mp3gpu_pcm_to_f32_planar(mp3gpu_ctx *ctx, const int16_t *d_pcm, size_t n_samples, float *d_left, float *d_right);

/* Copies `bytes` from a device scratch buffer to host_dst (pinned host memory) `reps` times with plain cudaMemcpyAsync
 * on the context's output stream and returns the device-timed seconds: the ceiling of the end-to-end path, whose cost
 * is the PCM going home over PCIe (4 bytes per stereo sample).  bench.py runs it on all devices of an engine at once. */
int mp3gpu_measure_d2h(mp3gpu_ctx *ctx, void *host_dst, size_t bytes, int reps, double *seconds);

/* Measures the FP32 FMA issue peak of the device with a register-resident FFMA loop (TFLOP/s).
 * Used by bench.py as the fp32 roofline denominator (MEASURED_PEAKS.json has no fp32 entry). */
int mp3gpu_measure_fp32_peak(mp3gpu_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
