/*
 * mp3_oracle.h — CPU restatement of llehouerou/go-mp3's decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the PCM-value level.  The reference ships
 * no golden PCM, no checksums and no Huffman known-answer vectors, and no Go
 * toolchain exists in this image, so the reference itself cannot be run.  What
 * IS pinned against the reference's own tests (see tests/test_oracle_*.py):
 *   internal/bits/bits_test.go:23-112        (bit reader values + OOB behaviour)
 *   internal/frameheader/frameheader_test.go (frame sizes, sync limit, resync, L1/L2 reject)
 *   internal/maindata/huffman_test.go:14-46  (region clamp does not error)
 *   trailing_tags_test.go:101-550            (Length()/ReadAll lengths, Seek, error kinds)
 *   fuzzing_test.go:22-107                   (crasher inputs do not crash)
 *   fixture invariants: classic_lame.mp3 -> 385 frames / 1,774,080 PCM bytes,
 *                       mpeg2.mp3 -> 2,872 frames / 6,617,088 PCM bytes @ 22,050 Hz
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).
 */
#ifndef MP3_ORACLE_H
#define MP3_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes (reference error values / types they stand for). */
enum {
    ORC_OK = 0,
    ORC_EOF = 1,                 /* io.EOF */
    ORC_ERR_UNEXPECTED_EOF = -1, /* *consts.UnexpectedEOFError     consts.go:17-23 */
    ORC_ERR_SYNC_LIMIT = -2,     /* *frameheader.SyncSearchLimitError frameheader.go:265-273 */
    ORC_ERR_FREE_FORMAT = -3,    /* frameheader.go:323-326 */
    ORC_ERR_MPEG25 = -4,         /* frame.go:79-81 */
    ORC_ERR_LAYER = -5,          /* frame.go:82-84 */
    ORC_ERR_FRAMESIZE = -6,      /* sideinfo.go:72-74, maindata.go:92-94 */
    ORC_ERR_MAINDATA_SIZE = -7,  /* maindata.go:291-293 */
    ORC_ERR_ISPOS = -8,          /* maindata/huffman.go:68-70 */
    ORC_ERR_REGION_INDEX = -9,   /* maindata/huffman.go:49-56 */
    ORC_ERR_HUFFMAN = -10,       /* huffman.go:382-386 (unreachable) */
    ORC_ERR_SEEK_UNSUPPORTED = -11, /* decode.go:291,323 ; source.go:31 */
    ORC_ERR_WHENCE = -12,        /* decode.go:104 */
    ORC_ERR_SAMPLE_RATE = -13,   /* frameheader.go:67-69 */
    ORC_ERR_REF_PANIC = -14,     /* reference would panic (Go runtime error), e.g. LSF mixed blocks: maindata.go:172-178 indexes scaleFactors[38] of a 38-element slice */
    ORC_ERR_INTERNAL = -99
};

/* ---- internal/bits ---------------------------------------------------- */
typedef struct {
    const uint8_t *vec;
    int len;
    int bit_pos;
    int byte_pos;
    int err; /* sticky ErrOutOfBounds (bits.go:20) */
} orc_bits;
void orc_bits_init(orc_bits *b, const uint8_t *vec, int len);
int orc_bits_bit(orc_bits *b);           /* bits.go:45-56 */
int orc_bits_bits(orc_bits *b, int num); /* bits.go:58-77 */
int orc_bits_pos(const orc_bits *b);     /* bits.go:79-81 */
void orc_bits_set_pos(orc_bits *b, int pos); /* bits.go:83-86 */

/* ---- internal/huffman -------------------------------------------------- */
/* One code word (huffman.go:348-419).  out = {x, y, v, w}. Returns ORC_OK or ORC_ERR_HUFFMAN. */
int orc_huffman_decode(orc_bits *m, int table_num, int out[4]);
/* Code-table introspection for LUT / encoder construction in tests and the
 * synthesiser check: number of symbols, linbits, and the i-th code word. */
int orc_huffman_table_info(int table_num, int *n_symbols, int *linbits);
int orc_huffman_table_code(int table_num, int i, int *x, int *y, int *hlen, uint32_t *hcod);

/* ---- internal/frameheader ---------------------------------------------- */
int orc_header_is_valid(uint32_t h);                  /* frameheader.go:168-189 */
int orc_header_frame_size(uint32_t h, int *size);     /* frameheader.go:223-232 */
int orc_header_bitrate(uint32_t h);                   /* frameheader.go:191-221 */
int orc_header_sampling_frequency_value(uint32_t h);  /* frameheader.go:56-71 (0 on error) */
int orc_header_side_info_size(uint32_t h);            /* frameheader.go:234-251 */
int orc_header_bytes_per_frame(uint32_t h);           /* frameheader.go:142-144 */
int orc_header_samples_per_frame(uint32_t h);         /* frameheader.go:153-155 */
int64_t orc_header_frame_duration_ns(uint32_t h);     /* frameheader.go:157-164 */
int orc_header_bytes_per_second(uint32_t h);          /* frameheader.go:166-174 */
/* frameheader.Read over a memory source (frameheader.go:279-328).
 * Returns ORC_OK and fills header, start_pos, new_pos; else an error code. */
int orc_frameheader_read_mem(const uint8_t *data, size_t len, size_t pos,
                             uint32_t *header, int64_t *start_pos, size_t *new_pos,
                             int64_t *bytes_searched);

/* ---- taps: every intermediate of the hot path, per decoded frame -------- */
typedef struct {
    int capacity_frames;      /* arrays below hold this many frames */
    int n_frames;             /* out: frames decoded */
    uint32_t *header;         /* [F] */
    int32_t *main_data_begin; /* [F] */
    int64_t *position;        /* [F] byte offset of the frame header */
    int16_t *is;              /* [F][2][2][576] Huffman integers (maindata/huffman.go) */
    int32_t *count1;          /* [F][2][2] */
    uint8_t *scalefac_l;      /* [F][2][2][22] */
    uint8_t *scalefac_s;      /* [F][2][2][39]  ([13][3]) */
    int32_t *part2_start;     /* [F][2][2] bit position in the frame's logical buffer */
    float *xr_requant;        /* [F][2][2][576] after requantize   frame.go:140-255 */
    float *xr_reorder;        /* after reorder                     frame.go:257-302 */
    float *xr_stereo;         /* after stereo                      frame.go:304-420 */
    float *xr_alias;          /* after antialias                   frame.go:422-452 */
    float *hybrid;            /* after hybridSynthesis+freqInversion frame.go:454-486, [sb*18+i] */
} orc_taps;

/* ---- package mp3: Decoder (decode.go) ---------------------------------- */
typedef struct orc_decoder orc_decoder;

/* mp3.NewDecoder over an in-memory reader (decode.go:361-388).  seekable=0 models a
 * plain io.Reader (Length() == -1, seeks fail). Returns NULL and sets *err on failure. */
orc_decoder *orc_new_decoder(const uint8_t *data, size_t len, int seekable, int *err);
void orc_free_decoder(orc_decoder *d);
void orc_set_taps(orc_decoder *d, orc_taps *taps); /* taps recorded for frames decoded after this call */
/* Decoder.Read (decode.go:70-80): returns bytes copied (>0) or 0 with *err = ORC_EOF / fatal code. */
long orc_read(orc_decoder *d, uint8_t *buf, size_t n, int *err);
/* io.ReadAll(d): allocates *out (free with orc_free). Returns total bytes; *err = ORC_OK on clean EOF. */
long orc_read_all(orc_decoder *d, uint8_t **out, int *err);
void orc_free(void *p);
int64_t orc_seek(orc_decoder *d, int64_t offset, int whence, int *err); /* decode.go:89-145 */
int orc_sample_rate(const orc_decoder *d);      /* decode.go:150-152 */
int64_t orc_length(const orc_decoder *d);       /* decode.go:224-226 */
int64_t orc_bytes_per_frame(const orc_decoder *d); /* decode.go:230-232 */
int64_t orc_duration_ns(const orc_decoder *d);  /* decode.go:236-241 */
int64_t orc_position_ns(const orc_decoder *d);  /* decode.go:244-246 */
int64_t orc_remaining_ns(const orc_decoder *d); /* decode.go:250-256 */
double orc_progress(const orc_decoder *d);      /* decode.go:260-268 */
int64_t orc_sample_position(const orc_decoder *d); /* decode.go:272-274 */
int64_t orc_sample_count(const orc_decoder *d);    /* decode.go:278-283 */
int orc_seek_to_sample(orc_decoder *d, int64_t sample); /* decode.go:288-307 */
int orc_skip(orc_decoder *d, int64_t delta_ns);         /* decode.go:313-315 */
int orc_seek_to_time(orc_decoder *d, int64_t t_ns);     /* decode.go:320-341 */
int orc_num_frame_starts(const orc_decoder *d);
int64_t orc_frame_start(const orc_decoder *d, int i);
const char *orc_error_string(int err);

/* ---- CPU baseline: bench_test.go:40-55 semantics, one thread per stream -- */
/* Decodes n streams (NewDecoder + io.ReadAll each) on `threads` pthreads.
 * pcm_bytes[i] receives the decoded length, checksum[i] a 64-bit FNV-1a of the PCM.
 * Returns wall seconds. */
double orc_decode_streams_mt(const uint8_t *const *data, const size_t *lens, int n, int threads,
                             int64_t *pcm_bytes, uint64_t *checksum);

/* Table accessors so tests can pin table generation (imdct.go:21-79, frame.go:488-628). */
const float *orc_table_imdct_win(void);   /* [4][36] */
const float *orc_table_cos_n12(void);     /* [6][12] */
const float *orc_table_cos_n36(void);     /* [18][36] */
const float *orc_table_synth_nwin(void);  /* [64][32] */
const float *orc_table_synth_dtbl(void);  /* [512] */
const double *orc_table_powtab34(void);   /* [8207] */

#ifdef __cplusplus
}
#endif
#endif
