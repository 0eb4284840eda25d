// stream_parser.h — the serial, host-side half of go-mp3's frame.Read: tag skipping, frame-header
// sync, side-info parsing and bit-reservoir resolution into independent per-granule bit-slices.
//
// Reference being mirrored (paths relative to the reference root):
//   source.go:42-122                          tag skipping, ReadFull/Unread semantics
//   internal/frameheader/frameheader.go:27-328 header fields, validity, frame size, resync
//   internal/sideinfo/sideinfo.go:66-156       side info
//   internal/maindata/maindata.go:85-117,290-323  main-data size, reservoir assembly
//   internal/frame/frame.go:56-115             frame.Read orchestration and error order
//
// Data model (SURVEY.md 8b'): per stream, M = concatenation of every frame's own main-data bytes.
// The reference's logical buffer of frame k is always a contiguous window of M that ends at the end
// of frame k's own bytes: Tail(prev, main_data_begin) ++ own  (maindata.go:310-322), or all of
// prev ++ own on reservoir underflow (maindata.go:295-308), or just own when prev == nil.  So a
// unit needs only an absolute bit position into M and the window end.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/mp3gpu.h"
#include "../../../include/mp3host.h"

namespace mp3host {

// ---- frame header accessors (frameheader.go:27-258) ---------------------------------------------
struct Header {
    uint32_t v = 0;
    int id() const { return (int)((v & 0x00180000u) >> 19); }
    int layer() const { return (int)((v & 0x00060000u) >> 17); }
    int protection_bit() const { return (int)((v & 0x00010000u) >> 16); }
    int bitrate_index() const { return (int)((v & 0x0000f000u) >> 12); }
    int sampling_frequency() const { return (int)((v & 0x00000c00u) >> 10); }
    int padding_bit() const { return (int)((v & 0x00000200u) >> 9); }
    int mode() const { return (int)((v & 0x000000c0u) >> 6); }
    int mode_extension() const { return (int)((v & 0x00000030u) >> 4); }
    int emphasis() const { return (int)(v & 3u); }
    int lsf() const { return id() == 3 ? 0 : 1; }
    int granules() const { return 2 >> lsf(); }
    int nch() const { return mode() == 3 ? 1 : 2; }
    int bytes_per_frame() const { return 576 * granules() * 4; }
    int sampling_frequency_value() const {
        switch (sampling_frequency()) {
        case 0: return 44100 >> lsf();
        case 1: return 48000 >> lsf();
        case 2: return 32000 >> lsf();
        }
        return 0;
    }
    bool is_valid() const {  // frameheader.go:168-189
        if ((v & 0xffe00000u) != 0xffe00000u) return false;
        if (id() == 1) return false;
        if (bitrate_index() == 15) return false;
        if (sampling_frequency() == 3) return false;
        if (layer() != 1) return false;
        if (emphasis() == 2) return false;
        return true;
    }
    int bitrate() const {  // frameheader.go:191-221 (Layer III rows only: is_valid() guarantees layer 3)
        static const int br[2][16] = {
            {0, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 160000, 192000, 224000, 256000, 320000, 0},
            {0, 8000, 16000, 24000, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 144000, 160000, 0}};
        return br[lsf()][bitrate_index()];
    }
    int frame_size() const { return ((144 * bitrate()) / sampling_frequency_value() + padding_bit()) >> lsf(); }  // :223-232
    int side_info_size() const {  // :234-251
        bool mono = mode() == 3;
        if (lsf()) return mono ? 9 : 17;
        return mono ? 17 : 32;
    }
};

// ---- in-memory source with the reference's ReadFull/Unread behaviour (source.go:94-122) ---------
struct Source {
    const uint8_t *data = nullptr;
    size_t len = 0;
    int64_t pos = 0;
    // Returns bytes available (<= n); advances by that many.
    int read_full(const uint8_t **p, int n) {
        int64_t avail = (int64_t)len - pos;
        if (avail < 0) avail = 0;
        int got = (int64_t)n <= avail ? n : (int)avail;
        *p = data + pos;
        pos += got;
        return got;
    }
    int skip_tags() {  // source.go:42-83
        for (;;) {
            const uint8_t *b;
            if (read_full(&b, 3) < 3) return MP3_EOF;
            if (b[0] == 'T' && b[1] == 'A' && b[2] == 'G') {
                if (read_full(&b, 125) < 125) return MP3_EOF;
            } else if (b[0] == 'I' && b[1] == 'D' && b[2] == '3') {
                if (read_full(&b, 3) < 3) return MP3_EOF;
                if (read_full(&b, 4) < 4) return MP3_EOF;
                uint32_t size = ((uint32_t)b[0] << 21) | ((uint32_t)b[1] << 14) | ((uint32_t)b[2] << 7) | (uint32_t)b[3];
                int64_t avail = (int64_t)len - pos;
                if ((int64_t)size > avail) {
                    pos = (int64_t)len;
                    return MP3_EOF;
                }
                pos += size;
            } else {
                pos -= 3;  // Unread
                return MP3_OK;
            }
        }
    }
};

// frameheader.Read (frameheader.go:279-328).
inline int read_frame_header(Source &s, Header *h, int64_t *start_pos) {
    const uint8_t *b;
    int64_t position = s.pos;
    int n = s.read_full(&b, 4);
    if (n < 4) return n == 0 ? MP3_EOF : MP3_ERR_UNEXPECTED_EOF;
    uint32_t v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    int64_t searched = 4;
    Header hh;
    hh.v = v;
    while (!hh.is_valid()) {
        if (searched >= 64 * 1024) return MP3_ERR_SYNC_LIMIT;
        if (s.read_full(&b, 1) < 1) return MP3_ERR_UNEXPECTED_EOF;
        hh.v = (hh.v << 8) | b[0];
        position++;
        searched++;
    }
    if (hh.bitrate_index() == 0) return MP3_ERR_FREE_FORMAT;
    *h = hh;
    *start_pos = position;
    return MP3_OK;
}

// ---- MSB-first reader for the 9/17/32 side-info bytes (bits.go:58-77; never out of bounds here) --
struct SideBits {  // MSB-first reader over a private, zero-padded copy of the side info (at most 32 bytes)
    uint8_t buf[40];
    int pos = 0;
    SideBits(const uint8_t *q, int n) {
        memset(buf, 0, sizeof buf);
        memcpy(buf, q, (size_t)(n < 32 ? n : 32));
    }
    int get(int n) {  // 1 <= n <= 25
        const uint8_t *b = buf + (pos >> 3);
        const uint32_t w = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
        const int v = (int)((w << (pos & 7)) >> (32 - n));
        pos += n;
        return v;
    }
};

struct GrCh {  // sideinfo.go:33-55, one (gr, ch)
    int part2_3_length, big_values, global_gain, scalefac_compress, win_switch, block_type, mixed;
    int table_select[3], subblock_gain[3], region0, region1, preflag, scalefac_scale, count1table;
};

// maindata.go:39-50
static const int kSlenMpeg1[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {3, 0}, {1, 1}, {1, 2}, {1, 3},
                                      {2, 1}, {2, 2}, {2, 3}, {3, 1}, {3, 2}, {3, 3}, {4, 2}, {4, 3}};
static const int kSfSizeMpeg2[3][6][4] = {
    {{6, 5, 5, 5}, {6, 5, 7, 3}, {11, 10, 0, 0}, {7, 7, 7, 0}, {6, 6, 6, 3}, {8, 8, 5, 0}},
    {{9, 9, 9, 9}, {9, 9, 12, 6}, {18, 18, 0, 0}, {12, 12, 12, 0}, {12, 9, 9, 6}, {15, 12, 9, 0}},
    {{6, 9, 9, 9}, {6, 9, 12, 6}, {15, 18, 0, 0}, {6, 15, 12, 0}, {6, 12, 9, 6}, {6, 18, 9, 0}}};

inline int nslen2_value(int sfc) {  // maindata.go:54-81, evaluated instead of tabulated
    if (sfc < 400) return (sfc / 80) | (((sfc / 16) % 5) << 3) | (((sfc / 4) % 4) << 6) | ((sfc % 4) << 9);
    if (sfc < 500) {
        int n = sfc - 400;
        return (n / 20) | (((n / 4) % 5) << 3) | ((n % 4) << 6) | (1 << 12);
    }
    int n = sfc - 500;  // 500..511
    return (n / 3) | ((n % 3) << 3) | (2 << 12) | (1 << 15);
}

// Bit cursor movement of the scalefactor reads of one unit, honouring Bits(n)'s rule that a read
// crossing the buffer end does not advance (bits.go:65-68).  Only needed when part2_3_length == 0
// (quirk Q1/Q6): otherwise readHuffman repositions the cursor to part2Start + part2_3_length.
struct ScalefacCursor {
    int64_t pos, total;
    void read(int n) {
        if (n > 0 && pos + n <= total) pos += n;
    }
};

// A fixed-capacity stand-in for std::vector over caller-owned memory (DecodeBatch parses straight into its pinned arenas).
// An append past the capacity is dropped and recorded; the caller sized the span from a bound and treats that as a bug.
template <class T>
struct SpanVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool overflow = false;
    size_t size() const { return n; }
    T *end() { return p + n; }
    T &operator[](size_t i) { return p[i]; }
    void insert(T *at, const T *b, const T *e) {  // only ever called with at == end()
        (void)at;
        const size_t k = (size_t)(e - b);
        if (n + k > cap) { overflow = true; return; }
        if (k) memcpy(p + n, b, k * sizeof(T));
        n += k;
    }
    void resize(size_t m) {
        if (m > cap) { overflow = true; return; }
        n = m;
    }
};

// One parsed stream: appended-to by StreamParser.
struct ParsedStream {
    std::vector<uint8_t> main_data;   // M
    std::vector<mp3gpu_unit> units;   // 2 per granule
    int64_t frames = 0;
    int sample_rate = 0;
    int status = MP3_OK;      // terminal status: MP3_OK = clean EOF-family end, < 0 fatal
    bool opened = false;      // NewDecoder would have succeeded
    void clear() {
        main_data.clear();
        units.clear();
        frames = 0;
        sample_rate = 0;
        status = MP3_OK;
        opened = false;
    }
};

// Incremental parser: next_frame() performs the host half of frame.Read for one frame and appends
// its main-data bytes and units.  State = the previous frame's logical-buffer window (the reservoir).
class StreamParser {
public:
    Source src;
    bool have_prev = false;   // prev != nil (frame.go:93-99)
    int64_t win_start = 0;    // previous frame's logical buffer = M[win_start, m_end)
    int64_t m_base = 0;       // absolute byte offset of out_main[0] in the (virtual) whole-stream M
    Header last_header;

    void reset_state() { have_prev = false; }  // Seek / fatal error: d.frame = nil (decode.go:47,108)

    // Returns MP3_OK and appends; MP3_EOF for io.EOF at a frame boundary; or a negative error.
    // `zero_state` marks the frame's first granule (first frame after reset).
    template <class MV, class UV>
    int next_frame(MV &M, UV &units, int64_t bit_base = 0) {
        Header h;
        int64_t fpos;
        int rc = read_frame_header(src, &h, &fpos);
        if (rc != MP3_OK) return rc;
        const uint8_t *b;
        if (h.protection_bit() == 0) {  // readCRC, frame.go:56-65,73-77
            if (src.read_full(&b, 2) < 2) return MP3_ERR_UNEXPECTED_EOF;
        }
        if (h.id() == 0) return MP3_ERR_MPEG25;  // frame.go:79-81
        if (h.layer() != 1) return MP3_ERR_LAYER;
        // sideinfo.Read (sideinfo.go:66-156)
        const int nch = h.nch();
        const int framesize = h.frame_size();
        if (framesize > 2000) return MP3_ERR_FRAMESIZE;
        const int si_size = h.side_info_size();
        if (src.read_full(&b, si_size) < si_size) return MP3_ERR_UNEXPECTED_EOF;
        SideBits sb(b, si_size);
        const bool mpeg1 = h.lsf() == 0;
        const int main_data_begin = sb.get(mpeg1 ? 9 : 8);
        if (h.mode() == 3) sb.get(mpeg1 ? 5 : 1); else sb.get(mpeg1 ? 3 : 2);
        int scfsi[2] = {0, 0};
        if (mpeg1)
            for (int ch = 0; ch < nch; ch++)
                for (int band = 0; band < 4; band++) scfsi[ch] |= sb.get(1) << band;
        GrCh gc[2][2];
        memset(gc, 0, sizeof gc);
        const int ngr = h.granules();
        for (int gr = 0; gr < ngr; gr++)
            for (int ch = 0; ch < nch; ch++) {
                GrCh &g = gc[gr][ch];
                g.part2_3_length = sb.get(12);
                g.big_values = sb.get(9);
                g.global_gain = sb.get(8);
                g.scalefac_compress = sb.get(mpeg1 ? 4 : 9);
                g.win_switch = sb.get(1);
                if (g.win_switch == 1) {
                    g.block_type = sb.get(2);
                    g.mixed = sb.get(1);
                    for (int r = 0; r < 2; r++) g.table_select[r] = sb.get(5);
                    for (int w = 0; w < 3; w++) g.subblock_gain[w] = sb.get(3);
                    g.region0 = (g.block_type == 2 && g.mixed == 0) ? 8 : 7;  // sideinfo.go:128-136
                    g.region1 = 20 - g.region0;
                } else {
                    for (int r = 0; r < 3; r++) g.table_select[r] = sb.get(5);
                    g.region0 = sb.get(4);
                    g.region1 = sb.get(3);
                    g.block_type = 0;
                }
                if (mpeg1) g.preflag = sb.get(1);
                g.scalefac_scale = sb.get(1);
                g.count1table = sb.get(1);
            }
        // maindata.Read (maindata.go:85-117)
        int md_size = framesize - si_size - 4;
        if (h.protection_bit() == 0) md_size -= 2;
        if (md_size > 1500) return MP3_ERR_MAINDATA_SIZE;
        if (md_size < 0) return MP3_ERR_UNEXPECTED_EOF;
        const int64_t m_end = m_base + (int64_t)M.size();  // == end of previous window
        int64_t new_start;
        if (have_prev && main_data_begin > (int)(m_end - win_start)) {
            new_start = win_start;  // underflow: bits.Append(prev, buf), parse from bit 0 (maindata.go:295-308)
        } else {
            new_start = have_prev ? m_end - main_data_begin : m_end;
        }
        if (src.read_full(&b, md_size) < md_size) return MP3_ERR_UNEXPECTED_EOF;
        M.insert(M.end(), b, b + md_size);
        const int64_t new_end = m_end + md_size;
        // Units.  The bit cursor starts at bit 0 of the window (maindata.go:133,202).
        const int64_t total_bits = (new_end - new_start) * 8;
        int64_t cursor = 0;
        int err = MP3_OK;
        const size_t u0 = units.size();
        units.resize(u0 + (size_t)ngr * 2);
        if (units.size() != u0 + (size_t)ngr * 2) return MP3_ERR_INVALID;  // a SpanVec that is full (never: the caller's bound)
        memset(&units[u0], 0, sizeof(mp3gpu_unit) * (size_t)ngr * 2);
        for (int gr = 0; gr < ngr && err == MP3_OK; gr++)
            for (int ch = 0; ch < nch; ch++) {
                const GrCh &g = gc[gr][ch];
                mp3gpu_unit &u = units[u0 + (size_t)gr * 2 + ch];
                u.bit_start = (uint64_t)(bit_base + new_start * 8 + cursor);
                u.buf_end_rel = (int32_t)(total_bits - cursor);
                u.w0 = (uint32_t)g.part2_3_length | ((uint32_t)g.big_values << 12) | ((uint32_t)g.global_gain << 21) |
                       ((uint32_t)g.win_switch << 29) | ((uint32_t)g.block_type << 30);
                u.w1 = (uint32_t)g.scalefac_compress | ((uint32_t)g.table_select[0] << 9) | ((uint32_t)g.table_select[1] << 14) |
                       ((uint32_t)g.table_select[2] << 19) | ((uint32_t)g.region0 << 24) | ((uint32_t)g.region1 << 28);
                u.w2 = (uint32_t)g.subblock_gain[0] | ((uint32_t)g.subblock_gain[1] << 3) | ((uint32_t)g.subblock_gain[2] << 6) |
                       ((uint32_t)g.preflag << 9) | ((uint32_t)g.scalefac_scale << 10) | ((uint32_t)g.count1table << 11) |
                       ((uint32_t)scfsi[ch] << 12) | ((uint32_t)h.lsf() << 16) | ((uint32_t)h.sampling_frequency() << 17) |
                       ((uint32_t)h.mode() << 19) | ((uint32_t)h.mode_extension() << 21) | ((uint32_t)gr << 23) |
                       ((uint32_t)ch << 24) | MP3GPU_W2_VALID | ((uint32_t)g.mixed << MP3GPU_W2_MIXED_SHIFT);
                if (!have_prev && gr == 0) u.w2 |= MP3GPU_W2_ZERO_STATE;
                // Errors the reference raises while reading this unit, in its order.
                if (!mpeg1 && g.block_type == 2 && g.mixed != 0) { err = MP3_ERR_REF_PANIC; break; }  // maindata.go:172-178
                if (g.part2_3_length != 0 && g.big_values * 2 > 576) { err = MP3_ERR_ISPOS; break; }  // huffman.go:66-70
                if (g.part2_3_length != 0) {
                    cursor += g.part2_3_length;  // SetPos(bitPosEnd + 1), huffman.go:136
                } else {
                    // Q1/Q6: cursor stays where the scalefactor reads left it.
                    ScalefacCursor sc{cursor, total_bits};
                    if (mpeg1) {
                        const int slen1 = kSlenMpeg1[g.scalefac_compress][0], slen2 = kSlenMpeg1[g.scalefac_compress][1];
                        if (g.win_switch == 1 && g.block_type == 2) {
                            int sfb0 = 0;
                            if (g.mixed) {
                                for (int i = 0; i < 8; i++) sc.read(slen1);
                                sfb0 = 3;
                            }
                            for (int sfb = sfb0; sfb < 12; sfb++)
                                for (int w = 0; w < 3; w++) sc.read(sfb < 6 ? slen1 : slen2);
                        } else {
                            static const int cnt[4] = {6, 5, 5, 5};
                            for (int band = 0; band < 4; band++)
                                if (((scfsi[ch] >> band) & 1) == 0 || gr == 0)
                                    for (int i = 0; i < cnt[band]; i++) sc.read(band < 2 ? slen1 : slen2);
                        }
                    } else {
                        int slen = nslen2_value(g.scalefac_compress);
                        int n = 0;
                        if (g.block_type == 2) { n++; if (g.mixed) n++; }
                        int d = (slen >> 12) & 7;
                        for (int i = 0; i < 4; i++) {
                            int num = slen & 7;
                            slen >>= 3;
                            if (num > 0)
                                for (int k = 0; k < kSfSizeMpeg2[n][d][i]; k++) sc.read(num);
                        }
                    }
                    cursor = sc.pos;
                }
            }
        if (err != MP3_OK) {
            units.resize(u0);  // the frame is not decoded (frame.Read returned an error)
            // its main-data bytes were consumed from the source; keep M consistent for a caller that continues
            return err;
        }
        have_prev = true;
        win_start = new_start;
        last_header = h;
        return MP3_OK;
    }
};

// Whole-stream parse with the reference's open/read loop semantics:
// NewDecoder = skipTags + first readFrame (decode.go:361-376); then Read until EOF (io.ReadAll).
struct StreamMeta {
    int64_t frames = 0;
    int sample_rate = 0;
    int status = MP3_OK;   // as ParsedStream::status
    bool opened = false;
};
template <class MV, class UV>
inline void parse_whole_stream_into(const uint8_t *data, size_t len, MV &M, UV &units, StreamMeta &out, int64_t bit_base = 0) {
    out = StreamMeta();
    StreamParser p;
    p.src.data = data;
    p.src.len = len;
    int rc = p.src.skip_tags();
    if (rc != MP3_OK) {
        out.status = rc == MP3_EOF ? MP3_EOF : rc;  // NewDecoder returns the error (io.EOF for short input)
        return;
    }
    for (;;) {
        rc = p.next_frame(M, units, bit_base);
        if (rc != MP3_OK) break;
        if (out.frames == 0) {
            out.sample_rate = p.last_header.sampling_frequency_value();
            out.opened = true;
        }
        out.frames++;
    }
    // readFrame maps the EOF family to io.EOF (decode.go:48-63)
    bool eof_family = rc == MP3_EOF || rc == MP3_ERR_UNEXPECTED_EOF || rc == MP3_ERR_SYNC_LIMIT;
    if (!out.opened)
        out.status = eof_family ? MP3_EOF : rc;  // NewDecoder fails with io.EOF / the error
    else
        out.status = eof_family ? MP3_OK : rc;   // io.ReadAll: nil on EOF, else the error (PCM so far is kept)
}
inline void parse_whole_stream(const uint8_t *data, size_t len, ParsedStream &out, int64_t bit_base = 0) {
    out.clear();
    StreamMeta m;
    parse_whole_stream_into(data, len, out.main_data, out.units, m, bit_base);
    out.frames = m.frames;
    out.sample_rate = m.sample_rate;
    out.status = m.status;
    out.opened = m.opened;
}

}  // namespace mp3host
