// lameinfo.h — host-side mirror of go-mp3's `package lameinfo` (LAME / Xing / Info tag of the first frame).
//
// Mirrors lameinfo/lameinfo.go: Info and its Has*/Total* helpers (:20-111), Parse (:139-270), isLAMEVersion (:273-282),
// ParseFromReader (:288-328) and the frame-size arithmetic it uses (:331-386).  Pure metadata: the decoder never
// consults it (quirk Q13: the tag frame is decoded as an ordinary, silent frame); DecodeBatch can use it to trim the
// encoder delay and padding the way the reference's README example does (README.md:110-195).
#pragma once
#include <cstdint>
#include <cstring>

#include "../../../include/mp3host.h"

namespace mp3host {

inline uint32_t be32_at(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

// lameinfo.go:273-282: "LAME", "L3.9" (older format), "Gogo", "GOGO"
inline bool is_lame_version(const uint8_t *s, size_t n) {
    if (n < 4) return false;
    return memcmp(s, "LAME", 4) == 0 || memcmp(s, "L3.9", 4) == 0 || memcmp(s, "Gogo", 4) == 0 || memcmp(s, "GOGO", 4) == 0;
}

// lameinfo.go:139-270.  Returns MP3_OK or MP3_ERR_NO_XING_HEADER; *out is zeroed first.
inline int lameinfo_parse(const uint8_t *frame, size_t len, mp3_lame_info *out) {
    memset(out, 0, sizeof *out);
    if (len < 4) return MP3_ERR_NO_XING_HEADER;
    const uint32_t header = be32_at(frame);
    if ((header & 0xFFE00000u) != 0xFFE00000u) return MP3_ERR_NO_XING_HEADER;  // 11 sync bits
    const uint32_t version_bits = (header >> 19) & 3u;
    if (version_bits == 1) return MP3_ERR_NO_XING_HEADER;  // reserved
    const bool mpeg1 = version_bits == 3;                    // 0 = MPEG 2.5, 2 = MPEG 2: both use the LSF side-info sizes
    const bool mono = ((header >> 6) & 3u) == 3;
    const size_t side_info = mpeg1 ? (mono ? 17 : 32) : (mono ? 9 : 17);  // lameinfo.go:118-130
    size_t pos = 4 + side_info;
    if (len < pos + 4) return MP3_ERR_NO_XING_HEADER;
    const bool xing = memcmp(frame + pos, "Xing", 4) == 0;
    if (!xing && memcmp(frame + pos, "Info", 4) != 0) return MP3_ERR_NO_XING_HEADER;
    out->is_xing = xing ? 1 : 0;
    pos += 4;
    if (len < pos + 4) return MP3_ERR_NO_XING_HEADER;
    out->flags = be32_at(frame + pos);
    pos += 4;
    if (out->flags & MP3_LAME_FLAG_FRAME_COUNT) {
        if (len < pos + 4) return MP3_ERR_NO_XING_HEADER;
        out->frame_count = be32_at(frame + pos);
        pos += 4;
    }
    if (out->flags & MP3_LAME_FLAG_BYTE_COUNT) {
        if (len < pos + 4) return MP3_ERR_NO_XING_HEADER;
        out->byte_count = be32_at(frame + pos);
        pos += 4;
    }
    if (out->flags & MP3_LAME_FLAG_TOC) {
        if (len < pos + 100) return MP3_ERR_NO_XING_HEADER;
        memcpy(out->toc, frame + pos, 100);
        pos += 100;
    }
    if (out->flags & MP3_LAME_FLAG_VBR_SCALE) {
        if (len < pos + 4) return MP3_ERR_NO_XING_HEADER;
        out->vbr_scale = be32_at(frame + pos);
        pos += 4;
    }
    // the 9-byte encoder version string, then 12 bytes of LAME fields, then 12 + 12 bits of delay / padding (:231-266)
    if (len >= pos + 9 && is_lame_version(frame + pos, 9)) {
        memcpy(out->lame_version, frame + pos, 9);
        out->has_lame_info = 1;
        pos += 9;
        const size_t d = pos + 12;
        if (len >= d + 3) {
            out->encoder_delay = (uint16_t)(((uint16_t)frame[d] << 4) | ((uint16_t)frame[d + 1] >> 4));
            out->encoder_padding = (uint16_t)(((uint16_t)(frame[d + 1] & 0x0F) << 8) | (uint16_t)frame[d + 2]);
        }
    }
    return MP3_OK;
}

// lameinfo.go:331-386: the size of the first frame from its header, 0 if the header names no bitrate / sampling rate.
inline int lameinfo_frame_size(uint32_t version_bits, uint32_t layer_bits, uint32_t bitrate_index, uint32_t srate_index, uint32_t padding) {
    static const int kLsf[3][16] = {{0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0},        // layer III
                                    {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0},        // layer II
                                    {0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256, 0}};  // layer I
    static const int kMpeg1[3][16] = {{0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0},
                                      {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 0},
                                      {0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448, 0}};
    static const int kRate[4][4] = {{11025, 12000, 8000, 0}, {0, 0, 0, 0}, {22050, 24000, 16000, 0}, {44100, 48000, 32000, 0}};
    if (layer_bits < 1 || layer_bits > 3 || version_bits == 1) return 0;
    const int bitrate = (version_bits == 3 ? kMpeg1 : kLsf)[layer_bits - 1][bitrate_index & 15] * 1000;
    const int rate = kRate[version_bits & 3][srate_index & 3];
    if (bitrate == 0 || rate == 0) return 0;
    if (layer_bits == 3) return (12 * bitrate / rate + (int)padding) * 4;  // layer I
    return (version_bits == 3 ? 144 : 72) * bitrate / rate + (int)padding;
}

// lameinfo.go:288-328 over an in-memory reader positioned at `data` (the start of an MP3 frame).  io.ReadFull's errors
// are kept: nothing to read -> MP3_EOF (io.EOF), a short read -> MP3_ERR_UNEXPECTED_EOF (io.ErrUnexpectedEOF).
inline int lameinfo_parse_from_reader(const uint8_t *data, size_t len, mp3_lame_info *out) {
    memset(out, 0, sizeof *out);
    if (len == 0) return MP3_EOF;
    if (len < 4) return MP3_ERR_UNEXPECTED_EOF;
    const uint32_t h = be32_at(data);
    if ((h & 0xFFE00000u) != 0xFFE00000u) return MP3_ERR_NO_XING_HEADER;
    const uint32_t version_bits = (h >> 19) & 3u, layer_bits = (h >> 17) & 3u, bitrate_index = (h >> 12) & 15u;
    const uint32_t srate_index = (h >> 10) & 3u, padding = (h >> 9) & 1u;
    if (version_bits == 1 || layer_bits == 0 || bitrate_index == 0 || bitrate_index == 15 || srate_index == 3) return MP3_ERR_NO_XING_HEADER;
    const int frame_size = lameinfo_frame_size(version_bits, layer_bits, bitrate_index, srate_index, padding);
    if (frame_size < 4) return MP3_ERR_NO_XING_HEADER;
    if ((size_t)frame_size > len) return frame_size > 4 && len == 4 ? MP3_EOF : MP3_ERR_UNEXPECTED_EOF;  // ReadFull(frame[4:])
    return lameinfo_parse(data, (size_t)frame_size, out);
}

inline int lameinfo_total_delay(const mp3_lame_info *i) {  // lameinfo.go:88-93
    return i->has_lame_info ? (int)i->encoder_delay + MP3_LAME_DECODER_DELAY : MP3_LAME_DECODER_DELAY;
}
inline int lameinfo_total_padding(const mp3_lame_info *i) {  // lameinfo.go:97-108
    if (!i->has_lame_info) return 0;
    const int padding = (int)i->encoder_padding - MP3_LAME_DECODER_DELAY;
    return padding < 0 ? 0 : padding;
}

// Coarse seek for sources without a frame index (the reference parses the TOC but never uses it): the
// byte offset, from the start of the audio frames, at which `fraction` (0..1) of the playing time has passed, by linear
// interpolation between the TOC's 100 entries (entry k = 256 * offset / size at k percent).  -1 without a TOC.
inline int64_t lameinfo_toc_offset(const mp3_lame_info *i, double fraction, uint64_t stream_bytes) {
    if (!(i->flags & MP3_LAME_FLAG_TOC)) return -1;
    const uint64_t total = (i->flags & MP3_LAME_FLAG_BYTE_COUNT) ? (uint64_t)i->byte_count : stream_bytes;
    double pct = fraction * 100.0;
    if (!(pct > 0.0)) pct = 0.0;   // also NaN
    if (pct > 100.0) pct = 100.0;
    int a = (int)pct;
    if (a > 99) a = 99;
    const double fa = (double)i->toc[a], fb = a < 99 ? (double)i->toc[a + 1] : 256.0;
    const double fx = fa + (fb - fa) * (pct - (double)a);
    return (int64_t)(fx / 256.0 * (double)total);
}

}  // namespace mp3host
