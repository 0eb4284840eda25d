// mp3synth.cc — seeded synthesiser of MPEG-1 / MPEG-2-LSF Layer III byte streams.
//
// Not an audio encoder: it draws random side info and spectral integers, Huffman-ENCODES them with the
// same code tables the decoder uses, lays the granule data out through a real bit reservoir
// (main_data_begin), and writes ordinary .mp3 bytes (header, optional CRC, side info, main data), so
// that host parsing and every decode stage are exercised.  It produces the workloads of
// BASELINE.json configs 3-5 and the feature coverage the shipped fixtures lack:
// intensity stereo, mixed blocks, MPEG-2 stereo, CRC frames, big linbits, count1 overshoot,
// zero-length units, region clamp, reservoir underflow, empty tables 0/4/14.
//
// Frame sizes follow the reference's own formula (frameheader.go:223-232), including its LSF padding
// quirk, because "valid" here means "parsed by the reference as intended".
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../go-mp3_b200/csrc/tables.h"

namespace {

struct Rng {  // xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto &v : s) v = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }  // inclusive
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    bool chance(double p) { return uni() < p; }
};

struct BitWriter {  // sequential MSB-first writer starting at a byte boundary; whole bytes are overwritten
    std::vector<uint8_t> &buf;
    size_t byte;    // next byte to write
    uint64_t acc;   // pending bits (low nacc bits)
    int nacc;
    BitWriter(std::vector<uint8_t> &b, size_t start_bit) : buf(b), byte(start_bit >> 3), acc(0), nacc(0) {}
    size_t bitpos() const { return byte * 8 + (size_t)nacc; }
    void put(uint32_t v, int n) {  // n <= 32
        if (n <= 0) return;
        acc = (acc << n) | (uint64_t)(n == 32 ? v : (v & ((1u << n) - 1u)));
        nacc += n;
        if (byte + 8 > buf.size()) buf.resize(byte + 64, 0);
        while (nacc >= 8) {
            buf[byte++] = (uint8_t)(acc >> (nacc - 8));
            nacc -= 8;
        }
    }
    void flush() {  // write the pending partial byte (zero padded) without advancing
        if (nacc > 0) {
            if (byte + 1 > buf.size()) buf.resize(byte + 64, 0);
            buf[byte] = (uint8_t)((acc << (8 - nacc)) & 0xff);
        }
    }
};

// -log(u) for u = (k + 0.5) / 1024: exponential variates by table instead of a log() per spectral value
static double g_neglog[1024];
static std::once_flag g_neglog_once;
static inline double exp_variate(uint64_t r) {
    return g_neglog[r >> 54];
}

// consts.go:68-97 (long-block sfb boundaries), [lsf][sfreq]
const int kSfbLong[2][3][23] = {
    {{0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576},
     {0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576},
     {0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576}},
    {{0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
     {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 114, 136, 162, 194, 232, 278, 332, 394, 464, 540, 576},
     {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576}}};
const int kSlen1[16] = {0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4};
const int kSlen2[16] = {0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3};
const int kSfSizeMpeg2[3][6][4] = {
    {{6, 5, 5, 5}, {6, 5, 7, 3}, {11, 10, 0, 0}, {7, 7, 7, 0}, {6, 6, 6, 3}, {8, 8, 5, 0}},
    {{9, 9, 9, 9}, {9, 9, 12, 6}, {18, 18, 0, 0}, {12, 12, 12, 0}, {12, 9, 9, 6}, {15, 12, 9, 0}},
    {{6, 9, 9, 9}, {6, 9, 12, 6}, {15, 18, 0, 0}, {6, 15, 12, 0}, {6, 12, 9, 6}, {6, 18, 9, 0}}};
const int kBitrate[2][16] = {
    {0, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 160000, 192000, 224000, 256000, 320000, 0},
    {0, 8000, 16000, 24000, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 144000, 160000, 0}};
const int kSfreq[3] = {44100, 48000, 32000};

int nslen2_value(int sfc) {  // maindata.go:54-81
    if (sfc < 400) return (sfc / 80) | (((sfc / 16) % 5) << 3) | (((sfc / 4) % 4) << 6) | ((sfc % 4) << 9);
    if (sfc < 500) {
        int n = sfc - 400;
        return (n / 20) | (((n / 4) % 5) << 3) | ((n % 4) << 6) | (1 << 12);
    }
    int n = sfc - 500;
    return (n / 3) | ((n % 3) << 3) | (2 << 12) | (1 << 15);
}

// Encoder-side view of one Huffman table: code per (x, y), max symbol value.
struct EncTable {
    bool empty = true;
    int linbits = 0, maxv = 0;
    uint32_t cod[16][16];
    uint8_t len[16][16];
};
EncTable g_enc[34];
void ensure_enc() {
    std::call_once(g_neglog_once, [] {
        for (int k = 0; k < 1024; k++) g_neglog[k] = -std::log(((double)k + 0.5) / 1024.0);
    });
    static std::once_flag once;
    std::call_once(once, [] {
        for (int t = 0; t < 34; t++) {
            const mp3gpu::HuffCode *c;
            int lb = 0;
            int n = mp3gpu::huff_table_codes(t, &c, &lb);
            EncTable &e = g_enc[t];
            memset(e.cod, 0, sizeof e.cod);
            memset(e.len, 0, sizeof e.len);
            e.linbits = lb;
            if (n <= 0 || c == nullptr) continue;
            e.empty = false;
            for (int i = 0; i < n; i++) {
                e.cod[c[i].x][c[i].y] = c[i].hcod;
                e.len[c[i].x][c[i].y] = c[i].hlen;
                e.maxv = std::max(e.maxv, (int)std::max(c[i].x, c[i].y));
            }
        }
    });
}

}  // namespace

extern "C" {

typedef struct synth_cfg {
    uint64_t seed;
    int32_t n_frames;
    int32_t lsf;          // 0 = MPEG-1, 1 = MPEG-2 LSF
    int32_t sfreq;        // sampling-frequency index 0..2
    int32_t mode;         // 0 stereo, 1 joint stereo, 2 dual channel, 3 mono
    int32_t crc;          // 1: frames carry the 2 CRC bytes (protection_bit = 0)
    int32_t bitrate_lo, bitrate_hi;  // bitrate index range per frame (equal = CBR)
    int32_t padding;      // 0 never, 1 always, 2 random, 3 = CBR-style accumulate (44.1 kHz fractional slots)
    int32_t blocks;       // 0 long only, 1 long/start/short/stop, 2 also mixed blocks
    int32_t mode_ext_mask;// joint stereo: bit i set = mode_extension i allowed (bit1 of i = MS, bit0 = intensity)
    int32_t reservoir;    // 0 none (main_data_begin 0), 1 shallow (<= 64 bytes), 2 deep (driven to the 9/8-bit maximum)
    int32_t gain_lo, gain_hi;  // global_gain range
    int32_t wild;         // 0 tame; 1 quirk coverage (empty tables, zero-length units, overshoot bits, ws with bt 0,
                          //   region clamp, big linbits, reservoir underflow); 2 fuzz (random side info over random bits)
    int32_t id3v2_bytes;  // > 0: prepend an ID3v2 tag with this many payload bytes
    int32_t trailer;      // 0 none, 1 ID3v1 ("TAG" + 125 bytes), 2 APE-like tag, 3 200 garbage bytes without sync words
    int32_t fill;         // percent of each frame's bit budget to spend (tame default 92)
} synth_cfg;

// Upper bound on the stream size for cfg.
size_t synth_bound(const synth_cfg *c) {
    return (size_t)c->n_frames * 1500 + (size_t)(c->id3v2_bytes > 0 ? c->id3v2_bytes + 10 : 0) + 512;
}

}  // extern "C"

namespace {

static inline void put_random(BitWriter &w, Rng &r, int nbits) {
    while (nbits > 0) {
        const int n = nbits > 32 ? 32 : nbits;
        w.put((uint32_t)(r.next() >> 20) & (n == 32 ? 0xffffffffu : ((1u << n) - 1u)), n);
        nbits -= n;
    }
}

struct UnitSide {
    int p23 = 0, bigv = 0, ggain = 0, sfc = 0, ws = 0, bt = 0, mixed = 0, tsel[3] = {0, 0, 0}, sbg[3] = {0, 0, 0};
    int r0 = 0, r1 = 0, preflag = 0, sfscale = 0, c1t = 0;
};

int part2_bits(const synth_cfg &c, const UnitSide &u, int gr, int scfsi) {
    if (!c.lsf) {
        int s1 = kSlen1[u.sfc], s2 = kSlen2[u.sfc];
        if (u.ws && u.bt == 2) return u.mixed ? 8 * s1 + 9 * s1 + 18 * s2 : 18 * s1 + 18 * s2;
        int bits = 0;
        static const int cnt[4] = {6, 5, 5, 5};
        for (int b = 0; b < 4; b++)
            if (gr == 0 || !((scfsi >> b) & 1)) bits += cnt[b] * (b < 2 ? s1 : s2);
        return bits;
    }
    int slen = nslen2_value(u.sfc);
    int n = 0;
    if (u.bt == 2) { n++; if (u.mixed) n++; }
    int d = (slen >> 12) & 7, bits = 0;
    for (int i = 0; i < 4; i++) {
        bits += (slen & 7) * kSfSizeMpeg2[n][d][i];
        slen >>= 3;
    }
    return bits;
}

// Choose a table whose range covers `maxval`; tame streams avoid the empty tables.
int pick_table(Rng &r, int maxval, bool wild) {
    static const int by_max[6][6] = {{1, 1, 1, 1, 1, 1},      // <= 1
                                     {2, 3, 2, 3, 2, 3},      // <= 2
                                     {5, 6, 5, 6, 5, 6},      // <= 3
                                     {7, 8, 9, 7, 8, 9},      // <= 5
                                     {10, 11, 12, 10, 11, 12},// <= 7
                                     {13, 15, 13, 15, 13, 15}};// <= 15
    if (wild && r.chance(0.04)) { static const int e[3] = {0, 4, 14}; return e[r.range(0, 2)]; }
    if (maxval <= 0) return wild ? 0 : 1;
    int cls = maxval <= 1 ? 0 : maxval <= 2 ? 1 : maxval <= 3 ? 2 : maxval <= 5 ? 3 : maxval <= 7 ? 4 : maxval <= 15 ? 5 : 6;
    if (cls < 6) {
        if (r.chance(0.25) && cls < 5) cls++;  // a roomier table than needed is legal
        return by_max[cls][r.range(0, 5)];
    }
    // linbits tables: need 15 + 2^linbits - 1 >= maxval
    int cand[16], n = 0;
    for (int t = 16; t < 32; t++)
        if (15 + (1 << g_enc[t].linbits) - 1 >= maxval) cand[n++] = t;
    return cand[r.range(0, n - 1)];
}

struct StreamGen {
    const synth_cfg &c;
    Rng r;
    std::vector<uint8_t> M;  // concatenated main-data slots of all frames
    explicit StreamGen(const synth_cfg &cfg) : c(cfg), r(cfg.seed) {}

    int nch() const { return c.mode == 3 ? 1 : 2; }
    int ngr() const { return c.lsf ? 1 : 2; }

    // Emit one granule-channel at w; returns side info.  `budget` = bits available for part2+part3.
    UnitSide gen_unit(BitWriter &w, int gr, int scfsi, int bt, int ws, int mixed, int budget) {
        UnitSide u;
        u.ws = ws; u.bt = bt; u.mixed = mixed;
        u.ggain = r.range(c.gain_lo, c.gain_hi);
        u.sfscale = r.chance(0.3);
        u.preflag = (!c.lsf && !(ws && bt == 2)) ? r.chance(0.3) : 0;
        u.c1t = r.range(0, 1);
        if (ws) for (int k = 0; k < 3; k++) u.sbg[k] = r.chance(0.5) ? r.range(0, 7) : 0;
        u.sfc = c.lsf ? r.range(0, 511) : r.range(0, 15);
        if (c.lsf && c.wild == 0 && u.sfc >= 500) u.sfc = r.range(0, 499);  // keep preflag-implying codes for wild streams
        const int p2 = part2_bits(c, u, gr, scfsi);
        if (ws) {  // implicit region counts (sideinfo.go:128-136); the bitstream carries none
            u.r0 = (bt == 2 && !mixed) ? 8 : 7;
            u.r1 = 20 - u.r0;
        } else {
            u.r0 = r.range(0, 15);
            u.r1 = r.range(0, 7);
            if (c.wild == 0) {  // tame: keep region2 non-empty most of the time
                u.r0 = r.range(2, 9);
                u.r1 = r.range(1, 6);
            }
        }
        if (budget > 4095) budget = 4095;
        if (c.wild && r.chance(0.03)) {  // Q1/Q6: zero-length unit; its scalefactor bits are still walked over
            put_random(w, r, p2);
            u.p23 = 0;
            u.bigv = r.range(0, 288);
            for (int k = 0; k < 3; k++) u.tsel[k] = r.range(0, 31);
            return u;
        }
        if (p2 > budget) {  // not enough room for the scalefactors: choose a cheaper scalefac_compress
            u.sfc = 0;
            if (c.lsf) u.sfc = 0;
        }
        const int p2b = part2_bits(c, u, gr, scfsi);
        const size_t start = w.bitpos();
        put_random(w, r, p2b);
        // region boundaries (maindata/huffman.go:41-64)
        int r1s, r2s;
        if (ws && bt == 2) { r1s = 36; r2s = 576; }
        else {
            const int *l = kSfbLong[c.lsf][c.sfreq];
            r1s = l[u.r0 + 1];
            int j = u.r0 + u.r1 + 2;
            r2s = j >= 23 ? 576 : l[j];
        }
        // spectral envelope: magnitudes decay with frequency
        const int want_big = std::min(288, r.range(c.wild ? 0 : 60, c.wild ? 288 : 250));
        const double a0 = c.wild ? (r.chance(0.1) ? 400.0 : 6.0 * r.uni() + 0.5) : 1.5 + 3.0 * r.uni();
        const double tau = 60.0 + 300.0 * r.uni();
        int region_max[3];
        for (int k = 0; k < 3; k++) {
            int lo = k == 0 ? 0 : (k == 1 ? r1s : r2s);
            double a = a0 * std::exp(-(double)lo / tau);
            int mv = (int)(a * 4.0) + 1;
            if (c.wild && r.chance(0.05)) mv = r.range(16, 8206);
            if (mv > 15 && !c.wild) mv = r.chance(0.5) ? 15 : std::min(mv, 40);
            region_max[k] = mv;
            u.tsel[k] = pick_table(r, mv, c.wild != 0);
        }
        int used = p2b;
        int nbig = 0;
        const double decay = std::exp(-2.0 / tau);
        double env = a0;  // a0 * exp(-pos / tau), updated per pair
        for (; nbig < want_big; nbig++, env *= decay) {
            int pos = nbig * 2;
            int k = pos < r1s ? 0 : (pos < r2s ? 1 : 2);
            const EncTable &e = g_enc[u.tsel[k]];
            int x = 0, y = 0, bits = 0;
            uint32_t code = 0;
            int lin_x = 0, lin_y = 0;
            if (!e.empty) {
                const double a = env;
                int cap = e.linbits ? 15 + (1 << e.linbits) - 1 : e.maxv;
                cap = std::min(cap, region_max[k]);
                auto draw = [&]() {
                    double v = exp_variate(r.next()) * a;  // exponential magnitude
                    int iv = (int)v;
                    return iv > cap ? cap : iv;
                };
                x = draw(); y = draw();
                int cx = std::min(x, 15), cy = std::min(y, 15);
                if (!e.linbits) { cx = std::min(x, e.maxv); cy = std::min(y, e.maxv); x = cx; y = cy; }
                code = e.cod[cx][cy];
                bits = e.len[cx][cy];
                if (e.linbits && cx == 15) { lin_x = 1; bits += e.linbits; }
                if (x) bits++;
                if (e.linbits && cy == 15) { lin_y = 1; bits += e.linbits; }
                if (y) bits++;
                if (used + bits > budget) break;
                w.put(code, e.len[cx][cy]);
                if (lin_x) w.put((uint32_t)(x - 15), e.linbits);
                if (x) w.put((uint32_t)r.next() & 1, 1);
                if (lin_y) w.put((uint32_t)(y - 15), e.linbits);
                if (y) w.put((uint32_t)r.next() & 1, 1);
                used += bits;
            }
        }
        u.bigv = nbig;
        // count1 quadruples
        const EncTable &q = g_enc[32 + u.c1t];
        int pos = nbig * 2;
        int want_q = r.range(0, std::max(0, (576 - pos) / 4));
        if (!c.wild) want_q = std::min(want_q, r.range(10, 60));
        for (int k = 0; k < want_q && pos <= 572; k++) {
            int sym = 0;
            for (int b = 0; b < 4; b++) sym = (sym << 1) | (r.chance(0.35) ? 1 : 0);
            int bits = q.len[0][sym] + __builtin_popcount((unsigned)sym);
            if (used + bits > budget) break;
            w.put(q.cod[0][sym], q.len[0][sym]);
            for (int b = 0; b < __builtin_popcount((unsigned)sym); b++) w.put((uint32_t)r.next() & 1, 1);
            used += bits;
            pos += 4;
        }
        if (c.wild && r.chance(0.15)) {  // Q4: a few stuffing bits inside part2_3_length -> the count1 loop runs on / overshoots
            int extra = r.range(1, 9);
            if (used + extra <= budget) {
                put_random(w, r, extra);
                used += extra;
            }
        }
        u.p23 = (int)(w.bitpos() - start);
        return u;
    }

    size_t generate(uint8_t *out, size_t cap) {
        ensure_enc();
        const int nf = c.n_frames, NCH = nch(), NGR = ngr();
        const int si_size = c.lsf ? (NCH == 1 ? 9 : 17) : (NCH == 1 ? 17 : 32);
        const int mdb_max = c.lsf ? 255 : 511;
        // ---- pass 1: frame geometry ----------------------------------------------------------
        struct Fr { int bri, pad, size, md_size, mode_ext; size_t slot_start; };
        std::vector<Fr> fr((size_t)nf);
        size_t slot = 0;
        long pad_acc = 0;
        const int sfv = kSfreq[c.sfreq] >> c.lsf;
        for (int f = 0; f < nf; f++) {
            Fr &F = fr[(size_t)f];
            F.bri = r.range(c.bitrate_lo, c.bitrate_hi);
            const int br = kBitrate[c.lsf][F.bri];
            switch (c.padding) {
            case 0: F.pad = 0; break;
            case 1: F.pad = 1; break;
            case 2: F.pad = r.range(0, 1); break;
            default: {  // accumulate the fractional slot like a CBR encoder
                long rem = (144L * br) % kSfreq[c.sfreq];
                pad_acc += rem;
                F.pad = 0;
                if (pad_acc >= kSfreq[c.sfreq]) { pad_acc -= kSfreq[c.sfreq]; F.pad = 1; }
            }
            }
            F.size = ((144 * br) / sfv + F.pad) >> c.lsf;  // frameheader.go:223-232
            F.md_size = F.size - 4 - si_size - (c.crc ? 2 : 0);
            F.mode_ext = 0;
            if (c.mode == 1) {
                int allowed[4], n = 0;
                for (int m = 0; m < 4; m++) if ((c.mode_ext_mask >> m) & 1) allowed[n++] = m;
                F.mode_ext = n ? allowed[r.range(0, n - 1)] : 0;
            }
            F.slot_start = slot;
            slot += (size_t)F.md_size;
        }
        M.assign(slot + 8, 0);
        if (c.wild == 2) for (auto &b : M) b = (uint8_t)r.next();
        // ---- pass 2: granule data through the reservoir ----------------------------------------
        std::vector<uint8_t> side((size_t)nf * (size_t)si_size, 0);
        size_t write_pos = 0;  // next free byte of M
        int bt_state[2] = {0, 0}, short_left[2] = {0, 0};
        bool filling = true;   // deep reservoir: under-fill until the gap reaches the maximum, then burst
        for (int f = 0; f < nf; f++) {
            const Fr &F = fr[(size_t)f];
            const size_t slot_end = F.slot_start + (size_t)F.md_size;
            if (c.reservoir == 0 || f == 0) write_pos = std::max(write_pos, F.slot_start);
            int gap = (int)(F.slot_start - write_pos);
            const int gap_cap = c.reservoir == 1 ? 64 : mdb_max;
            if (gap > gap_cap) { write_pos = F.slot_start - (size_t)gap_cap; gap = gap_cap; }
            int mdb = gap;
            const long avail_bits = (long)(slot_end - write_pos) * 8;
            // frame budget
            long budget;
            const int fillpct = c.fill > 0 ? c.fill : 92;
            if (c.reservoir == 2) {
                if (filling) { budget = (long)F.md_size * 8 * 55 / 100; if (gap >= mdb_max - 8) filling = false; }
                else { budget = avail_bits * 97 / 100; if (gap < 40) filling = true; }
            } else if (c.reservoir == 1) {
                budget = (long)F.md_size * 8 * r.range(fillpct - 8, std::min(100, fillpct + 6)) / 100;
            } else {
                budget = (long)F.md_size * 8 * fillpct / 100;
            }
            budget = std::min(budget, avail_bits);
            // block types for this frame
            int scfsi[2] = {0, 0};
            if (!c.lsf) for (int ch = 0; ch < NCH; ch++) scfsi[ch] = r.chance(0.5) ? r.range(0, 15) : 0;
            UnitSide us[2][2];
            BitWriter w(M, write_pos * 8);
            const int n_units = NGR * NCH;
            int unit_i = 0;
            for (int gr = 0; gr < NGR; gr++) {
                bool common_bt = r.chance(0.7);
                for (int ch = 0; ch < NCH; ch++) {
                    int s = common_bt ? 0 : ch;  // channel whose state machine drives this unit
                    int bt = 0, ws = 0, mixed = 0;
                    if (c.blocks > 0) {
                        int &st = bt_state[s];
                        if (!(common_bt && ch == 1)) {  // advance the machine once per granule when shared
                            if (st == 0) { if (r.chance(0.12)) st = 1; }
                            else if (st == 1) { st = 2; short_left[s] = r.range(1, 3); }
                            else if (st == 2) { if (--short_left[s] <= 0) st = 3; }
                            else st = 0;
                        }
                        bt = st; ws = bt != 0;
                        if (bt == 2 && c.blocks > 1 && !c.lsf) mixed = r.chance(0.25);  // LSF mixed blocks panic the reference
                        if (c.wild == 1 && bt != 2 && ws && r.chance(0.1)) mixed = 1;   // Q15: mixed flag on start/stop windows
                        if (c.wild == 1 && bt == 0 && r.chance(0.03)) ws = 1;           // Q15: window switching with block type 0
                    }
                    long share = (budget - (long)(w.bitpos() - write_pos * 8)) / (n_units - unit_i);
                    if (share < 0) share = 0;
                    share = share * r.range(70, 130) / 100;
                    long remaining = avail_bits - (long)(w.bitpos() - write_pos * 8);
                    if (share > remaining) share = remaining;
                    us[gr][ch] = gen_unit(w, gr, scfsi[ch], bt, ws, mixed, (int)share);
                    unit_i++;
                }
            }
            w.flush();
            size_t end_byte = (w.bitpos() + 7) >> 3;
            if (end_byte > slot_end) end_byte = slot_end;  // never triggers: budgets are bounded by avail_bits
            write_pos = end_byte;
            if (c.wild == 1 && f > 0 && r.chance(0.01)) mdb = r.range(mdb, mdb_max);  // Q8: reservoir underflow / misaligned start
            if (c.wild == 2) mdb = r.range(0, mdb_max);
            // ---- side info bits ----------------------------------------------------------------
            std::vector<uint8_t> sib((size_t)si_size, 0);
            BitWriter sw(sib, 0);
            if (c.wild == 2) {
                for (int i = 0; i < si_size; i++) sib[(size_t)i] = (uint8_t)r.next();
                // keep the stream decodable end to end: big_values <= 288, and no LSF mixed blocks
                BitWriter fw(sib, 0);
                std::vector<uint8_t> clean((size_t)si_size, 0);
                // re-emit field by field from the random bits, patching the two fatal cases
                struct Rd { const std::vector<uint8_t> &b; size_t p = 0; int get(int n) { int v = 0; for (int i = 0; i < n; i++, p++) v = (v << 1) | ((b[p >> 3] >> (7 - (p & 7))) & 1); return v; } } rd{sib};
                BitWriter cw(clean, 0);
                cw.put((uint32_t)rd.get(c.lsf ? 8 : 9), c.lsf ? 8 : 9);
                int pv = c.lsf ? (NCH == 1 ? 1 : 2) : (NCH == 1 ? 5 : 3);
                cw.put((uint32_t)rd.get(pv), pv);
                if (!c.lsf) cw.put((uint32_t)rd.get(4 * NCH), 4 * NCH);
                for (int gr = 0; gr < NGR; gr++)
                    for (int ch = 0; ch < NCH; ch++) {
                        cw.put((uint32_t)rd.get(12), 12);
                        int bv = rd.get(9);
                        if (bv > 288) bv = r.chance(0.02) ? bv : bv % 289;  // a few fatal ones stay (isPos error path)
                        cw.put((uint32_t)bv, 9);
                        cw.put((uint32_t)rd.get(8), 8);
                        int sl = c.lsf ? 9 : 4;
                        cw.put((uint32_t)rd.get(sl), sl);
                        int ws = rd.get(1);
                        cw.put((uint32_t)ws, 1);
                        if (ws) {
                            int bt = rd.get(2), mx = rd.get(1);
                            if (c.lsf && bt == 2 && mx && !r.chance(0.02)) mx = 0;
                            cw.put((uint32_t)bt, 2); cw.put((uint32_t)mx, 1);
                            cw.put((uint32_t)rd.get(19), 19);
                        } else cw.put((uint32_t)rd.get(22), 22);
                        int tail = c.lsf ? 2 : 3;
                        cw.put((uint32_t)rd.get(tail), tail);
                    }
                sib = clean;
            } else {
                sw.put((uint32_t)mdb, c.lsf ? 8 : 9);
                int pv = c.lsf ? (NCH == 1 ? 1 : 2) : (NCH == 1 ? 5 : 3);
                sw.put(0, pv);
                if (!c.lsf) for (int ch = 0; ch < NCH; ch++) for (int b = 0; b < 4; b++) sw.put((uint32_t)(scfsi[ch] >> b) & 1, 1);
                for (int gr = 0; gr < NGR; gr++)
                    for (int ch = 0; ch < NCH; ch++) {
                        const UnitSide &u = us[gr][ch];
                        sw.put((uint32_t)u.p23, 12);
                        sw.put((uint32_t)u.bigv, 9);
                        sw.put((uint32_t)u.ggain, 8);
                        sw.put((uint32_t)u.sfc, c.lsf ? 9 : 4);
                        sw.put((uint32_t)u.ws, 1);
                        if (u.ws) {
                            sw.put((uint32_t)u.bt, 2);
                            sw.put((uint32_t)u.mixed, 1);
                            sw.put((uint32_t)u.tsel[0], 5); sw.put((uint32_t)u.tsel[1], 5);
                            for (int k = 0; k < 3; k++) sw.put((uint32_t)u.sbg[k], 3);
                        } else {
                            for (int k = 0; k < 3; k++) sw.put((uint32_t)u.tsel[k], 5);
                            sw.put((uint32_t)u.r0, 4);
                            sw.put((uint32_t)u.r1, 3);
                        }
                        if (!c.lsf) sw.put((uint32_t)u.preflag, 1);
                        sw.put((uint32_t)u.sfscale, 1);
                        sw.put((uint32_t)u.c1t, 1);
                    }
            }
            sw.flush();
            memcpy(&side[(size_t)f * (size_t)si_size], sib.data(), (size_t)si_size);
        }
        // ---- pass 3: byte stream -------------------------------------------------------------------
        size_t o = 0;
        auto put8 = [&](uint8_t b) { if (o < cap) out[o] = b; o++; };
        if (c.id3v2_bytes > 0) {
            put8('I'); put8('D'); put8('3'); put8(4); put8(0); put8(0);
            uint32_t n = (uint32_t)c.id3v2_bytes;
            put8((uint8_t)((n >> 21) & 0x7f)); put8((uint8_t)((n >> 14) & 0x7f)); put8((uint8_t)((n >> 7) & 0x7f)); put8((uint8_t)(n & 0x7f));
            for (uint32_t i = 0; i < n; i++) put8((uint8_t)(i * 7));
        }
        for (int f = 0; f < nf; f++) {
            const Fr &F = fr[(size_t)f];
            put8(0xff);
            put8((uint8_t)(0xe0 | ((c.lsf ? 2 : 3) << 3) | (1 << 1) | (c.crc ? 0 : 1)));
            put8((uint8_t)((F.bri << 4) | (c.sfreq << 2) | (F.pad << 1)));
            put8((uint8_t)((c.mode << 6) | (F.mode_ext << 4) | 0x4));  // copyright bit set like the reference's test frame (FF FB 90 44)
            if (c.crc) { put8((uint8_t)r.next()); put8((uint8_t)r.next()); }
            for (int i = 0; i < si_size; i++) put8(side[(size_t)f * (size_t)si_size + (size_t)i]);
            for (int i = 0; i < F.md_size; i++) put8(M[F.slot_start + (size_t)i]);
        }
        if (c.trailer == 1) { put8('T'); put8('A'); put8('G'); for (int i = 0; i < 125; i++) put8((uint8_t)('a' + i % 26)); }
        else if (c.trailer == 2) { const char *a = "APETAGEX"; for (int i = 0; i < 8; i++) put8((uint8_t)a[i]); for (int i = 0; i < 56; i++) put8((uint8_t)(i & 0x7f)); }
        else if (c.trailer == 3) { for (int i = 0; i < 200; i++) put8((uint8_t)(i % 0x7f)); }
        return o;
    }
};

}  // namespace

extern "C" {

// Writes the stream for cfg into out (capacity cap); returns the stream length (<= synth_bound).
size_t synth_stream(const synth_cfg *cfg, uint8_t *out, size_t cap) {
    StreamGen g(*cfg);
    return g.generate(out, cap);
}

// n streams on `threads` threads.  Stream i is written at out + offsets[i]; lens[i] receives its length.
void synth_batch(const synth_cfg *cfgs, int n, int threads, uint8_t *out, const size_t *offsets, const size_t *caps, size_t *lens) {
    ensure_enc();
    std::atomic<int> next{0};
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            lens[i] = synth_stream(&cfgs[i], out + offsets[i], caps[i]);
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(work);
    for (auto &t : th) t.join();
}

}  // extern "C"
