// tables.cc — see tables.h.
#include "tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace mp3gpu {

namespace {
typedef HuffCode huff_code_t;
struct huff_table_desc_t {
    const huff_code_t *codes;
    int n;
    int linbits;
};
#include "huff_codes.inc"
#include "synth_window_k.inc"

// consts.go:68-97, order [lsf][sfreq].
const int kSfbLong[2][3][23] = {
    {{0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576},
     {0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576},
     {0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576}},
    {{0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
     {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 114, 136, 162, 194, 232, 278, 332, 394, 464, 540, 576},
     {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576}}};
const int kSfbShort[2][3][14] = {{{0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192},
                                  {0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192},
                                  {0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192}},
                                 {{0, 4, 8, 12, 18, 24, 32, 42, 56, 74, 100, 132, 174, 192},
                                  {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 136, 180, 192},
                                  {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 134, 174, 192}}};
const int kPretab[22] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0};
const int kSfSizeMpeg2[3][6][4] = {
    {{6, 5, 5, 5}, {6, 5, 7, 3}, {11, 10, 0, 0}, {7, 7, 7, 0}, {6, 6, 6, 3}, {8, 8, 5, 0}},
    {{9, 9, 9, 9}, {9, 9, 12, 6}, {18, 18, 0, 0}, {12, 12, 12, 0}, {12, 9, 9, 6}, {15, 12, 9, 0}},
    {{6, 9, 9, 9}, {6, 9, 12, 6}, {15, 18, 0, 0}, {6, 15, 12, 0}, {6, 12, 9, 6}, {6, 18, 9, 0}}};

const long double kPiL = 3.14159265358979323846264338327950288L;

// ---- Huffman LUT construction ------------------------------------------------
// Entry formats: tables.h.
// Pair trees: every code word is extended by its sign bits (x's first, then y's: huffman.go:404-416) into 1, 2 or 4
// "extended codes" whose leaves hold the signed pair; code words that take the linbits escape stay as they are.
struct ExtCode {
    uint32_t bits;   // the extended code, right-aligned
    int len;
    uint16_t entry;  // the leaf / escape entry
};
struct PairLutBuilder {
    std::vector<uint16_t> &lut;
    size_t base;
    std::vector<ExtCode> codes;
    PairLutBuilder(std::vector<uint16_t> &l, const huff_code_t *c, int n, bool esc_tree) : lut(l), base(l.size()) {
        for (int i = 0; i < n; i++) {
            const int x = c[i].x, y = c[i].y, tlen = c[i].hlen;
            if (esc_tree && (x == 15 || y == 15)) {
                const int other = x == 15 ? y : x;
                codes.push_back({c[i].hcod, tlen,
                                 (uint16_t)(0xc000u | (uint32_t)tlen | (x == 15 ? 0x20u : 0u) | (y == 15 ? 0x40u : 0u) |
                                            ((uint32_t)(other & 15) << 7))});
                continue;
            }
            const int nsx = x != 0, nsy = y != 0;
            for (int sx = 0; sx <= nsx; sx++)
                for (int sy = 0; sy <= nsy; sy++) {
                    const int xs = sx ? -x : x, ys = sy ? -y : y;
                    uint32_t bits = c[i].hcod;
                    int len = tlen;
                    if (nsx) { bits = (bits << 1) | (uint32_t)sx; len++; }
                    if (nsy) { bits = (bits << 1) | (uint32_t)sy; len++; }
                    codes.push_back({bits, len, (uint16_t)(((uint32_t)xs & 31u) | (((uint32_t)ys & 31u) << 5) | ((uint32_t)len << 10))});
                }
        }
    }
    // Fill a table of `bits` index bits for all extended codes whose first `plen` bits equal `prefix`.
    void fill(size_t off, int bits, uint32_t prefix, int plen) {
        for (uint32_t idx = 0; idx < (1u << bits); idx++) {
            const uint32_t full = (prefix << bits) | idx;  // plen + bits bits
            int found = -1;
            for (size_t i = 0; i < codes.size(); i++) {
                const int L = codes[i].len;
                if (L <= plen || L > plen + bits) continue;
                if ((full >> (plen + bits - L)) == codes[i].bits) { found = (int)i; break; }
            }
            if (found >= 0) {
                lut[base + off + idx] = codes[(size_t)found].entry;
                continue;
            }
            int maxrem = 0;
            for (size_t i = 0; i < codes.size(); i++) {
                const int L = codes[i].len;
                if (L <= plen + bits) continue;
                if ((codes[i].bits >> (L - plen - bits)) == full) maxrem = std::max(maxrem, L - plen - bits);
            }
            if (maxrem == 0) throw std::runtime_error("huffman tree not complete");
            const int sb = std::min(maxrem, kHuffSubBits);  // longer codes go through another link
            while ((lut.size() - base) % 16) lut.push_back(0);  // a link addresses its sub-table in units of 16 entries
            const size_t sub_off = lut.size() - base;
            if (sub_off / 16 > 0x3ff) throw std::runtime_error("huffman LUT too large");
            lut.resize(lut.size() + ((size_t)1 << sb), 0);
            lut[base + off + idx] = (uint16_t)(0x8000u | ((uint32_t)sb << 10) | (uint32_t)(sub_off / 16));
            fill(sub_off, sb, full, plen + bits);
        }
    }
};

// Count1 trees: at most 6 tree bits, one root table each (32-bit entries).
uint32_t quad_leaf(const huff_code_t &c) {
    const int q = c.y & 0xf;  // count1 tables are stored as (x = 0, y = v<<3 | w<<2 | x<<1 | y)
    const int signs = ((q >> 3) & 1) + ((q >> 2) & 1) + ((q >> 1) & 1) + (q & 1);
    return (uint32_t)q | ((uint32_t)c.hlen << 16) | ((uint32_t)(c.hlen + signs) << 26);
}
void fill_quad_root(uint32_t *root, const huff_code_t *codes, int n) {
    for (uint32_t idx = 0; idx < (1u << kQuadRootBits); idx++) {
        int found = -1;
        for (int i = 0; i < n; i++)
            if (codes[i].hlen <= kQuadRootBits && (idx >> (kQuadRootBits - codes[i].hlen)) == codes[i].hcod) { found = i; break; }
        if (found < 0) throw std::runtime_error("count1 tree not complete within the root table");
        root[idx] = quad_leaf(codes[found]);
    }
}

}  // namespace

int huff_table_codes(int table_num, const HuffCode **codes, int *linbits) {
    if (table_num < 0 || table_num > 33) return -1;
    *codes = HUFF_TABLES[table_num].codes;
    *linbits = HUFF_TABLES[table_num].linbits;
    return HUFF_TABLES[table_num].n;
}

void build_host_tables(HostTables &t) {
    // imdct.go:23-57
    const double pi36 = (double)(kPiL / 36.0L), pi12 = (double)(kPiL / 12.0L);
    float(*w)[36] = reinterpret_cast<float(*)[36]>(t.imdct_win);
    for (int i = 0; i < 36; i++) w[0][i] = (float)std::sin(pi36 * ((double)i + 0.5));
    for (int i = 0; i < 18; i++) w[1][i] = (float)std::sin(pi36 * ((double)i + 0.5));
    for (int i = 18; i < 24; i++) w[1][i] = 1.0f;
    for (int i = 24; i < 30; i++) w[1][i] = (float)std::sin(pi12 * ((double)i + 0.5 - 18.0));
    for (int i = 30; i < 36; i++) w[1][i] = 0.0f;
    for (int i = 0; i < 12; i++) w[2][i] = (float)std::sin(pi12 * ((double)i + 0.5));
    for (int i = 12; i < 36; i++) w[2][i] = 0.0f;
    for (int i = 0; i < 6; i++) w[3][i] = 0.0f;
    for (int i = 6; i < 12; i++) w[3][i] = (float)std::sin(pi12 * ((double)i + 0.5 - 6.0));
    for (int i = 12; i < 18; i++) w[3][i] = 1.0f;
    for (int i = 18; i < 36; i++) w[3][i] = (float)std::sin(pi36 * ((double)i + 0.5));
    // imdct.go:61-79
    const double pi24 = (double)(kPiL / 24.0L), pi72 = (double)(kPiL / 72.0L);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 12; j++)
            t.cos12[i * 12 + j] = (float)std::cos(pi24 * (2.0 * (double)j + 1.0 + 6.0) * (2.0 * (double)i + 1.0));
    for (int i = 0; i < 18; i++)
        for (int j = 0; j < 36; j++)
            t.cos36[i * 36 + j] = (float)std::cos(pi72 * (2.0 * (double)j + 1.0 + 18.0) * (2.0 * (double)i + 1.0));
    // frame.go:490-497
    const double pi64 = (double)(kPiL / 64.0L);
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 32; j++) t.synth_n[i * 32 + j] = (float)std::cos((double)((16 + i) * (2 * j + 1)) * pi64);
    // frame.go:499-628: 9-decimal literals of k/65536
    for (int i = 0; i < 512; i++) {
        long long k = SYNTH_WINDOW_K[i];
        long long a = k < 0 ? -k : k;
        long long n = (a * 1000000000LL + 32768) / 65536;
        char lit[40];
        snprintf(lit, sizeof lit, "%s%lld.%09lld", k < 0 ? "-" : "", n / 1000000000LL, n % 1000000000LL);
        t.synth_d[i] = strtof(lit, nullptr);
    }
    // frame.go:146-148,161-166: math.Pow(2.0, idx) with idx a multiple of 0.25
    for (int k = 0; k < kPow2N; k++) t.pow2q[k] = std::pow(2.0, (double)(k - kPow2Off) * 0.25);
    // frame.go:36-40
    t.powtab34.resize(8207);
    for (int i = 0; i < 8207; i++) t.powtab34[i] = std::pow((double)i, 4.0 / 3.0);
    // Requantisation (frame.go:146-155) is float32(2^(k/4) * |is|^(4/3)) with the product taken in float64.  With
    // k = 4e + q the first factor is 2^e * 2^(q/4) exactly (checked here), and a power of two commutes with both
    // roundings, so the value is float32(2^(q/4) * |is|^(4/3)) * 2^e: one table row per q and an exact float scale.
    for (int k = 0; k < kPow2N; k++) {
        const int k4 = k - kPow2Off, e = k4 >> 2, q = k4 & 3;
        if (t.pow2q[k] != std::ldexp(t.pow2q[kPow2Off + q], e)) throw std::runtime_error("2^(k/4) is not 2^e * 2^(q/4)");
    }
    t.powq4.assign((size_t)4 * kPowRow, 0.0f);
    for (int q = 0; q < 4; q++)
        for (int i = 0; i < 8207; i++) t.powq4[(size_t)q * kPowRow + i] = (float)(t.pow2q[kPow2Off + q] * t.powtab34[i]);
    // frame.go:422-425 (decimal literals rounded once to float32)
    const float cs[8] = {0.857493f, 0.881742f, 0.949629f, 0.983315f, 0.995518f, 0.999161f, 0.999899f, 0.999993f};
    const float ca[8] = {-0.514496f, -0.471732f, -0.313377f, -0.181913f, -0.094574f, -0.040966f, -0.014199f, -0.003700f};
    memcpy(t.cs, cs, sizeof cs);
    memcpy(t.ca, ca, sizeof ca);
    // frame.go:304-327: float32 division, as the reference does per band
    const float isr[6] = {0.000000f, 0.267949f, 0.577350f, 1.000000f, 1.732051f, 3.732051f};
    for (int p = 0; p < 6; p++) {
        volatile float den = 1.0f + isr[p];
        t.is_ratio_l[p] = isr[p] / den;
        t.is_ratio_r[p] = 1.0f / den;
    }
    t.is_ratio_l[6] = 1.0f;
    t.is_ratio_r[6] = 0.0f;
    t.is_ratio_l[7] = 1.0f;  // unused (is_pos >= 7: no intensity processing)
    t.is_ratio_r[7] = 1.0f;
    memset(t.pretab, 0, sizeof t.pretab);
    for (int i = 0; i < 22; i++) t.pretab[i] = (uint8_t)kPretab[i];
    t.pretab_pack = 0;
    for (int i = 0; i < 22; i++) {
        if (kPretab[i] > 3) throw std::runtime_error("pretab entry does not fit 2 bits");
        t.pretab_pack |= (uint64_t)kPretab[i] << (2 * i);
    }
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 6; b++)
            for (int c = 0; c < 4; c++) t.sfsize_mpeg2[a][b][c] = (uint8_t)kSfSizeMpeg2[a][b][c];

    for (int lsf = 0; lsf < 2; lsf++)
        for (int sf = 0; sf < 3; sf++) {
            int cfg = lsf * 3 + sf;
            memset(t.sfb_long[cfg], 0, sizeof t.sfb_long[cfg]);
            memset(t.sfb_short[cfg], 0, sizeof t.sfb_short[cfg]);
            for (int i = 0; i < 23; i++) t.sfb_long[cfg][i] = (uint16_t)kSfbLong[lsf][sf][i];
            for (int i = 0; i < 14; i++) t.sfb_short[cfg][i] = (uint16_t)kSfbShort[lsf][sf][i];
            for (int sfb = 0; sfb < 22; sfb++)
                for (int i = kSfbLong[lsf][sf][sfb]; i < kSfbLong[lsf][sf][sfb + 1]; i++) t.line_sfb_long[cfg][i] = (uint8_t)sfb;
            for (int sfb = 0; sfb < 13; sfb++) {
                int s = kSfbShort[lsf][sf][sfb] * 3;
                int wl = kSfbShort[lsf][sf][sfb + 1] - kSfbShort[lsf][sf][sfb];
                for (int win = 0; win < 3; win++)
                    for (int j = 0; j < wl; j++) {
                        int i = s + win * wl + j;
                        t.line_sfb_short[cfg][i] = (uint8_t)sfb;
                        t.line_win_short[cfg][i] = (uint8_t)win;
                        t.reorder_dst[cfg][i] = (uint16_t)(s + j * 3 + win);  // frame.go:291-296
                    }
            }
        }
    // pair tables: lines 2p and 2p+1 always share their band (every band boundary is even)
    for (int cfg = 0; cfg < kNumCfg; cfg++)
        for (int p = 0; p < 288; p++) {
            if (t.line_sfb_long[cfg][2 * p] != t.line_sfb_long[cfg][2 * p + 1] ||
                t.line_sfb_short[cfg][2 * p] != t.line_sfb_short[cfg][2 * p + 1] ||
                t.line_win_short[cfg][2 * p] != t.line_win_short[cfg][2 * p + 1] ||
                t.reorder_dst[cfg][2 * p + 1] != t.reorder_dst[cfg][2 * p] + 3)
                throw std::runtime_error("scalefactor band boundary is not even");
            t.pair_long[cfg][p] = t.line_sfb_long[cfg][2 * p];
            t.pair_short[cfg][p] = (uint8_t)(t.line_sfb_short[cfg][2 * p] * 3 + t.line_win_short[cfg][2 * p]);
            t.pair_dst[cfg][p] = t.reorder_dst[cfg][2 * p];
        }
    // maindata.go:54-81
    memset(t.nslen2, 0, sizeof t.nslen2);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 3; j++) t.nslen2[j + i * 3 + 500] = (uint16_t)(i | (j << 3) | (2 << 12) | (1 << 15));
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++) t.nslen2[l + k * 4 + j * 16 + i * 80] = (uint16_t)(i | (j << 3) | (k << 6) | (l << 9));
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 4; k++) t.nslen2[k + j * 4 + i * 20 + 400] = (uint16_t)(i | (j << 3) | (k << 6) | (1 << 12));

    // Huffman LUTs: one per distinct tree; tables sharing a tree share the LUT.  Every root table is indexed by the
    // same kHuffRootBits bits, so the kernel's root lookup needs no per-table shift.
    t.huff_lut.clear();
    // entries 0..2^root-1: the "empty table" LUT (zero-length zero leaves: nothing consumed, x = y = 0)
    t.huff_lut.resize((size_t)1 << kHuffRootBits, (uint16_t)0);
    const uint32_t empty_desc = 0u;
    const huff_code_t *seen[34];
    uint32_t seen_desc[34];
    int nseen = 0;
    for (int tab = 0; tab < 32; tab++) {
        const huff_table_desc_t &d = HUFF_TABLES[tab];
        if (d.codes == nullptr) {
            t.huff_desc[tab] = empty_desc;
            continue;
        }
        uint32_t desc = 0;
        bool have = false;
        for (int s = 0; s < nseen; s++)
            if (seen[s] == d.codes) { desc = seen_desc[s]; have = true; }
        if (!have) {
            while (t.huff_lut.size() % 16) t.huff_lut.push_back(0);
            PairLutBuilder b(t.huff_lut, d.codes, d.n, d.linbits != 0);
            if (b.base * 2 > 0xffffff) throw std::runtime_error("huffman LUT base overflow");
            t.huff_lut.resize(t.huff_lut.size() + ((size_t)1 << kHuffRootBits), 0);
            b.fill(0, kHuffRootBits, 0, 0);
            desc = (uint32_t)b.base * 2u;
            seen[nseen] = d.codes;
            seen_desc[nseen] = desc;
            nseen++;
        }
        t.huff_desc[tab] = desc | ((uint32_t)d.linbits << 24);
    }
    while (t.huff_lut.size() % 8) t.huff_lut.push_back(0);  // staged 16 bytes at a time (k_huffman)
    for (int q = 0; q < 2; q++) {
        fill_quad_root(t.quad_lut + q * 256, HUFF_TABLES[32 + q].codes, HUFF_TABLES[32 + q].n);
        t.huff_desc[32 + q] = (uint32_t)(q * 256 * 4);
    }
    // count1 sign expansion: index = pattern << 4 | the four bits after the tree bits; value = the two output words
    // (v | w << 16) and (x | y << 16) << 32 with the sign bits dealt to the non-zero values in order (huffman.go:387-403)
    for (int q = 0; q < 16; q++)
        for (int four = 0; four < 16; four++) {
            int val[4], used = 0;
            for (int i = 0; i < 4; i++) {
                val[i] = (q >> (3 - i)) & 1;
                if (val[i]) {
                    if ((four >> (3 - used)) & 1) val[i] = -1;
                    used++;
                }
            }
            const uint64_t vw = ((uint32_t)val[0] & 0xffffu) | ((uint32_t)val[1] << 16);
            const uint64_t xy = ((uint32_t)val[2] & 0xffffu) | ((uint32_t)val[3] << 16);
            t.quad_signs[q * 16 + four] = vw | (xy << 32);
        }
}

}  // namespace mp3gpu
