// hostemu.cc — TEST INFRASTRUCTURE ONLY.
// Compiles the __host__ __device__ unit logic of K1/K2 (go-mp3_b200/csrc/unit_logic.h) for the CPU so
// that `-m "not gpu"` tests can check the bit-level logic and the table construction against the oracle
// without a GPU.  Nothing in the product links or loads this file; the C ABI has no CPU path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../go-mp3_b200/csrc/tables.h"
#include "../../go-mp3_b200/csrc/unit_logic.h"

using namespace mp3gpu;

namespace {
HostTables *g_h = nullptr;
DeviceTables g_T;
const uint8_t kSlen[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {3, 0}, {1, 1}, {1, 2}, {1, 3},
                              {2, 1}, {2, 2}, {2, 3}, {3, 1}, {3, 2}, {3, 3}, {4, 2}, {4, 3}};
void ensure() {
    if (g_h) return;
    g_h = new HostTables();
    build_host_tables(*g_h);
    HostTables &h = *g_h;
    g_T.pow2q = h.pow2q;
    g_T.powtab34 = h.powtab34.data();
    g_T.powq4 = h.powq4.data();
    g_T.pretab_pack = h.pretab_pack;
    g_T.line_sfb_long = &h.line_sfb_long[0][0];
    g_T.line_sfb_short = &h.line_sfb_short[0][0];
    g_T.line_win_short = &h.line_win_short[0][0];
    g_T.reorder_dst = &h.reorder_dst[0][0];
    g_T.pair_long = &h.pair_long[0][0];
    g_T.pair_short = &h.pair_short[0][0];
    g_T.pair_dst = &h.pair_dst[0][0];
    g_T.sfb_long = &h.sfb_long[0][0];
    g_T.sfb_short = &h.sfb_short[0][0];
    g_T.nslen2 = h.nslen2;
    g_T.huff_lut = h.huff_lut.data();
    g_T.quad_lut = h.quad_lut;
    g_T.huff_desc = h.huff_desc;
    g_T.quad_signs = h.quad_signs;
    g_T.is_ratio_l = h.is_ratio_l;
    g_T.is_ratio_r = h.is_ratio_r;
    g_T.pretab = h.pretab;
    g_T.sfsize_mpeg2 = &h.sfsize_mpeg2[0][0][0];
    g_T.slen_mpeg1 = &kSlen[0][0];
    g_T.cs = h.cs;
    g_T.ca = h.ca;
    g_T.huff_lut_n = (int)h.huff_lut.size();
    g_T.pow2_off = kPow2Off;
}
}  // namespace

// tree length of code word (x, y) in pair table `t` (0 for the empty tables, -1 if the table has no such code)
static int emu_code_len(int t, int x, int y) {
    const mp3gpu::HuffCode *codes = nullptr;
    int lin = 0;
    const int n = mp3gpu::huff_table_codes(t, &codes, &lin);
    if (n <= 0 || codes == nullptr) return (x == 0 && y == 0) ? 0 : -1;
    for (int i = 0; i < n; i++)
        if (codes[i].x == x && codes[i].y == y) return codes[i].hlen;
    return -1;
}
extern "C" {

// K1 on the CPU: is16 [n][576] (zero-filled above count1), meta [n], scalefac [n][64] (as MP3GPU_TAP_SCALEFAC).
void emu_huffman(const uint8_t *main_data, unsigned long long main_bits, const mp3gpu_unit *units, long long n_units, int16_t *is16,
                 uint32_t *meta, uint8_t *scalefac) {
    ensure();
    for (long long u = 0; u < n_units; u++) {
        memset(is16 + u * 576, 0, 576 * sizeof(int16_t));
        memset(scalefac + u * 64, 0, 64);
        meta[u] = 0;
        if (!u_valid(units[u].w2)) continue;
        uint32_t pk[8];
        alignas(16) uint32_t out[288 + 4];
        memset(out, 0, sizeof out);
        uint32_t m = huffman_unit(g_T, SmemRef::of(g_T.huff_lut), g_T.quad_lut, g_T.huff_desc, g_T.quad_signs, main_data, main_bits, units, u, pk, out);
        meta[u] = m;
        int c1 = (int)(m & 0x3ff);
        for (int i = 0; i < c1; i++) is16[u * 576 + i] = (int16_t)((out[i >> 1] >> (16 * (i & 1))) & 0xffff);
        for (int k = 0; k < 64; k++) scalefac[u * 64 + k] = (uint8_t)sf_nib(pk, k);
        scalefac[u * 64 + 61] = (uint8_t)((m >> 10) & 1);
    }
}

// K1 on the CPU the way k_huffman runs it: tiles of `tile` consecutive units, the stretch of main data a tile reads
// staged (byte-swapped) into a buffer of at most cap16 16-byte chunks, every unit decoded through a StagedCursor.
// main_data must be followed by 64 readable bytes, like the device buffer.
void emu_huffman_staged(const uint8_t *main_data, unsigned long long main_bits, const mp3gpu_unit *units, long long n_units, int tile,
                        int cap16, int16_t *is16, uint32_t *meta, uint8_t *scalefac) {
    ensure();
    const uint32_t main16 = (uint32_t)(((main_bits >> 3) + 48) >> 4);
    std::vector<uint32_t> stage;
    for (long long base = 0; base < n_units; base += tile) {
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (long long u = base; u < n_units && u < base + tile; u++) {
            if (!u_valid(units[u].w2)) continue;
            uint32_t l, h;
            stage_reach(units[u], main_bits, &l, &h);
            lo = l < lo ? l : lo;
            hi = h > hi ? h : hi;
        }
        if (hi > main16) hi = main16;
        uint32_t n16 = hi > lo ? hi - lo : 0u;
        if (n16 > (uint32_t)cap16) n16 = (uint32_t)cap16;
        stage.assign((size_t)n16 * 4 + 4, 0xdeadbeefu);  // poisoned past the end: nothing may read it
        for (size_t i = 0; i < (size_t)n16 * 4; i++) {
            uint32_t v;
            memcpy(&v, main_data + (size_t)lo * 16 + i * 4, 4);
            stage[i] = be32(v);
        }
        StageCtx S;
        S.sw = SmemRef::of(stage.data());
        S.n_words = (int)(n16 * 4);
        S.lo_word = (unsigned long long)lo * 4ull;
        S.gw = reinterpret_cast<const uint32_t *>(main_data);
        S.main_bits = main_bits;
        for (long long u = base; u < n_units && u < base + tile; u++) {
            memset(is16 + u * 576, 0, 576 * sizeof(int16_t));
            memset(scalefac + u * 64, 0, 64);
            meta[u] = 0;
            if (!u_valid(units[u].w2)) continue;
            uint32_t pk[8];
            alignas(16) uint32_t out[288 + 4];
            memset(out, 0, sizeof out);
            uint32_t m = huffman_unit_staged(g_T, SmemRef::of(g_T.huff_lut), g_T.quad_lut, g_T.huff_desc, g_T.quad_signs, S, units, u, pk, out);
            meta[u] = m;
            int c1 = (int)(m & 0x3ff);
            for (int i = 0; i < c1; i++) is16[u * 576 + i] = (int16_t)((out[i >> 1] >> (16 * (i & 1))) & 0xffff);
            for (int k = 0; k < 64; k++) scalefac[u * 64 + k] = (uint8_t)sf_nib(pk, k);
            scalefac[u * 64 + 61] = (uint8_t)((m >> 10) & 1);
        }
    }
}

// K2 on the CPU, same order of operations as k_requant: xr [n_granules][2][576] after
// requantise + reorder + stereo + alias reduction, index sb*18+i.
void emu_requant(const mp3gpu_unit *units, long long n_granules, const int16_t *is16, const uint32_t *meta,
                 const uint8_t *scalefac, float *xr) {
    ensure();
    for (long long g = 0; g < n_granules; g++) {
        const mp3gpu_unit *ug = units + g * 2;
        float *x0 = xr + (g * 2) * 576, *x1 = x0 + 576;
        memset(x0, 0, 2 * 576 * sizeof(float));
        if (!u_valid(ug[0].w2)) continue;
        const bool valid_b = u_valid(ug[1].w2);
        const int cfg = u_lsf(ug[0].w2) * 3 + u_sfreq(ug[0].w2);
        uint32_t pk[2][8];
        memset(pk, 0, sizeof pk);
        for (int ch = 0; ch < 2; ch++)
            for (int k = 0; k < 61; k++) sf_put(pk[ch], k, scalefac[(g * 2 + ch) * 64 + k]);
        GranuleChan c[2];
        c[0] = make_chan(ug[0].w0, ug[0].w1, ug[0].w2, meta[g * 2]);
        c[1] = make_chan(ug[1].w0, ug[1].w1, ug[1].w2, valid_b ? meta[g * 2 + 1] : 0u);
        float *x[2] = {x0, x1};
        ScaleEnt scale[2][64];
        for (int ch = 0; ch < 2; ch++) {
            if (ch == 1 && !valid_b) break;
            for (int e = 0; e < 64; e++) scale[ch][e] = scale_entry(g_T, c[ch], pk[ch], e);
            const int16_t *is = is16 + (g * 2 + ch) * 576;
            const int npair = c[ch].cnt1 >> 1;
            for (int p = 0; p < 288; p++) {
                int d0, d1;
                const int e = pair_lookup(g_T, cfg, c[ch], p, &d0, &d1);
                float v0 = 0.0f, v1 = 0.0f;
                if (p < npair) {
                    v0 = requant_value(g_T, scale[ch][e], is[2 * p]);
                    v1 = requant_value(g_T, scale[ch][e], is[2 * p + 1]);
                }
                x[ch][d0] = v0;
                x[ch][d1] = v1;
            }
        }
        if (valid_b && u_mode(c[0].w2) == 1) {
            const int mode_ext = u_modeext(c[0].w2);
            if (mode_ext & 2) {
                const int max_pos = c[0].cnt1 > c[1].cnt1 ? c[0].cnt1 : c[1].cnt1;
                const float inv_sqrt2 = 0.70710678118654752440f;
                for (int i = 0; i < max_pos; i++) {
                    float a = x0[i], b = x1[i];
                    x0[i] = f_mul(f_add(a, b), inv_sqrt2);
                    x1[i] = f_mul(f_sub(a, b), inv_sqrt2);
                }
            }
            if (mode_ext & 1) {
                int isp[64];
                for (int e = 0; e < 64; e++) isp[e] = intensity_entry(g_T, cfg, c[0], pk[0], c[1].cnt1, e);
                for (int p = 0; p < 288; p++) {
                    int d0, d1;
                    const int is_pos = isp[pair_lookup(g_T, cfg, c[0], p, &d0, &d1)];
                    if (is_pos < 7) {
                        for (int i = 2 * p; i < 2 * p + 2; i++) {
                            x0[i] = f_mul(x0[i], g_T.is_ratio_l[is_pos]);
                            x1[i] = f_mul(x1[i], g_T.is_ratio_r[is_pos]);
                        }
                    }
                }
            }
        }
        for (int ch = 0; ch < 2; ch++) {
            if (ch == 1 && !valid_b) break;
            const int nb = alias_butterflies(c[ch]);
            for (int b = 0; b < nb; b++) alias_butterfly(g_T.cs, g_T.ca, x[ch], b);
        }
    }
}

// Table access for pinning the device tables against the oracle's.
const float *emu_table(int which, int *n) {
    ensure();
    switch (which) {
    case 0: *n = 18 * 36; return g_h->cos36;
    case 1: *n = 6 * 12; return g_h->cos12;
    case 2: *n = 4 * 36; return g_h->imdct_win;
    case 3: *n = 64 * 32; return g_h->synth_n;
    case 4: *n = 512; return g_h->synth_d;
    }
    *n = 0;
    return nullptr;
}
const float *emu_powq4(int *n) {
    ensure();
    *n = (int)g_h->powq4.size();
    return g_h->powq4.data();
}
const double *emu_powtab34(int *n) {
    ensure();
    *n = (int)g_h->powtab34.size();
    return g_h->powtab34.data();
}
int emu_huff_lut_size() {
    ensure();
    return (int)g_h->huff_lut.size();
}
// Decode one code word of `table` from a bit string with the LUT path; returns bits consumed.
int emu_huff_one(int table, const uint8_t *buf, int len_bytes, int *out4) {
    ensure();
    std::vector<uint8_t> padded((size_t)((len_bytes + 3) & ~3) + 64, 0);
    memcpy(padded.data(), buf, (size_t)len_bytes);
    BitCursor bc;
    bc.init(padded.data(), (uint64_t)len_bytes * 8, 0, len_bytes * 8);
    if (table < 32) {
        const uint32_t desc = g_T.huff_desc[table];
        uint32_t r = huff_pair(SmemRef::of(g_T.huff_lut).plus(desc & 0xffffffu), [&] { return (int)(desc >> 24); }, bc);
        out4[0] = (int16_t)(r & 0xffff);
        out4[1] = (int16_t)(r >> 16);
        out4[2] = out4[3] = 0;
    } else {
        uint32_t vw, xy;
        huff_quad(g_T.quad_lut, g_T.quad_signs, g_T.huff_desc[table] & 0xffffffu, bc, vw, xy);
        out4[0] = (int16_t)(xy & 0xffff); out4[1] = (int16_t)(xy >> 16);
        out4[2] = (int16_t)(vw & 0xffff); out4[3] = (int16_t)(vw >> 16);
    }
    return bc.pos();
}

// Diagnostic (tools/k1_pair_stats.py): code-length statistics of the big_values pairs of the given units.
// hist[l] = pairs whose code (tree bits + sign bits + linbits) is l bits long (l < 64); two[b] = number of positions
// where this pair and the next one (same region, both without linbits escape) together take <= b bits, b < 17;
// n_lookups[b] = lookups a decoder needs that takes two such pairs at once whenever they fit into b bits.
void emu_pair_stats(const uint8_t *main_data, unsigned long long main_bits, const mp3gpu_unit *units, long long n_units,
                    long long *hist, long long *n_lookups, long long *n_pairs_out) {
    ensure();
    long long n_pairs = 0;
    for (long long ui = 0; ui < n_units; ui++) {
        const mp3gpu_unit u = units[ui];
        if (!u_valid(u.w2) || u_p23len(u.w0) == 0) continue;
        // position of the first big_values bit: decode the scalefactors by running the unit and re-deriving — cheaper: rerun the
        // unit logic with a cursor and stop before part 3 is not exposed, so walk part 2 with the same helpers
        uint32_t pk[8];
        alignas(16) uint32_t out[288 + 4];
        (void)huffman_unit(g_T, SmemRef::of(g_T.huff_lut), g_T.quad_lut, g_T.huff_desc, g_T.quad_signs, main_data, main_bits, units, ui, pk, out);
        const HuffRegions R = huff_regions(g_T, SmemRef::of(g_T.huff_lut), g_T.huff_desc, u.w0, u.w1, u.w2);
        // code lengths from the decoded values: tree length from a table walk over all code words of the region's table
        std::vector<int> lens((size_t)R.nbig);
        for (int k = 0; k < R.nbig; k++) {
            const int reg = k < R.r1h ? 0 : (k < R.r2h ? 1 : 2);
            const int tsel = u_tsel(u.w1, reg);
            const int lin = (int)((R.lin >> (4 * reg)) & 0xf);
            const int x = (int16_t)(out[k] & 0xffff), y = (int16_t)(out[k] >> 16);
            const int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
            const int cx = ax > 15 ? 15 : ax, cy = ay > 15 ? 15 : ay;
            int l = emu_code_len(tsel, lin ? cx : ax, lin ? cy : ay);
            if (l < 0) { lens[(size_t)k] = 63; continue; }
            l += (ax != 0) + (ay != 0);
            if (lin && cx == 15) l += lin;
            if (lin && cy == 15) l += lin;
            lens[(size_t)k] = l | (((lin && (cx == 15 || cy == 15)) ? 1 : 0) << 8) | (reg << 12);
        }
        for (int k = 0; k < R.nbig; k++) hist[std::min(lens[(size_t)k] & 0xff, 63)]++;
        n_pairs += R.nbig;
        for (int b = 0; b < 17; b++) {
            long long n = 0;
            for (int k = 0; k < R.nbig;) {
                n++;
                const int l0 = lens[(size_t)k] & 0xff;
                if (k + 1 < R.nbig && (lens[(size_t)k] >> 8) == (lens[(size_t)k + 1] >> 8) && !((lens[(size_t)k] >> 8) & 1) &&
                    l0 + (lens[(size_t)k + 1] & 0xff) <= b)
                    k += 2;
                else
                    k += 1;
            }
            n_lookups[b] += n;
        }
    }
    *n_pairs_out = n_pairs;
}
}
