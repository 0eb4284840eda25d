// unit_logic.h — per-unit / per-line logic of K1 (scalefactors + Huffman) and K2 (requantise,
// reorder, stereo, alias reduction), written as __host__ __device__ functions.
//
// The CUDA kernels in kernels.cuh call these from device code.  tests/hostemu compiles the very
// same functions for the host so that `-m "not gpu"` tests can check the bit-level logic against
// the oracle without a GPU.  The host build is test infrastructure only: no product entry point
// calls it, and the C ABI has no CPU path.
//
// Reference being restated (paths relative to the reference root):
//   internal/bits/bits.go:45-86, internal/huffman/huffman.go:348-419,
//   internal/maindata/huffman.go:27-138, internal/maindata/maindata.go:119-288,
//   internal/frame/frame.go:140-452
#pragma once
#include <stdint.h>

#include "../../include/mp3gpu.h"

#if defined(__CUDACC__)
#define MP3_HD __host__ __device__ __forceinline__
#else
#define MP3_HD inline
struct alignas(16) uint4 { uint32_t x, y, z, w; };  // host-emulation build only (tests/hostemu)
#endif

namespace mp3gpu {

// Tables indexed per lane (divergent): global memory on the device, plain memory in hostemu.
struct DeviceTables {
    const double *pow2q;            // [kPow2N], index 4*idx + pow2_off
    const double *powtab34;         // [8207]
    const float *powq4;             // [4][kPowRow] float32(2^(q/4) * powtab34[v]), tables.cc
    const uint8_t *line_sfb_long;   // [6][576]
    const uint8_t *line_sfb_short;  // [6][576]
    const uint8_t *line_win_short;  // [6][576]
    const uint16_t *reorder_dst;    // [6][576]
    const uint8_t *pair_long;       // [6][288]  sfb of lines 2p, 2p+1 (all sfb boundaries are even)
    const uint8_t *pair_short;      // [6][288]  sfb*3 + win of lines 2p, 2p+1 in the window-major short layout
    const uint16_t *pair_dst;       // [6][288]  reorder destination of line 2p (line 2p+1 goes to +3)
    const uint16_t *sfb_long;       // [6][24]
    const uint16_t *sfb_short;      // [6][16]
    const uint16_t *nslen2;         // [512]
    const uint32_t *huff_lut;       // LUT entries (see tables.h)
    const uint32_t *huff_desc;      // [34]
    const uint64_t *quad_signs;     // [256] count1 sign expansion
    const float *is_ratio_l;        // [8]
    const float *is_ratio_r;        // [8]
    const uint8_t *pretab;          // [24]
    const uint8_t *sfsize_mpeg2;    // [3][6][4]
    const uint8_t *slen_mpeg1;      // [16][2]
    const float *cs;                // [8]
    const float *ca;                // [8]
    uint64_t pretab_pack;           // pretab[sfb] in bits 2*sfb .. 2*sfb+1
    int huff_lut_n;
    int pow2_off;
};

MP3_HD int imin(int a, int b) { return a < b ? a : b; }

// ---- side-info field accessors ---------------------------------------------------------------
MP3_HD int u_p23len(uint32_t w0) { return (int)(w0 & 0xfff); }
MP3_HD int u_bigval(uint32_t w0) { return (int)((w0 >> 12) & 0x1ff); }
MP3_HD int u_ggain(uint32_t w0) { return (int)((w0 >> 21) & 0xff); }
MP3_HD int u_winsw(uint32_t w0) { return (int)((w0 >> 29) & 1); }
MP3_HD int u_btype(uint32_t w0) { return (int)((w0 >> 30) & 3); }
MP3_HD int u_sfcomp(uint32_t w1) { return (int)(w1 & 0x1ff); }
MP3_HD int u_tsel(uint32_t w1, int r) { return (int)((w1 >> (9 + 5 * r)) & 0x1f); }
MP3_HD int u_reg0(uint32_t w1) { return (int)((w1 >> 24) & 0xf); }
MP3_HD int u_reg1(uint32_t w1) { return (int)((w1 >> 28) & 0xf); }
MP3_HD int u_sbg(uint32_t w2, int w) { return (int)((w2 >> (3 * w)) & 7); }
MP3_HD int u_preflag(uint32_t w2) { return (int)((w2 >> 9) & 1); }
MP3_HD int u_sfscale(uint32_t w2) { return (int)((w2 >> 10) & 1); }
MP3_HD int u_c1tsel(uint32_t w2) { return (int)((w2 >> 11) & 1); }
MP3_HD int u_scfsi(uint32_t w2) { return (int)((w2 >> 12) & 0xf); }
MP3_HD int u_lsf(uint32_t w2) { return (int)((w2 >> 16) & 1); }
MP3_HD int u_sfreq(uint32_t w2) { return imin((int)((w2 >> 17) & 3), 2); }  // 3 is reserved (frameheader.go:178); clamped so table indices stay in range
MP3_HD int u_mode(uint32_t w2) { return (int)((w2 >> 19) & 3); }
MP3_HD int u_modeext(uint32_t w2) { return (int)((w2 >> 21) & 3); }
MP3_HD int u_gr(uint32_t w2) { return (int)((w2 >> 23) & 1); }
MP3_HD bool u_valid(uint32_t w2) { return (w2 & MP3GPU_W2_VALID) != 0; }
MP3_HD bool u_zero(uint32_t w2) { return (w2 & MP3GPU_W2_ZERO_STATE) != 0; }
MP3_HD int u_mixed(uint32_t w2) { return (int)((w2 >> 27) & 1); }


// ---- rounding-exact float helpers (no contraction on either side) -----------------------------
#if defined(__CUDA_ARCH__)
MP3_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
MP3_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
MP3_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
MP3_HD float d_mul_to_f(double a, double b) { return __double2float_rn(__dmul_rn(a, b)); }
MP3_HD uint32_t load_raw32(const uint32_t *p) { return __ldg(p); }
MP3_HD uint32_t be32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
// 0xffffffff >> min(n, 32) with n taken as unsigned: n >= 32 and negative n give 0
MP3_HD uint32_t ones_shr_clamp(int n) { return __funnelshift_rc(0xffffffffu, 0u, (uint32_t)n); }
#else
MP3_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
MP3_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
MP3_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
MP3_HD float d_mul_to_f(double a, double b) { volatile double r = a * b; return (float)r; }
MP3_HD uint32_t load_raw32(const uint32_t *p) { return *p; }
MP3_HD uint32_t be32(uint32_t v) { return __builtin_bswap32(v); }
MP3_HD uint32_t ones_shr_clamp(int n) { return (uint32_t)n >= 32u ? 0u : 0xffffffffu >> n; }
#endif

// ---- bit cursor with bits.go semantics ----------------------------------------------------------
// Window = two consecutive big-endian 32-bit words [w0:w1] of main_data and a bit offset into w0; peek32() is one
// funnel shift.  Bits at/after the frame's logical buffer end read as 0 (bits.go:46-49,65-68).
#if defined(__CUDA_ARCH__)
MP3_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, int s) { return __funnelshift_l(lo, hi, s); }  // (hi:lo << s) >> 32, s in 0..31
#else
MP3_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, int s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
#endif

struct BitCursor {
    const uint32_t *wp;     // next 32-bit word of main_data to prefetch
    uint32_t w0, w1;        // current window (big-endian bit order, masked at the buffer end)
    uint32_t nxt;           // the word after w1 as loaded (raw byte order, unmasked; 0 when at/after the buffer end)
    int off;                // bit offset of the cursor inside w0 (0..31)
    int rem;                // bits from the MSB of the nxt word to the end of the frame's logical buffer (<= 0: past the end)
    int end_rel;            // buf_end_rel
    int lim;                // max(buf_end_rel, 0): the logical position never advances past it (bits.go:46-49)

    // main_bits = 8 * main_data_len: a descriptor that points outside main_data (a caller's bug, never produced by the
    // host stage) is clipped to it instead of reading out of bounds.
    MP3_HD void init(const uint8_t *main_data, uint64_t main_bits, uint64_t bit_start, int buf_end_rel) {
        if (bit_start > main_bits) bit_start = main_bits;
        const uint64_t room = main_bits - bit_start;
        if (buf_end_rel > 0 && (uint64_t)buf_end_rel > room) buf_end_rel = (int)room;
        wp = reinterpret_cast<const uint32_t *>(main_data) + (bit_start >> 5);
        off = (int)(bit_start & 31);
        rem = off + buf_end_rel;
        end_rel = buf_end_rel;
        lim = buf_end_rel > 0 ? buf_end_rel : 0;
        w1 = 0;
        fetch();
        shift_in();
        shift_in();
    }
    // Prefetch: nxt <- the word at wp (raw), or 0 at/after the buffer end.  The load is not consumed before the next
    // refill, a full word of code bits later, so its latency stays off the decode chain.
    MP3_HD void fetch() {
        nxt = 0;
        if (rem > 0) nxt = load_raw32(wp);
        wp++;
    }
    // w0 <- w1, w1 <- nxt in big-endian bit order with the bits at/after the buffer end forced to zero, then prefetch.
    MP3_HD void shift_in() {
        w0 = w1;
        w1 = be32(nxt) & ~ones_shr_clamp(rem);
        rem -= 32;
        fetch();
    }
    MP3_HD uint32_t peek32() const { return funnel_l(w0, w1, off); }  // next 32 bits, MSB first
    // Logical position relative to bit_start (BitPos() rebased to part2Start).  The window has consumed
    // end_rel - 64 - rem + off bits since init (rem drops by 32 per word shifted in, two of them in init); the
    // reference's cursor is that, clamped at the buffer end: Bit() past the end returns 0 without advancing, and a
    // refused Bits(n) moves neither.
    MP3_HD int pos() const { return imin(end_rel - 64 - rem + off, lim); }
    // Bit()-style consumption of n <= 32 bits (tree bits, sign bits).  Written without a branch: lanes of a warp refill
    // at different code words, and a divergent refill branch cost more issue slots than the predicated form.
    MP3_HD void skip(int n) {
        off += n;
        if (off >= 32) {
            off -= 32;
            shift_in();
        }
    }
    // Bits(n), bits.go:58-77: returns 0 WITHOUT advancing when the read would cross the end.
    MP3_HD int bits(int n) {
        if (n == 0) return 0;
        if (pos() + n > lim) return 0;
        int v = (int)(peek32() >> (32 - n));
        skip(n);
        return v;
    }
    MP3_HD int bit() {
        int v = (int)(peek32() >> 31);
        skip(1);
        return v;
    }
};

// ---- Huffman code words (LUT entry layout: tables.h) -------------------------------------------------------------
// the top (n mod 32) bits of x, as a number
#if defined(__CUDA_ARCH__)
MP3_HD uint32_t hi_bits_mod32(uint32_t x, uint32_t n) { return __funnelshift_l(x, 0u, n); }
#else
MP3_HD uint32_t hi_bits_mod32(uint32_t x, uint32_t n) { n &= 31; return n ? x >> (32 - n) : 0u; }
#endif
constexpr int kRootBits = 8;  // == tables.h kHuffRootBits
MP3_HD uint32_t lut_at(const uint32_t *lut, uint32_t byte_off) {
    return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(lut) + byte_off);
}
// Leaf entry of the code word at the head of w (MSB first); d = byte offset of the tree's root table.
MP3_HD uint32_t huff_lookup(const uint32_t *lut, uint32_t d, uint32_t w) {
    uint32_t e = lut_at(lut, d + ((w >> (32 - kRootBits)) << 2));
    if ((int32_t)e < 0) {  // code longer than the root index: one sub-table resolves the rest (tables.cc)
        const uint32_t idx = (e & 0xffff) + hi_bits_mod32(w << kRootBits, e >> 16);
        e = lut_at(lut, d + (idx << 2));
    }
    return e;
}
// x << (n mod 32): the LUT's shift-count fields are used without masking their neighbours off
#if defined(__CUDA_ARCH__)
MP3_HD uint32_t shl_mod32(uint32_t x, uint32_t n) { return __funnelshift_l(0u, x, n); }  // the funnel shift wraps its count itself
#else
MP3_HD uint32_t shl_mod32(uint32_t x, uint32_t n) { return funnel_l(x, 0u, (int)(n & 31)); }
#endif

// One big_values pair (huffman.go:404-416).  Returns x | y<<16 (int16 halves).  linbits() is only evaluated for an
// escape (x or y == 15 in a table with linbits).
template <class LinbitsFn>
MP3_HD uint32_t huff_pair(const uint32_t *lut, uint32_t d, LinbitsFn linbits_of, BitCursor &bc) {
    const uint32_t w = bc.peek32();
    const uint32_t e = huff_lookup(lut, d, w);
    int x = (int)(e & 0xf), y = (int)((e >> 8) & 0xf);
    if (e & 0x10u) {
        // Escape: x-linbits, x-sign, y-linbits, y-sign (huffman.go:405-416) are at most 2 * 13 + 2 = 28 bits, one window.
        // The reference reads them with Bits(n) / Bit(), which refuse to read (return 0, do not advance) at the end of
        // the frame's buffer (bits.go:45-77); p is that logical cursor.
        const int linbits = linbits_of();
        bc.skip((int)((e >> 16) & 0x1f));
        const uint32_t v = bc.peek32();
        const int p0 = bc.pos(), lim = bc.lim;
        int p = p0;
        if (x == 15 && p + linbits <= lim) { x += (int)((v << (p - p0)) >> (32 - linbits)); p += linbits; }
        if (x != 0 && p < lim) { if ((v << (p - p0)) >> 31) x = -x; p++; }
        if (y == 15 && p + linbits <= lim) { y += (int)((v << (p - p0)) >> (32 - linbits)); p += linbits; }
        if (y != 0 && p < lim) { if ((v << (p - p0)) >> 31) y = -y; p++; }
        bc.skip(p - p0);
    } else {
        // x's sign bit (if x != 0) follows the tree bits, y's (if y != 0) is the last of the `total` bits; the tree is
        // at most 19 bits, so both lie inside w.  A zero value ignores the mask it gets: (0 ^ m) - m == 0.
        const int mx = (int32_t)shl_mod32(w, e >> 16) >> 31;
        const int my = (int32_t)shl_mod32(w, e >> 21) >> 31;
        x = (x ^ mx) - mx;
        y = (y ^ my) - my;
        bc.skip((int)(e >> 26));
    }
    return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16);
}

// One count1 quadruple (huffman.go:387-403): v, w, x, y each in {-1, 0, 1}; returns (v | w<<16), (x | y<<16).
// quad_signs[pattern << 4 | next four bits] holds both words with the sign bits dealt out (tables.cc).
MP3_HD void huff_quad(const uint32_t *lut, const uint64_t *quad_signs, uint32_t d, BitCursor &bc, uint32_t &vw, uint32_t &xy) {
    const uint32_t wd = bc.peek32();
    const uint32_t e = huff_lookup(lut, d, wd);                   // <= 6 tree bits, then up to 4 sign bits
    const uint32_t four = shl_mod32(wd, e >> 16) >> 28;
    const uint64_t r = quad_signs[((e & 0xf) << 4) | four];
    bc.skip((int)(e >> 26));
    vw = (uint32_t)r;
    xy = (uint32_t)(r >> 32);
}

// Collects decoded line pairs and writes them 16 bytes at a time: units sit 1,152 bytes apart, so every lane of a
// store instruction hits its own sector and the instruction count is what costs.
struct PairSink {
    uint4 *dst;
    uint32_t a0, a1, a2, a3;
    int n;
    MP3_HD void init(uint32_t *out, int n0 = 0) { dst = reinterpret_cast<uint4 *>(out); a0 = a1 = a2 = a3 = 0; n = n0; }  // n0 % 4 == 0
    MP3_HD void put(uint32_t w) {
        a0 = a1; a1 = a2; a2 = a3; a3 = w;
        n++;
        if ((n & 3) == 0) {
            uint4 v; v.x = a0; v.y = a1; v.z = a2; v.w = a3;
            dst[(n >> 2) - 1] = v;
        }
    }
    MP3_HD void flush() {  // pad with zero pairs up to the next multiple of four (they lie above count1)
        while (n & 3) put(0);
    }
};

MP3_HD void sf_put(uint32_t *pk, int n, int v) { pk[n >> 3] |= (uint32_t)v << (4 * (n & 7)); }
MP3_HD int sf_nib(const uint32_t *pk, int n) { return (int)((pk[n >> 3] >> (4 * (n & 7))) & 0xf); }
// Sequential nibble writer: a word is stored once, when it is complete (or at flush), so that K1's scalefactor loops
// carry no load-modify-store chain through the (local-memory) pk array.  pk must start zeroed, and every word is
// written by one run of consecutive nibbles only (true for all callers: long 0..20, mixed 0..7 then 31..57, short 22..57).
struct NibWriter {
    uint32_t *pk;
    uint32_t acc;
    int n;
    MP3_HD void init(uint32_t *p, int n0) { pk = p; acc = 0; n = n0; }
    MP3_HD void put(int v) {
        acc |= (uint32_t)v << (4 * (n & 7));
        n++;
        if ((n & 7) == 0) {
            pk[(n >> 3) - 1] = acc;
            acc = 0;
        }
    }
    MP3_HD void flush() {
        if (n & 7) pk[n >> 3] = acc;
        acc = 0;
    }
    MP3_HD void seek(int n1) { flush(); n = n1; }
};

// Scalefactors of an MPEG-1 unit that reads all of them itself: gr 0, or gr 1 short blocks, or
// gr 1 with no scfsi band set (maindata.go:204-232 and the read arms of :233-279).
MP3_HD void sf_mpeg1_read_all(const DeviceTables &T, BitCursor &bc, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t *pk) {
    int sfc = u_sfcomp(w1) & 15;
    int slen1 = T.slen_mpeg1[sfc * 2], slen2 = T.slen_mpeg1[sfc * 2 + 1];
    NibWriter nw;
    nw.init(pk, 0);
    if (u_winsw(w0) == 1 && u_btype(w0) == 2) {
        int sfb0 = 0;
        if (u_mixed(w2)) {
#pragma unroll 1
            for (int sfb = 0; sfb < 8; sfb++) nw.put(bc.bits(slen1));
            sfb0 = 3;
        }
        nw.seek(22 + sfb0 * 3);
#pragma unroll 1
        for (int n = sfb0 * 3; n < 36; n++) nw.put(bc.bits(n < 18 ? slen1 : slen2));
    } else {
#pragma unroll 1
        for (int sfb = 0; sfb < 21; sfb++) nw.put(bc.bits(sfb < 11 ? slen1 : slen2));
    }
    nw.flush();
}

// K1 for one unit, in two stages so that a CTA can re-deal its units to threads in between (kernels.cuh):
//   stage A  scalefactors + the big_values pairs in whole groups of four   (cost ~ big_values)
//   stage B  the last big_values % 4 pairs + the count1 quadruples         (cost ~ bits left in part 3)
// `units` is the whole submission (absolute indexing): a gr-1 unit with scfsi set re-reads gr 0's scalefactor
// bits (maindata.go:239-278, quirk Q14).
// Outputs: pk[8] scalefactor nibbles (n = sfb for scalefac_l, 22 + sfb*3 + win for scalefac_s), is_out[0..count1/2)
// packed int16 pairs, meta = count1 | preflag << 10.  Lines >= count1 are zero by definition; K2 masks them
// instead of K1 writing zeros.
struct HuffRegions {  // region starts in PAIRS (all sfb boundaries are even) and the three trees
    int r1h, r2h, nbig;
    uint32_t d0, d1, d2;  // byte offset of the root table of each region's tree
    uint32_t lin;         // linbits of the three regions, 4 bits each
};
MP3_HD HuffRegions huff_regions(const DeviceTables &T, const uint32_t *huff_desc, uint32_t w0, uint32_t w1, uint32_t w2) {
    HuffRegions R;
    if (u_winsw(w0) == 1 && u_btype(w0) == 2) {
        R.r1h = 18;
        R.r2h = 288;
    } else {
        const uint16_t *l = T.sfb_long + (u_lsf(w2) * 3 + u_sfreq(w2)) * 24;
        int i = u_reg0(w1) + 1;  // <= 16 < 23
        R.r1h = l[i] >> 1;
        int j = u_reg0(w1) + u_reg1(w1) + 2;
        R.r2h = j >= 23 ? 288 : (l[j] >> 1);
    }
    R.nbig = u_bigval(w0);
    if (R.nbig > 288) R.nbig = 288;  // the host rejects such frames (huffman.go:68-70); never reached
    const uint32_t e0 = huff_desc[u_tsel(w1, 0)], e1 = huff_desc[u_tsel(w1, 1)], e2 = huff_desc[u_tsel(w1, 2)];
    R.d0 = e0 & 0xffffffu;
    R.d1 = e1 & 0xffffffu;
    R.d2 = e2 & 0xffffffu;
    R.lin = (e0 >> 24) | ((e1 >> 24) << 4) | ((e2 >> 24) << 8);
    return R;
}
MP3_HD uint32_t huff_pair_at(const uint32_t *lut, const HuffRegions &R, int k, BitCursor &bc) {
    const bool in0 = k < R.r1h, in1 = k < R.r2h;
    return huff_pair(lut, in0 ? R.d0 : (in1 ? R.d1 : R.d2),
                     [&] { return (int)((R.lin >> (in0 ? 0 : (in1 ? 4 : 8))) & 0xf); }, bc);
}

// State handed from stage A to stage B: bits 0..29 logical cursor position (relative to bit_start), bit 30 preflag,
// bit 31 = nothing left to do (part2_3_length == 0, quirk Q1: nothing decoded, Count1 stays 0).
constexpr uint32_t kHuffDone = 0x80000000u;
MP3_HD int huff_state_pos(uint32_t st) { return (int)(st & 0x3fffffffu); }
MP3_HD int huff_state_preflag(uint32_t st) { return (int)((st >> 30) & 1); }

MP3_HD uint32_t huffman_stage_a(const DeviceTables &T, const uint32_t *lut, const uint32_t *huff_desc,
                                const uint8_t *main_data, uint64_t main_bits, const mp3gpu_unit *units, long long unit_index,
                                uint32_t *pk, uint32_t *is_out) {
    const mp3gpu_unit u = units[unit_index];
    const uint32_t w0 = u.w0, w1 = u.w1, w2 = u.w2;
    BitCursor bc;
    bc.init(main_data, main_bits, u.bit_start, u.buf_end_rel);
    for (int i = 0; i < 8; i++) pk[i] = 0;

    // ---- part 2: scalefactors -------------------------------------------------------------
    int preflag = u_preflag(w2);
    if (u_lsf(w2)) {
        // maindata.go:132-179
        int slen = T.nslen2[u_sfcomp(w1)];
        preflag = (slen >> 15) & 1;
        int n = 0;
        if (u_btype(w0) == 2) {
            n++;
            if (u_mixed(w2)) n++;
        }
        int d = (slen >> 12) & 7;
        // long blocks fill scalefac_l[idx]; short fill scalefac_s[idx/3][idx%3] (maindata.go:169-179)
        NibWriter nw;
        nw.init(pk, n == 0 ? 0 : 22);
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            int num = slen & 7;
            slen >>= 3;
            int cnt = T.sfsize_mpeg2[(n * 6 + d) * 4 + i];
#pragma unroll 1
            for (int k = 0; k < cnt; k++) {
                int v = num > 0 ? bc.bits(num) : 0;
                if (nw.n < 64) nw.put(v);
            }
        }
        if (nw.n < 64) nw.flush();
    } else if (u_gr(w2) == 0 || (u_winsw(w0) == 1 && u_btype(w0) == 2) || u_scfsi(w2) == 0) {
        sf_mpeg1_read_all(T, bc, w0, w1, w2, pk);
    } else {
        // gr 1, long-type block, at least one scfsi band set: those bands copy ScalefacL[0][ch] as gr 0's parse left it
        // (21 values if gr 0 was a long-type block, sfb 0-7 only if it was mixed, zeros if it was short).  gr 0's
        // scalefactor bits are re-read by a second cursor in step with this unit's own.
        int n0 = 0, s1_0 = 0, s2_0 = 0;
        BitCursor b0;
        b0.init(main_data, main_bits, 0, 0);
        if (unit_index >= 2) {  // a submission always starts on a frame boundary; guard against one that does not
            const mp3gpu_unit u0 = units[unit_index - 2];
            const bool short0 = u_winsw(u0.w0) == 1 && u_btype(u0.w0) == 2;
            n0 = short0 ? (u_mixed(u0.w2) ? 8 : 0) : 21;
            const int sfc0 = u_sfcomp(u0.w1) & 15;
            s1_0 = T.slen_mpeg1[sfc0 * 2];
            s2_0 = T.slen_mpeg1[sfc0 * 2 + 1];
            b0.init(main_data, main_bits, u0.bit_start, u0.buf_end_rel);
        }
        int sfc = u_sfcomp(w1) & 15;
        int slen1 = T.slen_mpeg1[sfc * 2], slen2 = T.slen_mpeg1[sfc * 2 + 1];
        int scfsi = u_scfsi(w2);
        NibWriter nw;
        nw.init(pk, 0);
#pragma unroll 1
        for (int sfb = 0; sfb < 21; sfb++) {
            int band = sfb < 6 ? 0 : (sfb < 11 ? 1 : (sfb < 16 ? 2 : 3));
            const int v0 = sfb < n0 ? b0.bits(sfb < 11 ? s1_0 : s2_0) : 0;
            int v;
            if ((scfsi >> band) & 1)
                v = v0;
            else
                v = bc.bits(sfb < 11 ? slen1 : slen2);
            nw.put(v);
        }
        nw.flush();
    }

    // ---- part 3: Huffman (maindata/huffman.go:27-138), whole groups of four pairs ---------------
    if (u_p23len(w0) == 0) return kHuffDone | ((uint32_t)preflag << 30);
    const HuffRegions R = huff_regions(T, huff_desc, w0, w1, w2);
    uint4 *dst4 = reinterpret_cast<uint4 *>(is_out);
    for (int k = 0; k + 4 <= R.nbig; k += 4) {  // four pairs per 16-byte store
        uint4 v;
        v.x = huff_pair_at(lut, R, k, bc);
        v.y = huff_pair_at(lut, R, k + 1, bc);
        v.z = huff_pair_at(lut, R, k + 2, bc);
        v.w = huff_pair_at(lut, R, k + 3, bc);
        dst4[k >> 2] = v;
    }
    return (uint32_t)bc.pos() | ((uint32_t)preflag << 30);
}

// Bits of part 3 left after stage A: what stage B's cost grows with.
MP3_HD int huff_bits_left(uint32_t w0, uint32_t st) {
    if (st & kHuffDone) return 0;
    const int left = u_p23len(w0) - huff_state_pos(st);
    return left > 0 ? left : 0;
}

MP3_HD uint32_t huffman_stage_b(const DeviceTables &T, const uint32_t *lut, const uint32_t *huff_desc, const uint64_t *quad_signs,
                                const uint8_t *main_data, uint64_t main_bits, const mp3gpu_unit *units, long long unit_index,
                                uint32_t st, uint32_t *is_out) {
    const uint32_t preflag = (uint32_t)huff_state_preflag(st);
    if (st & kHuffDone) return preflag << 10;
    const mp3gpu_unit u = units[unit_index];
    const uint32_t w0 = u.w0, w1 = u.w1, w2 = u.w2;
    // Restart the cursor at the logical position stage A reached.  A cursor that had run past the buffer end reads
    // zeros and reports the end position; so does one restarted exactly there.
    const int pos0 = huff_state_pos(st);
    BitCursor bc;
    bc.init(main_data, main_bits, u.bit_start + (uint64_t)pos0, u.buf_end_rel - pos0);
    const int bit_pos_end = u_p23len(w0) - 1 - pos0;  // part2Start is position -pos0 of this cursor
    const HuffRegions R = huff_regions(T, huff_desc, w0, w1, w2);
    int k = R.nbig & ~3;
    PairSink sink;
    sink.init(is_out, k);
    for (; k < R.nbig; k++) sink.put(huff_pair_at(lut, R, k, bc));
    int is_pos = R.nbig * 2;
    {
        const uint32_t dq = huff_desc[32 + u_c1tsel(w2)] & 0xffffffu;
        while (is_pos <= 572 && bc.pos() <= bit_pos_end) {
            uint32_t vw, xy;
            huff_quad(lut, quad_signs, dq, bc, vw, xy);
            sink.put(vw);
            sink.put(xy);
            is_pos += 4;
        }
    }
    sink.flush();
    if (bc.pos() > bit_pos_end + 1) is_pos -= 4;  // overshoot: drop the last quadruple (huffman.go:119-122)
    if (is_pos < 0) is_pos = 0;
    return (uint32_t)is_pos | (preflag << 10);
}

// Both stages back to back (host emulation and tests).
MP3_HD uint32_t huffman_unit(const DeviceTables &T, const uint32_t *lut, const uint32_t *huff_desc, const uint64_t *quad_signs,
                             const uint8_t *main_data, uint64_t main_bits, const mp3gpu_unit *units, long long unit_index,
                             uint32_t *pk, uint32_t *is_out) {
    const uint32_t st = huffman_stage_a(T, lut, huff_desc, main_data, main_bits, units, unit_index, pk, is_out);
    return huffman_stage_b(T, lut, huff_desc, quad_signs, main_data, main_bits, units, unit_index, st, is_out);
}

// ---- K2 per-line logic ---------------------------------------------------------------------------
struct GranuleChan {  // decoded side info K2 needs for one channel
    uint32_t w0, w1, w2;
    int cnt1, preflag;
    bool is_short, mixed;
};
MP3_HD GranuleChan make_chan(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t meta) {
    GranuleChan c;
    c.w0 = w0; c.w1 = w1; c.w2 = w2;
    c.cnt1 = (int)(meta & 0x3ff);
    c.preflag = (int)((meta >> 10) & 1);
    c.is_short = u_winsw(w0) == 1 && u_btype(w0) == 2;
    c.mixed = c.is_short && u_mixed(w2) != 0;
    return c;
}

// ---- K2, scale-table form ------------------------------------------------------------------------
// The requantisation exponent is constant per scalefactor band (long) or per band x window (short), so K2
// first tabulates 2^(k/4) per band (64 doubles: [0..21] long sfb, [24..62] short sfb*3+win) and then handles
// the lines two at a time: every band boundary is even, so lines 2p and 2p+1 share their table entry.
constexpr int kScaleShortBase = 24;

// Table entry e of one channel (frame.go:146-148 long, :161-166 short): the scale 2^(k4/4), k4 = 4e + q, as the exact
// float 2^e and the row q of powq4 (tables.cc); {0, row 0} where the entry is not used.
constexpr int kPowRowLen = 8208;  // == tables.h kPowRow
struct ScaleEnt {
    float s;
    uint32_t row;
};
MP3_HD float exp2_int(int e) {  // 2^e, -126 <= e <= 127
#if defined(__CUDA_ARCH__)
    return __int_as_float((e + 127) << 23);
#else
    union { uint32_t u; float f; } c;
    c.u = (uint32_t)(e + 127) << 23;
    return c.f;
#endif
}
MP3_HD ScaleEnt scale_entry(const DeviceTables &T, const GranuleChan &c, const uint32_t *pk, int e) {
    const int mult = u_sfscale(c.w2) ? 4 : 2;  // 4 * sfMult
    const int gg = u_ggain(c.w0) - 210;
    ScaleEnt r;
    r.s = 0.0f;
    r.row = 0;
    int k4;
    if (e < 22) {
        if (c.is_short && !c.mixed) return r;
        const int pre = (int)((T.pretab_pack >> (2 * e)) & 3);
        k4 = gg - mult * (sf_nib(pk, e) + c.preflag * pre);
    } else if (e >= kScaleShortBase && e < kScaleShortBase + 39) {
        if (!c.is_short) return r;
        const int code = e - kScaleShortBase, win = code % 3;
        k4 = gg - 8 * u_sbg(c.w2, win) - mult * sf_nib(pk, 22 + code);
    } else {
        return r;
    }
    r.s = exp2_int(k4 >> 2);  // k4 >= -338: far above the float exponent range's lower end
    r.row = (uint32_t)(k4 & 3) * kPowRowLen;
    return r;
}

// Table index and reorder destinations of the pair of lines (2p, 2p+1) (frame.go:257-302).  The three table rows of
// the granule's sampling-rate configuration are passed in (k_hybrid keeps them in shared memory).
struct PairRows {
    const uint8_t *pair_long;    // [288]
    const uint8_t *pair_short;   // [288]
    const uint16_t *pair_dst;    // [288]
};
MP3_HD PairRows pair_rows(const DeviceTables &T, int cfg) {
    PairRows r;
    r.pair_long = T.pair_long + cfg * 288;
    r.pair_short = T.pair_short + cfg * 288;
    r.pair_dst = T.pair_dst + cfg * 288;
    return r;
}
MP3_HD int pair_lookup(const PairRows &R, const GranuleChan &c, int p, int *dst0, int *dst1) {
    if (c.is_short && (!c.mixed || p >= 18)) {
        *dst0 = R.pair_dst[p];
        *dst1 = *dst0 + 3;
        return kScaleShortBase + R.pair_short[p];
    }
    *dst0 = 2 * p;
    *dst1 = 2 * p + 1;
    return R.pair_long[p];
}
MP3_HD int pair_lookup(const DeviceTables &T, int cfg, const GranuleChan &c, int p, int *dst0, int *dst1) {
    return pair_lookup(pair_rows(T, cfg), c, p, dst0, dst1);
}

// One requantised line: sign(is) * float32(|is|^(4/3) * 2^(k4/4)) (frame.go:146-155; the reference multiplies in
// float64 and rounds once — the table row holds that product for 2^(q/4), the power of two is exact).
MP3_HD float requant_value(const DeviceTables &T, ScaleEnt sc, int v) {
    // the sign is applied after the rounding: round-to-nearest is symmetric, so float32(s * -a) == -float32(s * a)
    const float r = f_mul(T.powq4[sc.row + (uint32_t)(v < 0 ? -v : v)], sc.s);
    return v < 0 ? -r : r;
}

// Intensity-stereo table entry e (same indexing as the scale table, channel 0's block type and scalefactors,
// frame.go:312,340,385-419): is_pos if the band is intensity coded, else 7.
MP3_HD int intensity_entry(const DeviceTables &T, int cfg, const GranuleChan &c0, const uint32_t *pk0, int cnt1_r, int e) {
    if (e < 22) {
        if (c0.is_short && !c0.mixed) return 7;
        if (e < (c0.mixed ? 8 : 21) && (int)T.sfb_long[cfg * 24 + e] >= cnt1_r) return sf_nib(pk0, e);
        return 7;
    }
    if (e >= kScaleShortBase && e < kScaleShortBase + 39 && c0.is_short) {
        const int code = e - kScaleShortBase, sfb = code / 3;
        if (sfb < 12 && (int)T.sfb_short[cfg * 16 + sfb] * 3 >= cnt1_r) return sf_nib(pk0, 22 + code);
    }
    return 7;
}

// Intensity-stereo position of line i, or 7 if the line is not intensity coded.
// Channel 0's block type and scalefactors decide (frame.go:312,340,385-419); the short-block
// window index is taken from the pre-reorder layout although the data is already reordered
// (frame.go:341-357) — kept as is.
MP3_HD int intensity_pos(const DeviceTables &T, int cfg, const GranuleChan &c0, const uint32_t *pk0, int cnt1_r, int i) {
    if (c0.is_short && (!c0.mixed || i >= 36)) {
        int sfb = T.line_sfb_short[cfg * 576 + i];
        if (sfb < 12 && (int)T.sfb_short[cfg * 16 + sfb] * 3 >= cnt1_r)
            return sf_nib(pk0, 22 + sfb * 3 + T.line_win_short[cfg * 576 + i]);
    } else {
        int sfb = T.line_sfb_long[cfg * 576 + i];
        if (sfb < (c0.mixed ? 8 : 21) && (int)T.sfb_long[cfg * 24 + sfb] >= cnt1_r) return sf_nib(pk0, sfb);
    }
    return 7;
}

// Number of alias-reduction butterflies for a channel (frame.go:427-440): 0, 8 or 248.
MP3_HD int alias_butterflies(const GranuleChan &c) {
    const bool mixed_flag = u_mixed(c.w2) == 1;
    if (c.is_short && !mixed_flag) return 0;
    return ((c.is_short && mixed_flag) ? 1 : 31) * 8;
}
MP3_HD void alias_butterfly(const float *cs, const float *ca, float *x, int b) {
    int sb = (b >> 3) + 1, i = b & 7;
    int li = 18 * sb - 1 - i, ui = 18 * sb + i;
    float xl = x[li], xu = x[ui];
    x[li] = f_sub(f_mul(xl, cs[i]), f_mul(xu, ca[i]));
    x[ui] = f_add(f_mul(xu, cs[i]), f_mul(xl, ca[i]));
}

}  // namespace mp3gpu
