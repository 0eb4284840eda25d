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

// ---- checked build (-DMP3GPU_CHECKED=1) -------------------------------------------------------------------------
// compute-sanitizer is closed on the GPU pool this was developed on, so memory safety has its own build: every global
// load / store whose index depends on the input is guarded by MP3_CHECK(condition, index).  In the checked build a failed
// guard records the source line and the index (first fault wins) and the access is skipped; the host side turns a
// recorded fault into an error after the call (mp3gpu.cu).  In the product build the guard compiles to `true`.
#ifndef MP3GPU_CHECKED
#define MP3GPU_CHECKED 0
#endif
#if MP3GPU_CHECKED && defined(__CUDACC__)
namespace mp3gpu {
__device__ unsigned int g_fault[4];  // [0] faults, [1] source line of the first, [2..3] its index
__host__ __device__ __forceinline__ bool check_fail(int line, long long idx) {
#if defined(__CUDA_ARCH__)
    if (atomicAdd(&g_fault[0], 1u) == 0u) {
        g_fault[1] = (unsigned int)line;
        g_fault[2] = (unsigned int)(unsigned long long)idx;
        g_fault[3] = (unsigned int)((unsigned long long)idx >> 32);
    }
#else
    (void)line;
    (void)idx;
#endif
    return false;
}
}  // namespace mp3gpu
#define MP3_CHECK(cond, idx) ((cond) ? true : ::mp3gpu::check_fail(__LINE__, (long long)(idx)))
#else
#define MP3_CHECK(cond, idx) (true)
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
    const uint16_t *huff_lut;       // pair-tree LUT entries (see tables.h)
    const uint32_t *quad_lut;       // [2][256] count1-tree LUT entries
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
    int huff_lut_n;                 // uint16 entries, a multiple of 8
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
    static constexpr bool kFast = false;  // no unchecked fast path (see StagedCursor)
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

// ---- staged cursor: the same bits.go semantics over a stretch of main data staged in shared memory ----------------
// k_huffman copies the stretch of main_data its tile of units refers to into shared memory (coalesced 16-byte
// loads, byte-swapped once into big-endian bit order) and every unit reads its code stream from there.  The cursor is
// then nothing but a bit position: peek32() is two shared-memory loads and one funnel shift, skip() is an addition.
// There is no window to refill, which in the register-window cursor above was a divergent branch taken by one
// lane in four at every code word (26 % of K1's issue slots in ncu, profiles/).  Positions outside the staged stretch
// (malformed descriptors, a unit whose buffer reaches further than its tile's stretch, the scfsi look-back of a
// tile's first units) fall back to global loads word by word: rare, slow and exact.
// A read-only array in the kernel's shared memory.  On the device it is held as the 32-bit shared-window address and read
// with explicit ld.shared: handed a generic pointer, the compiler re-derives that address at every use (four instructions
// per code word in K1's inner loop).  In the host emulation it is a plain pointer.
#if defined(__CUDA_ARCH__)
struct SmemRef {
    uint32_t a;
    MP3_HD static SmemRef of(const void *p) { SmemRef r; r.a = (uint32_t)__cvta_generic_to_shared(p); return r; }
    MP3_HD SmemRef plus(uint32_t bytes) const { SmemRef r; r.a = a + bytes; return r; }
    MP3_HD int lds16(uint32_t off) const {  // sign-extended
        int v;
        asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a + off));
        return v;
    }
    MP3_HD uint32_t ld32(uint32_t off) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + off));
        return v;
    }
    MP3_HD void ld32x2(uint32_t off, uint32_t &w0, uint32_t &w1) const {
        asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "r"(a + off));
    }
};
#else
struct SmemRef {
    const uint8_t *a;
    MP3_HD static SmemRef of(const void *p) { SmemRef r; r.a = static_cast<const uint8_t *>(p); return r; }
    MP3_HD SmemRef plus(uint32_t bytes) const { SmemRef r; r.a = a + bytes; return r; }
    MP3_HD int lds16(uint32_t off) const { int16_t v; __builtin_memcpy(&v, a + off, 2); return v; }
    MP3_HD uint32_t ld32(uint32_t off) const { uint32_t v; __builtin_memcpy(&v, a + off, 4); return v; }
    MP3_HD void ld32x2(uint32_t off, uint32_t &w0, uint32_t &w1) const { w0 = ld32(off); w1 = ld32(off + 4); }
};
#endif

struct StageCtx {
    SmemRef sw;                   // staged words: word i = bits [32 i, 32 i + 32) of the stretch, MSB first
    int n_words;                  // staged words
    unsigned long long lo_word;   // index (in 32-bit words of main_data) of staged word 0
    const uint32_t *gw;           // main_data as words (fallback)
    unsigned long long main_bits;
};
// words of main_data (+ its 64 bytes of tail padding) a cursor may load
MP3_HD unsigned long long stage_readable_words(unsigned long long main_bits) { return ((main_bits >> 3) + 64) >> 2; }
// Stretch of main data one unit may touch, in 16-byte chunks [lo16, hi16): from its first part2 bit to the end of
// part 3 (a part2_3_length of 0 still reads its scalefactors, quirk Q1), never beyond the frame's buffer end.  A unit
// that runs past that (big_values never checks the bit budget, quirk Q4) leaves the stretch and takes the fallback.
MP3_HD void stage_reach(const mp3gpu_unit &u, unsigned long long main_bits, uint32_t *lo16, uint32_t *hi16) {
    unsigned long long start = u.bit_start;
    if (start > main_bits) start = main_bits;
    const int ber = u.buf_end_rel > 0 ? u.buf_end_rel : 0;
    const int want = u_p23len(u.w0) > 256 ? u_p23len(u.w0) : 256;
    const int reach = (ber < want ? ber : want) + 64;
    *lo16 = (uint32_t)(start >> 7);
    *hi16 = (uint32_t)((start + (unsigned)reach + 127) >> 7);
}

struct StagedCursor {
    static constexpr bool kFast = true;
    SmemRef sw;             // staged words, already offset to the word holding the unit's first bit (valid if sidx0 >= 0)
    int n_words_m1;         // a window needs words si and si + 1
    const uint32_t *gbase;  // main_data word holding the unit's first bit (fallback)
    int sidx0;              // staged index of that word; far negative when the unit starts outside the stretch
    int p;                  // cursor: bits from the MSB of that word (virtual: keeps counting past the buffer end)
    int off0;               // p at logical position 0
    int end;                // p of the frame's buffer end; bits at/after it read as 0 (bits.go:46-49,65-68)
    int lim;                // max(buf_end_rel, 0): the logical position never advances past it
    int fast_end;           // min(buffer end, end of the staged words) as a p; far negative when the unit is not staged.  A step
                            // that consumes at most n bits from p <= fast_end - n reads only staged bits in front of the buffer
                            // end: FastWindow is exact there, pos() needs no clamp and no Bits() can be refused
#if MP3GPU_CHECKED
    unsigned long long gword0, gwords;  // index of the word at gbase, and the words that may be loaded
    int staged_words_left;              // staged words from the word at sw on
#endif

    MP3_HD void init(const StageCtx &S, unsigned long long bit_start, int buf_end_rel) {
        if (bit_start > S.main_bits) bit_start = S.main_bits;  // same clipping as BitCursor::init
        const unsigned long long room = S.main_bits - bit_start;
        if (buf_end_rel > 0 && (unsigned long long)buf_end_rel > room) buf_end_rel = (int)room;
        const unsigned long long word = bit_start >> 5;
        n_words_m1 = S.n_words > 0 ? S.n_words - 1 : 0;
        gbase = S.gw + word;
        const long long d = (long long)word - (long long)S.lo_word;
        sidx0 = (d >= 0 && d < (long long)S.n_words) ? (int)d : -(1 << 30);
        sw = S.sw.plus(sidx0 >= 0 ? (uint32_t)sidx0 * 4u : 0u);
        off0 = (int)(bit_start & 31);
        p = off0;
        end = off0 + buf_end_rel;
        lim = buf_end_rel > 0 ? buf_end_rel : 0;
#if MP3GPU_CHECKED
        gword0 = word;
        gwords = stage_readable_words(S.main_bits);
        staged_words_left = sidx0 >= 0 ? S.n_words - sidx0 : 0;
#endif
        fast_end = -(1 << 30);
        if (sidx0 >= 0) {
            const int staged_end = (S.n_words - sidx0) * 32;  // p of the first bit behind the staged words
            fast_end = end < staged_end ? end : staged_end;
        }
    }
    MP3_HD uint32_t peek32() const {  // next 32 bits, MSB first
        const int idx = p >> 5, si = sidx0 + idx;
        uint32_t w0 = 0, w1 = 0;
        if ((uint32_t)si < (uint32_t)n_words_m1) {
            sw.ld32x2((uint32_t)idx * 4u, w0, w1);
        } else if (p < end) {  // outside the stretch: the two words straight from main_data (a word holding a bit below the buffer end lies inside main_data; the next one inside its tail padding)
            if (MP3_CHECK(idx >= 0 && gword0 + (unsigned long long)idx + 1 < gwords, gword0 + (unsigned long long)idx)) {
                w0 = be32(load_raw32(gbase + idx));
                w1 = be32(load_raw32(gbase + idx + 1));
            }
        }
        const int rem = end - p;  // bits left before the buffer end
        return funnel_l(w0, w1, p) & ~ones_shr_clamp(rem > 0 ? rem : 0);
    }
    MP3_HD int pos() const { return imin(p - off0, lim); }
    MP3_HD void skip(int n) { p += n; }
    MP3_HD int bits(int n) {  // Bits(n), bits.go:58-77: returns 0 WITHOUT advancing when the read would cross the end
        if (n == 0) return 0;
        if (pos() + n > lim) return 0;
        const int v = (int)(peek32() >> (32 - n));
        p += n;
        return v;
    }
    MP3_HD int bit() {
        const int v = (int)(peek32() >> 31);
        p += 1;
        return v;
    }
};

// Register window over the staged stretch for the unchecked fast loops (see StagedCursor::fast_end): three consecutive
// words, the third a prefetch, so that no shared-memory load sits on the code-word-to-code-word dependency chain —
// the next code word's position depends on this one's length, and with the window read from shared memory at every
// step K1 was bound by that latency (ncu: issue slots 51 % busy, short-scoreboard 3.9 stall cycles per issue).
// A step consumes fewer than 32 bits, so one refill per advance() is enough; the refill is predicated, not a branch.
// The stretch is followed by four words of padding: with the cursor inside the stretch, the window and its prefetch
// reach at most that far past it (loaded, never consumed).
struct FastWindow {
    uint32_t w0, w1, w2;
    SmemRef next;  // the word after w2
    int off;       // cursor inside w0
    int p;         // the StagedCursor's p, kept alongside for the loop bounds
#if MP3GPU_CHECKED
    int words_left;  // staged words + padding from `next` on
#endif
    MP3_HD void open(const StagedCursor &bc) {
        const uint32_t byte = (uint32_t)(bc.p >> 5) * 4u;
#if MP3GPU_CHECKED
        words_left = bc.staged_words_left + 4 - (bc.p >> 5) - 3;
        if (!MP3_CHECK(bc.p >= 0 && words_left >= 0, bc.p)) { w0 = w1 = w2 = 0; next = bc.sw; off = 0; p = bc.p; words_left = 0; return; }
#endif
        bc.sw.ld32x2(byte, w0, w1);
        w2 = bc.sw.ld32(byte + 8u);
        next = bc.sw.plus(byte + 12u);
        off = bc.p & 31;
        p = bc.p;
    }
    MP3_HD uint32_t peek() const { return funnel_l(w0, w1, off); }
    MP3_HD void advance(int n) {  // 0 <= n < 32
        off += n;
        p += n;
        if (off >= 32) {
            off -= 32;
            w0 = w1;
            w1 = w2;
#if MP3GPU_CHECKED
            if (!MP3_CHECK(words_left > 0, p)) return;
            words_left--;
#endif
            w2 = next.ld32(0u);
            next = next.plus(4u);
        }
    }
};

// ---- Huffman code words (LUT entry layout: tables.h) -------------------------------------------------------------
// the top (n mod 32) bits of x, as a number
#if defined(__CUDA_ARCH__)
MP3_HD uint32_t hi_bits_mod32(uint32_t x, uint32_t n) { return __funnelshift_l(x, 0u, n); }
#else
MP3_HD uint32_t hi_bits_mod32(uint32_t x, uint32_t n) { n &= 31; return n ? x >> (32 - n) : 0u; }
#endif
#ifndef MP3_HUFF_ROOT_BITS
#define MP3_HUFF_ROOT_BITS 8
#endif
constexpr int kRootBits = MP3_HUFF_ROOT_BITS;  // == tables.h kHuffRootBits (pair trees)
constexpr int kQuadBits = 8;                   // == tables.h kQuadRootBits (count1 trees)
// x << (n mod 32): the LUT's shift-count fields are used without masking their neighbours off
#if defined(__CUDA_ARCH__)
MP3_HD uint32_t shl_mod32(uint32_t x, uint32_t n) { return __funnelshift_l(0u, x, n); }  // the funnel shift wraps its count itself
#else
MP3_HD uint32_t shl_mod32(uint32_t x, uint32_t n) { return funnel_l(x, 0u, (int)(n & 31)); }
#endif

MP3_HD uint32_t lut_at(const uint32_t *lut, uint32_t byte_off) {
    return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(lut) + byte_off);
}
// Pair trees: entry (as int16, tables.h) of the code word — sign bits included — at the head of w (MSB first); `tree` is the
// tree's root table.  The result is a leaf (>= 0) or an escape entry (< 0).
MP3_HD int huff_lookup16(SmemRef tree, uint32_t w) {
    int e = tree.lds16((w >> (31 - kRootBits)) & (((1u << kRootBits) - 1u) << 1));
    uint32_t rest = w << kRootBits;
    while (e < -16384) {  // link: the code (with its signs) is longer than the bits indexed so far; three levels at most
        const uint32_t sb = ((uint32_t)e >> 10) & 15u;
        e = tree.lds16((((uint32_t)e & 0x3ffu) << 5) + (hi_bits_mod32(rest, sb) << 1));
        rest = shl_mod32(rest, sb);
    }
    return e;
}
// Signed 5-bit field of a leaf at bit `pos`, and the packed output word (x | y << 16 as int16 halves).
#if defined(__CUDA_ARCH__)
MP3_HD uint32_t leaf_pair(int e) {
    int x, y;
    asm("bfe.s32 %0, %1, 0, 5;" : "=r"(x) : "r"(e));
    asm("bfe.s32 %0, %1, 5, 5;" : "=r"(y) : "r"(e));
    return __byte_perm((uint32_t)x, (uint32_t)y, 0x5410);
}
#else
MP3_HD uint32_t leaf_pair(int e) {
    const int x = (int)((uint32_t)e << 27) >> 27, y = (int)((uint32_t)e << 22) >> 27;
    return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16);
}
#endif

// One big_values pair (huffman.go:404-416).  Returns x | y<<16 (int16 halves).  linbits() is only evaluated for an
// escape (x or y == 15 in a table with linbits).
template <class BC, class LinbitsFn>
MP3_HD uint32_t huff_pair(SmemRef tree, LinbitsFn linbits_of, BC &bc) {
    const int e = huff_lookup16(tree, bc.peek32());
    if (e < 0) {
        // Escape: x-linbits, x-sign, y-linbits, y-sign (huffman.go:405-416) are at most 2 * 13 + 2 = 28 bits, one window.
        // The reference reads them with Bits(n) / Bit(), which refuse to read (return 0, do not advance) at the end of
        // the frame's buffer (bits.go:45-77); p is that logical cursor.
        const int other = (e >> 7) & 15;
        int x = (e & 0x20) ? 15 : other, y = (e & 0x40) ? 15 : other;
        const int linbits = linbits_of();
        bc.skip(e & 31);
        const uint32_t v = bc.peek32();
        const int p0 = bc.pos(), lim = bc.lim;
        int p = p0;
        if (x == 15 && p + linbits <= lim) { x += (int)((v << (p - p0)) >> (32 - linbits)); p += linbits; }
        if (x != 0 && p < lim) { if ((v << (p - p0)) >> 31) x = -x; p++; }
        if (y == 15 && p + linbits <= lim) { y += (int)((v << (p - p0)) >> (32 - linbits)); p += linbits; }
        if (y != 0 && p < lim) { if ((v << (p - p0)) >> 31) y = -y; p++; }
        bc.skip(p - p0);
        return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16);
    }
    // The sign bits were part of the index (bits at/after the buffer end read as 0, like the reference's Bit()); the
    // cursor's logical position clamps at the end by itself.
    bc.skip(e >> 10);
    return leaf_pair(e);
}

// The same pair where nothing can touch the buffer end (FastWindow: the cursor is at p <= fast_end - 47): no refused reads, no
// clamp, no load on the dependency chain.
template <class LinbitsFn>
MP3_HD uint32_t huff_pair_fast(SmemRef tree, LinbitsFn linbits_of, FastWindow &fw) {
    const int e = huff_lookup16(tree, fw.peek());
    if (e < 0) {
        const int other = (e >> 7) & 15;
        int x = (e & 0x20) ? 15 : other, y = (e & 0x40) ? 15 : other;
        const int linbits = linbits_of();  // >= 1 in every table that has escapes
        fw.advance(e & 31);                // at most 19 tree bits
        uint32_t v = fw.peek();            // still in front of fast_end
        int n = 0;
        if (x == 15) { x += (int)(v >> (32 - linbits)); v <<= linbits; n = linbits; }
        if (x != 0) { if ((int32_t)v < 0) x = -x; v <<= 1; n++; }
        if (y == 15) { y += (int)(v >> (32 - linbits)); v <<= linbits; n += linbits; }
        if (y != 0) { if ((int32_t)v < 0) y = -y; n++; }
        fw.advance(n);                     // at most 2 * 13 + 2 = 28 bits
        return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16);
    }
    fw.advance(e >> 10);
    return leaf_pair(e);
}

// One count1 quadruple (huffman.go:387-403): v, w, x, y each in {-1, 0, 1}; returns (v | w<<16), (x | y<<16).
// quad_signs[pattern << 4 | next four bits] holds both words with the sign bits dealt out (tables.cc).
template <class BC>
MP3_HD void huff_quad(const uint32_t *qlut, const uint64_t *quad_signs, uint32_t d, BC &bc, uint32_t &vw, uint32_t &xy) {
    const uint32_t wd = bc.peek32();
    const uint32_t e = lut_at(qlut, d + ((wd >> (32 - kQuadBits)) << 2));  // <= 6 tree bits, then up to 4 sign bits
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
            if (MP3_CHECK((n >> 2) - 1 < 72, n)) dst[(n >> 2) - 1] = v;  // 576 lines = 288 pairs = 72 stores per unit
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

// A run of `count` scalefactors of `slen` bits each (maindata.go:146-162, 207-279).  Bits(n) refuses a read that would
// cross the buffer end (returns 0, does not advance), value by value; when the whole run lies before the end — every
// well-formed stream — nothing can be refused and the values are peeled off 32-bit windows, several per peek.
MP3_HD int sf_per_window(int slen) { return (int)((0x040506080a102000ull >> (8 * slen)) & 0xff); }  // 32 / slen for slen 1..7: 32, 16, 10, 8, 6, 5, 4
template <class BC>
MP3_HD void sf_run(BC &bc, NibWriter &nw, int count, int slen) {
    if (slen == 0 || bc.pos() + count * slen > bc.lim) {
#pragma unroll 1
        for (int i = 0; i < count; i++) nw.put(bc.bits(slen));
        return;
    }
    const int per = sf_per_window(slen);
#pragma unroll 1
    while (count > 0) {
        const int m = count < per ? count : per;
        uint32_t w = bc.peek32();
#pragma unroll 1
        for (int j = 0; j < m; j++) {
            nw.put((int)(w >> (32 - slen)));
            w <<= slen;
        }
        bc.skip(m * slen);
        count -= m;
    }
}
// The same run read and thrown away (the scfsi look-back cursor passing over a band its unit reads itself).
template <class BC>
MP3_HD void sf_skip_run(BC &bc, int count, int slen) {
    if (slen == 0 || count <= 0) return;
    if (bc.pos() + count * slen <= bc.lim) {
        bc.skip(count * slen);
        return;
    }
#pragma unroll 1
    for (int i = 0; i < count; i++) (void)bc.bits(slen);
}

// MPEG-1, long-type block (maindata.go:233-279), lane-uniform form: 21 scalefactors in four bands (sfb 0-5, 6-10, 11-15,
// 16-20), slen1 bits each in the first two bands and slen2 in the last two; in gr 1 a band whose scfsi bit is set is copied
// from what gr 0's parse left in ScalefacL[0][ch] — re-read here from gr 0's bits through the look-back cursor b0, which
// passes over all of gr 0's values (n0 of them, s1_0 / s2_0 bits each).  Every lane runs the same 21 steps whatever its
// slen values and scfsi bits are (a width of 0 reads nothing), where the run-by-run form above diverged into a handful of
// lanes (ncu: 12 % of K1's issue slots at 6 of 32 lanes).  Requires that no read can be refused: the caller checks that
// both cursors have all their bits in front of the buffer end.
template <class BC>
MP3_HD void sf_mpeg1_long_uniform(BC &bc, BC &b0, int scfsi, int slen1, int slen2, int n0, int s1_0, int s2_0, uint32_t *pk) {
    uint32_t acc[3] = {0u, 0u, 0u};
    uint32_t w = 0, w0 = 0;
    int used = 0, used0 = 0;
#pragma unroll
    for (int sfb = 0; sfb < 21; sfb++) {
        if ((sfb & 7) == 0) {  // eight values of at most four bits per window
            bc.skip(used);
            b0.skip(used0);
            used = used0 = 0;
            w = bc.peek32();
            w0 = b0.peek32();
        }
        const int band = sfb < 6 ? 0 : (sfb < 11 ? 1 : (sfb < 16 ? 2 : 3));
        const bool copied = ((scfsi >> band) & 1) != 0;
        const uint32_t s = copied ? 0u : (uint32_t)(sfb < 11 ? slen1 : slen2);
        const uint32_t s0 = sfb < n0 ? (uint32_t)(sfb < 11 ? s1_0 : s2_0) : 0u;
        const uint32_t v = hi_bits_mod32(w, s), v0 = hi_bits_mod32(w0, s0);
        w = shl_mod32(w, s);
        w0 = shl_mod32(w0, s0);
        used += (int)s;
        used0 += (int)s0;
        acc[sfb >> 3] |= (copied ? v0 : v) << (4 * (sfb & 7));
    }
    bc.skip(used);
    b0.skip(used0);
    pk[0] = acc[0];
    pk[1] = acc[1];
    pk[2] = acc[2];
}

// Scalefactors of an MPEG-1 unit that reads all of them itself: gr 0, or gr 1 short blocks, or
// gr 1 with no scfsi band set (maindata.go:204-232 and the read arms of :233-279).
template <class BC>
MP3_HD void sf_mpeg1_read_all(const DeviceTables &T, BC &bc, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t *pk) {
    int sfc = u_sfcomp(w1) & 15;
    int slen1 = T.slen_mpeg1[sfc * 2], slen2 = T.slen_mpeg1[sfc * 2 + 1];
    NibWriter nw;
    nw.init(pk, 0);
    if (u_winsw(w0) == 1 && u_btype(w0) == 2) {
        int sfb0 = 0;
        if (u_mixed(w2)) {
            sf_run(bc, nw, 8, slen1);
            sfb0 = 3;
        }
        nw.seek(22 + sfb0 * 3);
        sf_run(bc, nw, 18 - sfb0 * 3, slen1);  // n = sfb*3 + win < 18: slen1
        sf_run(bc, nw, 18, slen2);
    } else {
        sf_run(bc, nw, 11, slen1);
        sf_run(bc, nw, 10, slen2);
    }
    nw.flush();
}

// K1 for one unit: scalefactors (part 2), then the big_values pairs and the count1 quadruples (part 3), with one
// cursor from the first part2 bit to the end.  `units` is the whole submission (absolute indexing): a gr-1 unit with
// scfsi set re-reads gr 0's scalefactor bits (maindata.go:239-278, quirk Q14).  `mk(cursor, bit_start, buf_end_rel)`
// positions a cursor of type BC (k_huffman: a StagedCursor over the tile's staged stretch of main data).
// Outputs: pk[8] scalefactor nibbles (n = sfb for scalefac_l, 22 + sfb*3 + win for scalefac_s), is_out[0..count1/2)
// packed int16 pairs, return value = count1 | preflag << 10.  Lines >= count1 are zero by definition; K2 masks them
// instead of K1 writing zeros.
struct HuffRegions {  // region starts in PAIRS (all sfb boundaries are even) and the three trees
    int r1h, r2h, nbig;
    SmemRef t0, t1, t2;   // root table of each region's tree
    uint32_t lin;         // linbits of the three regions, 4 bits each
};
MP3_HD HuffRegions huff_regions(const DeviceTables &T, SmemRef lut, const uint32_t *huff_desc, uint32_t w0, uint32_t w1, uint32_t w2) {
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
    R.t0 = lut.plus(e0 & 0xffffffu);
    R.t1 = lut.plus(e1 & 0xffffffu);
    R.t2 = lut.plus(e2 & 0xffffffu);
    R.lin = (e0 >> 24) | ((e1 >> 24) << 4) | ((e2 >> 24) << 8);
    return R;
}
template <class BC>
MP3_HD uint32_t huff_pair_at(const HuffRegions &R, int k, BC &bc) {
    const bool in0 = k < R.r1h, in1 = k < R.r2h;
    return huff_pair(in0 ? R.t0 : (in1 ? R.t1 : R.t2),
                     [&] { return (int)((R.lin >> (in0 ? 0 : (in1 ? 4 : 8))) & 0xf); }, bc);
}
MP3_HD uint32_t huff_pair_fast_at(const HuffRegions &R, int k, FastWindow &fw) {
    const bool in0 = k < R.r1h, in1 = k < R.r2h;
    return huff_pair_fast(in0 ? R.t0 : (in1 ? R.t1 : R.t2),
                          [&] { return (int)((R.lin >> (in0 ? 0 : (in1 ? 4 : 8))) & 0xf); }, fw);
}

template <class BC, class MkCursor>
MP3_HD uint32_t huffman_unit_t(const DeviceTables &T, SmemRef lut, const uint32_t *qlut, const uint32_t *huff_desc, const uint64_t *quad_signs,
                               const mp3gpu_unit *units, long long unit_index, MkCursor mk, uint32_t *pk, uint32_t *is_out) {
    const mp3gpu_unit u = units[unit_index];  // the caller vouches for unit_index (kernels.cuh checks it against the submission)
    const uint32_t w0 = u.w0, w1 = u.w1, w2 = u.w2;
    BC bc;
    mk(bc, u.bit_start, u.buf_end_rel);
    for (int i = 0; i < 8; i++) pk[i] = 0;

    // ---- part 2: scalefactors -------------------------------------------------------------
    uint32_t preflag = (uint32_t)u_preflag(w2);
    if (u_lsf(w2)) {
        // maindata.go:132-179
        int slen = T.nslen2[u_sfcomp(w1)];
        preflag = (uint32_t)((slen >> 15) & 1);
        int n = 0;
        if (u_btype(w0) == 2) {
            n++;
            if (u_mixed(w2)) n++;
        }
        const int d = (slen >> 12) & 7;
        // long blocks fill scalefac_l[idx]; short fill scalefac_s[idx/3][idx%3] (maindata.go:169-179).  The four group
        // sizes of a row add up to at most 36 values (tables.cc kSfSizeMpeg2), so the nibble index stays below 22 + 36.
        NibWriter nw;
        nw.init(pk, n == 0 ? 0 : 22);
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            const int num = slen & 7;
            slen >>= 3;
            sf_run(bc, nw, (int)T.sfsize_mpeg2[(n * 6 + d) * 4 + i], num);
        }
        nw.flush();
    } else if (u_winsw(w0) == 1 && u_btype(w0) == 2) {
        sf_mpeg1_read_all(T, bc, w0, w1, w2, pk);  // short blocks read everything themselves (maindata.go:206-232)
    } else {
        // Long-type block.  In gr 1 the bands whose scfsi bit is set copy ScalefacL[0][ch] as gr 0's parse left it (21
        // values if gr 0 was a long-type block, sfb 0-7 only if it was mixed, zeros if it was short).  gr 0's scalefactor
        // bits are re-read by a second cursor, band by band: a band this unit copies is read from it, a band this unit
        // reads itself is passed over.
        const int scfsi = u_gr(w2) == 0 ? 0 : u_scfsi(w2);
        int n0 = 0, s1_0 = 0, s2_0 = 0;
        BC b0;
        mk(b0, 0, 0);
        if (scfsi != 0 && unit_index >= 2) {  // a submission always starts on a frame boundary; guard against one that does not
            const mp3gpu_unit u0 = units[unit_index - 2];
            const bool short0 = u_winsw(u0.w0) == 1 && u_btype(u0.w0) == 2;
            n0 = short0 ? (u_mixed(u0.w2) ? 8 : 0) : 21;
            const int sfc0 = u_sfcomp(u0.w1) & 15;
            s1_0 = T.slen_mpeg1[sfc0 * 2];
            s2_0 = T.slen_mpeg1[sfc0 * 2 + 1];
            mk(b0, u0.bit_start, u0.buf_end_rel);
        }
        const int sfc = u_sfcomp(w1) & 15;
        const int slen1 = T.slen_mpeg1[sfc * 2], slen2 = T.slen_mpeg1[sfc * 2 + 1];
        if (bc.pos() + 11 * slen1 + 10 * slen2 <= bc.lim && (n0 == 0 || b0.pos() + 11 * s1_0 + 10 * s2_0 <= b0.lim)) {
            sf_mpeg1_long_uniform(bc, b0, scfsi, slen1, slen2, n0, s1_0, s2_0, pk);
        } else {  // a read could be refused at the buffer end (bits.go:65-68): value by value
            NibWriter nw;
            nw.init(pk, 0);
#pragma unroll 1
            for (int band = 0; band < 4; band++) {
                const int first = band == 0 ? 0 : (band == 1 ? 6 : (band == 2 ? 11 : 16));
                const int len = band == 0 ? 6 : 5;
                int have0 = n0 - first;  // values of this band that gr 0's parse read (the rest stayed 0)
                have0 = have0 < 0 ? 0 : (have0 > len ? len : have0);
                const int s0 = band < 2 ? s1_0 : s2_0;
                if ((scfsi >> band) & 1) {
                    sf_run(b0, nw, have0, s0);
#pragma unroll 1
                    for (int i = have0; i < len; i++) nw.put(0);
                } else {
                    sf_skip_run(b0, have0, s0);
                    sf_run(bc, nw, len, band < 2 ? slen1 : slen2);
                }
            }
            nw.flush();
        }
    }

    // ---- part 3: Huffman (maindata/huffman.go:27-138) ------------------------------------------
    if (u_p23len(w0) == 0) return preflag << 10;  // quirk Q1: nothing decoded, the cursor is not moved, Count1 stays 0
    const int bit_pos_end = u_p23len(w0) - 1;
    const HuffRegions R = huff_regions(T, lut, huff_desc, w0, w1, w2);
    int k = 0;
    PairSink sink;
    if constexpr (BC::kFast) {
        // A pair takes at most 19 + 2 * 13 + 2 = 47 bits.  Four pairs per 16-byte store while four worst-case pairs still
        // end in front of fast_end; then pair by pair while one does; what is left (the last bits in front of the frame's
        // buffer end, or of the staged stretch) goes to the careful cursor.  No bit-budget check (quirk Q4).
        uint4 *dst4 = reinterpret_cast<uint4 *>(is_out);
        const int lim4 = bc.fast_end - 4 * 47, lim1 = bc.fast_end - 47;
        if (bc.p <= lim1 && R.nbig > 0) {
            FastWindow fw;
            fw.open(bc);
            while (k + 4 <= R.nbig && fw.p <= lim4) {
                uint4 v;
                v.x = huff_pair_fast_at(R, k, fw);
                v.y = huff_pair_fast_at(R, k + 1, fw);
                v.z = huff_pair_fast_at(R, k + 2, fw);
                v.w = huff_pair_fast_at(R, k + 3, fw);
                if (MP3_CHECK((k >> 2) < 72, k)) dst4[k >> 2] = v;  // 288 pairs = 72 stores per unit
                k += 4;
            }
            sink.init(is_out, k);
            while (k < R.nbig && fw.p <= lim1) {
                sink.put(huff_pair_fast_at(R, k, fw));
                k++;
            }
            bc.p = fw.p;
        } else {
            sink.init(is_out, 0);
        }
    } else {
        sink.init(is_out, 0);
    }
    for (; k < R.nbig; k++) sink.put(huff_pair_at(R, k, bc));  // the careful cursor: bits.go's rules at the buffer end
    int is_pos = R.nbig * 2;
    {
        const uint32_t dq = huff_desc[32 + u_c1tsel(w2)] & 0xffffffu;
        if constexpr (BC::kFast) {
            // a quadruple takes at most 6 + 4 = 10 bits; inside the fast range pos() is p - off0
            const int p_end = bc.off0 + bit_pos_end, limq = bc.fast_end - 10;
            if (is_pos <= 572 && bc.p <= p_end && bc.p <= limq) {
                FastWindow fw;
                fw.open(bc);
                do {
                    const uint32_t wd = fw.peek();
                    const uint32_t e = lut_at(qlut, dq + ((wd >> (32 - kQuadBits)) << 2));
                    const uint32_t four = shl_mod32(wd, e >> 16) >> 28;
                    const uint64_t r = quad_signs[((e & 0xf) << 4) | four];
                    fw.advance((int)(e >> 26));
                    sink.put((uint32_t)r);
                    sink.put((uint32_t)(r >> 32));
                    is_pos += 4;
                } while (is_pos <= 572 && fw.p <= p_end && fw.p <= limq);
                bc.p = fw.p;
            }
        }
        while (is_pos <= 572 && bc.pos() <= bit_pos_end) {
            uint32_t vw, xy;
            huff_quad(qlut, quad_signs, dq, bc, vw, xy);
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

// The register-window cursor over global memory (host emulation and tests; k_huffman uses the staged cursor).
MP3_HD uint32_t huffman_unit(const DeviceTables &T, SmemRef lut, const uint32_t *qlut, const uint32_t *huff_desc, const uint64_t *quad_signs,
                             const uint8_t *main_data, uint64_t main_bits, const mp3gpu_unit *units, long long unit_index,
                             uint32_t *pk, uint32_t *is_out) {
    return huffman_unit_t<BitCursor>(T, lut, qlut, huff_desc, quad_signs, units, unit_index,
                                     [&](BitCursor &c, uint64_t bit_start, int buf_end_rel) { c.init(main_data, main_bits, bit_start, buf_end_rel); },
                                     pk, is_out);
}
// The staged cursor over one stretch (k_huffman's per-unit call; tests/hostemu emulates the tiling).
MP3_HD uint32_t huffman_unit_staged(const DeviceTables &T, SmemRef lut, const uint32_t *qlut, const uint32_t *huff_desc, const uint64_t *quad_signs,
                                    const StageCtx &S, const mp3gpu_unit *units, long long unit_index, uint32_t *pk, uint32_t *is_out) {
    return huffman_unit_t<StagedCursor>(T, lut, qlut, huff_desc, quad_signs, units, unit_index,
                                        [&](StagedCursor &c, uint64_t bit_start, int buf_end_rel) { c.init(S, bit_start, buf_end_rel); },
                                        pk, is_out);
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
