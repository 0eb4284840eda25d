// tables.h — host-side construction of every constant table the device kernels use.
//
// Formulas follow the reference's table initialisers (paths relative to the reference root):
//   internal/imdct/imdct.go:21-79      IMDCT windows and cosine tables
//   internal/frame/frame.go:31-40      powtab34, pretab
//   internal/frame/frame.go:304-306    isRatios;  :422-425 cs/ca
//   internal/frame/frame.go:488-628    synthNWin, synthDtbl
//   internal/maindata/maindata.go:39-81 scalefactor size tables, nSlen2
//   internal/consts/consts.go:68-97    scalefactor band indices
//   internal/huffman/huffman.go:23-346 Huffman code tables (here: ISO (x,y,hlen,hcod) form,
//                                      expanded into multi-level lookup tables)
#pragma once
#include <cstdint>
#include <vector>

namespace mp3gpu {

constexpr int kPow2Off = 336;   // pow2q index = 4*idx + kPow2Off ; 4*idx in [-326, 45]
constexpr int kPow2N = 400;
constexpr int kPowRow = 8208;   // row length of powq4 (8,207 values of |is| + padding)
constexpr int kNumCfg = 6;      // cfg = lsf*3 + sampling_frequency index

// Huffman LUTs.  Every tree's root table is indexed by the first kHuffRootBits bits of the code stream.
//
// Pair trees (tables 1..31): uint16 entries, and the SIGN BITS that follow a code word are part of the index — a leaf is
// the finished, signed pair, so the decode step has no sign arithmetic (tables.cc, build_pair_lut):
//   leaf   : bit15 = 0, bits 0..4 = x, bits 5..9 = y (5-bit two's complement, -15..15), bits 10..14 = bits consumed
//            (tree bits + sign bits, 0..21; 0 only in the empty tables 0/4/14, which consume nothing: huffman.go:354-356)
//   link   : bits 15..14 = 10, bits 10..13 = sb in 1..kHuffSubBits: a sub-table of 2^sb entries, indexed by the next sb
//            bits, starts (bits 0..9) * 16 entries after the tree's base; its entries are leaves, escapes or further links
//   escape : bits 15..14 = 11: x == 15 or y == 15 in a tree used with linbits (tables 16..31), where linbits sit between
//            the code word and the signs; bits 0..4 = tree bits, bit 5 = x is 15, bit 6 = y is 15, bits 7..10 = the other
//            value (unused when both are 15).  Escape code words are indexed by tree bits only.
//   Read as int16: leaf >= 0, link < -16384 <= escape < 0 — one compare each in the decode loop.
// Count1 trees (tables 32/33, at most 6 tree bits): uint32 entries in a separate 2 x 256 table (quad_lut):
//          bits 0..3   the (v w x y) pattern
//          bits 16..20 tree bits of the code word
//          bits 26..30 tree bits + sign bits
// Table descriptor (uint32): bits 0..23 = byte offset of the tree's root table in its LUT, bits 24..27 = linbits.
#ifndef MP3_HUFF_ROOT_BITS
#define MP3_HUFF_ROOT_BITS 8
#endif
constexpr int kHuffRootBits = MP3_HUFF_ROOT_BITS;  // pair trees
constexpr int kQuadRootBits = 8;                   // count1 trees (at most 6 tree bits)
#ifndef MP3_HUFF_SUB_BITS
#define MP3_HUFF_SUB_BITS 8
#endif
constexpr int kHuffSubBits = MP3_HUFF_SUB_BITS;   // index bits of a sub-table at most (codes reach 21 bits with their signs: up to three levels)
struct HostTables {
    float cos36[18 * 36];
    float cos12[6 * 12];
    float imdct_win[4 * 36];
    float synth_n[64 * 32];
    float synth_d[512];
    double pow2q[kPow2N];
    std::vector<double> powtab34;  // 8207
    std::vector<float> powq4;      // [4][kPowRow]: float32(2^(q/4) * powtab34[v]) — requantisation without the f64 multiply
    uint64_t pretab_pack;          // pretab[sfb] in bits 2*sfb .. 2*sfb+1
    float cs[8], ca[8];
    float is_ratio_l[8], is_ratio_r[8];  // index = is_pos 0..6
    uint8_t pretab[24];
    uint16_t sfb_long[kNumCfg][24];
    uint16_t sfb_short[kNumCfg][16];
    uint8_t line_sfb_long[kNumCfg][576];
    uint8_t line_sfb_short[kNumCfg][576];  // for window-major index i within short sfbs
    uint8_t line_win_short[kNumCfg][576];
    uint16_t reorder_dst[kNumCfg][576];
    uint8_t pair_long[kNumCfg][288];
    uint8_t pair_short[kNumCfg][288];
    uint16_t pair_dst[kNumCfg][288];
    uint16_t nslen2[512];
    uint8_t sfsize_mpeg2[3][6][4];
    std::vector<uint16_t> huff_lut;   // pair trees (16-bit entries, see above); size is a multiple of 8 entries
    uint32_t quad_lut[2 * 256];       // count1 trees
    uint64_t quad_signs[256];         // count1 sign expansion, see tables.cc
    uint32_t huff_desc[34];
};

// Builds all tables.  Pure host code (libm); deterministic.
void build_host_tables(HostTables &t);

// Access to the raw ISO code list, for the stream synthesiser (encoder side).
struct HuffCode { uint8_t x, y, hlen; uint32_t hcod; };
int huff_table_codes(int table_num, const HuffCode **codes, int *linbits);

}  // namespace mp3gpu
