/*
 * mp3_oracle.c — CPU restatement of llehouerou/go-mp3 (MPEG-1/2 Layer III decoder).
 *
 * TEST INFRASTRUCTURE ONLY — see mp3_oracle.h.  "Parity unpinned" at the PCM-value
 * level (no golden PCM in the reference, no Go toolchain here); pinned against the
 * reference's own unit tests and fixture invariants by tests/test_oracle_*.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  float ops are
 * individually rounded and accumulate in the reference's order; float64 is used exactly
 * where the reference uses it (requantisation, table generation).
 *
 * All file:line citations are relative to the reference repository root.
 */
#define _GNU_SOURCE
#include "mp3_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------ */
/* Huffman code tables in ISO (x, y, hlen, hcod) form, derived from the      */
/* reference's packed tree array by tools/derive_tables.py.                  */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint8_t x, y, hlen;
    uint32_t hcod;
} huff_code_t;
typedef struct {
    const huff_code_t *codes;
    int n;
    int linbits;
} huff_table_desc_t;
#include "huff_codes.inc"
#include "synth_window_k.inc"

/* A binary tree per table, rebuilt from the code words, walked one bit at a time
 * exactly as huffman.go:361-381 walks its packed array. */
typedef struct {
    int16_t child[2]; /* index of child node for bit 0 / bit 1, -1 if none */
    int16_t leaf;     /* -1 for inner node, else (x<<4)|y */
} huff_node_t;
static huff_node_t *g_tree[34];
static int g_tree_n[34];

/* ------------------------------------------------------------------------ */
/* Constant tables                                                           */
/* ------------------------------------------------------------------------ */
#define SAMPLES_PER_GR 576 /* consts.go:53 */

/* consts.go:68-97.  Index order [lsf][sampling_frequency index][long/short]. */
static const int SFB_LONG[2][3][23] = {
    {{0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576},
     {0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576},
     {0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576}},
    {{0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
     {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 114, 136, 162, 194, 232, 278, 332, 394, 464, 540, 576},
     {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576}}};
static const int SFB_SHORT[2][3][14] = {
    {{0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192},
     {0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192},
     {0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192}},
    {{0, 4, 8, 12, 18, 24, 32, 42, 56, 74, 100, 132, 174, 192},
     {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 136, 180, 192},
     {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 134, 174, 192}}};

static double g_powtab34[8207]; /* frame.go:31-40 */
static const double PRETAB[22] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0}; /* frame.go:33 */
static float g_imdct_win[4][36];   /* imdct.go:21-57 */
static float g_cos_n12[6][12];     /* imdct.go:59-68 */
static float g_cos_n36[18][36];    /* imdct.go:70-79 */
static float g_synth_nwin[64][32]; /* frame.go:488-497 */
static float g_synth_dtbl[512];    /* frame.go:499-628 */
static int g_nslen2[512];          /* maindata.go:52-81 */

/* maindata.go:39-42 */
static const int SCALEFAC_SIZES_MPEG1[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {3, 0}, {1, 1}, {1, 2}, {1, 3},
                                                {2, 1}, {2, 2}, {2, 3}, {3, 1}, {3, 2}, {3, 3}, {4, 2}, {4, 3}};
/* maindata.go:44-50 */
static const int SCALEFAC_SIZES_MPEG2[3][6][4] = {
    {{6, 5, 5, 5}, {6, 5, 7, 3}, {11, 10, 0, 0}, {7, 7, 7, 0}, {6, 6, 6, 3}, {8, 8, 5, 0}},
    {{9, 9, 9, 9}, {9, 9, 12, 6}, {18, 18, 0, 0}, {12, 12, 12, 0}, {12, 9, 9, 6}, {15, 12, 9, 0}},
    {{6, 9, 9, 9}, {6, 9, 12, 6}, {15, 18, 0, 0}, {6, 15, 12, 0}, {6, 12, 9, 6}, {6, 18, 9, 0}}};

/* frame.go:304-306, 422-425: decimal literals, rounded once to float32 by the compiler. */
static const float IS_RATIOS[6] = {0.000000f, 0.267949f, 0.577350f, 1.000000f, 1.732051f, 3.732051f};
static const float CS[8] = {0.857493f, 0.881742f, 0.949629f, 0.983315f, 0.995518f, 0.999161f, 0.999899f, 0.999993f};
static const float CA[8] = {-0.514496f, -0.471732f, -0.313377f, -0.181913f,
                            -0.094574f, -0.040966f, -0.014199f, -0.003700f};

/* Go evaluates constant expressions such as math.Pi/36 exactly and rounds once to
 * float64; long double division gets the same float64 (rounded once from a 64-bit
 * mantissa quotient). */
#define PI_L 3.14159265358979323846264338327950288L

static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build_tree(int t) {
    const huff_table_desc_t *d = &HUFF_TABLES[t];
    if (d->codes == NULL) {
        g_tree[t] = NULL;
        g_tree_n[t] = 0;
        return;
    }
    int cap = 2 * d->n; /* complete binary tree: n leaves, n-1 inner nodes */
    huff_node_t *nodes = (huff_node_t *)malloc(sizeof(huff_node_t) * (size_t)cap);
    int n = 1;
    nodes[0].child[0] = nodes[0].child[1] = -1;
    nodes[0].leaf = -1;
    for (int i = 0; i < d->n; i++) {
        const huff_code_t *c = &d->codes[i];
        int p = 0;
        for (int b = c->hlen - 1; b >= 0; b--) {
            int bit = (int)((c->hcod >> b) & 1u);
            if (nodes[p].child[bit] < 0) {
                nodes[n].child[0] = nodes[n].child[1] = -1;
                nodes[n].leaf = -1;
                nodes[p].child[bit] = (int16_t)n;
                n++;
            }
            p = nodes[p].child[bit];
        }
        nodes[p].leaf = (int16_t)((c->x << 4) | c->y);
    }
    g_tree[t] = nodes;
    g_tree_n[t] = n;
}

static void init_tables(void) {
    for (int t = 0; t < 34; t++) build_tree(t);
    /* frame.go:36-40 */
    for (int i = 0; i < 8207; i++) g_powtab34[i] = pow((double)i, 4.0 / 3.0);
    /* imdct.go:23-57 */
    const double pi36 = (double)(PI_L / 36.0L), pi12 = (double)(PI_L / 12.0L);
    for (int i = 0; i < 36; i++) g_imdct_win[0][i] = (float)sin(pi36 * ((double)i + 0.5));
    for (int i = 0; i < 18; i++) g_imdct_win[1][i] = (float)sin(pi36 * ((double)i + 0.5));
    for (int i = 18; i < 24; i++) g_imdct_win[1][i] = 1.0f;
    for (int i = 24; i < 30; i++) g_imdct_win[1][i] = (float)sin(pi12 * ((double)i + 0.5 - 18.0));
    for (int i = 30; i < 36; i++) g_imdct_win[1][i] = 0.0f;
    for (int i = 0; i < 12; i++) g_imdct_win[2][i] = (float)sin(pi12 * ((double)i + 0.5));
    for (int i = 12; i < 36; i++) g_imdct_win[2][i] = 0.0f;
    for (int i = 0; i < 6; i++) g_imdct_win[3][i] = 0.0f;
    for (int i = 6; i < 12; i++) g_imdct_win[3][i] = (float)sin(pi12 * ((double)i + 0.5 - 6.0));
    for (int i = 12; i < 18; i++) g_imdct_win[3][i] = 1.0f;
    for (int i = 18; i < 36; i++) g_imdct_win[3][i] = (float)sin(pi36 * ((double)i + 0.5));
    /* imdct.go:61-68: cos(Pi/(2N) * (2j + 1 + N/2) * (2i + 1)), N = 12 */
    const double pi24 = (double)(PI_L / 24.0L), pi72 = (double)(PI_L / 72.0L);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 12; j++)
            g_cos_n12[i][j] = (float)cos(pi24 * (2.0 * (double)j + 1.0 + 6.0) * (2.0 * (double)i + 1.0));
    /* imdct.go:72-79: N = 36 */
    for (int i = 0; i < 18; i++)
        for (int j = 0; j < 36; j++)
            g_cos_n36[i][j] = (float)cos(pi72 * (2.0 * (double)j + 1.0 + 18.0) * (2.0 * (double)i + 1.0));
    /* frame.go:490-497 */
    const double pi64 = (double)(PI_L / 64.0L);
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 32; j++) g_synth_nwin[i][j] = (float)cos((double)((16 + i) * (2 * j + 1)) * pi64);
    /* frame.go:499-628: 9-decimal literals of k/65536, rounded once to float32. */
    for (int i = 0; i < 512; i++) {
        long long k = SYNTH_WINDOW_K[i];
        long long a = k < 0 ? -k : k;
        /* n = round(a * 1e9 / 65536), exact integer arithmetic (half cases do not occur, asserted by the generator) */
        long long n = (a * 1000000000LL + 32768) / 65536;
        char lit[40];
        snprintf(lit, sizeof lit, "%s%lld.%09lld", k < 0 ? "-" : "", n / 1000000000LL, n % 1000000000LL);
        g_synth_dtbl[i] = strtof(lit, NULL);
    }
    /* maindata.go:54-81 */
    memset(g_nslen2, 0, sizeof g_nslen2);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 3; j++) {
            int n = j + i * 3;
            g_nslen2[n + 500] = i | (j << 3) | (2 << 12) | (1 << 15);
        }
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++) {
                    int n = l + k * 4 + j * 16 + i * 80;
                    g_nslen2[n] = i | (j << 3) | (k << 6) | (l << 9) | (0 << 12);
                }
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 4; k++) {
                int n = k + j * 4 + i * 20;
                g_nslen2[n + 400] = i | (j << 3) | (k << 6) | (1 << 12);
            }
}
static void ensure_init(void) { pthread_once(&g_once, init_tables); }

const float *orc_table_imdct_win(void) { ensure_init(); return &g_imdct_win[0][0]; }
const float *orc_table_cos_n12(void) { ensure_init(); return &g_cos_n12[0][0]; }
const float *orc_table_cos_n36(void) { ensure_init(); return &g_cos_n36[0][0]; }
const float *orc_table_synth_nwin(void) { ensure_init(); return &g_synth_nwin[0][0]; }
const float *orc_table_synth_dtbl(void) { ensure_init(); return g_synth_dtbl; }
const double *orc_table_powtab34(void) { ensure_init(); return g_powtab34; }

/* ------------------------------------------------------------------------ */
/* internal/bits  (bits.go)                                                  */
/* ------------------------------------------------------------------------ */
void orc_bits_init(orc_bits *b, const uint8_t *vec, int len) {
    b->vec = vec;
    b->len = len;
    b->bit_pos = 0;
    b->byte_pos = 0;
    b->err = 0;
}

/* bits.go:45-56 — a read at or past the end returns 0, sets the sticky error and
 * does NOT advance the cursor. */
int orc_bits_bit(orc_bits *b) {
    if (b->len <= b->byte_pos) {
        b->err = 1;
        return 0;
    }
    unsigned tmp = (unsigned)b->vec[b->byte_pos] >> (7 - (unsigned)b->bit_pos);
    tmp &= 0x01;
    b->byte_pos += (b->bit_pos + 1) >> 3;
    b->bit_pos = (b->bit_pos + 1) & 0x07;
    return (int)tmp;
}

/* bits.go:58-77 — a read that would cross the end returns 0 without advancing. */
int orc_bits_bits(orc_bits *b, int num) {
    if (num == 0) return 0;
    int current = b->byte_pos * 8 + b->bit_pos;
    int total = b->len * 8;
    if (current + num > total) {
        b->err = 1;
        return 0;
    }
    uint8_t bb[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4 && b->byte_pos + i < b->len; i++) bb[i] = b->vec[b->byte_pos + i];
    uint32_t tmp = ((uint32_t)bb[0] << 24) | ((uint32_t)bb[1] << 16) | ((uint32_t)bb[2] << 8) | (uint32_t)bb[3];
    tmp <<= (unsigned)b->bit_pos;
    tmp >>= (32 - (unsigned)num);
    b->byte_pos += (b->bit_pos + num) >> 3;
    b->bit_pos = (b->bit_pos + num) & 0x07;
    return (int)tmp;
}
int orc_bits_pos(const orc_bits *b) { return (b->byte_pos << 3) + b->bit_pos; }
void orc_bits_set_pos(orc_bits *b, int pos) {
    b->byte_pos = pos >> 3;
    b->bit_pos = pos & 0x7;
}

/* ------------------------------------------------------------------------ */
/* internal/huffman  (huffman.go:348-419)                                    */
/* ------------------------------------------------------------------------ */
int orc_huffman_decode(orc_bits *m, int table_num, int out[4]) {
    ensure_init();
    int x = 0, y = 0, v = 0, w = 0;
    out[0] = out[1] = out[2] = out[3] = 0;
    const huff_node_t *tree = g_tree[table_num];
    int linbits = HUFF_TABLES[table_num].linbits;
    if (tree == NULL) return ORC_OK; /* huffman.go:354-356: tables 0, 4, 14 consume nothing */
    int point = 0, bitsleft = 32, decode_error = 1;
    for (;;) {
        if (tree[point].leaf >= 0) {
            decode_error = 0;
            x = (tree[point].leaf >> 4) & 0xf;
            y = tree[point].leaf & 0xf;
            break;
        }
        int nxt = tree[point].child[orc_bits_bit(m) != 0 ? 1 : 0];
        bitsleft--;
        if (nxt < 0 || bitsleft <= 0) break; /* unreachable: every tree is complete, depth <= 19 */
        point = nxt;
    }
    if (decode_error) return ORC_ERR_HUFFMAN;
    if (table_num > 31) { /* huffman.go:387-403 */
        v = (y >> 3) & 1;
        w = (y >> 2) & 1;
        x = (y >> 1) & 1;
        y &= 1;
        if (v != 0 && orc_bits_bit(m) == 1) v = -v;
        if (w != 0 && orc_bits_bit(m) == 1) w = -w;
        if (x != 0 && orc_bits_bit(m) == 1) x = -x;
        if (y != 0 && orc_bits_bit(m) == 1) y = -y;
    } else { /* huffman.go:404-416 */
        if (linbits != 0 && x == 15) x += orc_bits_bits(m, linbits);
        if (x != 0 && orc_bits_bit(m) == 1) x = -x;
        if (linbits != 0 && y == 15) y += orc_bits_bits(m, linbits);
        if (y != 0 && orc_bits_bit(m) == 1) y = -y;
    }
    out[0] = x;
    out[1] = y;
    out[2] = v;
    out[3] = w;
    return ORC_OK;
}

int orc_huffman_table_info(int table_num, int *n_symbols, int *linbits) {
    if (table_num < 0 || table_num > 33) return ORC_ERR_INTERNAL;
    *n_symbols = HUFF_TABLES[table_num].n;
    *linbits = HUFF_TABLES[table_num].linbits;
    return ORC_OK;
}
int orc_huffman_table_code(int table_num, int i, int *x, int *y, int *hlen, uint32_t *hcod) {
    if (table_num < 0 || table_num > 33 || i < 0 || i >= HUFF_TABLES[table_num].n) return ORC_ERR_INTERNAL;
    const huff_code_t *c = &HUFF_TABLES[table_num].codes[i];
    *x = c->x;
    *y = c->y;
    *hlen = c->hlen;
    *hcod = c->hcod;
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* internal/frameheader  (frameheader.go)                                    */
/* ------------------------------------------------------------------------ */
static int h_id(uint32_t f) { return (int)((f & 0x00180000u) >> 19); }        /* :30-32 */
static int h_layer(uint32_t f) { return (int)((f & 0x00060000u) >> 17); }     /* :35-37 */
static int h_protection_bit(uint32_t f) { return (int)(f & 0x00010000u) >> 16; } /* :40-42 */
static int h_bitrate_index(uint32_t f) { return (int)(f & 0x0000f000u) >> 12; }  /* :45-47 */
static int h_sampling_frequency(uint32_t f) { return (int)(f & 0x00000c00u) >> 10; } /* :50-52 */
static int h_padding_bit(uint32_t f) { return (int)(f & 0x00000200u) >> 9; }  /* :74-76 */
static int h_mode(uint32_t f) { return (int)((f & 0x000000c0u) >> 6); }       /* :85-87 */
static int h_mode_extension(uint32_t f) { return (int)(f & 0x00000030u) >> 4; } /* :90-92 */
static int h_emphasis(uint32_t f) { return (int)(f & 0x00000003u); }          /* :121-123 */
static int h_lsf(uint32_t f) { return h_id(f) == 3 ? 0 : 1; }                 /* :126-131 */
static int h_use_ms(uint32_t f) { return h_mode(f) == 1 && (h_mode_extension(f) & 0x2) != 0; }        /* :95-100 */
static int h_use_intensity(uint32_t f) { return h_mode(f) == 1 && (h_mode_extension(f) & 0x1) != 0; } /* :103-108 */
static int h_granules(uint32_t f) { return 2 >> h_lsf(f); }                   /* :137-140 */
static int h_nch(uint32_t f) { return h_mode(f) == 3 ? 1 : 2; }               /* :253-258 */

int orc_header_sampling_frequency_value(uint32_t f) { /* :56-71 */
    int lsf = h_lsf(f);
    switch (h_sampling_frequency(f)) {
    case 0: return 44100 >> lsf;
    case 1: return 48000 >> lsf;
    case 2: return 32000 >> lsf;
    }
    return 0; /* error */
}
int orc_header_bytes_per_frame(uint32_t f) { return SAMPLES_PER_GR * h_granules(f) * 4; } /* :133-135 */
int orc_header_samples_per_frame(uint32_t f) { return SAMPLES_PER_GR * h_granules(f); }   /* :145-147 */
int64_t orc_header_frame_duration_ns(uint32_t f) { /* :150-156 */
    int sr = orc_header_sampling_frequency_value(f);
    if (sr == 0) return 0;
    return (int64_t)1000000000 * (int64_t)orc_header_samples_per_frame(f) / (int64_t)sr;
}
int orc_header_bytes_per_second(uint32_t f) { return orc_header_sampling_frequency_value(f) * 4; } /* :160-166 */

int orc_header_is_valid(uint32_t f) { /* :168-189 */
    const uint32_t sync = 0xffe00000u;
    if ((f & sync) != sync) return 0;
    if (h_id(f) == 1) return 0;
    if (h_bitrate_index(f) == 15) return 0;
    if (h_sampling_frequency(f) == 3) return 0;
    if (h_layer(f) != 1) return 0;
    if (h_emphasis(f) == 2) return 0;
    return 1;
}

int orc_header_bitrate(uint32_t f) { /* :191-221 */
    static const int bitrates[2][3][16] = {
        {{0, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 160000, 192000, 224000, 256000, 320000, 0},
         {0, 32000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 160000, 192000, 224000, 256000, 320000, 384000, 0},
         {0, 32000, 64000, 96000, 128000, 160000, 192000, 224000, 256000, 288000, 320000, 352000, 384000, 416000, 448000, 0}},
        {{0, 8000, 16000, 24000, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 144000, 160000, 0},
         {0, 8000, 16000, 24000, 32000, 40000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 144000, 160000, 0},
         {0, 32000, 48000, 56000, 64000, 80000, 96000, 112000, 128000, 144000, 160000, 176000, 192000, 224000, 256000, 0}}};
    int layer = h_layer(f);
    if (layer < 1) return 0; /* Go would panic on index -1; never reached with a valid header */
    return bitrates[h_lsf(f)][layer - 1][h_bitrate_index(f)];
}

int orc_header_frame_size(uint32_t f, int *size) { /* :223-232 */
    int freq = orc_header_sampling_frequency_value(f);
    if (freq == 0) return ORC_ERR_SAMPLE_RATE;
    *size = ((144 * orc_header_bitrate(f)) / freq + h_padding_bit(f)) >> h_lsf(f);
    return ORC_OK;
}

int orc_header_side_info_size(uint32_t f) { /* :234-251 */
    int mono = h_mode(f) == 3;
    if (h_lsf(f) == 1) return mono ? 9 : 17;
    return mono ? 17 : 32;
}

/* ------------------------------------------------------------------------ */
/* source.go — buffered reader with Unread, over an in-memory reader.        */
/* Unread(buf) always returns the bytes just read, so a cursor rewind is      */
/* observationally identical to the reference's prepend (source.go:94-97).    */
/* ------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *data;
    size_t len;
    int64_t pos; /* source.pos */
    int seekable;
} source_t;

/* source.go:99-122: returns bytes read; *err = ORC_EOF when fewer than n were available. */
static int src_read_full(source_t *s, uint8_t *buf, int n, int *err) {
    int64_t avail = (int64_t)s->len - s->pos;
    if (avail < 0) avail = 0;
    int got = (int64_t)n <= avail ? n : (int)avail;
    if (got > 0) memcpy(buf, s->data + s->pos, (size_t)got);
    s->pos += got;
    *err = (got < n) ? ORC_EOF : ORC_OK;
    return got;
}
static void src_unread(source_t *s, int n) { s->pos -= n; } /* source.go:94-97 */

/* source.go:27-40 */
static int src_seek(source_t *s, int64_t position, int whence, int64_t *out) {
    if (!s->seekable) return ORC_ERR_SEEK_UNSUPPORTED;
    int64_t n;
    switch (whence) {
    case 0: n = position; break;
    case 1: n = s->pos + position; break;
    case 2: n = (int64_t)s->len + position; break;
    default: return ORC_ERR_WHENCE;
    }
    if (n < 0) return ORC_ERR_INTERNAL; /* bytes.Reader: negative position */
    s->pos = n;
    if (out) *out = n;
    return ORC_OK;
}

/* source.go:42-83 */
static int src_skip_tags(source_t *s) {
    for (;;) {
        uint8_t buf[4];
        int err;
        src_read_full(s, buf, 3, &err);
        if (err != ORC_OK) return err;
        if (buf[0] == 'T' && buf[1] == 'A' && buf[2] == 'G') {
            uint8_t skip[125];
            src_read_full(s, skip, 125, &err);
            if (err != ORC_OK) return err;
        } else if (buf[0] == 'I' && buf[1] == 'D' && buf[2] == '3') {
            uint8_t vf[3];
            src_read_full(s, vf, 3, &err);
            if (err != ORC_OK) return err;
            int n = src_read_full(s, buf, 4, &err);
            if (err != ORC_OK) return err;
            if (n != 4) return ORC_OK;
            uint32_t size = ((uint32_t)buf[0] << 21) | ((uint32_t)buf[1] << 14) | ((uint32_t)buf[2] << 7) | (uint32_t)buf[3];
            /* ReadFull of `size` bytes: only the cursor moves */
            int64_t avail = (int64_t)s->len - s->pos;
            if (avail < 0) avail = 0;
            if ((int64_t)size > avail) {
                s->pos += avail;
                return ORC_EOF;
            }
            s->pos += size;
        } else {
            src_unread(s, 3);
            return ORC_OK;
        }
    }
}

/* frameheader.go:279-328 */
static int frameheader_read(source_t *s, int64_t position, uint32_t *header, int64_t *start_position,
                            int64_t *bytes_searched_out) {
    uint8_t buf[4];
    int err;
    int n = src_read_full(s, buf, 4, &err);
    if (n < 4) {
        if (n == 0) return ORC_EOF;
        return ORC_ERR_UNEXPECTED_EOF; /* "readHeader (1)" */
    }
    uint32_t b1 = buf[0], b2 = buf[1], b3 = buf[2], b4 = buf[3];
    uint32_t h = (b1 << 24) | (b2 << 16) | (b3 << 8) | b4;
    int64_t bytes_searched = 4;
    while (!orc_header_is_valid(h)) {
        if (bytes_searched >= 64 * 1024) {
            if (bytes_searched_out) *bytes_searched_out = bytes_searched;
            return ORC_ERR_SYNC_LIMIT;
        }
        b1 = b2;
        b2 = b3;
        b3 = b4;
        uint8_t one;
        src_read_full(s, &one, 1, &err);
        if (err != ORC_OK) return ORC_ERR_UNEXPECTED_EOF; /* "readHeader (2)" */
        b4 = one;
        h = (b1 << 24) | (b2 << 16) | (b3 << 8) | b4;
        position++;
        bytes_searched++;
    }
    if (bytes_searched_out) *bytes_searched_out = bytes_searched;
    if (h_bitrate_index(h) == 0) return ORC_ERR_FREE_FORMAT;
    *header = h;
    *start_position = position;
    return ORC_OK;
}

int orc_frameheader_read_mem(const uint8_t *data, size_t len, size_t pos, uint32_t *header, int64_t *start_pos,
                             size_t *new_pos, int64_t *bytes_searched) {
    source_t s = {data, len, (int64_t)pos, 1};
    int64_t bs = 0;
    int rc = frameheader_read(&s, (int64_t)pos, header, start_pos, &bs);
    if (new_pos) *new_pos = (size_t)s.pos;
    if (bytes_searched) *bytes_searched = bs;
    return rc;
}

/* ------------------------------------------------------------------------ */
/* internal/sideinfo  (sideinfo.go)                                          */
/* ------------------------------------------------------------------------ */
typedef struct { /* sideinfo.go:33-55 */
    int main_data_begin;
    int private_bits;
    int scfsi[2][4];
    int part2_3_length[2][2];
    int big_values[2][2];
    int global_gain[2][2];
    int scalefac_compress[2][2];
    int win_switch_flag[2][2];
    int block_type[2][2];
    int mixed_block_flag[2][2];
    int table_select[2][2][3];
    int subblock_gain[2][2][3];
    int region0_count[2][2];
    int region1_count[2][2];
    int preflag[2][2];
    int scalefac_scale[2][2];
    int count1table_select[2][2];
    int count1[2][2];
} sideinfo_t;

static const int SIDEINFO_BITS_TO_READ[2][4] = {{9, 5, 3, 4}, {8, 1, 2, 9}}; /* sideinfo.go:57-64 */

/* sideinfo.go:66-156 */
static int sideinfo_read(source_t *s, uint32_t header, sideinfo_t *si) {
    int nch = h_nch(header);
    int framesize;
    int rc = orc_header_frame_size(header, &framesize);
    if (rc != ORC_OK) return rc;
    if (framesize > 2000) return ORC_ERR_FRAMESIZE;
    int sideinfo_size = orc_header_side_info_size(header);
    uint8_t buf[32];
    int err;
    int n = src_read_full(s, buf, sideinfo_size, &err);
    if (n < sideinfo_size) return ORC_ERR_UNEXPECTED_EOF; /* "sideinfo.Read" */
    orc_bits b;
    orc_bits_init(&b, buf, sideinfo_size);
    int mpeg1 = h_lsf(header) == 0;
    const int *btr = SIDEINFO_BITS_TO_READ[h_lsf(header)];
    memset(si, 0, sizeof *si);
    si->main_data_begin = orc_bits_bits(&b, btr[0]);
    if (h_mode(header) == 3)
        si->private_bits = orc_bits_bits(&b, btr[1]);
    else
        si->private_bits = orc_bits_bits(&b, btr[2]);
    if (mpeg1)
        for (int ch = 0; ch < nch; ch++)
            for (int band = 0; band < 4; band++) si->scfsi[ch][band] = orc_bits_bits(&b, 1);
    for (int gr = 0; gr < h_granules(header); gr++) {
        for (int ch = 0; ch < nch; ch++) {
            si->part2_3_length[gr][ch] = orc_bits_bits(&b, 12);
            si->big_values[gr][ch] = orc_bits_bits(&b, 9);
            si->global_gain[gr][ch] = orc_bits_bits(&b, 8);
            si->scalefac_compress[gr][ch] = orc_bits_bits(&b, btr[3]);
            si->win_switch_flag[gr][ch] = orc_bits_bits(&b, 1);
            if (si->win_switch_flag[gr][ch] == 1) {
                si->block_type[gr][ch] = orc_bits_bits(&b, 2);
                si->mixed_block_flag[gr][ch] = orc_bits_bits(&b, 1);
                for (int region = 0; region < 2; region++) si->table_select[gr][ch][region] = orc_bits_bits(&b, 5);
                for (int window = 0; window < 3; window++) si->subblock_gain[gr][ch][window] = orc_bits_bits(&b, 3);
                if (si->block_type[gr][ch] == 2 && si->mixed_block_flag[gr][ch] == 0)
                    si->region0_count[gr][ch] = 8;
                else
                    si->region0_count[gr][ch] = 7;
                si->region1_count[gr][ch] = 20 - si->region0_count[gr][ch];
            } else {
                for (int region = 0; region < 3; region++) si->table_select[gr][ch][region] = orc_bits_bits(&b, 5);
                si->region0_count[gr][ch] = orc_bits_bits(&b, 4);
                si->region1_count[gr][ch] = orc_bits_bits(&b, 3);
                si->block_type[gr][ch] = 0;
                if (!mpeg1) si->mixed_block_flag[0][ch] = 0;
            }
            if (mpeg1) si->preflag[gr][ch] = orc_bits_bits(&b, 1);
            si->scalefac_scale[gr][ch] = orc_bits_bits(&b, 1);
            si->count1table_select[gr][ch] = orc_bits_bits(&b, 1);
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* internal/maindata  (maindata.go, maindata/huffman.go)                     */
/* ------------------------------------------------------------------------ */
typedef struct { /* maindata.go:33-37 */
    int scalefac_l[2][2][22];
    int scalefac_s[2][2][13][3];
    float is[2][2][576];
} maindata_t;

/* The frame's logical main-data buffer (reservoir tail ++ own bytes), maindata.go:290-323. */
typedef struct {
    uint8_t *vec;
    int len;
    orc_bits bits;
} mdbits_t;

/* maindata/huffman.go:27-138 */
static int read_huffman(orc_bits *m, uint32_t header, sideinfo_t *si, maindata_t *md, int part2_start, int gr, int ch) {
    if (si->part2_3_length[gr][ch] == 0) {
        for (int i = 0; i < SAMPLES_PER_GR; i++) md->is[gr][ch][i] = 0.0f;
        return ORC_OK;
    }
    int bit_pos_end = part2_start + si->part2_3_length[gr][ch] - 1;
    int region1_start = 0, region2_start = 0;
    if (si->win_switch_flag[gr][ch] == 1 && si->block_type[gr][ch] == 2) {
        region1_start = 36;
        region2_start = SAMPLES_PER_GR;
    } else {
        int sfreq = h_sampling_frequency(header);
        int lsf = h_lsf(header);
        const int *l = SFB_LONG[lsf][sfreq];
        int i = si->region0_count[gr][ch] + 1;
        if (i < 0 || 23 <= i) return ORC_ERR_REGION_INDEX;
        region1_start = l[i];
        int j = si->region0_count[gr][ch] + si->region1_count[gr][ch] + 2;
        if (j < 0) return ORC_ERR_REGION_INDEX;
        if (j >= 23)
            region2_start = SAMPLES_PER_GR;
        else
            region2_start = l[j];
    }
    int out[4];
    for (int is_pos = 0; is_pos < si->big_values[gr][ch] * 2; is_pos++) {
        if (is_pos >= SAMPLES_PER_GR) return ORC_ERR_ISPOS;
        int table_num;
        if (is_pos < region1_start)
            table_num = si->table_select[gr][ch][0];
        else if (is_pos < region2_start)
            table_num = si->table_select[gr][ch][1];
        else
            table_num = si->table_select[gr][ch][2];
        int rc = orc_huffman_decode(m, table_num, out);
        if (rc != ORC_OK) return rc;
        md->is[gr][ch][is_pos] = (float)out[0];
        is_pos++;
        md->is[gr][ch][is_pos] = (float)out[1];
    }
    int table_num = si->count1table_select[gr][ch] + 32;
    int is_pos = si->big_values[gr][ch] * 2;
    while (is_pos <= 572 && orc_bits_pos(m) <= bit_pos_end) {
        int rc = orc_huffman_decode(m, table_num, out);
        if (rc != ORC_OK) return rc;
        md->is[gr][ch][is_pos] = (float)out[2]; /* v */
        is_pos++;
        if (is_pos >= SAMPLES_PER_GR) break;
        md->is[gr][ch][is_pos] = (float)out[3]; /* w */
        is_pos++;
        if (is_pos >= SAMPLES_PER_GR) break;
        md->is[gr][ch][is_pos] = (float)out[0]; /* x */
        is_pos++;
        if (is_pos >= SAMPLES_PER_GR) break;
        md->is[gr][ch][is_pos] = (float)out[1]; /* y */
        is_pos++;
    }
    if (orc_bits_pos(m) > bit_pos_end + 1) is_pos -= 4;
    if (is_pos < 0) is_pos = 0;
    si->count1[gr][ch] = is_pos;
    for (; is_pos < SAMPLES_PER_GR; is_pos++) md->is[gr][ch][is_pos] = 0.0f;
    orc_bits_set_pos(m, bit_pos_end + 1);
    return ORC_OK;
}

/* maindata.go:119-188 */
static int get_scale_factors_mpeg2(orc_bits *m, uint32_t header, sideinfo_t *si, maindata_t *md, int32_t part2_starts[2][2]) {
    int nch = h_nch(header);
    memset(md->scalefac_l, 0, sizeof md->scalefac_l);
    memset(md->scalefac_s, 0, sizeof md->scalefac_s);
    for (int ch = 0; ch < nch; ch++) {
        int part2_start = orc_bits_pos(m);
        part2_starts[0][ch] = part2_start;
        int slen = g_nslen2[si->scalefac_compress[0][ch]];
        si->preflag[0][ch] = (slen >> 15) & 0x1;
        int n = 0;
        if (si->block_type[0][ch] == 2) {
            n++;
            if (si->mixed_block_flag[0][ch] != 0) n++;
        }
        int scale_factors[64];
        int count = 0;
        int d = (slen >> 12) & 0x7;
        for (int i = 0; i < 4; i++) {
            int num = slen & 0x7;
            slen >>= 3;
            if (num > 0) {
                for (int k = 0; k < SCALEFAC_SIZES_MPEG2[n][d][i]; k++) scale_factors[count++] = orc_bits_bits(m, num);
            } else {
                for (int k = 0; k < SCALEFAC_SIZES_MPEG2[n][d][i]; k++) scale_factors[count++] = 0;
            }
        }
        n = (n << 1) + 1;
        for (int k = 0; k < n; k++) scale_factors[count++] = 0;
        if (count == 22) {
            for (int i = 0; i < 22; i++) md->scalefac_l[0][ch][i] = scale_factors[i];
        } else {
            /* Go indexes scaleFactors[x*3+i] for x<13, i<3 (39 values).  Pure-short rows of
             * SCALEFAC_SIZES_MPEG2 sum to 36 (+3 pad) = 39; mixed rows sum to 33 (+5 pad) = 38,
             * so the reference panics (index out of range) on LSF mixed blocks. */
            if (count < 39) return ORC_ERR_REF_PANIC;
            for (int x = 0; x < 13; x++)
                for (int i = 0; i < 3; i++) md->scalefac_s[0][ch][x][i] = scale_factors[x * 3 + i];
        }
        int rc = read_huffman(m, header, si, md, part2_start, 0, ch);
        if (rc != ORC_OK) return rc;
    }
    return ORC_OK;
}

/* maindata.go:190-288 */
static int get_scale_factors_mpeg1(int nch, orc_bits *m, uint32_t header, sideinfo_t *si, maindata_t *md,
                                   int32_t part2_starts[2][2]) {
    memset(md->scalefac_l, 0, sizeof md->scalefac_l);
    memset(md->scalefac_s, 0, sizeof md->scalefac_s);
    for (int gr = 0; gr < 2; gr++) {
        for (int ch = 0; ch < nch; ch++) {
            int part2_start = orc_bits_pos(m);
            part2_starts[gr][ch] = part2_start;
            int slen1 = SCALEFAC_SIZES_MPEG1[si->scalefac_compress[gr][ch]][0];
            int slen2 = SCALEFAC_SIZES_MPEG1[si->scalefac_compress[gr][ch]][1];
            if (si->win_switch_flag[gr][ch] == 1 && si->block_type[gr][ch] == 2) {
                if (si->mixed_block_flag[gr][ch] != 0) {
                    for (int sfb = 0; sfb < 8; sfb++) md->scalefac_l[gr][ch][sfb] = orc_bits_bits(m, slen1);
                    for (int sfb = 3; sfb < 12; sfb++) {
                        int nbits = sfb < 6 ? slen1 : slen2;
                        for (int win = 0; win < 3; win++) md->scalefac_s[gr][ch][sfb][win] = orc_bits_bits(m, nbits);
                    }
                } else {
                    for (int sfb = 0; sfb < 12; sfb++) {
                        int nbits = sfb < 6 ? slen1 : slen2;
                        for (int win = 0; win < 3; win++) md->scalefac_s[gr][ch][sfb][win] = orc_bits_bits(m, nbits);
                    }
                }
            } else {
                static const int band_lo[4] = {0, 6, 11, 16}, band_hi[4] = {6, 11, 16, 21};
                for (int band = 0; band < 4; band++) {
                    int nbits = band < 2 ? slen1 : slen2;
                    if (si->scfsi[ch][band] == 0 || gr == 0) {
                        for (int sfb = band_lo[band]; sfb < band_hi[band]; sfb++)
                            md->scalefac_l[gr][ch][sfb] = orc_bits_bits(m, nbits);
                    } else if (si->scfsi[ch][band] == 1 && gr == 1) {
                        for (int sfb = band_lo[band]; sfb < band_hi[band]; sfb++)
                            md->scalefac_l[1][ch][sfb] = md->scalefac_l[0][ch][sfb];
                    }
                }
            }
            int rc = read_huffman(m, header, si, md, part2_start, gr, ch);
            if (rc != ORC_OK) return rc;
        }
    }
    return ORC_OK;
}

/* maindata.go:290-323.  prev == NULL on the first frame or right after Seek. */
static int maindata_read_bits(source_t *s, const mdbits_t *prev, int size, int offset, mdbits_t *out) {
    if (size > 1500) return ORC_ERR_MAINDATA_SIZE;
    int err;
    if (prev != NULL && offset > prev->len) {
        /* reservoir underflow: keep everything of prev, append own bytes, parse from bit 0 */
        uint8_t *vec = (uint8_t *)malloc((size_t)(prev->len + size) + 4);
        memcpy(vec, prev->vec, (size_t)prev->len);
        int n = src_read_full(s, vec + prev->len, size, &err);
        if (n < size) {
            free(vec);
            return ORC_ERR_UNEXPECTED_EOF; /* "maindata.Read (1)" */
        }
        out->vec = vec;
        out->len = prev->len + size;
        orc_bits_init(&out->bits, out->vec, out->len);
        return ORC_OK;
    }
    int keep = prev != NULL ? offset : 0;
    uint8_t *vec = (uint8_t *)malloc((size_t)(keep + size) + 4);
    if (keep > 0) memcpy(vec, prev->vec + (prev->len - keep), (size_t)keep);
    int n = src_read_full(s, vec + keep, size, &err);
    if (n < size) {
        free(vec);
        return ORC_ERR_UNEXPECTED_EOF; /* "maindata.Read (2)" */
    }
    out->vec = vec;
    out->len = keep + size;
    orc_bits_init(&out->bits, out->vec, out->len);
    return ORC_OK;
}

/* maindata.go:85-117 */
static int maindata_read(source_t *s, const mdbits_t *prev, uint32_t header, sideinfo_t *si, maindata_t *md,
                         mdbits_t *out, int32_t part2_starts[2][2]) {
    int nch = h_nch(header);
    int framesize;
    int rc = orc_header_frame_size(header, &framesize);
    if (rc != ORC_OK) return rc;
    if (framesize > 2000) return ORC_ERR_FRAMESIZE;
    int sideinfo_size = orc_header_side_info_size(header);
    int main_data_size = framesize - sideinfo_size - 4;
    if (h_protection_bit(header) == 0) main_data_size -= 2;
    if (main_data_size < 0) return ORC_ERR_UNEXPECTED_EOF; /* Go: make([]byte, negative) panics; tiny frames cannot occur with valid bitrates */
    rc = maindata_read_bits(s, prev, main_data_size, si->main_data_begin, out);
    if (rc != ORC_OK) return rc;
    if (h_lsf(header) == 1)
        rc = get_scale_factors_mpeg2(&out->bits, header, si, md, part2_starts);
    else
        rc = get_scale_factors_mpeg1(nch, &out->bits, header, si, md, part2_starts);
    if (rc != ORC_OK) {
        free(out->vec);
        out->vec = NULL;
    }
    return rc;
}

/* ------------------------------------------------------------------------ */
/* internal/frame  (frame.go)                                                */
/* ------------------------------------------------------------------------ */
typedef struct { /* frame.go:42-50 */
    uint32_t header;
    sideinfo_t side_info;
    maindata_t *main_data; /* shared/reused across frames, frame.go:95-99 */
    mdbits_t main_data_bits;
    float store[2][32][18];
    float v_vec[2][1024];
    int32_t part2_starts[2][2];
    int64_t position;
} frame_t;

static void frame_free(frame_t *f, int free_maindata) {
    if (!f) return;
    free(f->main_data_bits.vec);
    if (free_maindata) free(f->main_data);
    free(f);
}

/* frame.go:67-115 */
static int frame_read(source_t *s, int64_t position, frame_t *prev, frame_t **out) {
    uint32_t h;
    int64_t pos;
    int rc = frameheader_read(s, position, &h, &pos, NULL);
    if (rc != ORC_OK) return rc;
    if (h_protection_bit(h) == 0) { /* readCRC, frame.go:56-65 */
        uint8_t crc[2];
        int err;
        int n = src_read_full(s, crc, 2, &err);
        if (n < 2) return ORC_ERR_UNEXPECTED_EOF;
    }
    if (h_id(h) == 0) return ORC_ERR_MPEG25;
    if (h_layer(h) != 1) return ORC_ERR_LAYER;
    frame_t *nf = (frame_t *)calloc(1, sizeof(frame_t));
    rc = sideinfo_read(s, h, &nf->side_info);
    if (rc != ORC_OK) {
        free(nf);
        return rc;
    }
    maindata_t *md = prev ? prev->main_data : (maindata_t *)calloc(1, sizeof(maindata_t));
    rc = maindata_read(s, prev ? &prev->main_data_bits : NULL, h, &nf->side_info, md, &nf->main_data_bits,
                       nf->part2_starts);
    if (rc != ORC_OK) {
        if (!prev) free(md);
        free(nf);
        return rc;
    }
    nf->header = h;
    nf->main_data = md;
    nf->position = pos;
    if (prev) {
        memcpy(nf->store, prev->store, sizeof nf->store);
        memcpy(nf->v_vec, prev->v_vec, sizeof nf->v_vec);
    }
    *out = nf;
    return ORC_OK;
}

/* frame.go:140-156 */
static void requantize_process_long(frame_t *f, int gr, int ch, int is_pos, int sfb) {
    double sf_mult = 0.5;
    if (f->side_info.scalefac_scale[gr][ch] != 0) sf_mult = 1.0;
    double pf_x_pt = (double)f->side_info.preflag[gr][ch] * PRETAB[sfb];
    double idx = -(sf_mult * ((double)f->main_data->scalefac_l[gr][ch][sfb] + pf_x_pt)) +
                 0.25 * ((double)f->side_info.global_gain[gr][ch] - 210);
    double tmp1 = pow(2.0, idx);
    double tmp2;
    float v = f->main_data->is[gr][ch][is_pos];
    if (v < 0.0f)
        tmp2 = -g_powtab34[(int)(-v)];
    else
        tmp2 = g_powtab34[(int)v];
    f->main_data->is[gr][ch][is_pos] = (float)(tmp1 * tmp2);
}

/* frame.go:158-174 */
static void requantize_process_short(frame_t *f, int gr, int ch, int is_pos, int sfb, int win) {
    double sf_mult = 0.5;
    if (f->side_info.scalefac_scale[gr][ch] != 0) sf_mult = 1.0;
    double idx = -(sf_mult * (double)f->main_data->scalefac_s[gr][ch][sfb][win]) +
                 0.25 * ((double)f->side_info.global_gain[gr][ch] - 210.0 -
                         8.0 * (double)f->side_info.subblock_gain[gr][ch][win]);
    double tmp1 = pow(2.0, idx);
    double tmp2;
    float v = f->main_data->is[gr][ch][is_pos];
    if (v < 0)
        tmp2 = -g_powtab34[(int)(-v)];
    else
        tmp2 = g_powtab34[(int)v];
    f->main_data->is[gr][ch][is_pos] = (float)(tmp1 * tmp2);
}

/* frame.go:184-255 */
static void requantize(frame_t *f, int gr, int ch) {
    const int *sfl = SFB_LONG[h_lsf(f->header)][h_sampling_frequency(f->header)];
    const int *sfs = SFB_SHORT[h_lsf(f->header)][h_sampling_frequency(f->header)];
    sideinfo_t *si = &f->side_info;
    if (si->win_switch_flag[gr][ch] == 1 && si->block_type[gr][ch] == 2) {
        if (si->mixed_block_flag[gr][ch] != 0) {
            int sfb = 0;
            int next_sfb = sfl[sfb + 1];
            for (int i = 0; i < 36; i++) {
                if (i == next_sfb) {
                    sfb++;
                    next_sfb = sfl[sfb + 1];
                }
                requantize_process_long(f, gr, ch, i, sfb);
            }
            sfb = 3;
            next_sfb = sfs[sfb + 1] * 3;
            int win_len = sfs[sfb + 1] - sfs[sfb];
            for (int i = 36; i < si->count1[gr][ch];) {
                if (i == next_sfb) {
                    sfb++;
                    next_sfb = sfs[sfb + 1] * 3;
                    win_len = sfs[sfb + 1] - sfs[sfb];
                }
                for (int win = 0; win < 3; win++)
                    for (int j = 0; j < win_len; j++) {
                        requantize_process_short(f, gr, ch, i, sfb, win);
                        i++;
                    }
            }
        } else {
            int sfb = 0;
            int next_sfb = sfs[sfb + 1] * 3;
            int win_len = sfs[sfb + 1] - sfs[sfb];
            for (int i = 0; i < si->count1[gr][ch];) {
                if (i == next_sfb) {
                    sfb++;
                    next_sfb = sfs[sfb + 1] * 3;
                    win_len = sfs[sfb + 1] - sfs[sfb];
                }
                for (int win = 0; win < 3; win++)
                    for (int j = 0; j < win_len; j++) {
                        requantize_process_short(f, gr, ch, i, sfb, win);
                        i++;
                    }
            }
        }
    } else {
        int sfb = 0;
        int next_sfb = sfl[sfb + 1];
        for (int i = 0; i < si->count1[gr][ch]; i++) {
            if (i == next_sfb) {
                sfb++;
                next_sfb = sfl[sfb + 1];
            }
            requantize_process_long(f, gr, ch, i, sfb);
        }
    }
}

/* frame.go:257-302 */
static void reorder(frame_t *f, int gr, int ch) {
    float re[SAMPLES_PER_GR];
    memset(re, 0, sizeof re);
    const int *sfs = SFB_SHORT[h_lsf(f->header)][h_sampling_frequency(f->header)];
    sideinfo_t *si = &f->side_info;
    float *is = f->main_data->is[gr][ch];
    if (si->win_switch_flag[gr][ch] == 1 && si->block_type[gr][ch] == 2) {
        int sfb = 0;
        if (si->mixed_block_flag[gr][ch] != 0) sfb = 3;
        int next_sfb = sfs[sfb + 1] * 3;
        int win_len = sfs[sfb + 1] - sfs[sfb];
        int i = 36;
        if (sfb == 0) i = 0;
        while (i < SAMPLES_PER_GR) {
            if (i == next_sfb) {
                int j = 3 * sfs[sfb];
                memcpy(&is[j], re, sizeof(float) * (size_t)(3 * win_len));
                if (i >= si->count1[gr][ch]) return;
                sfb++;
                next_sfb = sfs[sfb + 1] * 3;
                win_len = sfs[sfb + 1] - sfs[sfb];
            }
            for (int win = 0; win < 3; win++)
                for (int j = 0; j < win_len; j++) {
                    re[j * 3 + win] = is[i];
                    i++;
                }
        }
        int j = 3 * sfs[12];
        memcpy(&is[j], re, sizeof(float) * (size_t)(3 * win_len));
    }
}

/* frame.go:308-330 */
static void stereo_process_intensity_long(frame_t *f, int gr, int sfb) {
    float is_ratio_l = 0, is_ratio_r = 0;
    int is_pos = f->main_data->scalefac_l[gr][0][sfb];
    if (is_pos < 7) {
        const int *sfl = SFB_LONG[h_lsf(f->header)][h_sampling_frequency(f->header)];
        int sfb_start = sfl[sfb], sfb_stop = sfl[sfb + 1];
        if (is_pos == 6) {
            is_ratio_l = 1.0f;
            is_ratio_r = 0.0f;
        } else {
            is_ratio_l = IS_RATIOS[is_pos] / (1.0f + IS_RATIOS[is_pos]);
            is_ratio_r = 1.0f / (1.0f + IS_RATIOS[is_pos]);
        }
        for (int i = sfb_start; i < sfb_stop; i++) {
            f->main_data->is[gr][0][i] *= is_ratio_l;
            f->main_data->is[gr][1][i] *= is_ratio_r;
        }
    }
}

/* frame.go:332-360 */
static void stereo_process_intensity_short(frame_t *f, int gr, int sfb) {
    float is_ratio_l = 0, is_ratio_r = 0;
    const int *sfs = SFB_SHORT[h_lsf(f->header)][h_sampling_frequency(f->header)];
    int win_len = sfs[sfb + 1] - sfs[sfb];
    for (int win = 0; win < 3; win++) {
        int is_pos = f->main_data->scalefac_s[gr][0][sfb][win];
        if (is_pos < 7) {
            int sfb_start = sfs[sfb] * 3 + win_len * win;
            int sfb_stop = sfb_start + win_len;
            if (is_pos == 6) {
                is_ratio_l = 1.0f;
                is_ratio_r = 0.0f;
            } else {
                is_ratio_l = IS_RATIOS[is_pos] / (1.0f + IS_RATIOS[is_pos]);
                is_ratio_r = 1.0f / (1.0f + IS_RATIOS[is_pos]);
            }
            for (int i = sfb_start; i < sfb_stop; i++) {
                f->main_data->is[gr][0][i] *= is_ratio_l;
                f->main_data->is[gr][1][i] *= is_ratio_r;
            }
        }
    }
}

/* frame.go:362-420 */
static void stereo(frame_t *f, int gr) {
    sideinfo_t *si = &f->side_info;
    if (h_use_ms(f->header)) {
        int i = 1;
        if (si->count1[gr][0] > si->count1[gr][1]) i = 0;
        int max_pos = si->count1[gr][i];
        const float inv_sqrt2 = (float)(M_SQRT2 / 2); /* untyped constant math.Sqrt2/2 -> float32 */
        for (int k = 0; k < max_pos; k++) {
            float left = (f->main_data->is[gr][0][k] + f->main_data->is[gr][1][k]) * inv_sqrt2;
            float right = (f->main_data->is[gr][0][k] - f->main_data->is[gr][1][k]) * inv_sqrt2;
            f->main_data->is[gr][0][k] = left;
            f->main_data->is[gr][1][k] = right;
        }
    }
    if (h_use_intensity(f->header)) {
        const int *sfl = SFB_LONG[h_lsf(f->header)][h_sampling_frequency(f->header)];
        const int *sfs = SFB_SHORT[h_lsf(f->header)][h_sampling_frequency(f->header)];
        if (si->win_switch_flag[gr][0] == 1 && si->block_type[gr][0] == 2) {
            if (si->mixed_block_flag[gr][0] != 0) {
                for (int sfb = 0; sfb < 8; sfb++)
                    if (sfl[sfb] >= si->count1[gr][1]) stereo_process_intensity_long(f, gr, sfb);
                for (int sfb = 3; sfb < 12; sfb++)
                    if (sfs[sfb] * 3 >= si->count1[gr][1]) stereo_process_intensity_short(f, gr, sfb);
            } else {
                for (int sfb = 0; sfb < 12; sfb++)
                    if (sfs[sfb] * 3 >= si->count1[gr][1]) stereo_process_intensity_short(f, gr, sfb);
            }
        } else {
            for (int sfb = 0; sfb < 21; sfb++)
                if (sfl[sfb] >= si->count1[gr][1]) stereo_process_intensity_long(f, gr, sfb);
        }
    }
}

/* frame.go:427-452 */
static void antialias(frame_t *f, int gr, int ch) {
    sideinfo_t *si = &f->side_info;
    if (si->win_switch_flag[gr][ch] == 1 && si->block_type[gr][ch] == 2 && si->mixed_block_flag[gr][ch] == 0) return;
    int sblim = 32;
    if (si->win_switch_flag[gr][ch] == 1 && si->block_type[gr][ch] == 2 && si->mixed_block_flag[gr][ch] == 1) sblim = 2;
    float *is = f->main_data->is[gr][ch];
    for (int sb = 1; sb < sblim; sb++) {
        for (int i = 0; i < 8; i++) {
            int li = 18 * sb - 1 - i;
            int ui = 18 * sb + i;
            float lb = is[li] * CS[i] - is[ui] * CA[i];
            float ub = is[ui] * CS[i] + is[li] * CA[i];
            is[li] = lb;
            is[ui] = ub;
        }
    }
}

/* imdct.go:83-108 */
static void imdct_win(float out[36], const float in[18], int block_type) {
    memset(out, 0, sizeof(float) * 36);
    if (block_type == 2) {
        const float *iwd = g_imdct_win[block_type];
        const int N = 12;
        for (int i = 0; i < 3; i++) {
            for (int p = 0; p < N; p++) {
                float sum = 0.0f;
                for (int m = 0; m < N / 2; m++) sum += in[i + 3 * m] * g_cos_n12[m][p];
                out[6 * i + p + 6] += sum * iwd[p];
            }
        }
        return;
    }
    const int N = 36;
    const float *iwd = g_imdct_win[block_type];
    for (int p = 0; p < N; p++) {
        float sum = 0.0f;
        for (int m = 0; m < N / 2; m++) sum += in[m] * g_cos_n36[m][p];
        out[p] = sum * iwd[p];
    }
}

/* frame.go:454-478 */
static void hybrid_synthesis(frame_t *f, int gr, int ch) {
    float in[18], rawout[36];
    sideinfo_t *si = &f->side_info;
    float *is = f->main_data->is[gr][ch];
    for (int sb = 0; sb < 32; sb++) {
        int bt = si->block_type[gr][ch];
        if (si->win_switch_flag[gr][ch] == 1 && si->mixed_block_flag[gr][ch] == 1 && sb < 2) bt = 0;
        for (int i = 0; i < 18; i++) in[i] = is[sb * 18 + i];
        imdct_win(rawout, in, bt);
        for (int i = 0; i < 18; i++) {
            is[sb * 18 + i] = rawout[i] + f->store[ch][sb][i];
            f->store[ch][sb][i] = rawout[i + 18];
        }
    }
}

/* frame.go:480-486 */
static void frequency_inversion(frame_t *f, int gr, int ch) {
    float *is = f->main_data->is[gr][ch];
    for (int sb = 1; sb < 32; sb += 2)
        for (int i = 1; i < 18; i += 2) is[sb * 18 + i] = -is[sb * 18 + i];
}

/* Go's int(float32) on amd64 is CVTTSS2SQ: NaN and out-of-int64-range inputs give
 * INT64_MIN.  Spelt out so the oracle does not depend on C undefined behaviour. */
static int64_t go_int_from_f32(float v) {
    if (!(v < 9223372036854775808.0f && v >= -9223372036854775808.0f)) return INT64_MIN;
    return (int64_t)v;
}

/* frame.go:630-688 */
static void subband_synthesis(frame_t *f, int gr, int ch, uint8_t *out) {
    float u_vec[512], s_vec[32];
    int nch = h_nch(f->header);
    float *v = f->v_vec[ch];
    const float *d = f->main_data->is[gr][ch];
    for (int ss = 0; ss < 18; ss++) {
        memmove(&v[64], &v[0], sizeof(float) * (1024 - 64));
        for (int i = 0; i < 32; i++) s_vec[i] = d[i * 18 + ss];
        for (int i = 0; i < 64; i++) {
            float sum = 0;
            for (int j = 0; j < 32; j++) sum += g_synth_nwin[i][j] * s_vec[j];
            v[i] = sum;
        }
        for (int i = 0; i < 512; i += 64) {
            memcpy(&u_vec[i], &v[i << 1], sizeof(float) * 32);
            memcpy(&u_vec[i + 32], &v[(i << 1) + 96], sizeof(float) * 32);
        }
        for (int i = 0; i < 512; i++) u_vec[i] *= g_synth_dtbl[i];
        for (int i = 0; i < 32; i++) {
            float sum = 0;
            for (int j = 0; j < 512; j += 32) sum += u_vec[j + i];
            int64_t samp = go_int_from_f32(sum * 32767);
            if (samp > 32767)
                samp = 32767;
            else if (samp < -32767)
                samp = -32767;
            int16_t s = (int16_t)samp;
            int idx = 4 * (32 * ss + i);
            if (nch == 1) {
                out[idx] = (uint8_t)s;
                out[idx + 1] = (uint8_t)((uint16_t)s >> 8);
                out[idx + 2] = (uint8_t)s;
                out[idx + 3] = (uint8_t)((uint16_t)s >> 8);
                continue;
            }
            if (ch == 0) {
                out[idx] = (uint8_t)s;
                out[idx + 1] = (uint8_t)((uint16_t)s >> 8);
            } else {
                out[idx + 2] = (uint8_t)s;
                out[idx + 3] = (uint8_t)((uint16_t)s >> 8);
            }
        }
    }
}

static void tap_copy(float *dst, int frame_idx, int gr, int ch, const float *src) {
    if (dst) memcpy(dst + (((size_t)frame_idx * 2 + gr) * 2 + ch) * 576, src, sizeof(float) * 576);
}

/* frame.go:121-138.  taps (optional) records each stage. */
static void frame_decode(frame_t *f, uint8_t *out, orc_taps *taps) {
    int nch = h_nch(f->header);
    int fi = -1;
    if (taps && taps->n_frames < taps->capacity_frames) {
        fi = taps->n_frames++;
        if (taps->header) taps->header[fi] = f->header;
        if (taps->main_data_begin) taps->main_data_begin[fi] = f->side_info.main_data_begin;
        if (taps->position) taps->position[fi] = f->position;
        for (int gr = 0; gr < 2; gr++)
            for (int ch = 0; ch < 2; ch++) {
                size_t u = ((size_t)fi * 2 + gr) * 2 + ch;
                int live = gr < h_granules(f->header) && ch < nch;
                if (taps->is)
                    for (int i = 0; i < 576; i++) taps->is[u * 576 + i] = live ? (int16_t)f->main_data->is[gr][ch][i] : 0;
                if (taps->count1) taps->count1[u] = live ? f->side_info.count1[gr][ch] : 0;
                if (taps->part2_start) taps->part2_start[u] = live ? f->part2_starts[gr][ch] : 0;
                if (taps->scalefac_l)
                    for (int i = 0; i < 22; i++) taps->scalefac_l[u * 22 + i] = live ? (uint8_t)f->main_data->scalefac_l[gr][ch][i] : 0;
                if (taps->scalefac_s)
                    for (int i = 0; i < 39; i++)
                        taps->scalefac_s[u * 39 + i] = live ? (uint8_t)f->main_data->scalefac_s[gr][ch][i / 3][i % 3] : 0;
            }
    }
    for (int gr = 0; gr < h_granules(f->header); gr++) {
        for (int ch = 0; ch < nch; ch++) {
            requantize(f, gr, ch);
            if (fi >= 0) tap_copy(taps->xr_requant, fi, gr, ch, f->main_data->is[gr][ch]);
            reorder(f, gr, ch);
            if (fi >= 0) tap_copy(taps->xr_reorder, fi, gr, ch, f->main_data->is[gr][ch]);
        }
        stereo(f, gr);
        if (fi >= 0)
            for (int ch = 0; ch < nch; ch++) tap_copy(taps->xr_stereo, fi, gr, ch, f->main_data->is[gr][ch]);
        for (int ch = 0; ch < nch; ch++) {
            antialias(f, gr, ch);
            if (fi >= 0) tap_copy(taps->xr_alias, fi, gr, ch, f->main_data->is[gr][ch]);
            hybrid_synthesis(f, gr, ch);
            frequency_inversion(f, gr, ch);
            if (fi >= 0) tap_copy(taps->hybrid, fi, gr, ch, f->main_data->is[gr][ch]);
            subband_synthesis(f, gr, ch, out + SAMPLES_PER_GR * 4 * gr);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* package mp3: Decoder  (decode.go)                                         */
/* ------------------------------------------------------------------------ */
struct orc_decoder { /* decode.go:34-43 */
    source_t source;
    int sample_rate;
    int64_t length;
    int64_t *frame_starts;
    int n_frame_starts, cap_frame_starts;
    uint8_t *buf; /* d.buf: pending decoded bytes */
    size_t buf_off, buf_len, buf_cap;
    frame_t *frame;
    int64_t pos;
    int64_t bytes_per_frame;
    orc_taps *taps;
};

#define INVALID_LENGTH (-1)

static void dec_buf_append(orc_decoder *d, const uint8_t *p, size_t n) {
    if (d->buf_off > 0 && d->buf_off == d->buf_len) d->buf_off = d->buf_len = 0;
    if (d->buf_len + n > d->buf_cap) {
        size_t live = d->buf_len - d->buf_off;
        size_t cap = d->buf_cap ? d->buf_cap : 16384;
        while (cap < live + n) cap *= 2;
        uint8_t *nb = (uint8_t *)malloc(cap);
        if (live) memcpy(nb, d->buf + d->buf_off, live);
        free(d->buf);
        d->buf = nb;
        d->buf_cap = cap;
        d->buf_off = 0;
        d->buf_len = live;
    }
    memcpy(d->buf + d->buf_len, p, n);
    d->buf_len += n;
}

/* decode.go:45-67 */
static int dec_read_frame(orc_decoder *d) {
    frame_t *nf = NULL;
    int rc = frame_read(&d->source, d->source.pos, d->frame, &nf);
    if (rc != ORC_OK) {
        /* d.frame is overwritten with nil on error (decode.go:47) */
        if (d->frame) {
            frame_free(d->frame, 1);
            d->frame = NULL;
        }
        if (rc == ORC_EOF || rc == ORC_ERR_UNEXPECTED_EOF || rc == ORC_ERR_SYNC_LIMIT) return ORC_EOF;
        return rc;
    }
    if (d->frame) frame_free(d->frame, 0); /* main_data moved to nf */
    d->frame = nf;
    uint8_t out[4608];
    int nbytes = orc_header_bytes_per_frame(nf->header);
    memset(out, 0, sizeof out);
    frame_decode(nf, out, d->taps);
    dec_buf_append(d, out, (size_t)nbytes);
    return ORC_OK;
}

/* decode.go:154-216 */
static int dec_ensure_frame_starts_and_length(orc_decoder *d) {
    if (d->length != INVALID_LENGTH) return ORC_OK;
    if (!d->source.seekable) return ORC_OK;
    int64_t pos;
    int rc = src_seek(&d->source, 0, 1, &pos);
    if (rc != ORC_OK) return rc;
    d->source.pos = 0; /* rewind, source.go:85-92 */
    rc = src_skip_tags(&d->source);
    if (rc != ORC_OK) return rc;
    int64_t l = 0;
    for (;;) {
        uint32_t h;
        int64_t fpos;
        rc = frameheader_read(&d->source, d->source.pos, &h, &fpos, NULL);
        if (rc != ORC_OK) {
            if (rc == ORC_EOF || rc == ORC_ERR_UNEXPECTED_EOF || rc == ORC_ERR_SYNC_LIMIT) break;
            return rc;
        }
        if (d->n_frame_starts == d->cap_frame_starts) {
            d->cap_frame_starts = d->cap_frame_starts ? d->cap_frame_starts * 2 : 1024;
            d->frame_starts = (int64_t *)realloc(d->frame_starts, sizeof(int64_t) * (size_t)d->cap_frame_starts);
        }
        d->frame_starts[d->n_frame_starts++] = fpos;
        d->bytes_per_frame = orc_header_bytes_per_frame(h);
        l += d->bytes_per_frame;
        int framesize;
        rc = orc_header_frame_size(h, &framesize);
        if (rc != ORC_OK) return rc;
        rc = src_seek(&d->source, (int64_t)(framesize - 4), 1, NULL);
        if (rc != ORC_OK) return rc;
    }
    d->length = l;
    return src_seek(&d->source, pos, 0, NULL);
}

orc_decoder *orc_new_decoder(const uint8_t *data, size_t len, int seekable, int *err) { /* decode.go:361-388 */
    ensure_init();
    orc_decoder *d = (orc_decoder *)calloc(1, sizeof(orc_decoder));
    d->source.data = data;
    d->source.len = len;
    d->source.pos = 0;
    d->source.seekable = seekable;
    d->length = INVALID_LENGTH;
    int rc = src_skip_tags(&d->source);
    if (rc == ORC_OK) rc = dec_read_frame(d);
    if (rc == ORC_OK) {
        int freq = orc_header_sampling_frequency_value(d->frame->header);
        if (freq == 0)
            rc = ORC_ERR_SAMPLE_RATE;
        else
            d->sample_rate = freq;
    }
    if (rc == ORC_OK) rc = dec_ensure_frame_starts_and_length(d);
    if (rc != ORC_OK) {
        if (err) *err = rc;
        orc_free_decoder(d);
        return NULL;
    }
    if (err) *err = ORC_OK;
    return d;
}

void orc_free_decoder(orc_decoder *d) {
    if (!d) return;
    if (d->frame) frame_free(d->frame, 1);
    free(d->frame_starts);
    free(d->buf);
    free(d);
}
void orc_set_taps(orc_decoder *d, orc_taps *taps) { d->taps = taps; }
void orc_free(void *p) { free(p); }

long orc_read(orc_decoder *d, uint8_t *buf, size_t n, int *err) { /* decode.go:70-80 */
    while (d->buf_len - d->buf_off == 0) {
        int rc = dec_read_frame(d);
        if (rc != ORC_OK) {
            if (err) *err = rc;
            return 0;
        }
    }
    size_t live = d->buf_len - d->buf_off;
    size_t c = n < live ? n : live;
    memcpy(buf, d->buf + d->buf_off, c);
    d->buf_off += c;
    d->pos += (int64_t)c;
    if (err) *err = ORC_OK;
    return (long)c;
}

long orc_read_all(orc_decoder *d, uint8_t **out, int *err) {
    size_t cap = 1 << 20, len = 0;
    uint8_t *o = (uint8_t *)malloc(cap);
    for (;;) {
        if (cap - len < 65536) {
            cap *= 2;
            o = (uint8_t *)realloc(o, cap);
        }
        int rc;
        long n = orc_read(d, o + len, cap - len, &rc);
        if (n == 0) {
            if (err) *err = (rc == ORC_EOF) ? ORC_OK : rc;
            break;
        }
        len += (size_t)n;
    }
    *out = o;
    return (long)len;
}

int64_t orc_seek(orc_decoder *d, int64_t offset, int whence, int *err) { /* decode.go:89-145 */
    int e = ORC_OK;
    int64_t ret = 0;
    if (offset == 0 && whence == 1) {
        if (err) *err = ORC_OK;
        return d->pos;
    }
    int64_t npos = 0;
    switch (whence) {
    case 0: npos = offset; break;
    case 1: npos = d->pos + offset; break;
    case 2: npos = d->length + offset; break;
    default:
        if (err) *err = ORC_ERR_WHENCE;
        return 0;
    }
    d->pos = npos;
    d->buf_off = d->buf_len = 0;
    if (d->frame) {
        frame_free(d->frame, 1);
        d->frame = NULL;
    }
    if (d->pos < 0) d->pos = 0;
    if (d->length != INVALID_LENGTH && d->pos >= d->length) {
        if (err) *err = ORC_OK;
        return npos;
    }
    if (d->bytes_per_frame == 0 || d->frame_starts == NULL) { /* Go: divide by zero / index panic on non-seekable */
        if (err) *err = ORC_ERR_SEEK_UNSUPPORTED;
        return 0;
    }
    int64_t f = d->pos / d->bytes_per_frame;
    if (f > 0) {
        f--;
        e = src_seek(&d->source, d->frame_starts[f], 0, NULL);
        if (e == ORC_OK) e = dec_read_frame(d);
        if (e == ORC_OK) e = dec_read_frame(d);
        if (e == ORC_OK) {
            size_t drop = (size_t)(d->bytes_per_frame + (d->pos % d->bytes_per_frame));
            d->buf_off += drop; /* d.buf = d.buf[bytesPerFrame + pos%bytesPerFrame:] */
        }
    } else {
        e = src_seek(&d->source, d->frame_starts[f], 0, NULL);
        if (e == ORC_OK) e = dec_read_frame(d);
        if (e == ORC_OK) d->buf_off += (size_t)d->pos;
    }
    if (e != ORC_OK) {
        if (err) *err = e;
        return 0;
    }
    ret = npos;
    if (err) *err = ORC_OK;
    return ret;
}

int orc_sample_rate(const orc_decoder *d) { return d->sample_rate; }
int64_t orc_length(const orc_decoder *d) { return d->length; }
int64_t orc_bytes_per_frame(const orc_decoder *d) { return d->bytes_per_frame; }
static int64_t bytes_to_duration(const orc_decoder *d, int64_t bytes) { /* decode.go:344-348 */
    return (int64_t)1000000000 * bytes / (int64_t)(d->sample_rate * 4);
}
static int64_t duration_to_bytes(const orc_decoder *d, int64_t dur) { /* decode.go:351-354 */
    return dur * (int64_t)(d->sample_rate * 4) / (int64_t)1000000000;
}
int64_t orc_duration_ns(const orc_decoder *d) {
    if (d->length == INVALID_LENGTH) return -1;
    return bytes_to_duration(d, d->length);
}
int64_t orc_position_ns(const orc_decoder *d) { return bytes_to_duration(d, d->pos); }
int64_t orc_remaining_ns(const orc_decoder *d) {
    int64_t dur = orc_duration_ns(d);
    if (dur < 0) return -1;
    return dur - orc_position_ns(d);
}
double orc_progress(const orc_decoder *d) {
    if (d->length == INVALID_LENGTH) return -1;
    if (d->length == 0) return 0;
    return (double)d->pos / (double)d->length;
}
int64_t orc_sample_position(const orc_decoder *d) { return d->pos / 4; }
int64_t orc_sample_count(const orc_decoder *d) {
    if (d->length == INVALID_LENGTH) return -1;
    return d->length / 4;
}
int orc_seek_to_sample(orc_decoder *d, int64_t sample) {
    if (d->length == INVALID_LENGTH) return ORC_ERR_SEEK_UNSUPPORTED;
    if (sample < 0) sample = 0;
    int64_t max_samples = orc_sample_count(d);
    if (sample > max_samples) sample = max_samples;
    int err;
    orc_seek(d, sample * 4, 0, &err);
    return err;
}
int orc_seek_to_time(orc_decoder *d, int64_t t) {
    if (d->length == INVALID_LENGTH) return ORC_ERR_SEEK_UNSUPPORTED;
    if (t < 0) t = 0;
    int64_t max_dur = orc_duration_ns(d);
    if (t > max_dur) t = max_dur;
    int64_t bytes = duration_to_bytes(d, t);
    bytes &= ~(int64_t)3;
    int err;
    orc_seek(d, bytes, 0, &err);
    return err;
}
int orc_skip(orc_decoder *d, int64_t delta) { return orc_seek_to_time(d, orc_position_ns(d) + delta); }
int orc_num_frame_starts(const orc_decoder *d) { return d->n_frame_starts; }
int64_t orc_frame_start(const orc_decoder *d, int i) { return d->frame_starts[i]; }

const char *orc_error_string(int err) {
    switch (err) {
    case ORC_OK: return "ok";
    case ORC_EOF: return "EOF";
    case ORC_ERR_UNEXPECTED_EOF: return "mp3: unexpected EOF";
    case ORC_ERR_SYNC_LIMIT: return "mp3: no valid frame header found within 65536 bytes";
    case ORC_ERR_FREE_FORMAT: return "mp3: free bitrate format is not supported";
    case ORC_ERR_MPEG25: return "mp3: MPEG version 2.5 is not supported";
    case ORC_ERR_LAYER: return "mp3: only layer3 is supported";
    case ORC_ERR_FRAMESIZE: return "mp3: framesize too large";
    case ORC_ERR_MAINDATA_SIZE: return "mp3: main data size too large";
    case ORC_ERR_ISPOS: return "mp3: isPos was too big";
    case ORC_ERR_REGION_INDEX: return "mp3: readHuffman failed: invalid index";
    case ORC_ERR_HUFFMAN: return "mp3: illegal Huff code in data";
    case ORC_ERR_SEEK_UNSUPPORTED: return "mp3: seek not supported on non-seekable source";
    case ORC_ERR_WHENCE: return "mp3: invalid whence";
    case ORC_ERR_SAMPLE_RATE: return "mp3: frame header has invalid sample frequency";
    case ORC_ERR_REF_PANIC: return "mp3: reference implementation panics on this input";
    }
    return "mp3: internal error";
}

/* ------------------------------------------------------------------------ */
/* CPU baseline driver: one thread per stream, NewDecoder + io.ReadAll        */
/* (bench_test.go:40-55).                                                    */
/* ------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *const *data;
    const size_t *lens;
    int n;
    int next;
    pthread_mutex_t mu;
    int64_t *pcm_bytes;
    uint64_t *checksum;
} mt_job_t;

static void *mt_worker(void *arg) {
    mt_job_t *job = (mt_job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int i = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (i >= job->n) break;
        int err;
        int64_t total = 0;
        uint64_t h = 1469598103934665603ULL;
        orc_decoder *d = orc_new_decoder(job->data[i], job->lens[i], 1, &err);
        if (d) {
            uint8_t *out = NULL;
            long n = orc_read_all(d, &out, &err);
            total = n;
            for (long k = 0; k < n; k++) {
                h ^= out[k];
                h *= 1099511628211ULL;
            }
            free(out);
            orc_free_decoder(d);
        }
        if (job->pcm_bytes) job->pcm_bytes[i] = total;
        if (job->checksum) job->checksum[i] = h;
    }
    return NULL;
}

double orc_decode_streams_mt(const uint8_t *const *data, const size_t *lens, int n, int threads, int64_t *pcm_bytes,
                             uint64_t *checksum) {
    ensure_init();
    if (threads < 1) threads = 1;
    mt_job_t job = {data, lens, n, 0, PTHREAD_MUTEX_INITIALIZER, pcm_bytes, checksum};
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, mt_worker, &job);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    free(th);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
