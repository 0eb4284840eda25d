// kernels.cuh — sm_100a device kernels of the MP3 Layer III granule decode path.
//
//   k_huffman (K1)    : scalefactors + Huffman            maindata.go:119-288, maindata/huffman.go:27-138,
//                                                         huffman.go:348-419, bits.go:45-86
//   k_hybrid  (K2+K3) : requantise, reorder, stereo,      frame.go:140-452
//                       alias reduction;
//                       IMDCT + window + overlap-add      frame.go:454-486, imdct.go:83-108
//                       + frequency inversion
//   k_synth   (K4)    : polyphase synthesis + int16 store frame.go:630-688
//
// Design notes (see DESIGN.md for the full derivation):
//  * K1 is one THREAD per granule-channel: the code stream of a unit is serial, so the
//    parallelism is across units.  Code tables are multi-level LUTs staged in shared memory;
//    the bit cursor reproduces bits.go's out-of-bounds rule (reads at/after the frame's logical
//    buffer end return 0 and do not advance).
//  * K2 and K3 are fused into k_hybrid: one warp marches through a segment of consecutive granules
//    and keeps the IMDCT overlap itself; a one-granule halo in front of every segment rebuilds it.
//  * K4 is k_synth: thread-per-slot matrixing with compile-time (immediate) coefficients, then a
//    warp-per-30-slots window pass with the V history in registers; a 16-slot halo per CTA.
//    The reference's direct-form summation ORDER is kept (m / j / tap ascending).
//  * Both kernels are kept small enough for the 32 KB L1.5 instruction cache: a fully fused
//    K2+K3+K4 kernel was measured first and starved on instruction fetch (profiles/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mp3gpu.h"
#include "unit_logic.h"

#ifndef MP3GPU_EXACT
#define MP3GPU_EXACT 0
#endif

namespace mp3gpu {

// sum + a*b : fused (one rounding) in the fast build; two roundings (the reference's amd64
// arithmetic, Go does not fuse on amd64) in the exact build.
__device__ __forceinline__ float mac(float a, float b, float sum) {
#if MP3GPU_EXACT
    return __fadd_rn(sum, __fmul_rn(a, b));
#else
    return __fmaf_rn(a, b, sum);
#endif
}

// ------------------------------------------------------------------------------------------
// Tables.  The IMDCT cosine/window tables are lane-uniform and compile-time (const_tables.inc,
// generated at build time from tables.cc), so every IMDCT multiply-add is an FFMA with an
// immediate operand: no constant-cache or shared-memory traffic for coefficients.
// ------------------------------------------------------------------------------------------
#include "const_tables.inc"
__constant__ float c_win[4 * 36];  // IMDCT window rows, indexed by the run-time block type (warp-uniform except quirk Q15)
__constant__ float c_winz[4 * 36]; // fast build: window x output scale/sign of the fast 36-point IMDCT (imdct36_emit)

// Wave-local intermediate buffers.  Index j = local granule (g - wave_first).  k_hybrid looks one
// granule back (halo) and k_synth 16 time slots, so K1's outputs and hyb have valid look-back slots in
// front, carried over from the previous wave.
struct WaveBufs {
    int16_t *is16;     // [-2..nw)[2][576]
    uint32_t *meta;    // [-2..nw)[2]   bits 0..9 count1, bit 10 preflag (after LSF derivation)
    uint32_t *sfpack;  // [-2..nw)[2][8] scalefactors as nibbles: n = sfb (long), 22 + sfb*3+win (short)
    float *hyb;        // [-1..nw)[2][18][32] subband samples (k_hybrid -> k_synth), one look-back granule in front
    float *tap_xr;     // debug tap (opts.keep_intermediates): [nw][2][576] index sb*18+m, else nullptr
    const float *synth_d;  // [512]     frame.go:499-628
    unsigned int *work_counter;  // dynamic segment scheduler of k_hybrid; work_counter[1]: tile scheduler of k_huffman
    // extents, read by the guards of the checked build only (MP3_CHECK, unit_logic.h)
    long long units_total;       // unit slots of the whole submission (2 per granule)
    int n_gran;                  // granules of this wave: is16 / meta / sfpack hold granules [-2, n_gran), hyb [-1, n_gran)
#if MP3GPU_PROBE
    int probe;                   // diagnostic build (make probe): 1 k_hybrid does not store hyb, 2 k_hybrid reads K1's output from
                                 // 32 L2-resident granules, 4 k_synth does not store PCM, 8 k_synth reads hyb from 512 resident slots
#endif
};
#if MP3GPU_PROBE
#define MP3_PROBE(B, bit) ((B).probe & (bit))
#else
#define MP3_PROBE(B, bit) 0
#endif

// ------------------------------------------------------------------------------------------
// K1: scalefactors + Huffman.  One thread per unit slot; per-unit logic in unit_logic.h.
// ------------------------------------------------------------------------------------------
// One persistent CTA per SM; its warps share the code tables (30 KB, staged once) and otherwise never meet: every WARP
// pulls tiles of UPW (32 or 64) consecutive units from a global counter and handles a tile on its own —
//  1. Work order.  A unit's decode time is proportional to its number of code words, and the 32 lanes of a warp wait
//     for the slowest.  With UPW = 64 the warp sorts the tile's units by big_values (counting sort in its private bins)
//     and decodes the 32 longer ones first, then the 32 shorter ones, so that the lanes of a pass get similar lengths.
//  2. Staging.  The units of a tile are consecutive in stream order, so the bits they read are one contiguous stretch
//     of main_data (streams lie back to back).  The warp copies that stretch into its private staging area with
//     coalesced 16-byte loads, byte-swapped once into big-endian bit order; a unit's cursor is then a plain bit position
//     (StagedCursor / FastWindow, unit_logic.h) — no divergent refill from global memory.  What lies outside the stretch
//     (a tile whose stretch exceeds the staging capacity, malformed descriptors) is read from global memory word by word.
// The first version of this design did 1. and 2. per CTA (256-512 units sorted and staged at once).  It sorted better, but
// its four CTA barriers per tile left the warps with the shorter units and all warps during the staging loads idle: a
// third of the stall samples sat on the barriers (profiles/r02_k1_history.md); there is no __syncthreads in the loop now.
template <int UPW>
__global__ void __launch_bounds__(1024, 1)
k_huffman(const uint8_t *__restrict__ main_data, unsigned long long main_bits, const mp3gpu_unit *__restrict__ units,
          long long first_unit, int n_units, DeviceTables T, WaveBufs B, int stage_cap16, unsigned int *__restrict__ tile_counter) {
    extern __shared__ __align__(16) uint32_t s_dyn32[];  // pair-tree code tables (uint16 entries), then one block per warp
    __shared__ uint64_t s_quad[256];
    __shared__ uint32_t s_qlut[512];
    __shared__ uint32_t s_desc[34];
    constexpr int NP = UPW / 32;  // units per lane and tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SmemRef s_lut = SmemRef::of(s_dyn32);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.huff_lut);
        uint4 *dst = reinterpret_cast<uint4 *>(s_dyn32);
#pragma unroll 4
        for (int i = threadIdx.x; i < T.huff_lut_n / 8; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < 34) s_desc[threadIdx.x] = T.huff_desc[threadIdx.x];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_quad[i] = T.quad_signs[i];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_qlut[i] = T.quad_lut[i];
    __syncthreads();  // the only CTA barrier: the tables are staged
    // this warp's block: staging area (+ 16 bytes the FastWindow prefetch may touch), sort bins, work order
    const int warp_words = stage_cap16 * 4 + 4 + 40 + UPW / 4;
    uint32_t *const s_stage = s_dyn32 + T.huff_lut_n / 2 + warp * warp_words;  // huff_lut_n is a multiple of 8 entries: 16-byte aligned
    unsigned int *const s_bin = s_stage + stage_cap16 * 4 + 4;
    uint8_t *const s_order = reinterpret_cast<uint8_t *>(s_bin + 40);
    const uint32_t main16 = (uint32_t)(((main_bits >> 3) + 48) >> 4);  // 16-byte chunks that may be read: main_data is followed by 64 bytes of padding
    const int n_tiles = (n_units + UPW - 1) / UPW;
#pragma unroll 1
    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const int base = tile * UPW;
        // ---- the stretch of main data the tile reads, and the work order inside the tile ---------------------
        int key[NP];
        uint32_t lo16 = 0xffffffffu, hi16 = 0u;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            key[j] = 38;  // beyond the wave
            const int ul0 = base + j * 32 + lane;
            if (ul0 < n_units) {
                if (!MP3_CHECK(first_unit + ul0 >= 0 && first_unit + ul0 < B.units_total, first_unit + ul0)) continue;
                const mp3gpu_unit u = units[first_unit + ul0];
                key[j] = 37;
                if (u_valid(u.w2)) {
                    key[j] = 36 - ((u_p23len(u.w0) == 0 ? 0 : imin(u_bigval(u.w0), 288)) >> 3);
                    uint32_t l, h;
                    stage_reach(u, main_bits, &l, &h);
                    lo16 = l < lo16 ? l : lo16;
                    hi16 = h > hi16 ? h : hi16;
                }
            }
        }
        lo16 = __reduce_min_sync(0xffffffffu, lo16);
        hi16 = __reduce_max_sync(0xffffffffu, hi16);
        if (NP > 1) {
            s_bin[lane] = 0;
            if (lane < 8) s_bin[32 + lane] = 0;
            __syncwarp();
            unsigned int rank[NP];
#pragma unroll
            for (int j = 0; j < NP; j++) rank[j] = atomicAdd(&s_bin[key[j]], 1u);
            __syncwarp();
            {   // exclusive scan of the 39 bins: lanes 0..31 hold bins 0..31, lanes 0..6 also bins 32..38
                const unsigned int c0 = s_bin[lane];
                unsigned int x = c0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned int y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += y;
                }
                const unsigned int total32 = __shfl_sync(0xffffffffu, x, 31);
                const unsigned int c1 = lane < 7 ? s_bin[32 + lane] : 0u;
                unsigned int z = c1;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    const unsigned int y = __shfl_up_sync(0xffffffffu, z, d);
                    if (lane >= d) z += y;
                }
                __syncwarp();
                s_bin[lane] = x - c0;
                if (lane < 7) s_bin[32 + lane] = total32 + z - c1;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NP; j++) s_order[s_bin[key[j]] + rank[j]] = (uint8_t)(j * 32 + lane);
        }
        StageCtx S;
        S.sw = SmemRef::of(s_stage);
        S.gw = reinterpret_cast<const uint32_t *>(main_data);
        S.main_bits = main_bits;
        {
            const uint32_t hi = hi16 < main16 ? hi16 : main16;
            uint32_t n16 = hi > lo16 ? hi - lo16 : 0u;  // no valid unit in the tile: lo = ~0
            if (n16 > (uint32_t)stage_cap16) n16 = (uint32_t)stage_cap16;
            S.n_words = (int)(n16 * 4);
            S.lo_word = (unsigned long long)lo16 * 4ull;
            const uint4 *src = reinterpret_cast<const uint4 *>(main_data) + lo16;
            uint4 *dst = reinterpret_cast<uint4 *>(s_stage);
#pragma unroll 4
            for (uint32_t i = lane; i < n16; i += 32) {
                if (!MP3_CHECK(lo16 + i < main16, lo16 + i)) continue;
                uint4 v = __ldg(src + i);
                v.x = be32(v.x); v.y = be32(v.y); v.z = be32(v.z); v.w = be32(v.w);
                dst[i] = v;
            }
        }
        __syncwarp();
        // ---- decode: the longer units first -------------------------------------------------------------------
#pragma unroll 1
        for (int pass = 0; pass < NP; pass++) {
            const int ul = base + (NP > 1 ? (int)s_order[pass * 32 + lane] : lane);  // wave-local unit index
            if (ul < n_units && MP3_CHECK(ul >= 0 && ul < 2 * B.n_gran && first_unit + ul < B.units_total, ul)) {
                if (!u_valid(units[first_unit + ul].w2)) {
                    B.meta[ul] = 0;
                } else {
                    uint32_t pk[8];
                    uint32_t *out = reinterpret_cast<uint32_t *>(B.is16 + (size_t)ul * 576);
                    const uint32_t meta = huffman_unit_staged(T, s_lut, s_qlut, s_desc, s_quad, S, units, first_unit + ul, pk, out);
                    uint4 *dst = reinterpret_cast<uint4 *>(B.sfpack + (size_t)ul * 8);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    B.meta[ul] = meta;
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_hybrid = K2 + K3: one warp marches through a segment of consecutive granules (both channels).
//
//   K2  requantise + reorder + stereo + alias reduction      frame.go:140-452   lane = spectral line (mod 32)
//   K3  IMDCT + window + overlap-add + frequency inversion   frame.go:454-486   lane = subband
//
// The IMDCT overlap `store` (frame.go:48) stays with the warp in shared memory.  A segment starts by
// re-running the granule in front of it (halo) for its overlap output only, which recreates exactly
// the state a linear decode has at that point; a granule flagged ZERO_STATE clears the state.
// Output: subband samples hyb[g][ch][slot][sb] in global memory (2,304 B per granule-channel).
//
// The reference's summation orders are kept (m ascending), so the no-contraction build
// (MP3GPU_EXACT) is bit-identical to the reference and the FFMA build differs only by fused
// rounding.  Symmetries that hold BITWISE in the float32 tables are exploited:
// cos36[m][17-p] = -cos36[m][p], cos36[m][53-p] = cos36[m][p] (and the 12-point analogues), which
// halves the IMDCT multiply-adds with bit-identical results (negation is exact in IEEE 754).
// ------------------------------------------------------------------------------------------
constexpr int kXrStride = 19;              // padded subband stride of the spectrum staging (conflict-free lane = subband reads)
constexpr int kXrFloats = 32 * kXrStride;  // 608

__device__ __forceinline__ int xr_pad(int i) { return i + i / 18; }

// Value of a sum whose every term is the exact negation of the terms of `sum`, accumulated in the same order
// from +0: -sum, except that a zero sum stays +0 (accumulators start at +0, and +0 + -0 = +0).
__device__ __forceinline__ float mirror_neg(float sum) { return __fsub_rn(0.0f, sum); }

// sum_{m=0..17} in[m] * cos36[m][P], m ascending from 0 (imdct.go:101-107); three outputs at a time for ILP
template <int P0, int P1, int P2>
__device__ __forceinline__ void dot36x3(const float (&in)[18], float &r0, float &r1, float &r2) {
    float a = 0.0f, b = 0.0f, c = 0.0f;
#pragma unroll
    for (int m = 0; m < 18; m++) {
        a = mac(in[m], kCos36[m][P0], a);
        b = mac(in[m], kCos36[m][P1], b);
        c = mac(in[m], kCos36[m][P2], c);
    }
    r0 = a; r1 = b; r2 = c;
}
// sum_{m=0..5} in[i + 3m] * cos12[m][P]  (imdct.go:88-96)
template <int I, int P>
__device__ __forceinline__ float dot12(const float (&in)[18]) {
    float sum = 0.0f;
#pragma unroll
    for (int m = 0; m < 6; m++) sum = mac(in[I + 3 * m], kCos12[m][P], sum);
    return sum;
}

// Emits windowed IMDCT outputs of one subband: first-half outputs are overlap-added, frequency
// inverted and written as subband samples hyb[i][lane]; second-half outputs become the new overlap.
struct HybridSink {
    float *hyb;   // global [18][32] (this granule-channel)
    float *ov;    // shared [18][32] overlap state (this channel)
    int lane;
    bool first_half;  // false: only the overlap is wanted (halo granule)
#if MP3GPU_PROBE
    bool nostore;
#endif
    __device__ __forceinline__ void first(int i, float windowed) const {
        float v = __fadd_rn(windowed, ov[i * 32 + lane]);  // frame.go:474
        if ((i & 1) && (lane & 1)) v = -v;                  // frame.go:480-486
#if MP3GPU_PROBE
        if (nostore && v != 1.2345e-30f) return;
#endif
        hyb[i * 32 + lane] = v;
    }
    __device__ __forceinline__ void second(int i, float windowed) const { ov[i * 32 + lane] = windowed; }  // frame.go:475
};

// 36-point IMDCT + window (imdct.go:99-107).  The window row `bt` is a run-time value (warp-uniform except for
// the mixed-flag quirk of frame.go:462-466), looked up in the constant bank; the cosines are immediates.
#if MP3GPU_EXACT
__device__ __forceinline__ void imdct36_emit(const float (&in)[18], const HybridSink &k, int bt) {
    const float *w = c_win + bt * 36;
    if (k.first_half) {
        float u[9];
        dot36x3<0, 1, 2>(in, u[0], u[1], u[2]);
        dot36x3<3, 4, 5>(in, u[3], u[4], u[5]);
        dot36x3<6, 7, 8>(in, u[6], u[7], u[8]);
#pragma unroll
        for (int p = 0; p < 9; p++) {
            k.first(p, __fmul_rn(u[p], w[p]));
            k.first(17 - p, __fmul_rn(mirror_neg(u[p]), w[17 - p]));  // cos36[m][17-p] == -cos36[m][p] bitwise
        }
    }
    // the second half overwrites the overlap that first() has just consumed
    float v[9];
    dot36x3<18, 19, 20>(in, v[0], v[1], v[2]);
    dot36x3<21, 22, 23>(in, v[3], v[4], v[5]);
    dot36x3<24, 25, 26>(in, v[6], v[7], v[8]);
#pragma unroll
    for (int q = 0; q < 9; q++) {
        k.second(q, __fmul_rn(v[q], w[18 + q]));
        k.second(17 - q, __fmul_rn(v[q], w[35 - q]));      // cos36[m][53-p] == cos36[m][p] bitwise
    }
}
#else
// Fast build.  out[p] = sum_m in[m] cos(pi/72 (2p + 19)(2m + 1)) is an 18-point DCT-IV y[k] = sum_m in[m]
// cos(pi/72 (2k+1)(2m+1)) read out as out[p] = y[p+9] (p < 9), -y[26-p] (9 <= p < 27), -y[p-27] (p >= 27).
// With 2 cos(a) cos(b) = cos(a+b) + cos(a-b):
//   2 cos(pi (2k+1) / 72) y[k] = z[k] = sum_m x'[m] cos(pi m (2k+1) / 36),  x'[m] = in[m] + in[m-1]
//   z[k], z[17-k] = E[k] +- O[k]  with E the 9-point DCT-III of x'[even] and, by the same identity once more,
//   2 cos(pi (2k+1) / 36) O[k] = the 9-point DCT-III of x'[2j+1] + x'[2j-1].
// 176 operations instead of 378.  The factor 1 / (2 cos(pi (2k+1) / 72)), the read-out sign and the window are one
// table, c_winz (mp3gpu.cu).  Results differ from the direct sum by float32 rounding only (PCM +-1 LSB; tests).
__device__ __forceinline__ void dct3_9(const float (&a)[9], float (&D)[9]) {
    // D[k] = sum_j a[j] cos(pi j (2k+1) / 18); cos(pi j (2(8-k)+1) / 18) = (-1)^j cos(pi j (2k+1) / 18)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float e = a[0], o = a[1] * kDct9[1][k];
#pragma unroll
        for (int j = 2; j < 9; j += 2) e = fmaf(a[j], kDct9[j][k], e);
#pragma unroll
        for (int j = 3; j < 9; j += 2) o = fmaf(a[j], kDct9[j][k], o);
        D[k] = e + o;
        D[8 - k] = e - o;
    }
    D[4] = ((a[0] - a[2]) + (a[4] - a[6])) + a[8];
}
__device__ __forceinline__ void imdct36_emit(const float (&in)[18], const HybridSink &k, int bt) {
    const float *w = c_winz + bt * 36;
    float a[9], b[9], E[9], O[9];
    a[0] = in[0];
    b[0] = in[1] + in[0];
#pragma unroll
    for (int j = 1; j < 9; j++) {
        a[j] = in[2 * j] + in[2 * j - 1];                                      // x'[2j]
        b[j] = (in[2 * j + 1] + in[2 * j]) + (in[2 * j - 1] + in[2 * j - 2]);  // x'[2j+1] + x'[2j-1]
    }
    dct3_9(a, E);
    dct3_9(b, O);
    float z[18];
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const float o = O[q] * kSec36[0][q];
        z[q] = E[q] + o;
        z[17 - q] = E[q] - o;
    }
    if (k.first_half) {
#pragma unroll
        for (int p = 0; p < 9; p++) {
            k.first(p, z[9 + p] * w[p]);
            k.first(17 - p, z[9 + p] * w[17 - p]);
        }
    }
    // the second half overwrites the overlap that first() has just consumed
#pragma unroll
    for (int q = 0; q < 9; q++) {
        k.second(q, z[8 - q] * w[18 + q]);
        k.second(17 - q, z[8 - q] * w[35 - q]);
    }
}
#endif

// Short blocks (imdct.go:86-98): three 12-point transforms, windowed and overlapped into out[6..29];
// out[0..5] and out[30..35] stay 0.  raw[j] accumulates in window order i = 0, 1, 2 like the reference.
__device__ __forceinline__ void imdct12_emit(const float (&in)[18], const HybridSink &k) {
    float raw[36];
#pragma unroll
    for (int p = 0; p < 36; p++) raw[p] = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float s[12];
        // cos12[m][5-p] == -cos12[m][p], cos12[m][17-p] == cos12[m][p] bitwise
        const float a0 = i == 0 ? dot12<0, 0>(in) : i == 1 ? dot12<1, 0>(in) : dot12<2, 0>(in);
        const float a1 = i == 0 ? dot12<0, 1>(in) : i == 1 ? dot12<1, 1>(in) : dot12<2, 1>(in);
        const float a2 = i == 0 ? dot12<0, 2>(in) : i == 1 ? dot12<1, 2>(in) : dot12<2, 2>(in);
        const float b0 = i == 0 ? dot12<0, 6>(in) : i == 1 ? dot12<1, 6>(in) : dot12<2, 6>(in);
        const float b1 = i == 0 ? dot12<0, 7>(in) : i == 1 ? dot12<1, 7>(in) : dot12<2, 7>(in);
        const float b2 = i == 0 ? dot12<0, 8>(in) : i == 1 ? dot12<1, 8>(in) : dot12<2, 8>(in);
        s[0] = a0; s[1] = a1; s[2] = a2; s[3] = mirror_neg(a2); s[4] = mirror_neg(a1); s[5] = mirror_neg(a0);
        s[6] = b0; s[7] = b1; s[8] = b2; s[9] = b2; s[10] = b1; s[11] = b0;
#pragma unroll
        for (int p = 0; p < 12; p++) raw[6 * i + p + 6] = mac(s[p], kWin[2][p], raw[6 * i + p + 6]);
    }
    if (k.first_half) {
#pragma unroll
        for (int i = 0; i < 18; i++) k.first(i, raw[i]);
    }
#pragma unroll
    for (int i = 0; i < 18; i++) k.second(i, raw[18 + i]);
}

// IMDCT of one channel of one granule; lane = subband (frame.go:454-478).
__device__ __forceinline__ void hybrid_channel(const float (&in)[18], const HybridSink &k, uint32_t w0, uint32_t w2, int lane) {
    const bool winsw = u_winsw(w0) == 1;
    int bt = winsw ? u_btype(w0) : 0;
    // frame.go:462-466: with win_switch and mixed_block_flag set, subbands 0 and 1 use block type 0 whatever block_type says
    if (winsw && u_mixed(w2) == 1 && lane < 2) bt = 0;
    if (bt == 2) imdct12_emit(in, k);
    else imdct36_emit(in, k, bt);
}

constexpr int kHybWarps = 4;
constexpr int kHybSmemWords = 2 * kXrFloats + 2 * 18 * 32 + 16 + 2 * 64 * 2 + 288;  // per warp: staging, overlap, scalefactors, scale tables, pair-table rows
constexpr int kHybSmemBytes = kHybWarps * kHybSmemWords * 4;

// frame.go:422-425 (6-digit literals); checked against the host tables at mp3gpu_create
__device__ constexpr float kCs[8] = {0.857493f, 0.881742f, 0.949629f, 0.983315f, 0.995518f, 0.999161f, 0.999899f, 0.999993f};
__device__ constexpr float kCa[8] = {-0.514496f, -0.471732f, -0.313377f, -0.181913f, -0.094574f, -0.040966f, -0.014199f, -0.003700f};

// What k_hybrid loads one granule ahead: descriptors, K1's count1/preflag words and — for the long-block fast path,
// where lane = subband — the lane's own 18 lines (9 int16 pairs) of both channels.
struct GranulePre {
    uint32_t w0a, w1a, w2a, w0b, w1b, w2b, meta0, meta1;
    uint32_t sfw;        // lanes 0..15: word (lane & 7) of channel (lane >> 3)'s packed scalefactors
    uint32_t isw[2][9];
};
__device__ __forceinline__ void prefetch_granule(GranulePre &P, const mp3gpu_unit *__restrict__ units, long long first_granule, int g,
                                                 const WaveBufs &B, int lane) {
    const long long G = first_granule + g;
    if (G < 0) {  // in front of the submission: nothing there
        P.w2a = 0;
        return;
    }
    if (!MP3_CHECK(G * 2 + 1 < B.units_total && g >= -2 && g < B.n_gran, G)) {
        P.w2a = 0;
        return;
    }
    const mp3gpu_unit *ug = units + G * 2;
    P.w0a = __ldg(&ug[0].w0); P.w1a = __ldg(&ug[0].w1); P.w2a = __ldg(&ug[0].w2);
    P.w0b = __ldg(&ug[1].w0); P.w1b = __ldg(&ug[1].w1); P.w2b = __ldg(&ug[1].w2);
    P.meta0 = __ldg(B.meta + (long long)g * 2);
    P.meta1 = __ldg(B.meta + (long long)g * 2 + 1);
    P.sfw = __ldg(B.sfpack + (long long)g * 16 + (lane & 15));
    const uint32_t *is2 = reinterpret_cast<const uint32_t *>(B.is16 + (long long)(MP3_PROBE(B, 2) ? (g & 31) : g) * 2 * 576) + lane * 9;
#pragma unroll
    for (int q = 0; q < 9; q++) {
        P.isw[0][q] = __ldg(is2 + q);
        P.isw[1][q] = __ldg(is2 + 288 + q);  // channel 1's slot exists even when the channel does not
    }
}

template <bool TAPS>
__global__ void __launch_bounds__(kHybWarps * 32, 4)
k_hybrid(const mp3gpu_unit *__restrict__ units, long long first_granule, int n_granules, int seg_len, int n_segs,
         DeviceTables T, WaveBufs B, int counter_index) {
    extern __shared__ __align__(16) float s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *const s_base = s_dyn + warp * kHybSmemWords;
    float(*s_x)[kXrFloats] = reinterpret_cast<float(*)[kXrFloats]>(s_base);                 // spectrum staging (short-block path)
    float(*s_ov)[18 * 32] = reinterpret_cast<float(*)[18 * 32]>(s_base + 2 * kXrFloats);   // IMDCT overlap (Frame.store)
    uint32_t(*s_pk)[8] = reinterpret_cast<uint32_t(*)[8]>(s_base + 2 * kXrFloats + 2 * 18 * 32);
    ScaleEnt(*s_scale)[64] = reinterpret_cast<ScaleEnt(*)[64]>(s_base + 2 * kXrFloats + 2 * 18 * 32 + 16);  // 2^(k/4) per band
    uint16_t *s_pd = reinterpret_cast<uint16_t *>(s_base + 2 * kXrFloats + 2 * 18 * 32 + 16 + 2 * 64 * 2);     // pair_dst, pair_long and
    uint8_t *s_pl = reinterpret_cast<uint8_t *>(s_pd + 288);                                                  // pair_short rows of the
    uint8_t *s_ps = s_pl + 288;                                                                               // warp's current cfg

    int sfb_cfg = -1;      // sampling-rate configuration the lane's band codes below belong to
    uint32_t sfb_q[9];     // long-block scalefactor band of the lane's pairs 9*lane .. 9*lane+8 (fast path)
#pragma unroll
    for (int q = 0; q < 9; q++) sfb_q[q] = 0;

    for (;;) {
        int seg = 0;
        if (lane == 0) seg = (int)atomicAdd(B.work_counter + counter_index, 1u);
        seg = __shfl_sync(0xffffffffu, seg, 0);
        if (seg >= n_segs) break;
        const int g0 = seg * seg_len;
        const int g1 = min(g0 + seg_len, n_granules);
        // overlap before the halo granule is irrelevant: the halo's second half replaces it
#pragma unroll 1  // rolled: code size (the kernel competes with itself for the instruction cache on mixed content)
        for (int i = lane; i < 2 * 18 * 32; i += 32) s_ov[0][i] = 0.f;  // s_ov[0] and s_ov[1] are contiguous

        GranulePre P;
        P.w2a = 0;  // "nothing there": the first trip of the loop (g = g0 - 2) only issues the first prefetch
#pragma unroll 1
        for (int g = g0 - 2; g < g1; g++) {
            const GranulePre C = P;                                                      // this granule
            if (g + 1 < g1) prefetch_granule(P, units, first_granule, g + 1, B, lane);   // next one, in flight during this one
            const uint32_t w0a = C.w0a, w1a = C.w1a, w2a = C.w2a, w0b = C.w0b, w1b = C.w1b, w2b = C.w2b;
            if (g < g0 - 1 || first_granule + g < 0 || !u_valid(w2a)) continue;  // a granule always has channel 0
            const bool valid_b = u_valid(w2b);
            const bool need_first = g >= g0;  // the halo granule only contributes its overlap
            if (u_zero(w2a)) {  // start of a stream / of a Seek: Frame.store is zero (frame.go:48)
#pragma unroll 1
                for (int i = lane; i < 2 * 18 * 32; i += 32) s_ov[0][i] = 0.f;
            }
            __syncwarp();

            const int cfg = u_lsf(w2a) * 3 + u_sfreq(w2a);
            if (lane < 16) s_pk[lane >> 3][lane & 7] = C.sfw;
            const GranuleChan c0 = make_chan(w0a, w1a, w2a, C.meta0);
            const GranuleChan c1 = make_chan(w0b, w1b, w2b, valid_b ? C.meta1 : 0u);
            __syncwarp();
            if (cfg != sfb_cfg) {
                // The pair-table rows of this sampling-rate configuration go through shared memory, so that K2 consumes
                // them by LDS: anything fed by a global load inside this loop shares its scoreboard with the granule
                // prefetch issued at the top, and its first use then waits for that prefetch to land (measured on the
                // lane's band codes: 13 % of the kernel's stall samples on one instruction).
                sfb_cfg = cfg;
                __syncwarp();
                for (int i = lane; i < 288; i += 32) {
                    s_pl[i] = T.pair_long[cfg * 288 + i];
                    s_ps[i] = T.pair_short[cfg * 288 + i];
                    s_pd[i] = T.pair_dst[cfg * 288 + i];
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 9; q++) sfb_q[q] = s_pl[lane * 9 + q];
            }
            PairRows rows;
            rows.pair_long = s_pl;
            rows.pair_short = s_ps;
            rows.pair_dst = s_pd;
            const bool fast = !c0.is_short && !(valid_b && c1.is_short);
            float x0[18], x1[18];  // fast path: the lane's subband, both channels
            if (fast) {
                // ---------------- K2, long blocks: lane = subband, everything in registers ---------------
                if (lane < 22) {
                    s_scale[0][lane] = scale_entry(T, c0, s_pk[0], lane);
                    if (valid_b) s_scale[1][lane] = scale_entry(T, c1, s_pk[1], lane);
                }
                __syncwarp();
                {   // requantise (frame.go:140-158); lines at or above count1 stay +0
                    const int np0 = c0.cnt1 >> 1, np1 = c1.cnt1 >> 1;
#pragma unroll
                    for (int q = 0; q < 9; q++) {
                        const int p = lane * 9 + q;
                        // a pair at or above count1 is read as (0, 0): powtab34[0] = 0 and scale * 0 = +0
                        const uint32_t wa = p < np0 ? C.isw[0][q] : 0u;
                        const ScaleEnt sa = s_scale[0][sfb_q[q]];
                        x0[2 * q] = requant_value(T, sa, (int)(int16_t)(wa & 0xffffu));
                        x0[2 * q + 1] = requant_value(T, sa, (int)(int16_t)(wa >> 16));
                        x1[2 * q] = x1[2 * q + 1] = 0.0f;
                        if (valid_b) {
                            const uint32_t wb = p < np1 ? C.isw[1][q] : 0u;
                            const ScaleEnt sb = s_scale[1][sfb_q[q]];
                            x1[2 * q] = requant_value(T, sb, (int)(int16_t)(wb & 0xffffu));
                            x1[2 * q + 1] = requant_value(T, sb, (int)(int16_t)(wb >> 16));
                        }
                    }
                }
                if (valid_b && u_mode(w2a) == 1) {  // stereo (frame.go:362-420)
                    const int mode_ext = u_modeext(w2a);
                    if (mode_ext & 2) {
                        const int max_pos = c0.cnt1 > c1.cnt1 ? c0.cnt1 : c1.cnt1;
                        const float inv_sqrt2 = 0.70710678118654752440f;
#pragma unroll
                        for (int i = 0; i < 18; i++) {
                            if (lane * 18 + i < max_pos) {
                                const float a = x0[i], b = x1[i];
                                x0[i] = f_mul(f_add(a, b), inv_sqrt2);
                                x1[i] = f_mul(f_sub(a, b), inv_sqrt2);
                            }
                        }
                    }
                    if (mode_ext & 1) {
                        // per-band intensity ratios (channel 0's scalefactors); a band that is not intensity coded gets
                        // (1, 1), and x * 1.0f is x exactly, so the lines need no test
                        float2 *s_isr = reinterpret_cast<float2 *>(s_scale[1]);  // channel 1's scale table is no longer needed
                        __syncwarp();
                        if (lane < 22) {
                            const int is_pos = intensity_entry(T, cfg, c0, s_pk[0], c1.cnt1, lane);
                            s_isr[lane] = is_pos < 7 ? make_float2(T.is_ratio_l[is_pos], T.is_ratio_r[is_pos]) : make_float2(1.0f, 1.0f);
                        }
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 9; q++) {
                            const float2 r = s_isr[sfb_q[q]];
                            x0[2 * q] = f_mul(x0[2 * q], r.x); x1[2 * q] = f_mul(x1[2 * q], r.y);
                            x0[2 * q + 1] = f_mul(x0[2 * q + 1], r.x); x1[2 * q + 1] = f_mul(x1[2 * q + 1], r.y);
                        }
                    }
                }
                // alias reduction (frame.go:427-452): butterfly i of the boundary below subband `lane` pairs this lane's
                // line i with line 17-i of lane-1; all 31 x 8 butterflies touch disjoint lines, so order is free.
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    float(&x)[18] = ch ? x1 : x0;
                    if (ch == 1 && !valid_b) break;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float lo_of_below = __shfl_up_sync(0xffffffffu, x[17 - i], 1);   // xr[18 sb - 1 - i] seen from sb
                        const float up_of_above = __shfl_down_sync(0xffffffffu, x[i], 1);      // xr[18 (sb+1) + i] seen from sb
                        const float nu = f_add(f_mul(x[i], kCs[i]), f_mul(lo_of_below, kCa[i]));
                        const float nl = f_sub(f_mul(x[17 - i], kCs[i]), f_mul(up_of_above, kCa[i]));
                        if (lane > 0) x[i] = nu;
                        if (lane < 31) x[17 - i] = nl;
                    }
                }
            } else {
                // ---------------- K2: requantise + reorder (frame.go:140-302) -------------------------------
                // Per-band scale table first, then the lines two at a time (unit_logic.h, "scale-table form").
    #pragma unroll 1
                for (int ch = 0; ch < 2; ch++) {
                    if (ch == 1 && !valid_b) break;
                    const GranuleChan &c = ch ? c1 : c0;
                    s_scale[ch][lane] = scale_entry(T, c, s_pk[ch], lane);
                    s_scale[ch][lane + 32] = scale_entry(T, c, s_pk[ch], lane + 32);
                }
                __syncwarp();
    #pragma unroll 1
                for (int ch = 0; ch < 2; ch++) {
                    if (ch == 1 && !valid_b) break;
                    const GranuleChan &c = ch ? c1 : c0;
                    const uint32_t *is2 = reinterpret_cast<const uint32_t *>(B.is16 + ((long long)g * 2 + ch) * 576);
                    int npair = c.cnt1 >> 1;  // count1 is even: big_values pairs + count1 quadruples
                    if (!MP3_CHECK(g >= -2 && g < B.n_gran && npair <= 288, g)) npair = 0;
                    float *xs = s_x[ch];
    #pragma unroll 3
                    for (int p = lane; p < 288; p += 32) {
                        int d0, d1;
                        const int e = pair_lookup(rows, c, p, &d0, &d1);
                        float x0 = 0.0f, x1 = 0.0f;  // lines at or above count1 stay +0 (maindata/huffman.go:130-134)
                        if (p < npair) {
                            const uint32_t w = __ldg(is2 + p);
                            const ScaleEnt sc = s_scale[ch][e];
                            x0 = requant_value(T, sc, (int)(int16_t)(w & 0xffffu));
                            x1 = requant_value(T, sc, (int)(int16_t)(w >> 16));
                        }
                        xs[xr_pad(d0)] = x0;
                        xs[xr_pad(d1)] = x1;
                    }
                }
                __syncwarp();
                // ---------------- stereo (frame.go:362-420) -------------------------------------------------
                if (valid_b && u_mode(w2a) == 1) {
                    const int mode_ext = u_modeext(w2a);
                    if (mode_ext & 2) {
                        const int max_pos = c0.cnt1 > c1.cnt1 ? c0.cnt1 : c1.cnt1;
                        const float inv_sqrt2 = 0.70710678118654752440f;
                        for (int i = lane; i < max_pos; i += 32) {
                            const int p = xr_pad(i);
                            float a = s_x[0][p], b = s_x[1][p];
                            s_x[0][p] = f_mul(f_add(a, b), inv_sqrt2);
                            s_x[1][p] = f_mul(f_sub(a, b), inv_sqrt2);
                        }
                        __syncwarp();
                    }
                    if (mode_ext & 1) {
                        // per-band intensity positions (channel 0's block type and scalefactors), then per line pair
                        uint8_t *s_isp = reinterpret_cast<uint8_t *>(s_scale[1]);  // channel 1's scale table is no longer needed
                        __syncwarp();
                        s_isp[lane] = (uint8_t)intensity_entry(T, cfg, c0, s_pk[0], c1.cnt1, lane);
                        s_isp[lane + 32] = (uint8_t)intensity_entry(T, cfg, c0, s_pk[0], c1.cnt1, lane + 32);
                        __syncwarp();
                        for (int p = lane; p < 288; p += 32) {
                            int d0, d1;
                            // the window is looked up at the PRE-reorder index although the data is reordered (frame.go:341-357)
                            const int is_pos = s_isp[pair_lookup(rows, c0, p, &d0, &d1)];
                            if (is_pos < 7) {
                                const float rl = T.is_ratio_l[is_pos], rr = T.is_ratio_r[is_pos];
                                const int q0 = xr_pad(2 * p), q1 = xr_pad(2 * p + 1);
                                s_x[0][q0] = f_mul(s_x[0][q0], rl);
                                s_x[1][q0] = f_mul(s_x[1][q0], rr);
                                s_x[0][q1] = f_mul(s_x[0][q1], rl);
                                s_x[1][q1] = f_mul(s_x[1][q1], rr);
                            }
                        }
                        __syncwarp();
                    }
                }
                // ---------------- alias reduction (frame.go:427-452) ----------------------------------------
    #pragma unroll 1
                for (int ch = 0; ch < 2; ch++) {
                    if (ch == 1 && !valid_b) break;
                    const int nb = alias_butterflies(ch ? c1 : c0);
                    for (int b = lane; b < nb; b += 32) {
                        const int sb = (b >> 3) + 1, i = b & 7;
                        const int li = 18 * sb - 1 - i + (sb - 1), ui = 18 * sb + i + sb;  // padded positions
                        const float xl = s_x[ch][li], xu = s_x[ch][ui];
                        s_x[ch][li] = f_sub(f_mul(xl, kCs[i]), f_mul(xu, kCa[i]));
                        s_x[ch][ui] = f_add(f_mul(xu, kCs[i]), f_mul(xl, kCa[i]));
                    }
                }
                __syncwarp();

            }

            // ---------------- K3: IMDCT, lane = subband (frame.go:454-486) ------------------------------
#pragma unroll 1
            for (int ch = 0; ch < 2; ch++) {
                if (ch == 1 && !valid_b) break;
                float in[18];
                if (fast) {
#pragma unroll
                    for (int m = 0; m < 18; m++) in[m] = ch ? x1[m] : x0[m];
                } else {
#pragma unroll
                    for (int m = 0; m < 18; m++) in[m] = s_x[ch][lane * kXrStride + m];
                }
                if (!MP3_CHECK(!need_first || (g >= 0 && g < B.n_gran), g)) continue;  // hyb / tap stores below
                if (TAPS && need_first && B.tap_xr) {
                    float *o = B.tap_xr + ((long long)g * 2 + ch) * 576 + lane * 18;
#pragma unroll
                    for (int m = 0; m < 18; m++) o[m] = in[m];
                }
#if MP3GPU_PROBE
                HybridSink sink{B.hyb + ((long long)g * 2 + ch) * 576, s_ov[ch], lane, need_first, MP3_PROBE(B, 1) != 0};
#else
                HybridSink sink{B.hyb + ((long long)g * 2 + ch) * 576, s_ov[ch], lane, need_first};
#endif
                hybrid_channel(in, sink, ch ? w0b : w0a, ch ? w2b : w2a, lane);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_synth = K4: polyphase synthesis + int16 clamp/interleave (frame.go:630-688).
//
// Every WARP owns a segment of kSynSegSlots consecutive time slots of the wave (slot = granule*18 + t) and walks it
// in blocks of 30 slots, alternating two phases on a private shared-memory buffer; warps never synchronise with each
// other, so one warp's loads overlap the other warps' arithmetic.
//   Phase A (matrixing, frame.go:644-650): lane = slot of the block, both channels.  Exact build: the 32 subband
//     samples of the slot sit in registers and the cosine matrix is compile-time, so every multiply-add is an FFMA
//     with an immediate coefficient; only the 33 rows that are unique are computed: synthNWin[32-i][j] ==
//     -synthNWin[i][j] (i = 1..16) and synthNWin[48+m][j] == synthNWin[48-m][j] (m = 1..15) hold BITWISE in the
//     float32 table, and a sum of negated terms in the same order is the exact negation.  Fast build: a 32-point
//     Lee DCT (below).  Row u of the slot goes to shared memory: U[0..16] = V[0..16], U[17..32] = V[33..48].
//   Phase B (window, frame.go:651-678): lane = output index i, marching in time with the 15-slot V history in
//     registers (circular, period 15) and D[32d + i] in registers; taps accumulate in the reference's order
//     (d ascending).  Both channels are done together so the int16 pair is packed and stored as one coalesced word.
// The history carries over from block to block; a segment starts by replaying the 15 slots in front of it
// (phase A + history fill only), which recreates exactly the state a linear decode has there.
// ------------------------------------------------------------------------------------------
constexpr int kSynWarps = 4;
constexpr int kSynThreads = kSynWarps * 32;
constexpr int kSynBlock = 30;                    // slots per block: two periods of the circular history
constexpr int kSynSegBlocks = 24;                // default segment: 24 blocks = 720 slots = 40 granules
constexpr int kURow = 36;                        // floats per U row: 16-byte aligned, conflict-free 128-bit stores
constexpr int kSynWarpWords = 4 * kSynBlock * kURow + 8;  // U[2][30][36], staging S[2][30][36], flags[30] (8 words)
constexpr int kSynSmemBytes = kSynWarps * kSynWarpWords * 4;

__device__ __forceinline__ int pcm_from_float(float sum) {
    // frame.go:663-668: int(sum * 32767) truncates toward zero, then clamps to +-32767.  The clamp is done in float
    // first (NaN -> -32767 like Go's INT64_MIN from CVTTSS2SQ; fmaxf returns the non-NaN operand).  The only inputs
    // on which this differs from the reference are f >= 2^63 or +Inf (Go: INT64_MIN -> -32767, here +32767); they
    // are unreachable: |xr| <= 8206^(4/3) * 2^(45/4) < 4.1e8, and 18 IMDCT terms x 32 x 16 synthesis terms x 32767
    // bound |f| by 1.3e17 < 2^63.
    float f = __fmul_rn(sum, 32767.0f);
    f = fminf(fmaxf(f, -32767.0f), 32767.0f);
    return __float2int_rz(f);
}

// NR consecutive rows of V for one slot: V[i] = sum_j N[i][j] * s[j], j ascending (frame.go:644-650)
template <int R0, int NR>
__device__ __forceinline__ void matrix_rows(const float (&s)[32], float *urow, int u0) {
    float acc[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) acc[r] = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j++) {
#pragma unroll
        for (int r = 0; r < NR; r++) acc[r] = mac(kSynthN[R0 + r][j], s[j], acc[r]);
    }
    if (NR == 4) {
        *reinterpret_cast<float4 *>(urow + u0) = make_float4(acc[0], acc[1], acc[2], acc[3]);  // u0 % 4 == 0
    } else {
#pragma unroll
        for (int r = 0; r < NR; r++) urow[u0 + r] = acc[r];
    }
}

#if !MP3GPU_EXACT
// Fast build: the 64 x 32 matrixing is a 32-point DCT-II, V[i] = +-c[n] with c[n] = sum_j s[j] cos(n (2j+1) pi / 64),
// computed by Lee's recursion (80 multiplies + 209 adds instead of 1,056 multiply-adds): split into the symmetric and
// the antisymmetric half, the latter scaled by 1 / (2 cos), two half-size DCTs, odd outputs W[k] + W[k+1].  The result
// differs from the reference's direct summation by float32 rounding only; the PCM stays within +-1 LSB with the same
// exact-match fraction as fused multiply-add alone gives (tests/test_gpu_*.py state both).  The exact build keeps the
// direct form below and is bit-identical to the reference.
template <int N, int LVL>
__device__ __forceinline__ void lee_dct(const float (&x)[N], float (&X)[N]) {
    if constexpr (N == 1) {
        X[0] = x[0];
    } else {
        constexpr int H = N / 2;
        float u[H], v[H], E[H], W[H];
#pragma unroll
        for (int n = 0; n < H; n++) {
            u[n] = x[n] + x[N - 1 - n];
            v[n] = (x[n] - x[N - 1 - n]) * kLeeSec[LVL][n];
        }
        lee_dct<H, LVL + 1>(u, E);
        lee_dct<H, LVL + 1>(v, W);
#pragma unroll
        for (int k = 0; k < H; k++) {
            X[2 * k] = E[k];
            X[2 * k + 1] = k < H - 1 ? W[k] + W[k + 1] : W[H - 1];
        }
    }
}
#endif

__device__ __forceinline__ void load_slot(const float *__restrict__ src, float (&s)[32]) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + q);
        s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
    }
}

__device__ __forceinline__ void matrix_slot(const float (&s)[32], float *urow) {
#if MP3GPU_EXACT
    // rows 0..16 -> U[0..16]; rows 33..48 -> U[17..32]
    matrix_rows<0, 4>(s, urow, 0);   matrix_rows<4, 4>(s, urow, 4);   matrix_rows<8, 4>(s, urow, 8);
    matrix_rows<12, 4>(s, urow, 12); matrix_rows<16, 1>(s, urow, 16);
    matrix_rows<33, 3>(s, urow, 17); matrix_rows<36, 4>(s, urow, 20); matrix_rows<40, 4>(s, urow, 24);
    matrix_rows<44, 4>(s, urow, 28); matrix_rows<48, 1>(s, urow, 32);
#else
    // V[0..15] = c[16..31], V[16] = c[32] = 0 (the reference's row 16 is float32(cos(odd * pi/2)) ~ 1e-17 times the
    // samples: below half an ulp of any sum it joins), V[33 + m] = -c[15 - m], V[48] = -c[0].  Stored without the
    // signs: U[17..32] = c[15..0]; phase B folds the minus into its window coefficients.
    float c[32];
    lee_dct<32, 0>(s, c);
    float4 *u4 = reinterpret_cast<float4 *>(urow);
    u4[0] = make_float4(c[16], c[17], c[18], c[19]);
    u4[1] = make_float4(c[20], c[21], c[22], c[23]);
    u4[2] = make_float4(c[24], c[25], c[26], c[27]);
    u4[3] = make_float4(c[28], c[29], c[30], c[31]);
    u4[4] = make_float4(0.0f, c[15], c[14], c[13]);
    u4[5] = make_float4(c[12], c[11], c[10], c[9]);
    u4[6] = make_float4(c[8], c[7], c[6], c[5]);
    u4[7] = make_float4(c[4], c[3], c[2], c[1]);
    urow[32] = c[0];
#endif
}

// V history of one lane (phase B) and its window coefficients.
//   out[i] = sum_d V_{t-d}[(d odd ? 32 : 0) + i] * D[32 d + i], d = 0..15 (U construction, frame.go:651-661).
// Slot t-d lives at circular position (P - d) mod 15: when slot t arrives at position P, its V[i] part is stored
// before the sum and its V[32+i] part after it, because that place still holds slot t-15's, the d = 15 tap.
struct SynHist {
    float A[2][15], B[2][15];  // [channel][position]: V[i] and V[32+i] of the slot at that position
    float dw[16];
    __device__ __forceinline__ void set_coef(int d, float v) { dw[d] = v; }
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 15; i++) { A[0][i] = B[0][i] = A[1][i] = B[1][i] = 0.f; }
    }
    template <int P, int CH> __device__ __forceinline__ void put_a(float a) { A[CH][P] = a; }
    template <int P, int CH> __device__ __forceinline__ void put_b(float b) { B[CH][P] = b; }
    template <int P, int CH> __device__ __forceinline__ float sum() const {  // d ascending, one chain: the reference's order
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < 16; d++) {
            const int h = (P - d + 30) % 15;
            acc = mac((d & 1) ? B[CH][h] : A[CH][h], dw[d], acc);
        }
        return acc;
    }
};
// (Measured and dropped: pairing taps (d, d+1) into packed FFMA2, and two scalar chains per channel.  Both leave the
// kernel time unchanged within 1 %: with three warps per scheduler the window phase waits on the FMA pipe and on
// fixed-latency dependencies, not on issue slots.)

struct SynLane {  // where lane i finds V[i] and V[32+i] in a U row
    int ai, bi;
#if MP3GPU_PROBE
    int nostore;
#endif
};

// One slot of phase B at circular position P.  FAST: the caller has checked that this and the other 14 rows of the
// block are plain stereo slots (both channels present, no state reset), so there are no flag tests at all.
template <int P, bool FAST, bool WARMUP>
__device__ __forceinline__ void synth_window_slot(const float *U0, const float *U1, const uint8_t *flags, int row, const SynLane &L,
                                                  SynHist &h, uint32_t *pcm32, long long sigma, long long n_slots, int lane) {
    const int f = FAST ? 3 : flags[row];
    if (!FAST && (f & 4)) h.clear();  // first slot of a ZERO_STATE granule: Frame.vVec is zero (frame.go:49)
    uint32_t pl = 0, pr = 0;
    if (f & 1) {
        const float a = U0[row * kURow + L.ai];
        const float b = U0[row * kURow + L.bi];
        h.template put_a<P, 0>(a);
        if (!WARMUP) pl = (uint32_t)pcm_from_float(h.template sum<P, 0>()) & 0xffffu;
        h.template put_b<P, 0>(b);
    }
    if (f & 2) {
        const float a = U1[row * kURow + L.ai];
        const float b = U1[row * kURow + L.bi];
        h.template put_a<P, 1>(a);
        if (!WARMUP) pr = (uint32_t)pcm_from_float(h.template sum<P, 1>()) & 0xffffu;
        h.template put_b<P, 1>(b);
    } else {
        pr = pl;  // mono: both output channels carry channel 0 (frame.go:671-678)
    }
#if MP3GPU_PROBE
    if (L.nostore && (pl | (pr << 16)) != 0x12345678u) return;
#endif
    if (!WARMUP && (f & 1) && (FAST || sigma < n_slots) && MP3_CHECK(sigma >= 0 && sigma < n_slots, sigma)) pcm32[sigma * 32 + lane] = pl | (pr << 16);
}

// 15 consecutive rows starting at `rb` (circular positions 0..14).
template <bool FAST, bool WARMUP>
__device__ __forceinline__ void synth_window_block(const float *U0, const float *U1, const uint8_t *flags, int rb, const SynLane &L,
                                                   SynHist &h, uint32_t *pcm32, long long sb, long long n_slots, int lane) {
#define MP3_SLOT(Pp) synth_window_slot<Pp, FAST, WARMUP>(U0, U1, flags, rb + Pp, L, h, pcm32, sb + Pp, n_slots, lane);
    MP3_SLOT(0) MP3_SLOT(1) MP3_SLOT(2) MP3_SLOT(3) MP3_SLOT(4) MP3_SLOT(5) MP3_SLOT(6) MP3_SLOT(7)
    MP3_SLOT(8) MP3_SLOT(9) MP3_SLOT(10) MP3_SLOT(11) MP3_SLOT(12) MP3_SLOT(13) MP3_SLOT(14)
#undef MP3_SLOT
}

// Slot flags of wave-local slot sigma: bit 0 / 1 = channel 0 / 1 present, bit 2 = first slot of a ZERO_STATE granule;
// 0 outside the submission.  Split in two so that the descriptor loads can be issued a block ahead of their use:
// synth_flag_words() only loads (bit 31 of the result: slot inside the submission), synth_flags_of() decodes.
__device__ __forceinline__ uint2 synth_flag_words(const mp3gpu_unit *__restrict__ units, long long units_total, long long first_granule, int n_slots, int sigma) {
    if (sigma < -18 || sigma >= n_slots || first_granule * 18 + sigma < 0) return make_uint2(0u, 0u);
    const int g = (sigma + 18) / 18 - 1;
    const mp3gpu_unit *ug = units + (first_granule + g) * 2;
    if (!MP3_CHECK(first_granule + g >= 0 && (first_granule + g) * 2 + 1 < units_total, first_granule + g)) return make_uint2(0u, 0u);
    return make_uint2(__ldg(&ug[0].w2), __ldg(&ug[1].w2));
}
__device__ __forceinline__ int synth_flags_of(uint2 w, int sigma) {
    if (!u_valid(w.x)) return 0;
    const int g = (sigma + 18) / 18 - 1;
    return 1 | (u_valid(w.y) ? 2 : 0) | ((u_zero(w.x) && sigma == g * 18) ? 4 : 0);
}
// Subband samples of (slot sigma, channel ch) in hyb (one look-back granule in front: sigma >= -18).
// Row ((g*2 + ch)*18 + t) with g = sigma / 18, t = sigma % 18 is row 2*sigma - t + 18*ch.
__device__ __forceinline__ const float *synth_slot_src(const WaveBufs &B, int sigma, int ch) {
    const int t = (sigma + 18) % 18;
    if (MP3_PROBE(B, 8)) return B.hyb + (long long)(2 * (sigma & 511) - t + 18 * ch) * 32;
    return B.hyb + (long long)(2 * sigma - t + 18 * ch) * 32;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
// Starts the copy of the block's 2 x 30 input rows (128 B each) into the staging rows S[ch][row][0..31]: 480 chunks
// of 16 bytes, consecutive lanes on consecutive chunks.  Rows outside the submission are skipped (phase A does not
// read them); rows of an absent channel are copied as they are and ignored.
__device__ __forceinline__ void synth_stage_block(const WaveBufs &B, long long first_granule, int n_slots, float *S, int sigma0, int lane) {
#pragma unroll
    for (int it = 0; it < 15; it++) {
        const int c = it * 32 + lane;
        const int ch = c >= 8 * kSynBlock ? 1 : 0;
        const int rem = c - ch * 8 * kSynBlock;
        const int row = rem >> 3, q = rem & 7;
        const int sigma = sigma0 + row;
        if (sigma < n_slots && first_granule * 18 + sigma >= 0 && MP3_CHECK(sigma >= -18, sigma))
            cp_async16(S + (ch * kSynBlock + row) * kURow + q * 4, synth_slot_src(B, sigma, ch) + q * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kSynThreads, 3)
k_synth(const mp3gpu_unit *__restrict__ units, long long first_granule, int n_granules, WaveBufs B,
        int16_t *__restrict__ pcm /* wave-local: [n_granules][576][2] */, int seg_blocks /* blocks of 30 slots per warp segment */) {
    extern __shared__ __align__(16) float s_u[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *U0 = s_u + warp * kSynWarpWords, *U1 = U0 + kSynBlock * kURow;
    float *S = U0 + 2 * kSynBlock * kURow;  // staging rows of the next block, same shape as U
    uint8_t *flags = reinterpret_cast<uint8_t *>(S + 2 * kSynBlock * kURow);
    const int n_slots = n_granules * 18;  // < 2^31: a wave is at most a few million granules
    const long long seg_ll = ((long long)blockIdx.x * kSynWarps + warp) * (long long)(seg_blocks * kSynBlock);
    if (seg_ll >= n_slots) return;
    const int seg_first = (int)seg_ll;  // wave-local slot

    synth_stage_block(B, first_granule, n_slots, S, seg_first, lane);  // block 0, in flight during the warm-up

    // V[i] = i <= 16 ? U[i] : -U[32-i];  V[32+i] = i == 0 ? -U[0] : i <= 16 ? U[16+i] : U[48-i].  A lane's sign is the
    // same for all its even taps (V[i]) and for all its odd taps (V[32+i]), so it is folded into the window
    // coefficients: fma(-u, D, acc) == fma(u, -D, acc) exactly, and a zero's sign never reaches a sum that starts at +0.
    SynLane L;
    L.ai = lane <= 16 ? lane : 32 - lane;
    L.bi = lane == 0 ? 0 : (lane <= 16 ? 16 + lane : 48 - lane);
#if MP3GPU_PROBE
    L.nostore = MP3_PROBE(B, 4);
#endif
#if MP3GPU_EXACT
    const float sa = lane <= 16 ? 1.0f : -1.0f, sb = lane == 0 ? -1.0f : 1.0f;
#else
    const float sa = lane <= 16 ? 1.0f : -1.0f, sb = -1.0f;  // U[17..32] hold +c[15..0] (matrix_slot): V[32+i] = -U[bi] for every lane
#endif
    SynHist h;
#pragma unroll
    for (int d = 0; d < 16; d++) h.set_coef(d, __ldg(B.synth_d + 32 * d + lane) * ((d & 1) ? sb : sa));
    h.clear();
    uint32_t *pcm32 = reinterpret_cast<uint32_t *>(pcm);

    // warm-up: the 15 slots in front of the segment (slot seg_first - 15 + p sits at circular position p), straight
    // from global memory
    {
        const int f = lane < 15 ? synth_flags_of(synth_flag_words(units, B.units_total, first_granule, n_slots, seg_first - 15 + lane), seg_first - 15 + lane) : 0;
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++)
            if (f & (1 << ch)) {
                float s[32];
                load_slot(synth_slot_src(B, seg_first - 15 + lane, ch), s);
                matrix_slot(s, (ch ? U1 : U0) + lane * kURow);
            }
        if (lane < kSynBlock) flags[lane] = (uint8_t)f;
        __syncwarp();
        const bool plain = __all_sync(0xffffffffu, lane >= 15 || flags[lane] == 3);
        if (plain) synth_window_block<true, true>(U0, U1, flags, 0, L, h, pcm32, seg_first - 15, n_slots, lane);
        else synth_window_block<false, true>(U0, U1, flags, 0, L, h, pcm32, seg_first - 15, n_slots, lane);
        __syncwarp();
    }
    uint2 fw_next = synth_flag_words(units, B.units_total, first_granule, n_slots, lane < kSynBlock ? seg_first + lane : n_slots);
#pragma unroll 1
    for (int blk = 0; blk < seg_blocks; blk++) {
        const int sigma0 = seg_first + blk * kSynBlock;
        if (sigma0 >= n_slots) break;
        const int f = synth_flags_of(fw_next, sigma0 + lane);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- phase A: lane = slot sigma0 + lane, rows from the staging buffer ----
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++)
            if (f & (1 << ch)) {
                float s[32];
                const float4 *src = reinterpret_cast<const float4 *>(S + (ch * kSynBlock + lane) * kURow);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float4 v = src[q];
                    s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
                }
                matrix_slot(s, (ch ? U1 : U0) + lane * kURow);
            }
        if (lane < kSynBlock) flags[lane] = (uint8_t)f;
        __syncwarp();
        // ---- next block: start its copy and its flag loads, both consumed after phase B ----
        const bool more = blk + 1 < seg_blocks && sigma0 + kSynBlock < n_slots;
        if (more) {
            synth_stage_block(B, first_granule, n_slots, S, sigma0 + kSynBlock, lane);
            fw_next = synth_flag_words(units, B.units_total, first_granule, n_slots, lane < kSynBlock ? sigma0 + kSynBlock + lane : n_slots);
        }
        // ---- phase B ----
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            const int rb = half * 15;
            const bool plain = __all_sync(0xffffffffu, lane >= 15 || flags[rb + lane] == 3);
            if (plain) synth_window_block<true, false>(U0, U1, flags, rb, L, h, pcm32, sigma0 + rb, n_slots, lane);
            else synth_window_block<false, false>(U0, U1, flags, rb, L, h, pcm32, sigma0 + rb, n_slots, lane);
        }
        __syncwarp();
    }
}

// Look-back carry between sub-waves: a plain copy as a KERNEL.  (A cudaMemcpyAsync between two kernels goes through a
// copy engine: two engine hand-overs of ~15 us each per sub-wave, measured; a kernel stays on the compute queue.)
__global__ void __launch_bounds__(256) k_carry(float4 *__restrict__ dst, const float4 *__restrict__ src, int n4) {
    for (int i = threadIdx.x; i < n4; i += 256) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// Output side: PCM for a consumer on the same GPU.  The reference's only consumer takes s16le
// stereo (example/main.go:40-52) and so does Decoder.Read; a GPU audio pipeline (resampler, feature extractor, model
// front end) takes float planes.  k_pcm_to_f32_planar converts device-resident interleaved int16 to two float32 planes
// scaled by 1/32768 without the PCM ever crossing PCIe (the ceiling of the end-to-end path).  Pure streaming: 4 bytes
// in and 8 bytes out per stereo sample, 16-byte loads and stores, grid-stride over a few CTAs per SM.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pcm_to_f32_planar(const uint4 *__restrict__ pcm4 /* 4 stereo samples per element */, size_t n4, float4 *__restrict__ left4,
                    float4 *__restrict__ right4, const int16_t *__restrict__ pcm, size_t n, float *__restrict__ left, float *__restrict__ right) {
    const float k = 1.0f / 32768.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(pcm4 + i);
        float4 l, r;
        l.x = (float)(int16_t)(v.x & 0xffffu) * k; r.x = (float)(int16_t)(v.x >> 16) * k;
        l.y = (float)(int16_t)(v.y & 0xffffu) * k; r.y = (float)(int16_t)(v.y >> 16) * k;
        l.z = (float)(int16_t)(v.z & 0xffffu) * k; r.z = (float)(int16_t)(v.z >> 16) * k;
        l.w = (float)(int16_t)(v.w & 0xffffu) * k; r.w = (float)(int16_t)(v.w >> 16) * k;
        left4[i] = l;
        right4[i] = r;
    }
    // the last n % 4 samples
    const size_t tail0 = n4 * 4;
    if (blockIdx.x == 0 && threadIdx.x < n - tail0) {
        const size_t j = tail0 + threadIdx.x;
        left[j] = (float)pcm[2 * j] * k;
        right[j] = (float)pcm[2 * j + 1] * k;
    }
}

// ------------------------------------------------------------------------------------------
// FP32 FMA peak micro-benchmark (roofline denominator; MEASURED_PEAKS.json has no fp32 entry)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c); a3 = __fmaf_rn(a3, b, c);
            a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c); a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace mp3gpu
