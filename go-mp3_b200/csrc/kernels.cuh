// kernels.cuh — sm_100a device kernels of the MP3 Layer III granule decode path.
//
//   K1 k_huffman : scalefactors + Huffman         maindata.go:119-288, maindata/huffman.go:27-138,
//                                                 huffman.go:348-419, bits.go:45-86
//   K2 k_requant : requantise, reorder, stereo,   frame.go:140-452
//                  alias reduction
//   K3 k_imdct   : IMDCT + window + overlap-add   frame.go:454-486, imdct.go:83-108
//                  + frequency inversion
//   K4 k_synth   : polyphase synthesis + int16    frame.go:630-688
//
// Design notes (see DESIGN.md for the full derivation):
//  * K1 is one THREAD per granule-channel: the code stream of a unit is serial, so the
//    parallelism is across units.  Code tables are multi-level LUTs staged in shared memory;
//    the bit cursor reproduces bits.go's out-of-bounds rule (reads at/after the frame's logical
//    buffer end return 0 and do not advance).
//  * K3/K4 keep the reference's direct-form summation ORDER (m ascending / j ascending / tap
//    ascending) and take their cosine/window coefficients as constant-bank operands of FFMA, so
//    the inner loops are pure FFMA streams with no shared-memory or register traffic for the
//    coefficient matrices.
//  * Cross-granule state (IMDCT overlap `store`, synthesis `vVec`) is not carried serially:
//    K3 recomputes the previous granule's second IMDCT half at the start of each run of
//    granules; K4 recomputes the matrixing of the 15 preceding time slots (halo) per CTA.
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
// Constant-bank tables (uniform access only; coefficients become FFMA c[bank][imm] operands)
// ------------------------------------------------------------------------------------------
__constant__ float c_cos36[18 * 36];
__constant__ float c_cos12[6 * 12];
__constant__ float c_win[4 * 36];
__constant__ float c_synth_n[64 * 32];
__constant__ float c_synth_d[512];

// Wave-local intermediate buffers.  Index j = local granule (g - wave_first); arrays that a
// later kernel reads with a one-granule look-back (xr_t, hyb) have a valid slot at j = -1.
struct WaveBufs {
    int16_t *is16;     // [nw][2][576]
    uint32_t *meta;    // [nw][2]   bits 0..9 count1, bit 10 preflag (after LSF derivation)
    uint32_t *sfpack;  // [nw][2][8] scalefactors as nibbles: n = sfb (long), 22 + sfb*3+win (short)
    float *xr_t;       // [-1..nw)[2][18][32]  spectral lines, transposed: [m][sb] = xr[sb*18+m]
    float *hyb[2];     // per channel: [-18..nw*18)[32]  subband samples, slot-major
};

// ------------------------------------------------------------------------------------------
// K1: scalefactors + Huffman.  One thread per unit slot; per-unit logic in unit_logic.h.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_huffman(const uint8_t *__restrict__ main_data, const mp3gpu_unit *__restrict__ units, long long first_unit,
          int n_units, DeviceTables T, WaveBufs B) {
    extern __shared__ uint16_t s_lut[];
    __shared__ uint32_t s_desc[34];
    for (int i = threadIdx.x; i < T.huff_lut_n; i += blockDim.x) s_lut[i] = T.huff_lut[i];
    if (threadIdx.x < 34) s_desc[threadIdx.x] = T.huff_desc[threadIdx.x];
    __syncthreads();
    int ul = blockIdx.x * blockDim.x + threadIdx.x;  // wave-local unit index
    if (ul >= n_units) return;
    if (!u_valid(units[first_unit + ul].w2)) {
        B.meta[ul] = 0;
        return;
    }
    uint32_t pk[8];
    uint32_t *out = reinterpret_cast<uint32_t *>(B.is16 + (size_t)ul * 576);
    uint32_t meta = huffman_unit(T, s_lut, s_desc, main_data, units, first_unit + ul, pk, out);
    uint4 *dst = reinterpret_cast<uint4 *>(B.sfpack + (size_t)ul * 8);
    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    B.meta[ul] = meta;
}

// ------------------------------------------------------------------------------------------
// K2: requantise + reorder + stereo + alias reduction.  One warp per granule (both channels).
// ------------------------------------------------------------------------------------------
constexpr int kK2Warps = 8;

__global__ void __launch_bounds__(kK2Warps * 32)
k_requant(const mp3gpu_unit *__restrict__ units, long long first_granule, int n_granules, DeviceTables T, WaveBufs B) {
    __shared__ float s_x[kK2Warps][2][576];
    __shared__ uint32_t s_pk[kK2Warps][2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gl = blockIdx.x * kK2Warps + warp;
    if (gl >= n_granules) return;
    const mp3gpu_unit *ug = units + (first_granule + gl) * 2;
    const bool valid_b = u_valid(ug[1].w2);
    if (!u_valid(ug[0].w2)) return;  // a granule always has channel 0
    const int cfg = u_lsf(ug[0].w2) * 3 + u_sfreq(ug[0].w2);
    float(*x)[576] = s_x[warp];
    if (lane < 16) s_pk[warp][lane >> 3][lane & 7] = B.sfpack[((size_t)gl * 2 + (lane >> 3)) * 8 + (lane & 7)];
    GranuleChan c[2];
    c[0] = make_chan(ug[0].w0, ug[0].w1, ug[0].w2, B.meta[gl * 2]);
    c[1] = make_chan(ug[1].w0, ug[1].w1, ug[1].w2, valid_b ? B.meta[gl * 2 + 1] : 0u);
    __syncwarp();

    // ---- requantise (frame.go:140-255) + reorder (frame.go:257-302) ----------------------
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
        if (ch == 1 && !valid_b) break;
        const int16_t *is = B.is16 + ((size_t)gl * 2 + ch) * 576;
        for (int i = lane; i < 576; i += 32) {
            int dst;
            float r = requant_line(T, cfg, c[ch], s_pk[warp][ch], i, i < c[ch].cnt1 ? (int)is[i] : 0, &dst);
            x[ch][dst] = r;
        }
    }
    __syncwarp();

    // ---- stereo (frame.go:362-420) --------------------------------------------------------
    if (valid_b && u_mode(c[0].w2) == 1) {
        const int mode_ext = u_modeext(c[0].w2);
        if (mode_ext & 2) {
            const int max_pos = c[0].cnt1 > c[1].cnt1 ? c[0].cnt1 : c[1].cnt1;
            const float inv_sqrt2 = 0.70710678118654752440f;
            for (int i = lane; i < max_pos; i += 32) {
                float a = x[0][i], b = x[1][i];
                x[0][i] = f_mul(f_add(a, b), inv_sqrt2);
                x[1][i] = f_mul(f_sub(a, b), inv_sqrt2);
            }
            __syncwarp();
        }
        if (mode_ext & 1) {
            for (int i = lane; i < 576; i += 32) {
                int is_pos = intensity_pos(T, cfg, c[0], s_pk[warp][0], c[1].cnt1, i);
                if (is_pos < 7) {
                    x[0][i] = f_mul(x[0][i], T.is_ratio_l[is_pos]);
                    x[1][i] = f_mul(x[1][i], T.is_ratio_r[is_pos]);
                }
            }
            __syncwarp();
        }
    }

    // ---- alias reduction (frame.go:427-452) + transposed store ---------------------------
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
        if (ch == 1 && !valid_b) break;
        const int nb = alias_butterflies(c[ch]);
        for (int b = lane; b < nb; b += 32) alias_butterfly(T.cs, T.ca, x[ch], b);
        __syncwarp();
        float *o = B.xr_t + ((size_t)gl * 2 + ch) * 576;
#pragma unroll
        for (int m = 0; m < 18; m++) o[m * 32 + lane] = x[ch][lane * 18 + m];
    }
}

// ------------------------------------------------------------------------------------------
// K3: IMDCT + window + overlap-add + frequency inversion.
// One thread per subband, one warp per (run of kRun granules, channel).
// ------------------------------------------------------------------------------------------
constexpr int kRun = 8;
constexpr int kK3Warps = 4;

// out[p] = sum_{m} in[m] * cos36[m][p], m ascending from 0 (imdct.go:101-107).
template <int P0, int P1>
__device__ __forceinline__ void imdct36_range(const float (&in)[18], float (&raw)[36]) {
#pragma unroll
    for (int p = P0; p < P1; p++) {
        float sum = 0.0f;
#pragma unroll
        for (int m = 0; m < 18; m++) sum = mac(in[m], c_cos36[m * 36 + p], sum);
        raw[p] = sum;
    }
}

// Short blocks (imdct.go:86-98): three 12-point transforms, windowed and overlapped into out[6..29].
__device__ __forceinline__ void imdct12_win(const float (&in)[18], float (&raw)[36]) {
#pragma unroll
    for (int p = 0; p < 36; p++) raw[p] = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int p = 0; p < 12; p++) {
            float sum = 0.0f;
#pragma unroll
            for (int m = 0; m < 6; m++) sum = mac(in[i + 3 * m], c_cos12[m * 12 + p], sum);
            raw[6 * i + p + 6] = mac(sum, c_win[2 * 36 + p], raw[6 * i + p + 6]);
        }
    }
}

template <int BT, int P0, int P1>
__device__ __forceinline__ void win36(float (&raw)[36]) {
#pragma unroll
    for (int p = P0; p < P1; p++) raw[p] = __fmul_rn(raw[p], c_win[BT * 36 + p]);
}

// Full windowed IMDCT of one subband for a granule with side info (w0, w1); lane = subband.
// HALF = 0: all 36 outputs; HALF = 1: only raw[18..35] is needed (overlap halo).
template <int HALF>
__device__ __forceinline__ void imdct_granule(const float (&in)[18], float (&raw)[36], uint32_t w0, uint32_t w2, int lane) {
    const int bt = u_btype(w0);
    const bool winsw = u_winsw(w0) == 1, mixed = u_mixed(w2) == 1;
    constexpr int P0 = HALF ? 18 : 0;
    if (!(winsw && mixed)) {
        // uniform across the warp
        if (bt == 2) {
            imdct12_win(in, raw);
        } else {
            imdct36_range<P0, 36>(in, raw);
            if (bt == 0) win36<0, P0, 36>(raw);
            else if (bt == 1) win36<1, P0, 36>(raw);
            else win36<3, P0, 36>(raw);
        }
    } else {
        // frame.go:462-466: subbands 0,1 use block type 0 whatever block_type says.
        const int ebt = lane < 2 ? 0 : bt;
        if (__any_sync(0xffffffffu, ebt != 2)) {
            float r2[36];
            imdct36_range<P0, 36>(in, r2);
            if (ebt != 2) {
#pragma unroll
                for (int p = P0; p < 36; p++) raw[p] = __fmul_rn(r2[p], c_win[ebt * 36 + p]);
            }
        }
        if (__any_sync(0xffffffffu, ebt == 2)) {
            float r2[36];
            imdct12_win(in, r2);
            if (ebt == 2) {
#pragma unroll
                for (int p = 0; p < 36; p++) raw[p] = r2[p];
            }
        }
    }
}

__global__ void __launch_bounds__(kK3Warps * 32)
k_imdct(const mp3gpu_unit *__restrict__ units, long long first_granule, int n_granules, WaveBufs B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kK3Warps + warp;
    const int ch = item & 1;
    const int g0 = (item >> 1) * kRun;
    if (g0 >= n_granules) return;
    const int g1 = min(g0 + kRun, n_granules);
    float st[18];
#pragma unroll
    for (int i = 0; i < 18; i++) st[i] = 0.0f;
    {
        // Overlap state entering the run: second IMDCT half of the previous granule, unless
        // this granule starts from zero state or the previous one has no such channel.
        const mp3gpu_unit *uc = units + (first_granule + g0) * 2 + ch;
        if (!u_zero(uc->w2) && (first_granule + g0) > 0) {
            const mp3gpu_unit *up = uc - 2;
            if (u_valid(up->w2)) {
                float in[18], raw[36];
                const float *src = B.xr_t + ((long long)(g0 - 1) * 2 + ch) * 576;
#pragma unroll
                for (int m = 0; m < 18; m++) in[m] = src[m * 32 + lane];
                imdct_granule<1>(in, raw, up->w0, up->w2, lane);
#pragma unroll
                for (int i = 0; i < 18; i++) st[i] = raw[18 + i];
            }
        }
    }
    for (int g = g0; g < g1; g++) {
        const mp3gpu_unit *uc = units + (first_granule + g) * 2 + ch;
        const uint32_t w0 = uc->w0, w2 = uc->w2;
        if (u_zero(w2)) {
#pragma unroll
            for (int i = 0; i < 18; i++) st[i] = 0.0f;
        }
        if (!u_valid(w2)) {
            // channel absent in this granule (mono): state restarts from zero afterwards
#pragma unroll
            for (int i = 0; i < 18; i++) st[i] = 0.0f;
            continue;
        }
        float in[18], raw[36];
        const float *src = B.xr_t + ((long long)g * 2 + ch) * 576;
#pragma unroll
        for (int m = 0; m < 18; m++) in[m] = src[m * 32 + lane];
        imdct_granule<0>(in, raw, w0, w2, lane);
        float *dst = B.hyb[ch] + (long long)g * 18 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 18; i++) {
            float v = __fadd_rn(raw[i], st[i]);  // frame.go:474
            st[i] = raw[i + 18];                 // frame.go:475
            if ((i & 1) && (lane & 1)) v = -v;   // frame.go:480-486
            dst[i * 32] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: polyphase synthesis.  One thread per time slot; CTA = kK4Threads consecutive slots of which
// the first 16 are the V-history halo (matrixing only).
// ------------------------------------------------------------------------------------------
constexpr int kK4Threads = 128;
constexpr int kK4Halo = 16;
constexpr int kK4Slots = kK4Threads - kK4Halo;  // output slots per CTA
constexpr int kVStride = 68;                    // floats per V row: 16-byte aligned, conflict-free for 128-bit LDS

__device__ __forceinline__ int pcm_from_float(float sum) {
    // frame.go:663-668: int(sum * 32767) truncates; Go on amd64 (CVTTSS2SQ) yields INT64_MIN for NaN / out of
    // int64 range, which the clamp turns into -32767.
    float f = __fmul_rn(sum, 32767.0f);
    if (!(f < 9223372036854775808.0f && f >= -9223372036854775808.0f)) return -32767;
    int s = __float2int_rz(f);  // saturates beyond int32, which the clamp absorbs
    return max(-32767, min(32767, s));
}

__global__ void __launch_bounds__(kK4Threads)
k_synth(const mp3gpu_unit *__restrict__ units, long long first_granule, int n_granules, WaveBufs B,
        int16_t *__restrict__ pcm /* wave-local: [n_granules][576][2] */) {
    __shared__ __align__(16) float s_v[kK4Threads * kVStride];
    const int tid = threadIdx.x;
    const long long t = (long long)blockIdx.x * kK4Slots - kK4Halo + tid;  // wave-local slot index
    const long long n_slots = (long long)n_granules * 18;
    const bool in_range = t >= -18 && t < n_slots && (first_granule * 18 + t) >= 0;
    const bool is_out = tid >= kK4Halo && t < n_slots;
    long long g = 0;
    uint32_t w2a = 0, w2b = 0;
    int slot_in_gr = 0;
    if (in_range) {
        g = t >= 0 ? t / 18 : -1;
        slot_in_gr = (int)(t - g * 18);
        const mp3gpu_unit *ug = units + (first_granule + g) * 2;
        w2a = ug[0].w2;
        w2b = ug[1].w2;
    }
    // History available to this slot: a zero-state granule has nothing before its first slot.
    const int hist = (in_range && u_zero(w2a)) ? slot_in_gr : 15;
    uint32_t packed[32];  // per sample: L | R<<16
#pragma unroll
    for (int i = 0; i < 32; i++) packed[i] = 0;

#pragma unroll 1
    for (int ch = 0; ch < 2; ch++) {
        const bool valid = in_range && u_valid(ch ? w2b : w2a);
        float *vrow = s_v + tid * kVStride;
        if (valid) {
            float s[32];
            const float4 *src = reinterpret_cast<const float4 *>(B.hyb[ch] + t * 32);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float4 v = __ldg(src + q);
                s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
            }
            // Matrixing: V[i] = sum_j N[i][j] * s[j], j ascending (frame.go:644-650)
#pragma unroll
            for (int i4 = 0; i4 < 16; i4++) {
                float v4[4];
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const int i = i4 * 4 + ii;
                    float sum = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; j++) sum = mac(c_synth_n[i * 32 + j], s[j], sum);
                    v4[ii] = sum;
                }
                *reinterpret_cast<float4 *>(vrow + i4 * 4) = make_float4(v4[0], v4[1], v4[2], v4[3]);
            }
        } else {
#pragma unroll
            for (int i4 = 0; i4 < 16; i4++) *reinterpret_cast<float4 *>(vrow + i4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        if (is_out && valid) {
            // out[i] = sum over k of V[t-2k][i]*D[64k+i] then V[t-2k-1][32+i]*D[64k+32+i]
            // (U construction frame.go:651-661), taps ascending.
#pragma unroll
            for (int i4 = 0; i4 < 8; i4++) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int d = 0; d < 16; d++) {
                    // d even: row t-d, columns i ; d odd: row t-d, columns 32+i
                    const int col = (d & 1) ? 32 + i4 * 4 : i4 * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (d <= hist) v = *reinterpret_cast<const float4 *>(s_v + (tid - d) * kVStride + col);
                    const int dbase = (d >> 1) * 64 + ((d & 1) ? 32 : 0) + i4 * 4;
                    acc[0] = mac(v.x, c_synth_d[dbase + 0], acc[0]);
                    acc[1] = mac(v.y, c_synth_d[dbase + 1], acc[1]);
                    acc[2] = mac(v.z, c_synth_d[dbase + 2], acc[2]);
                    acc[3] = mac(v.w, c_synth_d[dbase + 3], acc[3]);
                }
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    uint32_t s16 = (uint32_t)pcm_from_float(acc[ii]) & 0xffffu;
                    packed[i4 * 4 + ii] |= ch ? (s16 << 16) : s16;
                }
            }
        }
        __syncthreads();
    }
    if (is_out && in_range && u_valid(w2a)) {
        if (!u_valid(w2b)) {  // mono: duplicate (frame.go:671-678)
#pragma unroll
            for (int i = 0; i < 32; i++) packed[i] |= packed[i] << 16;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(pcm + t * 64);
#pragma unroll
        for (int q = 0; q < 8; q++) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
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
