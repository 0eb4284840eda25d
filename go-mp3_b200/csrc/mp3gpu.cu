// mp3gpu.cu — C ABI (include/mp3gpu.h) of the B200 MP3 Layer III granule decode engine.
// Context management, wave scheduling, host<->device pipelining, timing and debug taps.
// The kernels are in kernels.cuh; constant tables are built by tables.cc.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mp3gpu.h"
#include "kernels.cuh"
#include "tables.h"

using namespace mp3gpu;

static_assert(sizeof(mp3gpu_unit) == 32, "mp3gpu_unit must be 32 bytes");

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            if (e__ == cudaErrorMemoryAllocation) {                                           \
                cudaGetLastError(); /* not sticky: the context stays usable */                \
                return MP3GPU_E_NOMEM;                                                        \
            }                                                                                 \
            return MP3GPU_E_CUDA;                                                             \
        }                                                                                     \
    } while (0)

namespace {
constexpr int kMaxSubWaves = 1024;  // dynamic-scheduler counters of k_hybrid in sub-wave mode
constexpr int kTimingSlots = 64;  // per-kernel event pairs kept per call (waves beyond this are not timed individually)
}

struct mp3gpu_ctx {
    int device = 0;
    std::string err;
    mp3gpu_opts opts{};
    uint32_t wave = 0;
    cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr;
    // tables
    void *d_tab_blob = nullptr;
    DeviceTables T{};
    int lut_bytes = 0;
    int smem_per_sm = 0, smem_per_cta_max = 0;  // shared-memory budget (bytes) of an SM / of one CTA (opt-in maximum)
    int huff_static_smem = 0;                   // static shared memory of k_huffman
    int k1_upw_override = 0, k1_warps_override = 0, k1_stage_pct_override = 0;  // experiments: MP3GPU_K1_UPW / _WARPS / _STAGE_PCT
    // workspace for one wave (+1 granule look-back where needed)
    int16_t *d_is16 = nullptr;
    uint32_t *d_meta = nullptr;
    uint32_t *d_sfpack = nullptr;
    float *d_hyb = nullptr;          // (wave+1) granules x [2][18][32] subband samples, look-back granule in front
    float *d_tap_xr = nullptr;       // debug tap (opts.keep_intermediates)
    float *d_synth = nullptr;        // synth_d[512]
    unsigned int *d_counter = nullptr;
    int sm_count = 0;
    int seg_len = 32;
    int sub_granules = 0;     // > 0: k_hybrid / k_synth alternate over sub-waves of this many granules (L2-resident hand-off)
    int sub_seg_len = 8;      // k_hybrid segment length in sub-wave mode
    int sub_syn_blocks = 6;   // k_synth blocks of 30 slots per warp segment in sub-wave mode
    size_t ws_granules = 0;   // granules the wave workspace currently holds
    // staging for host-buffer calls
    uint8_t *d_main = nullptr;
    size_t d_main_cap = 0;
    mp3gpu_unit *d_units = nullptr;
    size_t d_units_cap = 0;
    int16_t *d_pcm_ring[3] = {nullptr, nullptr, nullptr};
    size_t ring_granules = 0;  // granules each ring slot holds
    cudaEvent_t ev_in[3]{}, ev_k[3]{}, ev_out[3]{};
    // timing
    cudaEvent_t ev_t[kTimingSlots][4]{};
    cudaEvent_t ev_copy[4]{};
    cudaEvent_t ev_user[8]{};
    mp3gpu_timings last{};
    int last_slots = 0;          // timing slots recorded by the last call
    bool last_collected = true;  // per-kernel times of the last call already read back
    // taps
    size_t last_wave_granules = 0;
    long long last_wave_first = 0;
};

// Wave workspace, sized for min(wave, granules of the call) and grown on demand: a streaming Decoder that submits a
// few hundred frames at a time should not pay for a million-granule batch workspace.
static int ensure_workspace(mp3gpu_ctx *ctx, size_t granules) {
    const size_t W = std::min<size_t>(ctx->wave, std::max<size_t>(granules, 1));
    if (W <= ctx->ws_granules) return MP3GPU_OK;
    CK(cudaStreamSynchronize(ctx->s_compute));
    cudaFree(ctx->d_is16); cudaFree(ctx->d_meta); cudaFree(ctx->d_sfpack); cudaFree(ctx->d_hyb); cudaFree(ctx->d_tap_xr);

    ctx->d_is16 = nullptr; ctx->d_meta = nullptr; ctx->d_sfpack = nullptr; ctx->d_hyb = nullptr; ctx->d_tap_xr = nullptr;
    ctx->ws_granules = 0;
    // K1 outputs, with two look-back granules (halo of k_hybrid) in front
    CK(cudaMalloc(&ctx->d_is16, (W + 2) * 2 * 576 * sizeof(int16_t)));
    CK(cudaMalloc(&ctx->d_meta, (W + 2) * 2 * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->d_sfpack, (W + 2) * 2 * 8 * sizeof(uint32_t)));
    CK(cudaMemset(ctx->d_is16, 0, 2 * 2 * 576 * sizeof(int16_t)));
    CK(cudaMemset(ctx->d_meta, 0, 2 * 2 * sizeof(uint32_t)));
    CK(cudaMemset(ctx->d_sfpack, 0, 2 * 2 * 8 * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->d_hyb, (W + 1) * 2 * 576 * sizeof(float)));
    CK(cudaMemset(ctx->d_hyb, 0, 2 * 576 * sizeof(float)));
    if (ctx->opts.keep_intermediates) CK(cudaMalloc(&ctx->d_tap_xr, W * 2 * 576 * sizeof(float)));
    ctx->ws_granules = W;
    return MP3GPU_OK;
}

static int upload_tables(mp3gpu_ctx *ctx) {
    HostTables h;
    try {
        build_host_tables(h);
    } catch (const std::exception &e) {
        ctx->err = std::string("table build failed: ") + e.what();
        return MP3GPU_E_INVALID;
    }
    // The compile-time IMDCT tables (const_tables.inc) must be the ones this host computes.
    if (memcmp(kCos36_host, h.cos36, sizeof h.cos36) != 0 || memcmp(kCos12_host, h.cos12, sizeof h.cos12) != 0 ||
        memcmp(kWin_host, h.imdct_win, sizeof h.imdct_win) != 0 || memcmp(kSynthN_host, h.synth_n, sizeof h.synth_n) != 0) {
        ctx->err = "const_tables.inc does not match the host-built IMDCT tables (rebuild the library on this libm)";
        return MP3GPU_E_INVALID;
    }
    {
        float cs[8], ca[8];
        memcpy(cs, kCs, sizeof cs);
        memcpy(ca, kCa, sizeof ca);
        if (memcmp(cs, h.cs, sizeof cs) != 0 || memcmp(ca, h.ca, sizeof ca) != 0) {
            ctx->err = "alias-reduction constants differ between kernels.cuh and tables.cc";
            return MP3GPU_E_INVALID;
        }
    }
    CK(cudaMemcpyToSymbol(c_win, h.imdct_win, sizeof h.imdct_win));
    {   // fast IMDCT read-out (kernels.cuh, imdct36_emit): windowed out[p] = z[idx(p)] * winz[bt][p]
        float winz[4 * 36];
        const double pi = 3.14159265358979323846;
        for (int bt = 0; bt < 4; bt++)
            for (int p = 0; p < 36; p++) {
                const int kk = p < 9 ? p + 9 : (p < 27 ? 26 - p : p - 27);
                const double sec = 1.0 / (2.0 * cos(pi * (2 * kk + 1) / 72.0));
                winz[bt * 36 + p] = (float)((p < 9 ? sec : -sec) * (double)h.imdct_win[bt * 36 + p]);
            }
        CK(cudaMemcpyToSymbol(c_winz, winz, sizeof winz));
    }
    // the symmetric matrixing of k_synth relies on these identities holding bitwise in the float32 table
    for (int j = 0; j < 32; j++) {
        for (int i = 0; i <= 15; i++)
            if (h.synth_n[(32 - i) * 32 + j] != -h.synth_n[i * 32 + j]) { ctx->err = "synthNWin mirror-16 symmetry does not hold"; return MP3GPU_E_INVALID; }
        for (int k = 1; k <= 15; k++)
            if (h.synth_n[(48 + k) * 32 + j] != h.synth_n[(48 - k) * 32 + j]) { ctx->err = "synthNWin mirror-48 symmetry does not hold"; return MP3GPU_E_INVALID; }
    }
    CK(cudaMalloc(&ctx->d_synth, 512 * sizeof(float)));
    CK(cudaMemcpy(ctx->d_synth, h.synth_d, sizeof h.synth_d, cudaMemcpyHostToDevice));
    // maindata.go:39-42
    static const uint8_t slen[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {3, 0}, {1, 1}, {1, 2}, {1, 3},
                                        {2, 1}, {2, 2}, {2, 3}, {3, 1}, {3, 2}, {3, 3}, {4, 2}, {4, 3}};

    // one blob for the per-lane tables
    std::vector<uint8_t> blob;
    auto put = [&](const void *p, size_t n) {
        size_t off = (blob.size() + 255) & ~size_t(255);
        blob.resize(off + n);
        memcpy(blob.data() + off, p, n);
        return off;
    };
    size_t o_pow2 = put(h.pow2q, sizeof h.pow2q);
    size_t o_p34 = put(h.powtab34.data(), h.powtab34.size() * sizeof(double));
    size_t o_pq4 = put(h.powq4.data(), h.powq4.size() * sizeof(float));
    size_t o_lsl = put(h.line_sfb_long, sizeof h.line_sfb_long);
    size_t o_lss = put(h.line_sfb_short, sizeof h.line_sfb_short);
    size_t o_lws = put(h.line_win_short, sizeof h.line_win_short);
    size_t o_rd = put(h.reorder_dst, sizeof h.reorder_dst);
    size_t o_pl = put(h.pair_long, sizeof h.pair_long);
    size_t o_ps = put(h.pair_short, sizeof h.pair_short);
    size_t o_pd = put(h.pair_dst, sizeof h.pair_dst);
    size_t o_sl = put(h.sfb_long, sizeof h.sfb_long);
    size_t o_ss = put(h.sfb_short, sizeof h.sfb_short);
    size_t o_ns = put(h.nslen2, sizeof h.nslen2);
    size_t o_lut = put(h.huff_lut.data(), h.huff_lut.size() * sizeof(uint16_t));
    size_t o_ql = put(h.quad_lut, sizeof h.quad_lut);
    size_t o_qs = put(h.quad_signs, sizeof h.quad_signs);
    size_t o_rl = put(h.is_ratio_l, sizeof h.is_ratio_l);
    size_t o_rr = put(h.is_ratio_r, sizeof h.is_ratio_r);
    size_t o_pt = put(h.pretab, sizeof h.pretab);
    size_t o_hd = put(h.huff_desc, sizeof h.huff_desc);
    size_t o_s2 = put(h.sfsize_mpeg2, sizeof h.sfsize_mpeg2);
    size_t o_s1 = put(slen, sizeof slen);
    size_t o_cs = put(h.cs, sizeof h.cs);
    size_t o_ca = put(h.ca, sizeof h.ca);
    CK(cudaMalloc(&ctx->d_tab_blob, blob.size()));
    CK(cudaMemcpy(ctx->d_tab_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    uint8_t *b = (uint8_t *)ctx->d_tab_blob;
    ctx->T.pow2q = (const double *)(b + o_pow2);
    ctx->T.powtab34 = (const double *)(b + o_p34);
    ctx->T.powq4 = (const float *)(b + o_pq4);
    ctx->T.pretab_pack = h.pretab_pack;
    ctx->T.line_sfb_long = b + o_lsl;
    ctx->T.line_sfb_short = b + o_lss;
    ctx->T.line_win_short = b + o_lws;
    ctx->T.reorder_dst = (const uint16_t *)(b + o_rd);
    ctx->T.pair_long = b + o_pl;
    ctx->T.pair_short = b + o_ps;
    ctx->T.pair_dst = (const uint16_t *)(b + o_pd);
    ctx->T.sfb_long = (const uint16_t *)(b + o_sl);
    ctx->T.sfb_short = (const uint16_t *)(b + o_ss);
    ctx->T.nslen2 = (const uint16_t *)(b + o_ns);
    ctx->T.huff_lut = (const uint16_t *)(b + o_lut);
    ctx->T.quad_lut = (const uint32_t *)(b + o_ql);
    ctx->T.quad_signs = (const uint64_t *)(b + o_qs);
    ctx->T.is_ratio_l = (const float *)(b + o_rl);
    ctx->T.is_ratio_r = (const float *)(b + o_rr);
    ctx->T.pretab = b + o_pt;
    ctx->T.huff_desc = (const uint32_t *)(b + o_hd);
    ctx->T.sfsize_mpeg2 = b + o_s2;
    ctx->T.slen_mpeg1 = b + o_s1;
    ctx->T.cs = (const float *)(b + o_cs);
    ctx->T.ca = (const float *)(b + o_ca);
    ctx->T.huff_lut_n = (int)h.huff_lut.size();
    ctx->T.pow2_off = kPow2Off;
    ctx->lut_bytes = (int)(h.huff_lut.size() * sizeof(uint16_t));
    return MP3GPU_OK;
}

extern "C" int mp3gpu_create(int device, const mp3gpu_opts *opts, mp3gpu_ctx **out) {
    if (!out) return MP3GPU_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return MP3GPU_E_NO_DEVICE;
    if (opts && opts->abi_version != MP3GPU_ABI_VERSION) {
        fprintf(stderr, "mp3gpu_create: caller was built against ABI version %u, this library is version %u\n", opts->abi_version,
                (unsigned)MP3GPU_ABI_VERSION);
        return MP3GPU_E_INVALID;
    }
    mp3gpu_ctx *ctx = new mp3gpu_ctx();
    ctx->device = device;
    if (opts) ctx->opts = *opts;
    ctx->wave = ctx->opts.wave_granules ? ctx->opts.wave_granules : 2097152u;
    if (ctx->wave > (1u << 24)) ctx->wave = 1u << 24;  // the kernels index a wave's units and time slots with 32-bit integers
    auto fail = [&](int rc) {
        fprintf(stderr, "mp3gpu_create: %s\n", ctx->err.c_str());
        mp3gpu_destroy(ctx);
        return rc;
    };
    auto init = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        int rc = upload_tables(ctx);
        if (rc) return rc;
        CK(cudaMalloc(&ctx->d_counter, (2 + kMaxSubWaves) * sizeof(unsigned int)));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        for (int i = 0; i < 3; i++) {
            CK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < kTimingSlots; i++)
            for (int j = 0; j < 4; j++) CK(cudaEventCreate(&ctx->ev_t[i][j]));
        for (int i = 0; i < 4; i++) CK(cudaEventCreate(&ctx->ev_copy[i]));
        for (int i = 0; i < 8; i++) CK(cudaEventCreate(&ctx->ev_user[i]));
        ctx->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
        ctx->smem_per_cta_max = (int)prop.sharedMemPerBlockOptin;
        {
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, k_huffman<64>));
            ctx->huff_static_smem = (int)fa.sharedSizeBytes;
            CK(cudaFuncSetAttribute(k_huffman<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_per_cta_max - ctx->huff_static_smem));
            CK(cudaFuncSetAttribute(k_huffman<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_per_cta_max - ctx->huff_static_smem));
            if (const char *e = getenv("MP3GPU_K1_UPW")) ctx->k1_upw_override = atoi(e);
            if (const char *e = getenv("MP3GPU_SEG_LEN")) ctx->seg_len = std::max(2, atoi(e));  // experiments: k_hybrid segment length
            if (const char *e = getenv("MP3GPU_SUB")) ctx->sub_granules = std::max(0, atoi(e));
            if (const char *e = getenv("MP3GPU_SUB_SEG")) ctx->sub_seg_len = std::max(2, atoi(e));
            if (const char *e = getenv("MP3GPU_SUB_SYN")) ctx->sub_syn_blocks = std::max(1, atoi(e));
            if (const char *e = getenv("MP3GPU_K1_WARPS")) ctx->k1_warps_override = atoi(e);
            if (const char *e = getenv("MP3GPU_K1_STAGE_PCT")) ctx->k1_stage_pct_override = atoi(e);
        }
        CK(cudaFuncSetAttribute(k_hybrid<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHybSmemBytes));
        CK(cudaFuncSetAttribute(k_hybrid<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHybSmemBytes));
        CK(cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, kSynSmemBytes));
        return MP3GPU_OK;
    };
    int rc = init();
    if (rc) return fail(rc);
    *out = ctx;
    return MP3GPU_OK;
}

extern "C" void mp3gpu_destroy(mp3gpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_tab_blob);
    cudaFree(ctx->d_is16);
    cudaFree(ctx->d_meta);
    cudaFree(ctx->d_sfpack);
    cudaFree(ctx->d_hyb);
    cudaFree(ctx->d_tap_xr);
    cudaFree(ctx->d_synth);
    cudaFree(ctx->d_counter);
    cudaFree(ctx->d_main);
    cudaFree(ctx->d_units);
    for (int i = 0; i < 3; i++) {
        cudaFree(ctx->d_pcm_ring[i]);
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    for (int i = 0; i < kTimingSlots; i++)
        for (int j = 0; j < 4; j++)
            if (ctx->ev_t[i][j]) cudaEventDestroy(ctx->ev_t[i][j]);
    for (int i = 0; i < 4; i++)
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    for (int i = 0; i < 8; i++)
        if (ctx->ev_user[i]) cudaEventDestroy(ctx->ev_user[i]);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    delete ctx;
}

extern "C" const char *mp3gpu_last_error(const mp3gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

// Launch the four kernels for granules [first, first+n) of the submission on s_compute.
// d_pcm_wave points at the PCM of granule `first`.  `slot` selects the timing events (or -1).
static int launch_wave(mp3gpu_ctx *ctx, const uint8_t *d_main, size_t main_len, const mp3gpu_unit *d_units, long long first, int n,
                       int16_t *d_pcm_wave, int slot, double bytes_per_unit, size_t n_granules_total) {
    WaveBufs B;
    B.is16 = ctx->d_is16 + 2 * 2 * 576;  // granules -2, -1 live in front
    B.meta = ctx->d_meta + 2 * 2;
    B.sfpack = ctx->d_sfpack + 2 * 2 * 8;
    B.hyb = ctx->d_hyb + 2 * 576;        // granule -1 lives in front
    B.tap_xr = ctx->d_tap_xr;
    B.synth_d = ctx->d_synth;
    B.work_counter = ctx->d_counter;
    B.units_total = (long long)n_granules_total * 2;
    B.n_gran = n;
#if MP3GPU_PROBE
    B.probe = getenv("MP3GPU_PROBE") ? atoi(getenv("MP3GPU_PROBE")) : 0;
#endif
    cudaStream_t s = ctx->s_compute;
    if (slot >= 0) CK(cudaEventRecord(ctx->ev_t[slot][0], s));
    {
        // k_huffman: one persistent CTA per SM whose warps each stage the stretch of main data a tile of 32 or 64 units
        // reads in their own piece of shared memory.  The piece is sized from the call's average bytes per unit (x 1.25 for
        // tile-to-tile variation, VBR; a tile that needs more reads its tail from global memory), and the CTA gets as many
        // warps as fit next to the 30 KB of code tables.  64-unit tiles (sorted, two passes) keep the lanes of a warp
        // busier; 32-unit tiles allow twice the warps when the bitrate is high.
        const int nu = 2 * n;
        const unsigned long long main_bits = (unsigned long long)main_len * 8ull;
        const double avg = bytes_per_unit > 1.0 ? bytes_per_unit : 1.0;
        const int budget = ctx->smem_per_cta_max - ctx->huff_static_smem - ctx->lut_bytes;
        const int pct = ctx->k1_stage_pct_override ? ctx->k1_stage_pct_override : 125;
        int upw = 64, warps = 0, cap16 = 0;
        for (int pass = 0; pass < 2; pass++) {
            upw = pass == 0 ? 64 : 32;
            if (ctx->k1_upw_override) upw = ctx->k1_upw_override == 32 ? 32 : 64;
            int stage = ((int)(avg * upw * pct / 100.0) + 512 + 15) & ~15;
            const int max_stage = budget / 8 - (16 + 160 + upw);  // at least eight warps
            if (stage > max_stage) stage = max_stage & ~15;
            const int per_warp = stage + 16 + 160 + upw;
            warps = std::min(32, budget / per_warp);
            if (ctx->k1_warps_override) warps = std::min(warps, ctx->k1_warps_override);
            cap16 = stage / 16;
            if (warps >= 20 || pass == 1 || ctx->k1_upw_override) break;
        }
        if (warps < 1) warps = 1;
        const int tiles = (nu + upw - 1) / upw;
        const int grid = std::min((tiles + warps - 1) / warps, ctx->sm_count);
        const size_t dyn = (size_t)ctx->lut_bytes + (size_t)warps * (size_t)(cap16 * 16 + 16 + 160 + upw);
        CK(cudaMemsetAsync(ctx->d_counter + 1, 0, sizeof(unsigned int), s));
        if (upw == 64) k_huffman<64><<<grid, warps * 32, dyn, s>>>(d_main, main_bits, d_units, first * 2, nu, ctx->T, B, cap16, ctx->d_counter + 1);
        else k_huffman<32><<<grid, warps * 32, dyn, s>>>(d_main, main_bits, d_units, first * 2, nu, ctx->T, B, cap16, ctx->d_counter + 1);
    }
    if (slot >= 0) CK(cudaEventRecord(ctx->ev_t[slot][1], s));
    const int sub = (ctx->sub_granules > 0 && !ctx->d_tap_xr) ? ctx->sub_granules : 0;
    if (sub == 0) {
        {
            CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int), s));
            const int n_segs = (n + ctx->seg_len - 1) / ctx->seg_len;
            const int grid = std::min((n_segs + kHybWarps - 1) / kHybWarps, ctx->sm_count * 5);
            if (ctx->d_tap_xr) {
                CK(cudaMemsetAsync(ctx->d_tap_xr, 0, (size_t)n * 2 * 576 * sizeof(float), s));
                CK(cudaMemsetAsync(B.hyb, 0, (size_t)n * 2 * 576 * sizeof(float), s));  // taps of absent channels read as 0
                k_hybrid<true><<<grid, kHybWarps * 32, kHybSmemBytes, s>>>(d_units, first, n, ctx->seg_len, n_segs, ctx->T, B, 0);
            } else {
                k_hybrid<false><<<grid, kHybWarps * 32, kHybSmemBytes, s>>>(d_units, first, n, ctx->seg_len, n_segs, ctx->T, B, 0);
            }
        }
        if (slot >= 0) CK(cudaEventRecord(ctx->ev_t[slot][2], s));
        {
            const long long slots = (long long)n * 18;
            const long long segs = (slots + kSynSegBlocks * kSynBlock - 1) / (kSynSegBlocks * kSynBlock);
            const int grid = (int)((segs + kSynWarps - 1) / kSynWarps);
            k_synth<<<grid, kSynThreads, kSynSmemBytes, s>>>(d_units, first, n, B, d_pcm_wave, kSynSegBlocks);
        }
        if (slot >= 0) CK(cudaEventRecord(ctx->ev_t[slot][3], s));
        // Carry the last granule's subband samples into the look-back slot for the next wave's k_synth halo.
        CK(cudaMemcpyAsync(ctx->d_hyb, B.hyb + (size_t)(n - 1) * 2 * 576, 2 * 576 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else {
        // L2-resident hand-off: k_hybrid and k_synth alternate over sub-waves small enough for the subband samples between them
        // (hyb, 4,608 bytes per granule) to stay in the 126 MB L2: every sub-wave writes and reads the SAME hyb area, so the
        // samples are overwritten in L2 before they are ever evicted to HBM.  The segments of both kernels are shortened so
        // that one sub-wave still fills the machine.
        if (slot >= 0) CK(cudaEventRecord(ctx->ev_t[slot][2], s));  // k_hybrid_ms reads 0; k_synth_ms holds both kernels
        const int n_sub = (n + sub - 1) / sub;
        CK(cudaMemsetAsync(ctx->d_counter + 2, 0, (size_t)std::min(n_sub, kMaxSubWaves) * sizeof(unsigned int), s));
        for (int k = 0; k < n_sub; k++) {
            const int g0 = k * sub, ng = std::min(sub, n - g0);
            WaveBufs S = B;
            S.is16 = B.is16 + (size_t)g0 * 2 * 576;
            S.meta = B.meta + (size_t)g0 * 2;
            S.sfpack = B.sfpack + (size_t)g0 * 2 * 8;
            S.n_gran = ng;  // hyb stays: every sub-wave uses granule slots [-1, sub) of it
            if (k >= kMaxSubWaves) CK(cudaMemsetAsync(ctx->d_counter + 2 + k % kMaxSubWaves, 0, sizeof(unsigned int), s));
            const int n_segs = (ng + ctx->sub_seg_len - 1) / ctx->sub_seg_len;
            const int grid = std::min((n_segs + kHybWarps - 1) / kHybWarps, ctx->sm_count * 5);
            k_hybrid<false><<<grid, kHybWarps * 32, kHybSmemBytes, s>>>(d_units, first + g0, ng, ctx->sub_seg_len, n_segs, ctx->T, S, 2 + k % kMaxSubWaves);
            const long long slots = (long long)ng * 18;
            const long long segs = (slots + ctx->sub_syn_blocks * kSynBlock - 1) / (ctx->sub_syn_blocks * kSynBlock);
            k_synth<<<(int)((segs + kSynWarps - 1) / kSynWarps), kSynThreads, kSynSmemBytes, s>>>(d_units, first + g0, ng, S, d_pcm_wave + (size_t)g0 * 1152,
                                                                                                  ctx->sub_syn_blocks);
            k_carry<<<1, 256, 0, s>>>(reinterpret_cast<float4 *>(ctx->d_hyb), reinterpret_cast<const float4 *>(S.hyb + (size_t)(ng - 1) * 2 * 576), 2 * 576 / 4);
            ctx->last.launches += 2;
        }
        ctx->last.launches -= 2;  // the caller's count below adds three per wave
        if (slot >= 0) CK(cudaEventRecord(ctx->ev_t[slot][3], s));
    }
    // Carry K1's outputs of the last two granules into the look-back slots for the next wave.  In order k = 0, 1
    // so that a one-granule wave shifts slot -1 to slot -2 before overwriting it.
    for (int k = 0; k < 2; k++) {
        const long long src = (long long)n - 2 + k;  // wave-local granule index; negative = an old look-back slot
        CK(cudaMemcpyAsync(B.is16 + (k - 2) * 2 * 576, B.is16 + src * 2 * 576, 2 * 576 * sizeof(int16_t), cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(B.meta + (k - 2) * 2, B.meta + src * 2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(B.sfpack + (k - 2) * 2 * 8, B.sfpack + src * 2 * 8, 2 * 8 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    CK(cudaGetLastError());
    ctx->last.launches += 3;
    ctx->last_wave_first = first;
    ctx->last_wave_granules = (size_t)n;
    return MP3GPU_OK;
}

// Checked build: a guard that failed in a kernel (MP3_CHECK, unit_logic.h) becomes an error of the call that ran it.
static int check_faults(mp3gpu_ctx *ctx) {
#if MP3GPU_CHECKED
    unsigned int f[4] = {0, 0, 0, 0};
    CK(cudaMemcpyFromSymbol(f, g_fault, sizeof f));
    if (f[0]) {
        const unsigned int zero[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_fault, zero, sizeof zero);
        ctx->err = "checked build: " + std::to_string(f[0]) + " out-of-range access(es) stopped; first at source line " + std::to_string(f[1]) +
                   ", index " + std::to_string((long long)(((unsigned long long)f[3] << 32) | f[2]));
        return MP3GPU_E_CUDA;
    }
#else
    (void)ctx;
#endif
    return MP3GPU_OK;
}

static int collect_timings(mp3gpu_ctx *ctx, int nslots) {
    float k[3] = {0, 0, 0};
    for (int i = 0; i < nslots; i++)
        for (int j = 0; j < 3; j++) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev_t[i][j], ctx->ev_t[i][j + 1]));
            k[j] += ms;
        }
    ctx->last.k1_huffman_ms = k[0];
    ctx->last.k_hybrid_ms = k[1];
    ctx->last.k_synth_ms = k[2];
    if (nslots > 0) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev_t[0][0], ctx->ev_t[nslots - 1][3]));
        ctx->last.total_ms = ms;
    }
    return MP3GPU_OK;
}

extern "C" int mp3gpu_decode_device_async(mp3gpu_ctx *ctx, const uint8_t *d_main_data, size_t main_data_len,
                                          const mp3gpu_unit *d_units, size_t n_granules, int16_t *d_pcm_out) {
    if (!ctx) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->last = mp3gpu_timings{};
    ctx->last_slots = 0;
    ctx->last_collected = true;
    if (n_granules == 0) return MP3GPU_OK;
    if (!d_main_data || !d_units || !d_pcm_out) {
        ctx->err = "null device pointer";
        return MP3GPU_E_INVALID;
    }
    if (((uintptr_t)d_main_data & 15) || ((uintptr_t)d_units & 15) || ((uintptr_t)d_pcm_out & 3)) {
        // the kernels read main_data and the descriptors 16 bytes at a time and write PCM as 32-bit L|R words; a
        // misaligned access would fault, and a CUDA fault is sticky (it takes the whole context down)
        ctx->err = "device pointers must be aligned: main_data and units to 16 bytes, pcm to 4 bytes";
        return MP3GPU_E_INVALID;
    }
    {
        int rc = ensure_workspace(ctx, n_granules);
        if (rc) return rc;
    }
    int slot = 0;
    for (size_t first = 0; first < n_granules; first += ctx->ws_granules) {
        int n = (int)std::min<size_t>(ctx->ws_granules, n_granules - first);
        int rc = launch_wave(ctx, d_main_data, main_data_len, d_units, (long long)first, n, d_pcm_out + first * 1152,
                             slot < kTimingSlots ? slot : -1, (double)main_data_len / (2.0 * (double)n_granules), n_granules);
        if (rc) return rc;
        if (slot < kTimingSlots) slot++;
        ctx->last.waves++;
    }
    ctx->last_slots = slot;
    ctx->last_collected = false;
    return MP3GPU_OK;
}

extern "C" int mp3gpu_decode_device(mp3gpu_ctx *ctx, const uint8_t *d_main_data, size_t main_data_len,
                                    const mp3gpu_unit *d_units, size_t n_granules, int16_t *d_pcm_out) {
    int rc = mp3gpu_decode_device_async(ctx, d_main_data, main_data_len, d_units, n_granules, d_pcm_out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->s_compute));
    return check_faults(ctx);
}

extern "C" int mp3gpu_event_record(mp3gpu_ctx *ctx, int which) {
    if (!ctx || which < 0 || which >= 8) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_user[which], ctx->s_compute));
    return MP3GPU_OK;
}

extern "C" int mp3gpu_event_elapsed_ms(mp3gpu_ctx *ctx, int from, int to, float *ms) {
    if (!ctx || !ms || from < 0 || from >= 8 || to < 0 || to >= 8) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev_user[to]));
    CK(cudaEventElapsedTime(ms, ctx->ev_user[from], ctx->ev_user[to]));
    return MP3GPU_OK;
}

template <typename T>
static int ensure_cap(mp3gpu_ctx *ctx, T **p, size_t *cap, size_t need) {
    if (*cap >= need) return MP3GPU_OK;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 8 + 256;
    CK(cudaMalloc(p, want));
    *cap = want;
    return MP3GPU_OK;
}

extern "C" int mp3gpu_decode(mp3gpu_ctx *ctx, const uint8_t *main_data, size_t main_data_len, const mp3gpu_unit *units,
                             size_t n_granules, int16_t *pcm_out) {
    return mp3gpu_decode_range(ctx, main_data, main_data_len, units, n_granules, 0, pcm_out);
}

extern "C" int mp3gpu_decode_range(mp3gpu_ctx *ctx, const uint8_t *main_data, size_t main_data_len, const mp3gpu_unit *units,
                                   size_t n_granules, size_t first_out, int16_t *pcm_out) {
    if (!ctx) return MP3GPU_E_INVALID;
    if (first_out > n_granules) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->last = mp3gpu_timings{};
    ctx->last_slots = 0;
    ctx->last_collected = true;
    if (n_granules == 0) return MP3GPU_OK;
    if (!main_data || !units || !pcm_out) {
        ctx->err = "null pointer";
        return MP3GPU_E_INVALID;
    }

    int rc = ensure_cap(ctx, &ctx->d_main, &ctx->d_main_cap, main_data_len + 64);
    if (rc) return rc;
    rc = ensure_cap(ctx, (uint8_t **)&ctx->d_units, &ctx->d_units_cap, n_granules * 2 * sizeof(mp3gpu_unit));
    if (rc) return rc;
    // Host-buffer calls pipeline H2D / kernels / D2H wave by wave.  Only the first wave's upload + kernels and the last
    // wave's download are exposed, so a call is cut into 16 waves (of at least 16,384 granules to keep launches efficient).
    const size_t want = std::max<size_t>((n_granules + 15) / 16, std::min<size_t>(n_granules, 16384));
    rc = ensure_workspace(ctx, want);
    if (rc) return rc;
    const size_t W = std::min<size_t>(ctx->ws_granules, want);
    if (ctx->ring_granules < W) {  // sized by the wave this call uses, not by the (possibly much larger) workspace
        CK(cudaStreamSynchronize(ctx->s_out));
        for (int i = 0; i < 3; i++) {
            cudaFree(ctx->d_pcm_ring[i]);
            ctx->d_pcm_ring[i] = nullptr;
        }
        ctx->ring_granules = 0;
        for (int i = 0; i < 3; i++) CK(cudaMalloc(&ctx->d_pcm_ring[i], W * MP3GPU_PCM_BYTES_PER_GRANULE));
        ctx->ring_granules = W;
    }

    // Wave w needs main_data up to the end of its last unit's frame buffer.  Units are in stream
    // order, so the byte ranges are monotonic; each wave uploads only the bytes not yet resident.
    CK(cudaEventRecord(ctx->ev_copy[0], ctx->s_in));
    size_t main_resident = 0;
    int slot = 0;
    int widx = 0;
    for (size_t first = 0; first < n_granules; first += W, widx++) {
        const int n = (int)std::min<size_t>(W, n_granules - first);
        const int r = widx % 3;
        // -- H2D on s_in
        size_t need_end = main_resident;
        if (first + n >= n_granules) {
            need_end = main_data_len;
        } else {
            // buffer ends are monotonic in stream order: the last valid unit of the wave decides
            for (size_t u = (first + n) * 2; u-- > first * 2;) {
                if (!(units[u].w2 & MP3GPU_W2_VALID)) continue;
                long long e = (long long)units[u].bit_start + units[u].buf_end_rel;
                size_t eb = e > 0 ? (size_t)((e + 7) >> 3) : 0;
                need_end = std::max(need_end, eb);
                break;
            }
            need_end = std::min(need_end, main_data_len);
        }
        if (need_end > main_resident) {
            CK(cudaMemcpyAsync(ctx->d_main + main_resident, main_data + main_resident, need_end - main_resident,
                               cudaMemcpyHostToDevice, ctx->s_in));
            main_resident = need_end;
        }
        CK(cudaMemcpyAsync(ctx->d_units + first * 2, units + first * 2, (size_t)n * 2 * sizeof(mp3gpu_unit),
                           cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ctx->ev_in[r], ctx->s_in));
        // -- kernels on s_compute (wait for inputs and for the ring slot to be drained)
        CK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_in[r], 0));
        if (widx >= 3) CK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_out[r], 0));
        rc = launch_wave(ctx, ctx->d_main, main_data_len, ctx->d_units, (long long)first, n, ctx->d_pcm_ring[r],
                         slot < kTimingSlots ? slot : -1, (double)main_data_len / (2.0 * (double)n_granules), n_granules);
        if (rc) return rc;
        if (slot < kTimingSlots) slot++;
        CK(cudaEventRecord(ctx->ev_k[r], ctx->s_compute));
        // -- D2H on s_out
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[r], 0));
        if (widx == 0) CK(cudaEventRecord(ctx->ev_copy[2], ctx->s_out));
        if (first + n > first_out) {  // granules in front of first_out (a frame-range job's halo) are decoded but not returned
            const size_t skip = first_out > first ? first_out - first : 0;
            CK(cudaMemcpyAsync(pcm_out + (first + skip - first_out) * 1152, ctx->d_pcm_ring[r] + skip * 1152,
                               ((size_t)n - skip) * MP3GPU_PCM_BYTES_PER_GRANULE, cudaMemcpyDeviceToHost, ctx->s_out));
        }
        CK(cudaEventRecord(ctx->ev_out[r], ctx->s_out));
        ctx->last.waves++;
    }
    CK(cudaEventRecord(ctx->ev_copy[1], ctx->s_in));
    CK(cudaEventRecord(ctx->ev_copy[3], ctx->s_out));
    CK(cudaStreamSynchronize(ctx->s_in));
    CK(cudaStreamSynchronize(ctx->s_compute));
    CK(cudaStreamSynchronize(ctx->s_out));
    ctx->last_slots = slot;
    ctx->last_collected = false;
    CK(cudaEventElapsedTime(&ctx->last.h2d_ms, ctx->ev_copy[0], ctx->ev_copy[1]));
    CK(cudaEventElapsedTime(&ctx->last.d2h_ms, ctx->ev_copy[2], ctx->ev_copy[3]));
    return check_faults(ctx);
}

extern "C" void *mp3gpu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {  // portable: every device of a multi-GPU engine copies to / from it
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void mp3gpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" void *mp3gpu_device_alloc(mp3gpu_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void mp3gpu_device_free(mp3gpu_ctx *ctx, void *p) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}
extern "C" int mp3gpu_copy_to_device(mp3gpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return MP3GPU_OK;
}
extern "C" int mp3gpu_copy_to_host(mp3gpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return MP3GPU_OK;
}
extern "C" int mp3gpu_synchronize(mp3gpu_ctx *ctx) {
    if (!ctx) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    return check_faults(ctx);
}

extern "C" int mp3gpu_last_timings(mp3gpu_ctx *ctx, mp3gpu_timings *out) {
    if (!ctx || !out) return MP3GPU_E_INVALID;
    if (!ctx->last_collected) {
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->s_compute));
        int rc = collect_timings(ctx, ctx->last_slots);
        if (rc) return rc;
        ctx->last_collected = true;
    }
    *out = ctx->last;
    return MP3GPU_OK;
}

extern "C" int mp3gpu_debug_read(mp3gpu_ctx *ctx, int tap, size_t first, size_t n, void *host_out) {
    if (!ctx || !host_out) return MP3GPU_E_INVALID;
    if (first + n > ctx->last_wave_granules) {
        ctx->err = "debug_read: range outside the last wave";
        return MP3GPU_E_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> meta(n * 2);
    const int16_t *d_is = ctx->d_is16 + 2 * 2 * 576;
    const uint32_t *d_meta = ctx->d_meta + 2 * 2;
    const uint32_t *d_sf = ctx->d_sfpack + 2 * 2 * 8;
    CK(cudaMemcpy(meta.data(), d_meta + first * 2, n * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    switch (tap) {
    case MP3GPU_TAP_IS: {
        int16_t *o = (int16_t *)host_out;
        CK(cudaMemcpy(o, d_is + first * 2 * 576, n * 2 * 576 * sizeof(int16_t), cudaMemcpyDeviceToHost));
        for (size_t u = 0; u < n * 2; u++) {
            size_t c1 = meta[u] & 0x3ff;
            for (size_t i = c1; i < 576; i++) o[u * 576 + i] = 0;  // rzero region (huffman.go:130-134)
        }
        return MP3GPU_OK;
    }
    case MP3GPU_TAP_COUNT1: {
        int32_t *o = (int32_t *)host_out;
        for (size_t u = 0; u < n * 2; u++) o[u] = (int32_t)(meta[u] & 0x3ff);
        return MP3GPU_OK;
    }
    case MP3GPU_TAP_SCALEFAC: {
        std::vector<uint32_t> pk(n * 2 * 8);
        CK(cudaMemcpy(pk.data(), d_sf + first * 2 * 8, pk.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        uint8_t *o = (uint8_t *)host_out;
        for (size_t u = 0; u < n * 2; u++) {
            for (int k = 0; k < 64; k++) o[u * 64 + k] = (uint8_t)((pk[u * 8 + (k >> 3)] >> (4 * (k & 7))) & 0xf);
            o[u * 64 + 61] = (uint8_t)((meta[u] >> 10) & 1);
        }
        return MP3GPU_OK;
    }
    case MP3GPU_TAP_XR: {
        if (!ctx->d_tap_xr) {
            ctx->err = "debug_read: create the context with opts.keep_intermediates = 1";
            return MP3GPU_E_INVALID;
        }
        CK(cudaMemcpy(host_out, ctx->d_tap_xr + first * 2 * 576, n * 2 * 576 * sizeof(float), cudaMemcpyDeviceToHost));
        return MP3GPU_OK;
    }
    case MP3GPU_TAP_HYBRID: {
        if (!ctx->d_tap_xr) {
            ctx->err = "debug_read: create the context with opts.keep_intermediates = 1";
            return MP3GPU_E_INVALID;
        }
        std::vector<float> t(n * 2 * 576);
        CK(cudaMemcpy(t.data(), ctx->d_hyb + (first + 1) * 2 * 576, t.size() * sizeof(float), cudaMemcpyDeviceToHost));
        float *o = (float *)host_out;
        for (size_t u = 0; u < n * 2; u++)
            for (int i = 0; i < 18; i++)
                for (int sb = 0; sb < 32; sb++) o[u * 576 + sb * 18 + i] = t[u * 576 + i * 32 + sb];
        return MP3GPU_OK;
    }
    }
    ctx->err = "debug_read: unknown tap";
    return MP3GPU_E_INVALID;
}

extern "C" int mp3gpu_device_info(mp3gpu_ctx *ctx, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor) {
    if (!ctx) return MP3GPU_E_INVALID;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, ctx->device));
    if (name && name_len) {
        strncpy(name, p.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return MP3GPU_OK;
}

extern "C" int mp3gpu_device_pci_bus_id(mp3gpu_ctx *ctx, char *out, size_t out_len) {
    if (!ctx || !out || out_len < 16) return MP3GPU_E_INVALID;
    CK(cudaDeviceGetPCIBusId(out, (int)out_len, ctx->device));
    return MP3GPU_OK;
}

extern "C" int mp3gpu_pcm_to_f32_planar(mp3gpu_ctx *ctx, const int16_t *d_pcm, size_t n_samples, float *d_left, float *d_right) {
    if (!ctx) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (n_samples == 0) return MP3GPU_OK;
    if (!d_pcm || !d_left || !d_right || ((uintptr_t)d_pcm & 15) || ((uintptr_t)d_left & 15) || ((uintptr_t)d_right & 15)) {
        ctx->err = "pcm_to_f32_planar: device pointers must be non-null and 16-byte aligned";
        return MP3GPU_E_INVALID;
    }
    const size_t n4 = n_samples / 4;
    const int grid = (int)std::min<size_t>((n4 + 255) / 256 + 1, (size_t)ctx->sm_count * 8);
    k_pcm_to_f32_planar<<<grid, 256, 0, ctx->s_compute>>>(reinterpret_cast<const uint4 *>(d_pcm), n4, reinterpret_cast<float4 *>(d_left),
                                                          reinterpret_cast<float4 *>(d_right), d_pcm, n_samples, d_left, d_right);
    CK(cudaGetLastError());
    return MP3GPU_OK;
}

extern "C" int mp3gpu_measure_d2h(mp3gpu_ctx *ctx, void *host_dst, size_t bytes, int reps, double *seconds) {
    if (!ctx || !host_dst || !seconds || bytes == 0 || reps < 1) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    void *d = nullptr;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemsetAsync(d, 0x5a, bytes, ctx->s_out));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaMemcpyAsync(host_dst, d, bytes, cudaMemcpyDeviceToHost, ctx->s_out));  // warm-up: first touch of the host pages
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaEventRecord(a, ctx->s_out));
    for (int r = 0; r < reps; r++) CK(cudaMemcpyAsync(host_dst, d, bytes, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaEventRecord(b, ctx->s_out));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *seconds = (double)ms * 1e-3;
    return MP3GPU_OK;
}

extern "C" int mp3gpu_measure_fp32_peak(mp3gpu_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return MP3GPU_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, ctx->device));
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    float *d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a, ctx->s_compute));
        k_fp32_peak<<<blocks, threads, 0, ctx->s_compute>>>(d, iters);
        CK(cudaEventRecord(b, ctx->s_compute));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        double flops = (double)blocks * threads * iters * 16.0 * 8.0 * 2.0;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return MP3GPU_OK;
}
