// mp3host.cc — host-side mirror of go-mp3's `package mp3` above the mp3gpu C ABI (include/mp3host.h).
//
// What lives here is what the north star keeps on the host: tag skipping, header sync, side info,
// reservoir resolution (stream_parser.h), the Decoder's Read/Seek/time semantics (decode.go:45-388)
// and the DecodeBatch entry point.  Every PCM byte comes from the device engine (libmp3gpu.so,
// loaded with dlopen from this library's own directory); there is no CPU decode path.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/mp3host.h"
#include "lameinfo.h"
#include "stream_parser.h"

using namespace mp3host;

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// The device engine's entry points, resolved at mp3_engine_create.
struct GpuApi {
    void *handle = nullptr;
    int (*create)(int, const mp3gpu_opts *, mp3gpu_ctx **) = nullptr;
    void (*destroy)(mp3gpu_ctx *) = nullptr;
    const char *(*last_error)(const mp3gpu_ctx *) = nullptr;
    int (*decode)(mp3gpu_ctx *, const uint8_t *, size_t, const mp3gpu_unit *, size_t, int16_t *) = nullptr;
    int (*decode_range)(mp3gpu_ctx *, const uint8_t *, size_t, const mp3gpu_unit *, size_t, size_t, int16_t *) = nullptr;
    void *(*host_alloc)(size_t) = nullptr;
    void (*host_free)(void *) = nullptr;
    int (*last_timings)(mp3gpu_ctx *, mp3gpu_timings *) = nullptr;
};

std::string self_dir() {
    Dl_info info;
    if (dladdr((void *)&now_s, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t k = p.rfind('/');
        return k == std::string::npos ? std::string(".") : p.substr(0, k);
    }
    return ".";
}

bool load_gpu_api(GpuApi &api, unsigned which, std::string &err) {
    std::string path = self_dir() + (which == 1 ? "/libmp3gpu_exact.so" : (which == 2 ? "/libmp3gpu_checked.so" : "/libmp3gpu.so"));
    api.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) {
        err = std::string("dlopen ") + path + ": " + dlerror();
        return false;
    }
#define SYM(field, name)                                                    \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
    if (!api.field) {                                                       \
        err = std::string("missing symbol ") + name;                        \
        return false;                                                       \
    }
    SYM(create, "mp3gpu_create")
    SYM(destroy, "mp3gpu_destroy")
    SYM(last_error, "mp3gpu_last_error")
    SYM(decode, "mp3gpu_decode")
    SYM(decode_range, "mp3gpu_decode_range")
    SYM(host_alloc, "mp3gpu_host_alloc")
    SYM(host_free, "mp3gpu_host_free")
    SYM(last_timings, "mp3gpu_last_timings")
#undef SYM
    return true;
}

// A grow-only pinned host arena.
struct Pinned {
    void *p = nullptr;
    size_t cap = 0;
};

int hw_threads(int want) {
    if (want > 0) return want;
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// Parse n streams on `threads` threads, each stream into its own ParsedStream (bit_base 0).
void parse_all(const uint8_t *const *data, const size_t *lens, size_t n, int threads, std::vector<ParsedStream> &out) {
    out.resize(n);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            parse_whole_stream(data[i], lens[i], out[i], 0);
        }
    };
    int nt = (int)std::min<size_t>((size_t)hw_threads(threads), n ? n : 1);
    if (nt <= 1) {
        work();
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work);
    for (auto &t : th) t.join();
}

struct BatchLayout {
    std::vector<size_t> m_off, u_off;  // per stream: byte offset into main_data, unit offset
    size_t m_total = 0, u_total = 0;
};

// Streams are laid back to back; every stream's M starts on a 4-byte boundary so that the
// device's aligned 32-bit loads never straddle into a neighbour in a way that matters (reads are
// masked at the window end anyway) and offsets stay cheap to compute.
BatchLayout layout_batch(const std::vector<ParsedStream> &ps) {
    BatchLayout L;
    L.m_off.resize(ps.size());
    L.u_off.resize(ps.size());
    for (size_t i = 0; i < ps.size(); i++) {
        L.m_off[i] = L.m_total;
        L.u_off[i] = L.u_total;
        L.m_total += (ps[i].main_data.size() + 3) & ~size_t(3);
        L.u_total += ps[i].units.size();
    }
    return L;
}

void gather_batch(const std::vector<ParsedStream> &ps, const BatchLayout &L, uint8_t *main_data, mp3gpu_unit *units,
                  mp3_stream_result *res, int threads) {
    std::atomic<size_t> next{0};
    const size_t n = ps.size();
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            const ParsedStream &s = ps[i];
            uint8_t *m = main_data + L.m_off[i];
            if (!s.main_data.empty()) memcpy(m, s.main_data.data(), s.main_data.size());
            size_t padded = (s.main_data.size() + 3) & ~size_t(3);
            for (size_t k = s.main_data.size(); k < padded; k++) m[k] = 0;
            mp3gpu_unit *u = units + L.u_off[i];
            const uint64_t bit_base = (uint64_t)L.m_off[i] * 8;
            for (size_t k = 0; k < s.units.size(); k++) {
                u[k] = s.units[k];
                u[k].bit_start += bit_base;
            }
            res[i].pcm_offset = (int64_t)(L.u_off[i] / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
            res[i].pcm_bytes = (int64_t)(s.units.size() / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
            res[i].sample_rate = s.sample_rate;
            res[i].status = s.status;
            res[i].frames = s.frames;
        }
    };
    int nt = (int)std::min<size_t>((size_t)hw_threads(threads), n ? n : 1);
    if (nt <= 1) {
        work();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    memset(main_data + L.m_total, 0, 64);
}

}  // namespace

// One device of the engine: its device engine (libmp3gpu context) and the pinned staging arenas of calls that target
// this device alone (mp3_decode_frames, the streaming Decoder).
struct DeviceSlot {
    int device = 0;
    mp3gpu_ctx *gpu = nullptr;
    mp3gpu_ctx *gpu_b = nullptr;  // second device engine of the same GPU, created by the first chunked DecodeBatch: two calls in
                                  // flight hide each call's exposed head (first wave's upload + kernels) and tail (last wave's
                                  // download) behind the other's copies
    Pinned r_main, r_units;  // frame-range jobs
    std::mutex mu, mu_b;     // one call at a time per device engine (a device engine is single-owner)
};

struct mp3_engine {
    mp3_engine_opts opts{};
    GpuApi api;
    std::vector<DeviceSlot *> devs;
    mp3gpu_ctx *gpu = nullptr;  // devs[0]->gpu
    std::string err;
    std::mutex err_mu;
    Pinned a_main, a_units, a_pcm;  // DecodeBatch / stream-split arenas, shared by the devices (each works in its own region)

    void set_err(const std::string &m) {
        std::lock_guard<std::mutex> lk(err_mu);
        err = m;
    }
    int ensure(Pinned &a, size_t bytes) {
        if (a.cap >= bytes) return MP3_OK;
        if (a.p) api.host_free(a.p);
        a.p = nullptr;
        a.cap = 0;
        size_t want = bytes + bytes / 16 + 4096;
        a.p = api.host_alloc(want);
        if (!a.p) {
            set_err("pinned host allocation of " + std::to_string(want) + " bytes failed");
            return MP3_ERR_DEVICE;
        }
        a.cap = want;
        return MP3_OK;
    }
};

extern "C" int mp3_engine_create(const mp3_engine_opts *opts, mp3_engine **out) {
    if (!out) return MP3_ERR_INVALID;
    *out = nullptr;
    mp3_engine *e = new mp3_engine();
    if (opts) e->opts = *opts;
    if (e->opts.n_devices < 0 || e->opts.n_devices > MP3_MAX_DEVICES) {
        delete e;
        return MP3_ERR_INVALID;
    }
    if (!load_gpu_api(e->api, e->opts.use_exact_library, e->err)) {
        fprintf(stderr, "mp3_engine_create: %s\n", e->err.c_str());
        delete e;
        return MP3_ERR_DEVICE;
    }
    mp3gpu_opts go{};
    go.abi_version = MP3GPU_ABI_VERSION;
    go.wave_granules = e->opts.wave_granules;
    go.keep_intermediates = e->opts.keep_intermediates;
    const int n = e->opts.n_devices > 0 ? e->opts.n_devices : 1;
    for (int i = 0; i < n; i++) {
        DeviceSlot *d = new DeviceSlot();
        d->device = e->opts.n_devices > 0 ? e->opts.devices[i] : e->opts.device;
        int rc = e->api.create(d->device, &go, &d->gpu);
        if (rc != MP3GPU_OK) {
            fprintf(stderr, "mp3_engine_create: mp3gpu_create on device %d failed (%d): no such CUDA device, and there is no CPU decode path\n",
                    d->device, rc);
            delete d;
            mp3_engine_destroy(e);
            return MP3_ERR_DEVICE;
        }
        e->devs.push_back(d);
    }
    e->gpu = e->devs[0]->gpu;
    *out = e;
    return MP3_OK;
}

extern "C" void mp3_engine_destroy(mp3_engine *e) {
    if (!e) return;
    if (e->a_main.p) e->api.host_free(e->a_main.p);
    if (e->a_units.p) e->api.host_free(e->a_units.p);
    if (e->a_pcm.p) e->api.host_free(e->a_pcm.p);
    for (DeviceSlot *d : e->devs) {
        if (d->r_main.p) e->api.host_free(d->r_main.p);
        if (d->r_units.p) e->api.host_free(d->r_units.p);
        if (d->gpu) e->api.destroy(d->gpu);
        if (d->gpu_b) e->api.destroy(d->gpu_b);
        delete d;
    }
    if (e->api.handle) dlclose(e->api.handle);
    delete e;
}

extern "C" const char *mp3_engine_last_error(const mp3_engine *e) { return e ? e->err.c_str() : "no engine"; }
extern "C" mp3gpu_ctx *mp3_engine_gpu(mp3_engine *e) { return e ? e->gpu : nullptr; }
extern "C" int mp3_engine_device_count(const mp3_engine *e) { return e ? (int)e->devs.size() : 0; }
extern "C" mp3gpu_ctx *mp3_engine_gpu_at(mp3_engine *e, int slot) {
    return (e && slot >= 0 && (size_t)slot < e->devs.size()) ? e->devs[(size_t)slot]->gpu : nullptr;
}

extern "C" const char *mp3_error_string(int code) {
    switch (code) {
    case MP3_OK: return "";
    case MP3_EOF: return "EOF";
    case MP3_ERR_UNEXPECTED_EOF: return "mp3: unexpected EOF";
    case MP3_ERR_SYNC_LIMIT: return "mp3: no valid frame header found within 65536 bytes";
    case MP3_ERR_FREE_FORMAT: return "mp3: free bitrate format is not supported";
    case MP3_ERR_MPEG25: return "mp3: MPEG version 2.5 is not supported";
    case MP3_ERR_LAYER: return "mp3: only layer3 (want 1; got other) is supported";
    case MP3_ERR_FRAMESIZE: return "mp3: framesize too large";
    case MP3_ERR_MAINDATA_SIZE: return "mp3: main data size too large";
    case MP3_ERR_ISPOS: return "mp3: isPos was too big: 576";
    case MP3_ERR_SEEK_UNSUPPORTED: return "mp3: seek not supported on non-seekable source";
    case MP3_ERR_WHENCE: return "mp3: invalid whence";
    case MP3_ERR_REF_PANIC: return "mp3: input on which the reference decoder panics";
    case MP3_ERR_NO_XING_HEADER: return "lameinfo: no Xing/Info header found";
    case MP3_ERR_DEVICE: return "mp3: device engine failure";
    case MP3_ERR_INVALID: return "mp3: invalid argument";
    }
    return "mp3: unknown error";
}

// ---- lameinfo (lameinfo.h) ---------------------------------------------------------------------------
extern "C" int mp3_lameinfo_parse(const uint8_t *frame, size_t len, mp3_lame_info *out) {
    if (!out || (!frame && len)) return MP3_ERR_INVALID;
    return lameinfo_parse(frame, len, out);
}
extern "C" int mp3_lameinfo_parse_from_reader(const uint8_t *data, size_t len, mp3_lame_info *out) {
    if (!out || (!data && len)) return MP3_ERR_INVALID;
    return lameinfo_parse_from_reader(data, len, out);
}
extern "C" int mp3_lameinfo_total_delay(const mp3_lame_info *info) { return info ? lameinfo_total_delay(info) : MP3_LAME_DECODER_DELAY; }
extern "C" int mp3_lameinfo_total_padding(const mp3_lame_info *info) { return info ? lameinfo_total_padding(info) : 0; }
extern "C" int mp3_lameinfo_is_lame_version(const uint8_t *s, size_t n) { return is_lame_version(s, n) ? 1 : 0; }
extern "C" int64_t mp3_lameinfo_toc_offset(const mp3_lame_info *info, double fraction, uint64_t stream_bytes) {
    return info ? lameinfo_toc_offset(info, fraction, stream_bytes) : -1;
}

// ---- host-only parse (tests, bench staging) ---------------------------------------------------------
extern "C" int mp3_parse_streams(const uint8_t *const *data, const size_t *lens, size_t n, int host_threads, mp3_parsed **out) {
    if (!out || (n && (!data || !lens))) return MP3_ERR_INVALID;
    std::vector<ParsedStream> ps;
    parse_all(data, lens, n, host_threads, ps);
    BatchLayout L = layout_batch(ps);
    mp3_parsed *p = (mp3_parsed *)calloc(1, sizeof(mp3_parsed));
    p->main_data = (uint8_t *)aligned_alloc(64, (L.m_total + 64 + 63) & ~size_t(63));
    p->units = (mp3gpu_unit *)malloc(sizeof(mp3gpu_unit) * (L.u_total ? L.u_total : 1));
    p->streams = (mp3_stream_result *)calloc(n ? n : 1, sizeof(mp3_stream_result));
    if (!p->main_data || !p->units || !p->streams) {
        mp3_parsed_free(p);
        return MP3_ERR_INVALID;
    }
    gather_batch(ps, L, p->main_data, p->units, p->streams, host_threads);
    p->main_data_len = L.m_total;
    p->n_granules = L.u_total / 2;
    p->n_streams = n;
    *out = p;
    return MP3_OK;
}

extern "C" void mp3_parsed_free(mp3_parsed *p) {
    if (!p) return;
    free(p->main_data);
    free(p->units);
    free(p->streams);
    free(p);
}

// ---- DecodeBatch -----------------------------------------------------------------------------------
namespace {

// Upper bound of the unit slots a stream will produce: the frame walk of parse_whole_stream without side info
// and main data (a frame that later fails to parse only makes the real count smaller).
struct StreamBound {
    size_t slots = 0;  // unit slots (2 per granule)
    size_t m_bytes = 0;  // main-data bytes (frame size - header - CRC - side info of every walked frame)
};
StreamBound stream_upper_bound(const uint8_t *data, size_t len) {
    Source s;
    s.data = data;
    s.len = len;
    StreamBound b;
    if (s.skip_tags() != MP3_OK) return b;
    for (;;) {
        Header h;
        int64_t fpos;
        if (read_frame_header(s, &h, &fpos) != MP3_OK) break;
        if (h.id() == 0 || h.layer() != 1) break;
        const int fs = h.frame_size();
        if (fs > 2000 || fs < 4 || s.pos + (fs - 4) > (int64_t)len) break;
        s.pos += fs - 4;
        b.slots += (size_t)h.granules() * 2;
        const int md = fs - 4 - h.side_info_size() - (h.protection_bit() == 0 ? 2 : 0);
        if (md > 0) b.m_bytes += (size_t)md;
    }
    return b;
}
size_t unit_slots_upper_bound(const uint8_t *data, size_t len) { return stream_upper_bound(data, len).slots; }

struct DeviceJob {
    const uint8_t *main_data;
    size_t main_len;
    const mp3gpu_unit *units;
    size_t n_granules;
    int16_t *pcm;
};

}  // namespace

// Test hook (tests/test_host_and_emulation.py): the arena bound DecodeBatch relies on.
extern "C" size_t mp3_debug_unit_slots_upper_bound(const uint8_t *data, size_t len) { return unit_slots_upper_bound(data, len); }
extern "C" size_t mp3_debug_main_bytes_upper_bound(const uint8_t *data, size_t len) { return stream_upper_bound(data, len).m_bytes; }

namespace {

// A chunk = streams [i0, i1) of a batch with its fixed place in the shared pinned arenas (computed from the streams' bounds,
// so the place does not depend on which device ends up decoding the chunk).
struct Chunk {
    size_t i0 = 0, i1 = 0;
    size_t m_off = 0, m_total = 0;  // bytes into a_main (m_total excludes the 64 bytes of zero padding that follow)
    size_t u_off = 0, u_total = 0;  // unit slots into a_units (PCM: slot / 2 * 2304 bytes into a_pcm)
};
// What one device did in a DecodeBatch call.
struct Shard {
    double parse_s = 0, gather_s = 0, device_s = 0;
    size_t m_used = 0, u_used = 0, chunks = 0;
    int rc = MP3_OK;
    std::string err;
};

size_t shard_chunks(size_t n) {
    static const int forced = [] { const char *e = getenv("MP3HOST_CHUNKS"); return e ? atoi(e) : 0; }();  // experiments
    if (forced > 0) return std::min<size_t>((size_t)forced, n ? n : 1);
    return n >= 256 ? std::min<size_t>(32, n / 32) : 1;  // measured (1,024 x 30 s streams): 8 / 16 / 32 chunks = 0.83 / 0.86 / 0.90 of the plain-copy ceiling
}

// One device's share of a DecodeBatch call.  The batch is cut into chunks of streams and the devices PULL chunks from a
// common counter (a device behind a slower PCIe link simply takes fewer): while the device decodes chunk k (the PCIe-bound
// part), the host threads parse chunk k+1 straight behind it into the pinned arenas.
void decode_shard(mp3_engine *e, DeviceSlot *dev, const uint8_t *const *data, const size_t *lens, const StreamBound *bounds,
                  mp3_stream_result *results, const std::vector<Chunk> &chunks, std::atomic<size_t> &next_chunk, Shard &S, int threads) {
    const size_t n_chunks = chunks.size();
    if (n_chunks == 0) return;
    int dev_rc = MP3GPU_OK;
    std::string dev_err;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<DeviceJob> jobs;
    bool closed = false;
    if (n_chunks > 1 && !dev->gpu_b && !e->opts.keep_intermediates) {
        mp3gpu_opts go{};
        go.abi_version = MP3GPU_ABI_VERSION;
        go.wave_granules = e->opts.wave_granules;
        if (e->api.create(dev->device, &go, &dev->gpu_b) != MP3GPU_OK) dev->gpu_b = nullptr;  // one engine still works
    }
    std::mutex rc_mu;
    auto run_job = [&](const DeviceJob &j, int which) {
        if (j.n_granules == 0) return;
        {
            std::lock_guard<std::mutex> lk(rc_mu);
            if (dev_rc != MP3GPU_OK) return;
        }
        mp3gpu_ctx *ctx = which ? dev->gpu_b : dev->gpu;
        const double a = now_s();
        int rc;
        std::string msg;
        {
            std::lock_guard<std::mutex> lk(which ? dev->mu_b : dev->mu);
            rc = e->api.decode(ctx, j.main_data, j.main_len, j.units, j.n_granules, j.pcm);
            if (rc != MP3GPU_OK) msg = e->api.last_error(ctx);
        }
        std::lock_guard<std::mutex> lk(rc_mu);
        S.device_s += now_s() - a;
        if (rc != MP3GPU_OK && dev_rc == MP3GPU_OK) {
            dev_rc = rc;
            dev_err = msg;
        }
    };
    std::vector<std::thread> workers;
    if (n_chunks > 1)
        for (int w = 0; w < (dev->gpu_b ? 2 : 1); w++)
            workers.emplace_back([&, w]() {  // a worker owns one device engine for the duration of the shard
                for (;;) {
                    DeviceJob j;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return closed || !jobs.empty(); });
                        if (jobs.empty()) return;
                        j = jobs.front();
                        jobs.pop_front();
                    }
                    cv.notify_all();  // the producer may pull the next chunk
                    run_job(j, w);
                }
            });
    // Every stream of a chunk is parsed STRAIGHT into the pinned arenas, at the place its bounds give it: stream i's main
    // data starts where the bounds of the streams in front of it end (4-byte aligned), its units likewise.  For well-formed
    // streams the bounds are exact, so the streams lie back to back as the device kernels like them; a stream that ends
    // early (truncated, garbage) leaves unit slots that are marked invalid and a hole in the PCM that its result does not
    // cover.  No per-stream vectors, no gather copy.
    std::vector<size_t> m_at, u_at;
    for (;;) {
        if (n_chunks > 1) {  // do not pull chunks faster than this device decodes them: at most two parsed chunks waiting
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return jobs.size() < 2; });
        }
        const size_t c = next_chunk.fetch_add(1);
        if (c >= n_chunks) break;
        const Chunk &K = chunks[c];
        const size_t i0 = K.i0, i1 = K.i1;
        const double ta = now_s();
        const size_t nc = i1 - i0;
        m_at.assign(nc + 1, 0);
        u_at.assign(nc + 1, 0);
        for (size_t k = 0; k < nc; k++) {
            m_at[k + 1] = m_at[k] + ((bounds[i0 + k].m_bytes + 3) & ~size_t(3));
            u_at[k + 1] = u_at[k] + bounds[i0 + k].slots;
        }
        const size_t m_total = m_at[nc], u_total = u_at[nc];
        if (m_total != K.m_total || u_total != K.u_total) {
            S.rc = MP3_ERR_INVALID;  // cannot happen: the chunk table was built from the same bounds
            S.err = "DecodeBatch chunk table mismatch";
            break;
        }
        uint8_t *m_dst = (uint8_t *)e->a_main.p + K.m_off;
        mp3gpu_unit *u_dst = (mp3gpu_unit *)e->a_units.p + K.u_off;
        const int64_t pcm_off = (int64_t)(K.u_off / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
        std::atomic<size_t> next{0};
        std::atomic<bool> overflow{false};
        auto work = [&]() {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= nc) break;
                const size_t i = i0 + k;
                SpanVec<uint8_t> M;
                M.p = m_dst + m_at[k];
                M.cap = bounds[i].m_bytes;
                SpanVec<mp3gpu_unit> U;
                U.p = u_dst + u_at[k];
                U.cap = bounds[i].slots;
                StreamMeta meta;
                parse_whole_stream_into(data[i], lens[i], M, U, meta, (int64_t)m_at[k] * 8);
                if (M.overflow || U.overflow) overflow = true;
                for (size_t q = M.n; q < m_at[k + 1] - m_at[k]; q++) M.p[q] = 0;      // up to the next stream's start
                if (U.n < U.cap) memset(U.p + U.n, 0, (U.cap - U.n) * sizeof(mp3gpu_unit));  // invalid slots: no VALID bit
                results[i].pcm_offset = pcm_off + (int64_t)(u_at[k] / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
                results[i].pcm_bytes = (int64_t)(U.n / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
                results[i].sample_rate = meta.sample_rate;
                results[i].status = meta.status;
                results[i].frames = meta.frames;
            }
        };
        const int nt = (int)std::min<size_t>((size_t)hw_threads(threads), nc ? nc : 1);
        if (nt <= 1) {
            work();
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work);
            for (auto &t : th) t.join();
        }
        memset(m_dst + m_total, 0, 64);
        S.parse_s += now_s() - ta;
        if (overflow) {
            S.rc = MP3_ERR_INVALID;  // cannot happen (the bounds are bounds: tests/test_host_and_emulation.py); refuse rather than decode garbage
            S.err = "DecodeBatch stream bound exceeded";
            break;
        }
        S.m_used += m_total;
        S.u_used += u_total;
        S.chunks++;
        DeviceJob j{m_dst, m_total, u_dst, u_total / 2, (int16_t *)((uint8_t *)e->a_pcm.p + pcm_off)};
        if (n_chunks == 1) {
            run_job(j, 0);
        } else {
            {
                std::lock_guard<std::mutex> lk(mu);
                jobs.push_back(j);
            }
            cv.notify_all();
        }
    }
    if (!workers.empty()) {
        {
            std::lock_guard<std::mutex> lk(mu);
            closed = true;
        }
        cv.notify_all();
        for (auto &t : workers) t.join();
    }
    if (S.rc == MP3_OK && dev_rc != MP3GPU_OK) {
        S.rc = MP3_ERR_DEVICE;
        S.err = dev_err;
    }
}

// The reference's README example (README.md:110-195): skip TotalDelay() stereo samples at the start and TotalPadding()
// at the end when the first frame (behind the tags) carries a LAME tag.
void trim_gapless(const uint8_t *data, size_t len, mp3_stream_result &r) {
    if (r.pcm_bytes <= 0) return;
    Source s;
    s.data = data;
    s.len = len;
    if (s.skip_tags() != MP3_OK) return;
    mp3_lame_info info;
    if (lameinfo_parse_from_reader(data + s.pos, len - (size_t)s.pos, &info) != MP3_OK || !info.has_lame_info) return;
    const int64_t skip = (int64_t)lameinfo_total_delay(&info) * 4, trim = (int64_t)lameinfo_total_padding(&info) * 4;
    if (skip + trim > r.pcm_bytes) return;
    r.pcm_offset += skip;
    r.pcm_bytes -= skip + trim;
}

}  // namespace

// DecodeBatch.  The batch is cut into chunks of streams of about equal input bytes; the engine's devices pull chunks from a
// common counter (streams share nothing: no collective).  Every device runs the parse -> device pipeline above
// on its own host thread, with its share of the parsing threads; a chunk's place in the pinned arenas is fixed by the bounds.
extern "C" int mp3_decode_batch(mp3_engine *e, const uint8_t *const *data, const size_t *lens, size_t n,
                                mp3_stream_result *results, const uint8_t **pcm_base, mp3_batch_timings *timings) {
    if (!e || !results || !pcm_base || (n && (!data || !lens))) return MP3_ERR_INVALID;
    const double t0 = now_s();
    const size_t D = e->devs.size();
    const int threads = hw_threads(e->opts.host_threads);
    // ---- arena bounds from a header-only frame walk: unit slots and main-data bytes per stream (exact for well-formed streams) ----
    std::vector<StreamBound> ub(n);
    {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                ub[i] = stream_upper_bound(data[i], lens[i]);
            }
        };
        int nt = (int)std::min<size_t>((size_t)threads, n ? n : 1);
        if (nt > 1 && n >= 64) {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work);
            for (auto &t : th) t.join();
        } else {
            work();
        }
    }
    // ---- chunks: contiguous blocks of streams of about equal input bytes, each with its place in the arenas ----
    std::vector<Chunk> chunks;
    size_t u_end = 0;
    {
        size_t total = 0;
        for (size_t i = 0; i < n; i++) total += lens[i];
        size_t nck = D == 1 ? shard_chunks(n) : (n >= 256 * D ? std::min<size_t>(32 * D, n / 32) : std::min<size_t>(D, n));
        if (nck < 1) nck = 1;
        size_t i = 0, acc = 0, m_off = 0, u_off = 0;
        for (size_t c = 0; c < nck; c++) {
            Chunk K;
            K.i0 = i;
            const size_t target = (size_t)((double)total * (double)(c + 1) / (double)nck);
            while (i < n && (c + 1 == nck || acc + lens[i] / 2 < target)) acc += lens[i++];
            K.i1 = i;
            for (size_t k = K.i0; k < K.i1; k++) {
                K.m_total += (ub[k].m_bytes + 3) & ~size_t(3);
                K.u_total += ub[k].slots;
            }
            K.m_off = m_off;
            K.u_off = u_off;
            m_off += (K.m_total + 64 + 63) & ~size_t(63);  // 64 bytes of zero padding behind every chunk's main data
            u_off += K.u_total;
            if (K.i1 > K.i0) chunks.push_back(K);
        }
        u_end = u_off;
        int rc = e->ensure(e->a_main, m_off + 64);
        if (rc == MP3_OK) rc = e->ensure(e->a_units, sizeof(mp3gpu_unit) * u_off + 64);
        if (rc == MP3_OK) rc = e->ensure(e->a_pcm, (u_off / 2) * MP3GPU_PCM_BYTES_PER_GRANULE + 64);
        if (rc != MP3_OK) return rc;
    }
    std::vector<Shard> shards(D);
    std::atomic<size_t> next_chunk{0};
    if (D == 1 || chunks.size() <= 1) {
        decode_shard(e, e->devs[0], data, lens, ub.data(), results, chunks, next_chunk, shards[0], threads);
    } else {
        const int per = std::max(1, threads / (int)D);
        std::vector<std::thread> th;
        for (size_t d = 0; d < D; d++)
            th.emplace_back([&, d]() { decode_shard(e, e->devs[d], data, lens, ub.data(), results, chunks, next_chunk, shards[d], per); });
        for (auto &t : th) t.join();
    }
    const double t3 = now_s();
    for (size_t d = 0; d < D; d++)
        if (shards[d].rc != MP3_OK) {
            e->set_err(shards[d].err);
            return shards[d].rc;
        }
    if (e->opts.trim_gapless)
        for (size_t i = 0; i < n; i++) trim_gapless(data[i], lens[i], results[i]);
    *pcm_base = (const uint8_t *)e->a_pcm.p;
    if (timings) {
        memset(timings, 0, sizeof *timings);
        size_t granules = 0;
        for (size_t d = 0; d < D; d++) {  // the devices run side by side: the slowest one's sums
            timings->parse_s = std::max(timings->parse_s, shards[d].parse_s);   // summed over the chunks; chunks after the first overlap the device call
            timings->gather_s = std::max(timings->gather_s, shards[d].gather_s);
            timings->device_s = std::max(timings->device_s, shards[d].device_s);  // summed device-call time (worker thread)
            timings->main_data_bytes += shards[d].m_used;
            granules += shards[d].u_used / 2;
        }
        timings->total_s = t3 - t0;
        timings->n_granules = granules;
        // the span of the PCM buffer that holds results (regions of devices are laid out by their bounds: for well-formed
        // streams the bound is exact and the span is dense)
        timings->pcm_bytes = (uint64_t)(u_end / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
    }
    return MP3_OK;
}

// ---- Decoder ---------------------------------------------------------------------------------------
// Mirrors *mp3.Decoder.  The reference decodes one frame per readFrame; here readFrame's host half
// runs ahead by up to `chunk_frames` frames and the whole chunk goes to the GPU in one call.  What
// the caller can observe is unchanged: the same bytes, the same error at the same byte position,
// the same source position after Seek-to-end (tracked per buffered frame).
struct mp3_decoder {
    mp3_engine *eng = nullptr;
    DeviceSlot *dev = nullptr;
    StreamParser parser;
    bool seekable = false;
    int sample_rate = 0;
    int64_t length = -1;           // invalidLength (decode.go:218)
    std::vector<int64_t> frame_starts;
    int64_t bytes_per_frame = 0;
    int64_t pos = 0;               // d.pos

    // d.buf: decoded PCM not yet handed out, with the source position after each buffered frame
    std::vector<uint8_t> buf;
    size_t buf_off = 0;
    struct FrameMark { size_t pcm_end; int64_t src_end; };
    std::vector<FrameMark> marks;  // frames whose PCM is in buf (pcm_end = offset in buf one past the frame)
    int64_t ref_src_pos = 0;       // where the reference's source would stand (after the last frame *it* has read)
    int pending = MP3_OK;          // terminal status of the chunk, delivered once the buffer drains
    size_t eager_frames = 1;       // frames the reference reads before any byte is asked for (1 at open / refill, 2 in Seek)
    std::vector<int> frame_slots;  // unit slots per frame of `units` (halo frames first)

    // Where the reference's source stands: it reads a frame only when d.buf is empty, so it has read the
    // eager frames plus every buffered frame of which at least one byte was delivered.
    int64_t reference_source_pos() const {
        if (marks.empty()) return parser.src.pos;
        size_t nread = 0, prev_end = 0;
        for (const auto &m : marks) {
            if (buf_off > prev_end) nread++;
            prev_end = m.pcm_end;
        }
        nread = std::max(nread, std::min(eager_frames, marks.size()));
        return marks[nread - 1].src_end;
    }

    // chunk state: M window and units, with a halo of the last two granules kept for the next chunk
    std::vector<uint8_t> M;
    std::vector<mp3gpu_unit> units;   // halo units first, then the chunk's
    size_t halo_units = 0;
    std::vector<int16_t> pcm_tmp;

    void drop_state() {  // d.frame = nil
        parser.reset_state();
        M.clear();
        parser.m_base = 0;
        units.clear();
        frame_slots.clear();
        halo_units = 0;
    }
    void clear_buf() {
        buf.clear();
        buf_off = 0;
        marks.clear();
    }

    // Parse up to max_frames frames and decode them.  Returns MP3_OK if at least one frame was
    // appended to buf; otherwise the readFrame error (EOF family already mapped to MP3_EOF).
    int fill(int max_frames) {
        // compact buf
        if (buf_off == buf.size()) clear_buf();
        int got = 0, rc = MP3_OK;
        std::vector<int64_t> src_ends;
        std::vector<int> frame_granules;
        eager_frames = 1;
        while (got < max_frames) {
            size_t u_before = units.size();
            rc = parser.next_frame(M, units, 0);
            if (rc != MP3_OK) break;
            src_ends.push_back(parser.src.pos);
            frame_granules.push_back((int)((units.size() - u_before) / 2));
            frame_slots.push_back((int)(units.size() - u_before));
            got++;
        }
        if (rc != MP3_OK) {
            // d.frame = nil on any error (decode.go:47); the EOF family becomes io.EOF (decode.go:48-63)
            parser.reset_state();
            if (rc == MP3_ERR_UNEXPECTED_EOF || rc == MP3_ERR_SYNC_LIMIT) rc = MP3_EOF;
        }
        if (got > 0) {
            const size_t n_gr = units.size() / 2;
            // rebase bit positions to the retained window M (absolute byte m_base = M[0])
            std::vector<mp3gpu_unit> sub(units);
            const uint64_t base_bits = (uint64_t)parser.m_base * 8;
            for (auto &u : sub) u.bit_start -= base_bits;
            M.resize(M.size() + 64, 0);  // device reads whole words; keep the tail defined
            pcm_tmp.resize(n_gr * 1152);
            int grc;
            {
                std::lock_guard<std::mutex> lk(dev->mu);
                grc = eng->api.decode(dev->gpu, M.data(), M.size() - 64, sub.data(), n_gr, pcm_tmp.data());
                if (grc != MP3GPU_OK) eng->set_err(eng->api.last_error(dev->gpu));
            }
            M.resize(M.size() - 64);
            if (grc != MP3GPU_OK) return MP3_ERR_DEVICE;
            const uint8_t *pcm = reinterpret_cast<const uint8_t *>(pcm_tmp.data()) + (halo_units / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
            size_t off = 0;
            for (int f = 0; f < got; f++) {
                size_t nb = (size_t)frame_granules[f] * MP3GPU_PCM_BYTES_PER_GRANULE;
                buf.insert(buf.end(), pcm + off, pcm + off + nb);
                off += nb;
                marks.push_back({buf.size(), src_ends[f]});
            }
            if (rc == MP3_OK) {
                // keep whole frames covering the last two granules (and the bytes their windows reach back
                // to) as the next chunk's halo: MPEG-1 frames own 4 unit slots, LSF frames 2
                size_t keep_u = 0, keep_f = 0;
                while (keep_f < frame_slots.size() && keep_u < 4) keep_u += (size_t)frame_slots[frame_slots.size() - 1 - keep_f++];
                std::vector<mp3gpu_unit> halo(units.end() - keep_u, units.end());
                frame_slots.erase(frame_slots.begin(), frame_slots.end() - keep_f);
                uint64_t min_bit = UINT64_MAX;
                for (auto &u : halo)
                    if (u.w2 & MP3GPU_W2_VALID) min_bit = std::min(min_bit, u.bit_start);
                // the parser's own window (next frame's reservoir) must stay too
                min_bit = std::min<uint64_t>(min_bit, (uint64_t)parser.win_start * 8);
                int64_t keep_from = (int64_t)(min_bit / 8) & ~int64_t(3);
                if (keep_from > parser.m_base) {
                    M.erase(M.begin(), M.begin() + (keep_from - parser.m_base));
                    parser.m_base = keep_from;
                }
                units.swap(halo);
                halo_units = units.size();
            } else {
                M.clear();
                parser.m_base = 0;
                units.clear();
                frame_slots.clear();
                halo_units = 0;
            }
            pending = rc;  // delivered after the buffered PCM (MP3_OK: nothing pending)
            return MP3_OK;
        }
        M.clear();
        parser.m_base = 0;
        units.clear();
        frame_slots.clear();
        halo_units = 0;
        return rc;
    }

    // decode.go:45-67 as seen by the caller: make at least one more frame available.
    int read_frame(int max_frames) {
        if (pending != MP3_OK) {
            int rc = pending;
            pending = MP3_OK;
            return rc;
        }
        return fill(max_frames);
    }

    int chunk() const { return eng->opts.chunk_frames ? (int)eng->opts.chunk_frames : 256; }

    // decode.go:154-216
    int ensure_frame_starts_and_length() {
        if (length != -1) return MP3_OK;
        if (!seekable) return MP3_OK;
        Source s;
        s.data = parser.src.data;
        s.len = parser.src.len;
        s.pos = 0;
        int rc = s.skip_tags();
        if (rc != MP3_OK) return rc;
        int64_t l = 0;
        for (;;) {
            Header h;
            int64_t fpos;
            rc = read_frame_header(s, &h, &fpos);
            if (rc != MP3_OK) {
                if (rc == MP3_EOF || rc == MP3_ERR_UNEXPECTED_EOF || rc == MP3_ERR_SYNC_LIMIT) break;
                return rc;
            }
            frame_starts.push_back(fpos);
            bytes_per_frame = h.bytes_per_frame();
            l += bytes_per_frame;
            // source.Seek(framesize-4, io.SeekCurrent): bytes.Reader allows seeking past the end;
            // a negative resulting position is an error the reference returns.
            int64_t np = s.pos + (int64_t)(h.frame_size() - 4);
            if (np < 0) return MP3_ERR_INVALID;
            s.pos = np;
        }
        length = l;
        return MP3_OK;
    }
};

extern "C" mp3_decoder *mp3_new_decoder(mp3_engine *e, const uint8_t *data, size_t len, int seekable, int *err) {
    return mp3_new_decoder_on(e, 0, data, len, seekable, err);
}

extern "C" mp3_decoder *mp3_new_decoder_on(mp3_engine *e, int slot, const uint8_t *data, size_t len, int seekable, int *err) {
    int dummy;
    if (!err) err = &dummy;
    if (!e || (!data && len) || slot < 0 || (size_t)slot >= e->devs.size()) {
        *err = MP3_ERR_INVALID;
        return nullptr;
    }
    mp3_decoder *d = new mp3_decoder();
    d->eng = e;
    d->dev = e->devs[(size_t)slot];
    d->seekable = seekable != 0;
    d->parser.src.data = data;
    d->parser.src.len = len;
    int rc = d->parser.src.skip_tags();
    if (rc == MP3_OK) rc = d->read_frame(d->chunk());
    if (rc == MP3_OK) {
        // the sample rate comes from the first frame (decode.go:377-381)
        Source s = d->parser.src;
        s.pos = 0;
        s.skip_tags();
        Header h;
        int64_t fpos;
        if (read_frame_header(s, &h, &fpos) == MP3_OK) d->sample_rate = h.sampling_frequency_value();
        rc = d->ensure_frame_starts_and_length();
    }
    if (rc != MP3_OK) {
        *err = rc;
        delete d;
        return nullptr;
    }
    *err = MP3_OK;
    return d;
}

extern "C" void mp3_decoder_free(mp3_decoder *d) { delete d; }

extern "C" long mp3_decoder_read(mp3_decoder *d, uint8_t *out, size_t n, int *err) {
    int dummy;
    if (!err) err = &dummy;
    if (!d) {
        *err = MP3_ERR_INVALID;
        return 0;
    }
    while (d->buf.size() - d->buf_off == 0) {  // decode.go:71-75
        int rc = d->read_frame(d->chunk());
        if (rc != MP3_OK) {
            *err = rc;
            return 0;
        }
    }
    // d.buf of the reference never holds more than the rest of ONE frame (decode.go:65,76-77), so a Read returns
    // at most that; the decode-ahead buffer is cut at the same frame boundary.
    size_t live = d->buf.size() - d->buf_off;
    for (const auto &m : d->marks)
        if (m.pcm_end > d->buf_off) {
            live = m.pcm_end - d->buf_off;
            break;
        }
    size_t c = std::min(n, live);
    memcpy(out, d->buf.data() + d->buf_off, c);
    d->buf_off += c;
    d->pos += (int64_t)c;
    *err = MP3_OK;
    return (long)c;
}

extern "C" int64_t mp3_decoder_seek(mp3_decoder *d, int64_t offset, int whence, int *err) {
    int dummy;
    if (!err) err = &dummy;
    if (!d) {
        *err = MP3_ERR_INVALID;
        return 0;
    }
    *err = MP3_OK;
    if (offset == 0 && whence == 1) return d->pos;  // decode.go:90-93
    int64_t npos = 0;
    switch (whence) {
    case 0: npos = offset; break;
    case 1: npos = d->pos + offset; break;
    case 2: npos = d->length + offset; break;
    default: *err = MP3_ERR_WHENCE; return 0;
    }
    const int64_t ref_pos = d->reference_source_pos();
    d->pos = npos;
    d->clear_buf();
    d->pending = MP3_OK;
    d->drop_state();
    if (d->pos < 0) d->pos = 0;
    if (d->length != -1 && d->pos >= d->length) {
        d->parser.src.pos = ref_pos;  // the source is not touched (decode.go:115-118)
        return npos;
    }
    if (d->bytes_per_frame == 0 || d->frame_starts.empty()) {
        // Go: integer divide by zero / index out of range on a non-seekable source
        *err = MP3_ERR_SEEK_UNSUPPORTED;
        return 0;
    }
    int64_t f = d->pos / d->bytes_per_frame;
    if ((size_t)f >= d->frame_starts.size()) {
        *err = MP3_ERR_REF_PANIC;  // index out of range in the reference
        return 0;
    }
    int rc;
    if (f > 0) {
        f--;
        d->parser.src.pos = d->frame_starts[(size_t)f];
        // two readFrame calls (decode.go:123-133); the chunk decodes ahead from there
        rc = d->fill(std::max(2, d->chunk()));
        d->eager_frames = 2;
        if (rc == MP3_OK && d->marks.size() < 2) {
            // the second readFrame failed
            rc = d->pending != MP3_OK ? d->pending : MP3_EOF;
            d->pending = MP3_OK;
        }
        if (rc != MP3_OK) {
            d->clear_buf();
            *err = rc;
            return 0;
        }
        size_t drop = (size_t)(d->bytes_per_frame + (d->pos % d->bytes_per_frame));
        if (drop > d->marks[1].pcm_end) {
            *err = MP3_ERR_REF_PANIC;  // slice bounds out of range in the reference
            d->clear_buf();
            return 0;
        }
        d->buf_off = drop;
    } else {
        d->parser.src.pos = d->frame_starts[0];
        rc = d->fill(d->chunk());
        if (rc != MP3_OK) {
            *err = rc;
            return 0;
        }
        if ((size_t)d->pos > d->marks[0].pcm_end) {
            *err = MP3_ERR_REF_PANIC;
            d->clear_buf();
            return 0;
        }
        d->buf_off = (size_t)d->pos;
    }
    return npos;
}

extern "C" int mp3_decoder_sample_rate(const mp3_decoder *d) { return d->sample_rate; }
extern "C" int64_t mp3_decoder_length(const mp3_decoder *d) { return d->length; }
extern "C" int64_t mp3_decoder_bytes_per_frame(const mp3_decoder *d) { return d->bytes_per_frame; }
static int64_t bytes_to_duration(const mp3_decoder *d, int64_t bytes) {  // decode.go:344-348
    return (int64_t)1000000000 * bytes / (int64_t)(d->sample_rate * 4);
}
static int64_t duration_to_bytes(const mp3_decoder *d, int64_t dur) {  // decode.go:351-354
    return dur * (int64_t)(d->sample_rate * 4) / (int64_t)1000000000;
}
extern "C" int64_t mp3_decoder_duration_ns(const mp3_decoder *d) {
    if (d->length == -1) return -1;
    return bytes_to_duration(d, d->length);
}
extern "C" int64_t mp3_decoder_position_ns(const mp3_decoder *d) { return bytes_to_duration(d, d->pos); }
extern "C" int64_t mp3_decoder_remaining_ns(const mp3_decoder *d) {
    int64_t dur = mp3_decoder_duration_ns(d);
    if (dur < 0) return -1;
    return dur - mp3_decoder_position_ns(d);
}
extern "C" double mp3_decoder_progress(const mp3_decoder *d) {
    if (d->length == -1) return -1;
    if (d->length == 0) return 0;
    return (double)d->pos / (double)d->length;
}
extern "C" int64_t mp3_decoder_sample_position(const mp3_decoder *d) { return d->pos / 4; }
extern "C" int64_t mp3_decoder_sample_count(const mp3_decoder *d) {
    if (d->length == -1) return -1;
    return d->length / 4;
}
extern "C" int mp3_decoder_seek_to_sample(mp3_decoder *d, int64_t sample) {  // decode.go:288-307
    if (d->length == -1) return MP3_ERR_SEEK_UNSUPPORTED;
    if (sample < 0) sample = 0;
    int64_t max_samples = mp3_decoder_sample_count(d);
    if (sample > max_samples) sample = max_samples;
    int err;
    mp3_decoder_seek(d, sample * 4, 0, &err);
    return err;
}
extern "C" int mp3_decoder_seek_to_time(mp3_decoder *d, int64_t t) {  // decode.go:320-341
    if (d->length == -1) return MP3_ERR_SEEK_UNSUPPORTED;
    if (t < 0) t = 0;
    int64_t max_dur = mp3_decoder_duration_ns(d);
    if (t > max_dur) t = max_dur;
    int64_t bytes = duration_to_bytes(d, t);
    bytes &= ~(int64_t)3;
    int err;
    mp3_decoder_seek(d, bytes, 0, &err);
    return err;
}
extern "C" int mp3_decoder_skip(mp3_decoder *d, int64_t delta) {  // decode.go:313-315
    return mp3_decoder_seek_to_time(d, mp3_decoder_position_ns(d) + delta);
}

// ---- One long stream: frame index, frame-range decode, split over the devices (BASELINE.json configs[4]) ------------
struct mp3_stream_index {
    const uint8_t *data = nullptr;
    size_t len = 0;
    int sample_rate = 0;
    std::vector<int64_t> frame_pos;     // byte offset of every frame header (decode.go:154-216's frameStarts)
    std::vector<uint32_t> own_bytes;    // main-data bytes the frame owns (frame size - header - CRC - side info)
    std::vector<int64_t> gran_before;   // granules in front of frame f; one more entry at the end = all granules
};

extern "C" int mp3_stream_index_create(const uint8_t *data, size_t len, mp3_stream_index **out) {
    if (!out || (!data && len)) return MP3_ERR_INVALID;
    *out = nullptr;
    mp3_stream_index *ix = new mp3_stream_index();
    ix->data = data;
    ix->len = len;
    Source s;
    s.data = data;
    s.len = len;
    int rc = s.skip_tags();
    if (rc != MP3_OK) {
        delete ix;
        return rc;
    }
    int64_t gran = 0;
    for (;;) {
        Header h;
        int64_t fpos;
        rc = read_frame_header(s, &h, &fpos);
        if (rc != MP3_OK) break;
        if (h.id() == 0 || h.layer() != 1) break;  // frame.Read refuses these (frame.go:79-84): the linear decode ends here
        const int fs = h.frame_size();
        const int own = fs - 4 - h.side_info_size() - (h.protection_bit() == 0 ? 2 : 0);
        if (fs > 2000 || own < 0 || own > 1500 || s.pos + (fs - 4) > (int64_t)len) break;
        if (ix->frame_pos.empty()) ix->sample_rate = h.sampling_frequency_value();
        ix->frame_pos.push_back(fpos);
        ix->own_bytes.push_back((uint32_t)own);
        ix->gran_before.push_back(gran);
        gran += h.granules();
        s.pos += fs - 4;
    }
    ix->gran_before.push_back(gran);
    *out = ix;
    return MP3_OK;
}
extern "C" void mp3_stream_index_free(mp3_stream_index *ix) { delete ix; }
extern "C" int64_t mp3_stream_index_frames(const mp3_stream_index *ix) { return ix ? (int64_t)ix->frame_pos.size() : 0; }
extern "C" int mp3_stream_index_sample_rate(const mp3_stream_index *ix) { return ix ? ix->sample_rate : 0; }
extern "C" int64_t mp3_stream_index_frame_pos(const mp3_stream_index *ix, int64_t f) {
    return (ix && f >= 0 && f < (int64_t)ix->frame_pos.size()) ? ix->frame_pos[(size_t)f] : -1;
}
extern "C" int64_t mp3_stream_index_pcm_bytes(const mp3_stream_index *ix, int64_t f0, int64_t f1) {
    if (!ix || f0 < 0 || f1 < f0 || f1 > (int64_t)ix->frame_pos.size()) return -1;
    return (ix->gran_before[(size_t)f1] - ix->gran_before[(size_t)f0]) * MP3GPU_PCM_BYTES_PER_GRANULE;
}

namespace {

// Host half of a frame-range job: everything up to the device call.  Thread-safe (touches only the index and its own job).
struct RangeJob {
    std::vector<uint8_t> M;
    std::vector<mp3gpu_unit> units;  // lead-in frames dropped: the halo's units first, then the range's
    size_t halo_gr = 0, n_gr = 0;    // granules of the halo / of halo + range
    int rc = MP3_OK;                 // parse status of the range (EOF family already mapped)
    bool empty = true;               // nothing of the range itself could be parsed
};

void parse_range(const mp3_stream_index *ix, int64_t f0, int64_t f1, RangeJob &J) {
    // Halo: whole frames covering the two granules in front of f0 (PCM of a granule needs the IMDCT overlap of the one
    // before it and 15 slots of synthesis history, which need the overlap of the one before that: SURVEY.md 8e).
    int64_t fh = f0;
    while (fh > 0 && ix->gran_before[(size_t)f0] - ix->gran_before[(size_t)fh] < 2) fh--;
    // Lead-in: frames parsed (not decoded) so that the reservoir state at fh is the linear decode's.  A frame's logical
    // buffer reaches at most 511 bytes back (9-bit main_data_begin, maindata.go:290-323); once the frames parsed in front
    // own that many bytes, window starts and lengths no longer depend on how the parse began (prev == nil ignores
    // main_data_begin, and an underflow keeps all of prev: both only ever make the window start later than the
    // linear decode's until 511 bytes are covered).  One more frame for the first frame's own quirk (Q8).
    int64_t fl = fh;
    {
        size_t covered = 0;
        while (fl > 0 && covered < 512) covered += ix->own_bytes[(size_t)--fl];
        if (fl > 0) fl--;
    }
    StreamParser P;
    P.src.data = ix->data;
    P.src.len = ix->len;
    P.src.pos = ix->frame_pos[(size_t)fl];
    std::vector<mp3gpu_unit> units;
    J.M.reserve((size_t)(ix->frame_pos[(size_t)f1 - 1] - ix->frame_pos[(size_t)fl]) + 2048);
    units.reserve((size_t)(ix->gran_before[(size_t)f1] - ix->gran_before[(size_t)fl]) * 2);
    size_t units_at_fh = 0, units_at_f0 = 0;
    int rc = MP3_OK;
    int64_t f = fl;
    for (; f < f1; f++) {
        if (f == fh) units_at_fh = units.size();
        if (f == f0) units_at_f0 = units.size();
        rc = P.next_frame(J.M, units, 0);
        if (rc != MP3_OK) break;
    }
    const bool eof_family = rc == MP3_EOF || rc == MP3_ERR_UNEXPECTED_EOF || rc == MP3_ERR_SYNC_LIMIT;  // decode.go:48-63
    if (f <= f0) {
        J.rc = eof_family ? MP3_EOF : rc;
        return;
    }
    if (fl > 0)
        // the parse began mid-stream: nothing in front of the halo is decoded, and the halo starts from whatever state
        // the device has there (its output is dropped) — but never from a zero-state flag the lead-in's first frame got
        for (size_t k = units_at_fh; k < units.size(); k++) units[k].w2 &= ~MP3GPU_W2_ZERO_STATE;
    J.units.assign(units.begin() + (long)units_at_fh, units.end());
    J.halo_gr = (units_at_f0 - units_at_fh) / 2;
    J.n_gr = J.units.size() / 2;
    J.rc = eof_family ? MP3_OK : rc;  // a clean end of stream inside the range
    J.empty = false;
}

// Device half: pinned staging (makes the device call's copies asynchronous; it pipelines H2D / kernels / D2H wave by wave).
int run_range(mp3_engine *e, DeviceSlot *dev, const RangeJob &J, uint8_t *pcm_out, int64_t *pcm_bytes) {
    if (pcm_bytes) *pcm_bytes = 0;
    if (J.empty) return J.rc;
    int grc;
    {
        std::lock_guard<std::mutex> lk(dev->mu);
        int prc = e->ensure(dev->r_main, J.M.size() + 64);
        if (prc == MP3_OK) prc = e->ensure(dev->r_units, J.units.size() * sizeof(mp3gpu_unit) + 64);
        if (prc != MP3_OK) return prc;
        memcpy(dev->r_main.p, J.M.data(), J.M.size());
        memset((uint8_t *)dev->r_main.p + J.M.size(), 0, 64);
        memcpy(dev->r_units.p, J.units.data(), J.units.size() * sizeof(mp3gpu_unit));
        grc = e->api.decode_range(dev->gpu, (const uint8_t *)dev->r_main.p, J.M.size(), (const mp3gpu_unit *)dev->r_units.p, J.n_gr, J.halo_gr,
                                  (int16_t *)pcm_out);
        if (grc != MP3GPU_OK) e->set_err(e->api.last_error(dev->gpu));
    }
    if (grc != MP3GPU_OK) return MP3_ERR_DEVICE;
    if (pcm_bytes) *pcm_bytes = (int64_t)(J.n_gr - J.halo_gr) * MP3GPU_PCM_BYTES_PER_GRANULE;
    return J.rc;
}

}  // namespace

extern "C" int mp3_decode_frames(mp3_engine *e, int slot, const mp3_stream_index *ix, int64_t f0, int64_t f1, uint8_t *pcm_out,
                                 int64_t *pcm_bytes) {
    if (pcm_bytes) *pcm_bytes = 0;
    if (!e || !ix || slot < 0 || (size_t)slot >= e->devs.size() || f0 < 0 || f1 < f0 || f1 > (int64_t)ix->frame_pos.size() ||
        (f1 > f0 && !pcm_out))
        return MP3_ERR_INVALID;
    if (f1 == f0) return MP3_OK;
    RangeJob J;
    parse_range(ix, f0, f1, J);
    return run_range(e, e->devs[(size_t)slot], J, pcm_out, pcm_bytes);
}

// Every device gets one contiguous stretch of the stream; a stretch is cut further into ranges that the host threads parse
// side by side (the per-stream parse is serial: 0.7 us per frame), and the device's runner thread decodes them in order as
// they become ready — host parse, upload, kernels and download of neighbouring ranges overlap.
extern "C" int mp3_decode_stream_split(mp3_engine *e, const mp3_stream_index *ix, const uint8_t **pcm_base, int64_t *pcm_bytes,
                                       mp3_batch_timings *timings) {
    if (!e || !ix || !pcm_base || !pcm_bytes) return MP3_ERR_INVALID;
    const double t0 = now_s();
    const int64_t frames = (int64_t)ix->frame_pos.size();
    const size_t D = e->devs.size();
    const int64_t total = mp3_stream_index_pcm_bytes(ix, 0, frames);
    int rc = e->ensure(e->a_pcm, (size_t)total + 64);
    if (rc != MP3_OK) return rc;
    const int threads = hw_threads(e->opts.host_threads);
    // ranges per device: enough to keep the parsing threads busy, each at least ~4,096 frames
    size_t R = (size_t)std::max(1, std::min(16, (threads + (int)D - 1) / (int)D * 2));
    const char *mf = getenv("MP3HOST_SPLIT_MIN_FRAMES");  // tests lower it
    const int64_t min_frames = mf ? (int64_t)atoll(mf) : (int64_t)4096;
    while (R > 1 && frames / (int64_t)(D * R) < min_frames) R--;
    const size_t NR = D * R;
    std::vector<RangeJob> jobs(NR);
    std::vector<int64_t> lo(NR + 1), got(NR, 0), want(NR, 0);
    std::vector<int> rcs(NR, MP3_OK);
    for (size_t r = 0; r <= NR; r++) lo[r] = frames * (int64_t)r / (int64_t)NR;
    std::vector<char> ready(NR, 0);
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<size_t> next{0};
    double parse_s = 0;
    // ranges are parsed in the order the runners need them: range r of every device before range r + 1 of any
    auto order = [&](size_t i) { return (i % D) * R + i / D; };
    auto parser = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= NR) return;
            const size_t r = order(i);
            const double a = now_s();
            if (lo[r + 1] > lo[r]) parse_range(ix, lo[r], lo[r + 1], jobs[r]);
            const double b = now_s();
            std::lock_guard<std::mutex> lk(mu);
            parse_s += b - a;
            ready[r] = 1;
            cv.notify_all();
        }
    };
    std::vector<double> dev_s(D, 0.0);
    auto runner = [&](size_t d) {
        for (size_t k = 0; k < R; k++) {
            const size_t r = d * R + k;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return ready[r] != 0; });
            }
            want[r] = mp3_stream_index_pcm_bytes(ix, lo[r], lo[r + 1]);
            if (lo[r + 1] == lo[r]) continue;
            const double a = now_s();
            rcs[r] = run_range(e, e->devs[d], jobs[r], (uint8_t *)e->a_pcm.p + mp3_stream_index_pcm_bytes(ix, 0, lo[r]), &got[r]);
            dev_s[d] += now_s() - a;
            RangeJob().M.swap(jobs[r].M);  // release the host copy
            std::vector<mp3gpu_unit>().swap(jobs[r].units);
            if (rcs[r] != MP3_OK || got[r] != want[r]) {  // the linear decode stops here; later ranges of this device are moot
                for (size_t k2 = k + 1; k2 < R; k2++) want[d * R + k2] = -1;
                return;
            }
        }
    };
    {
        std::vector<std::thread> th;
        const int np = (int)std::min<size_t>((size_t)threads, NR);
        for (int t = 0; t < np; t++) th.emplace_back(parser);
        for (size_t d = 0; d < D; d++) th.emplace_back(runner, d);
        for (auto &t : th) t.join();
    }
    // what a linear decode returns: everything up to the first range that ended early
    int64_t out = 0;
    int status = MP3_OK;
    for (size_t r = 0; r < NR; r++) {
        if (want[r] < 0) break;
        out += got[r];
        if (rcs[r] != MP3_OK || got[r] != want[r]) {
            status = rcs[r];
            break;
        }
    }
    *pcm_base = (const uint8_t *)e->a_pcm.p;
    *pcm_bytes = out;
    if (timings) {
        memset(timings, 0, sizeof *timings);
        timings->parse_s = parse_s;  // summed over the parsing threads
        for (size_t d = 0; d < D; d++) timings->device_s = std::max(timings->device_s, dev_s[d]);
        timings->total_s = now_s() - t0;
        timings->n_granules = (uint64_t)(out / MP3GPU_PCM_BYTES_PER_GRANULE);
        timings->pcm_bytes = (uint64_t)out;
    }
    return status;
}
