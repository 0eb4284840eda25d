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

bool load_gpu_api(GpuApi &api, bool exact, std::string &err) {
    std::string path = self_dir() + (exact ? "/libmp3gpu_exact.so" : "/libmp3gpu.so");
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

struct mp3_engine {
    mp3_engine_opts opts{};
    GpuApi api;
    mp3gpu_ctx *gpu = nullptr;
    std::string err;
    Pinned a_main, a_units, a_pcm;

    int ensure(Pinned &a, size_t bytes) {
        if (a.cap >= bytes) return MP3_OK;
        if (a.p) api.host_free(a.p);
        a.p = nullptr;
        a.cap = 0;
        size_t want = bytes + bytes / 16 + 4096;
        a.p = api.host_alloc(want);
        if (!a.p) {
            err = "pinned host allocation of " + std::to_string(want) + " bytes failed";
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
    if (!load_gpu_api(e->api, e->opts.use_exact_library != 0, e->err)) {
        fprintf(stderr, "mp3_engine_create: %s\n", e->err.c_str());
        delete e;
        return MP3_ERR_DEVICE;
    }
    mp3gpu_opts go{};
    go.abi_version = MP3GPU_ABI_VERSION;
    go.wave_granules = e->opts.wave_granules;
    go.keep_intermediates = e->opts.keep_intermediates;
    int rc = e->api.create(e->opts.device, &go, &e->gpu);
    if (rc != MP3GPU_OK) {
        fprintf(stderr, "mp3_engine_create: mp3gpu_create failed (%d): no CUDA device, and there is no CPU decode path\n", rc);
        dlclose(e->api.handle);
        delete e;
        return MP3_ERR_DEVICE;
    }
    *out = e;
    return MP3_OK;
}

extern "C" void mp3_engine_destroy(mp3_engine *e) {
    if (!e) return;
    if (e->a_main.p) e->api.host_free(e->a_main.p);
    if (e->a_units.p) e->api.host_free(e->a_units.p);
    if (e->a_pcm.p) e->api.host_free(e->a_pcm.p);
    if (e->gpu) e->api.destroy(e->gpu);
    if (e->api.handle) dlclose(e->api.handle);
    delete e;
}

extern "C" const char *mp3_engine_last_error(const mp3_engine *e) { return e ? e->err.c_str() : "no engine"; }
extern "C" mp3gpu_ctx *mp3_engine_gpu(mp3_engine *e) { return e ? e->gpu : nullptr; }

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
    case MP3_ERR_DEVICE: return "mp3: device engine failure";
    case MP3_ERR_INVALID: return "mp3: invalid argument";
    }
    return "mp3: unknown error";
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
size_t unit_slots_upper_bound(const uint8_t *data, size_t len) {
    Source s;
    s.data = data;
    s.len = len;
    if (s.skip_tags() != MP3_OK) return 0;
    size_t slots = 0;
    for (;;) {
        Header h;
        int64_t fpos;
        if (read_frame_header(s, &h, &fpos) != MP3_OK) break;
        if (h.id() == 0 || h.layer() != 1) break;
        const int fs = h.frame_size();
        if (fs > 2000 || fs < 4 || s.pos + (fs - 4) > (int64_t)len) break;
        s.pos += fs - 4;
        slots += (size_t)h.granules() * 2;
    }
    return slots;
}

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

// Large batches are cut into chunks of streams: while the device decodes chunk k (the PCIe-bound part), the host
// threads parse and gather chunk k+1 straight behind it in the same pinned arenas.
extern "C" int mp3_decode_batch(mp3_engine *e, const uint8_t *const *data, const size_t *lens, size_t n,
                                mp3_stream_result *results, const uint8_t **pcm_base, mp3_batch_timings *timings) {
    if (!e || !results || !pcm_base || (n && (!data || !lens))) return MP3_ERR_INVALID;
    const double t0 = now_s();
    const size_t n_chunks = n >= 256 ? std::min<size_t>(8, n / 128) : 1;
    // ---- arena sizes: main data never exceeds the input bytes; unit slots from a header-only frame walk ----
    size_t main_ub = 64 * (n_chunks + 1), slots_ub = 0;
    {
        std::vector<size_t> ub(n);
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                ub[i] = n_chunks > 1 ? unit_slots_upper_bound(data[i], lens[i]) : 0;
            }
        };
        int nt = (int)std::min<size_t>((size_t)hw_threads(e->opts.host_threads), n ? n : 1);
        if (n_chunks > 1 && nt > 1) {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work);
            for (auto &t : th) t.join();
        } else {
            work();
        }
        for (size_t i = 0; i < n; i++) {
            main_ub += (lens[i] + 3) & ~size_t(3);
            slots_ub += ub[i];
        }
    }
    double parse_s = 0, gather_s = 0, device_s = 0;
    size_t m_cursor = 0, u_cursor = 0;  // bytes into a_main (64-byte aligned per chunk), unit slots into a_units
    int dev_rc = MP3GPU_OK;
    std::string dev_err;

    std::mutex mu;
    std::condition_variable cv;
    std::deque<DeviceJob> jobs;
    bool closed = false;
    std::thread worker;
    auto run_job = [&](const DeviceJob &j) {
        if (dev_rc != MP3GPU_OK || j.n_granules == 0) return;
        const double a = now_s();
        int rc = e->api.decode(e->gpu, j.main_data, j.main_len, j.units, j.n_granules, j.pcm);
        device_s += now_s() - a;
        if (rc != MP3GPU_OK) {
            dev_rc = rc;
            dev_err = e->api.last_error(e->gpu);
        }
    };

    std::vector<ParsedStream> ps;  // reused by every chunk: the per-stream vectors keep their (already touched) capacity
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t i0 = n * c / n_chunks, i1 = n * (c + 1) / n_chunks;
        const double ta = now_s();
        parse_all(data + i0, lens + i0, i1 - i0, e->opts.host_threads, ps);
        const double tb = now_s();
        parse_s += tb - ta;
        BatchLayout L = layout_batch(ps);
        if (n_chunks == 1) {  // sizes are exact here
            int rc = e->ensure(e->a_main, L.m_total + 64);
            if (rc == MP3_OK) rc = e->ensure(e->a_units, sizeof(mp3gpu_unit) * L.u_total + 64);
            if (rc == MP3_OK) rc = e->ensure(e->a_pcm, (L.u_total / 2) * MP3GPU_PCM_BYTES_PER_GRANULE + 64);
            if (rc != MP3_OK) return rc;
        } else if (c == 0) {  // before the worker exists: growing an arena moves it
            int rc = e->ensure(e->a_main, main_ub);
            if (rc == MP3_OK) rc = e->ensure(e->a_units, sizeof(mp3gpu_unit) * slots_ub + 64);
            if (rc == MP3_OK) rc = e->ensure(e->a_pcm, (slots_ub / 2) * MP3GPU_PCM_BYTES_PER_GRANULE + 64);
            if (rc != MP3_OK) return rc;
            worker = std::thread([&]() {
                for (;;) {
                    DeviceJob j;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return closed || !jobs.empty(); });
                        if (jobs.empty()) return;
                        j = jobs.front();
                        jobs.pop_front();
                    }
                    run_job(j);
                }
            });
        }
        if (n_chunks > 1 && (m_cursor + L.m_total + 64 > e->a_main.cap || u_cursor + L.u_total > slots_ub)) {
            // cannot happen (the bounds are bounds); refuse rather than overrun
            {
                std::lock_guard<std::mutex> lk(mu);
                closed = true;
            }
            cv.notify_all();
            worker.join();
            e->err = "DecodeBatch arena bound exceeded";
            return MP3_ERR_INVALID;
        }
        uint8_t *m_dst = (uint8_t *)e->a_main.p + m_cursor;
        mp3gpu_unit *u_dst = (mp3gpu_unit *)e->a_units.p + u_cursor;
        gather_batch(ps, L, m_dst, u_dst, results + i0, e->opts.host_threads);
        const int64_t pcm_off = (int64_t)(u_cursor / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
        for (size_t i = i0; i < i1; i++) results[i].pcm_offset += pcm_off;
        gather_s += now_s() - tb;
        DeviceJob j{m_dst, L.m_total, u_dst, L.u_total / 2, (int16_t *)((uint8_t *)e->a_pcm.p + pcm_off)};
        if (n_chunks == 1) {
            run_job(j);
        } else {
            {
                std::lock_guard<std::mutex> lk(mu);
                jobs.push_back(j);
            }
            cv.notify_one();
        }
        m_cursor += (L.m_total + 64 + 63) & ~size_t(63);
        u_cursor += L.u_total;
    }
    if (worker.joinable()) {
        {
            std::lock_guard<std::mutex> lk(mu);
            closed = true;
        }
        cv.notify_all();
        worker.join();
    }
    const double t3 = now_s();
    if (dev_rc != MP3GPU_OK) {
        e->err = dev_err;
        return MP3_ERR_DEVICE;
    }
    *pcm_base = (const uint8_t *)e->a_pcm.p;
    if (timings) {
        timings->parse_s = parse_s;     // summed over the chunks; chunks after the first overlap the device call
        timings->gather_s = gather_s;
        timings->device_s = device_s;   // summed device-call time (worker thread)
        timings->total_s = t3 - t0;
        timings->main_data_bytes = m_cursor;
        timings->n_granules = u_cursor / 2;
        timings->pcm_bytes = (uint64_t)(u_cursor / 2) * MP3GPU_PCM_BYTES_PER_GRANULE;
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
            int grc = eng->api.decode(eng->gpu, M.data(), M.size() - 64, sub.data(), n_gr, pcm_tmp.data());
            M.resize(M.size() - 64);
            if (grc != MP3GPU_OK) {
                eng->err = eng->api.last_error(eng->gpu);
                return MP3_ERR_DEVICE;
            }
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
    int dummy;
    if (!err) err = &dummy;
    if (!e || (!data && len)) {
        *err = MP3_ERR_INVALID;
        return nullptr;
    }
    mp3_decoder *d = new mp3_decoder();
    d->eng = e;
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
