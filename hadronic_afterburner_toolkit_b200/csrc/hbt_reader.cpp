// Fast reader for the gzipped particle samples that feed the HBT path.
//
// Replaces, for read_in_mode = 10 ("particle_samples.gz"), the chain
//   particleSamples::read_in_particle_samples_gzipped   src/particleSamples.cpp:1247-1286
//   gz_readline (one gzread per byte + a stringstream)   src/particleSamples.cpp:2209-2218
//   boostParticles (rap_shift)                           src/particleSamples.cpp:441-470
//   filter_particles / decide_to_pick_OSCAR              src/particleSamples.cpp:625-678,1329-1346
//   the HBT gather with its rapidity cut                 src/HBT_correlation.cpp:255-281
// with the same grouping rule (events are appended while the particle count of ALL species is
// below event_buffer_size; the end of the file ends the batch), the same doubles (correctly
// rounded decimal -> binary64, as operator>> gives) and the same order.  A background thread
// inflates the file in 4 MiB pieces, parses it with std::from_chars and keeps up to two batches
// ahead of the consumer, so that reading overlaps the pair kernels of the previous batch.
//
// Host code only (no CUDA); part of libhbt_b200.so, C ABI in include/hbt_b200.h.
#include <zlib.h>

#include <charconv>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hbt_b200.h"

namespace {

struct Batch {
    std::vector<double> p;       // particles of interest, 8 doubles each: px py pz E x y z t
    std::vector<int64_t> off;    // per-event offsets into p (in particles), nev + 1 entries
    int64_t all_particles = 0;   // particles of every species read for this batch
};

}  // namespace

struct hbt_reader {
    gzFile gz = nullptr;
    int32_t monval = 0;
    int64_t buffer_size = 0;
    double rap_shift = 0.0;
    bool cut = false;
    double cut_lo = 0.0, cut_hi = 0.0;  // tanh(HBTrap_min), tanh(HBTrap_max)

    // inflate buffer
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool at_eof = false;   // gzread returned fewer bytes than asked: nothing more in the file
    bool hit_eof = false;  // a read was attempted beyond the last byte (what gzeof() reports)

    // producer / consumer
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::unique_ptr<Batch>> ready;
    bool done = false, stop = false;
    std::string error;
    std::unique_ptr<Batch> current;
    uint64_t bytes_inflated = 0;

    bool fill() {
        if (at_eof) return false;
        if (pos > 0 && pos < end) std::memmove(buf.data(), buf.data() + pos, end - pos);
        end -= pos;
        pos = 0;
        if (end == buf.size()) buf.resize(buf.size() * 2);  // a line longer than the buffer
        const int want = static_cast<int>(buf.size() - end);
        const int got = gzread(gz, buf.data() + end, static_cast<unsigned>(want));
        if (got < 0) {
            int errnum = 0;
            error = gzerror(gz, &errnum);
            at_eof = true;
            return false;
        }
        end += static_cast<size_t>(got);
        bytes_inflated += static_cast<uint64_t>(got);
        if (got < want) at_eof = true;
        return got > 0;
    }

    // next line without its '\n' (gz_readline); sets hit_eof when the read ran past the end
    bool readline(const char **b, const char **e) {
        size_t scan = pos;
        for (;;) {
            const char *nl = static_cast<const char *>(std::memchr(buf.data() + scan, '\n', end - scan));
            if (nl) {
                *b = buf.data() + pos;
                *e = nl;
                pos = static_cast<size_t>(nl - buf.data()) + 1;
                return true;
            }
            const size_t have = end - pos;
            if (!fill()) {  // the file ends without a newline: what is left is the line
                hit_eof = true;
                *b = buf.data() + pos;
                *e = buf.data() + end;
                pos = end;
                return have > 0;
            }
            scan = have;  // fill() moved the partial line to the front
        }
    }

    static const char *skip_ws(const char *p, const char *e) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) p++;
        return p;
    }
    // operator>>(double): 0 on failure
    static const char *parse_double(const char *p, const char *e, double *v) {
        p = skip_ws(p, e);
        if (p < e && *p == '+') p++;
        const auto r = std::from_chars(p, e, *v);
        if (r.ec != std::errc()) { *v = 0.0; return e; }
        return r.ptr;
    }
    static const char *parse_int(const char *p, const char *e, long long *v) {
        p = skip_ws(p, e);
        if (p < e && *p == '+') p++;
        const auto r = std::from_chars(p, e, *v);
        if (r.ec != std::errc()) { *v = 0; return e; }
        return r.ptr;
    }

    // one batch, src/particleSamples.cpp:1256-1284; returns false when nothing at all could be read
    std::unique_ptr<Batch> read_batch() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        while (num_particles < buffer_size) {
            const char *lb, *le;
            readline(&lb, &le);
            if (hit_eof) break;  // gzeof() after the header read (:1259): the read ran past the last byte
            long long n_particle = 0;
            parse_int(lb, le, &n_particle);
            for (long long ip = 0; ip < n_particle; ip++) {
                if (!readline(&lb, &le) && hit_eof) {
                    error = "particle_samples.gz ends inside an event";
                    return b;
                }
                long long mv = 0;
                const char *q = parse_int(lb, le, &mv);
                double mass, t, x, y, z, E, px, py, pz;
                q = parse_double(q, le, &mass);
                q = parse_double(q, le, &t);
                q = parse_double(q, le, &x);
                q = parse_double(q, le, &y);
                q = parse_double(q, le, &z);
                q = parse_double(q, le, &E);
                q = parse_double(q, le, &px);
                q = parse_double(q, le, &py);
                q = parse_double(q, le, &pz);
                (void)mass;
                if (mv != monval) continue;  // decide_to_pick_OSCAR, single-species branch (:672-676)
                // boostParticles (:447-452), evaluated as written
                const double E_s = E * ch + pz * sh;
                const double pz_s = pz * ch + E * sh;
                if (cut) {  // the HBT gather's rapidity cut (src/HBT_correlation.cpp:261-266)
                    const double ratio = pz_s / E_s;
                    if (!(ratio > cut_lo && ratio < cut_hi)) continue;
                }
                const double rec[8] = {px, py, pz_s, E_s, x, y, z, t};
                b->p.insert(b->p.end(), rec, rec + 8);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    void run() {
        for (;;) {
            std::unique_ptr<Batch> b = read_batch();
            const bool last = b->off.size() == 1 || !error.empty();  // no event: end of the file
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready.size() < 2 || stop; });
            if (stop) return;
            if (b->off.size() > 1) ready.push_back(std::move(b));
            if (last) done = true;
            cv.notify_all();
            if (last) return;
        }
    }
};

extern "C" int hbt_reader_open(const char *path, int32_t read_in_mode, int32_t particle_monval, int64_t event_buffer_size,
                               double rap_shift, const hbt_params *rapidity_cut, hbt_reader **out) {
    if (!path || !out) return HBT_ERR_INVALID;
    *out = nullptr;
    if (read_in_mode != 10) return HBT_ERR_INVALID;  // the text format of read_in_particle_samples_gzipped only
    // species groups (9999, 9998, ... : all charged, ...) need the particle table; single species only
    const int32_t a = particle_monval < 0 ? -particle_monval : particle_monval;
    if (a >= 9996 && a <= 99999 && (a <= 9999 || a == 99999)) return HBT_ERR_INVALID;
    gzFile gz = gzopen(path, "rb");
    if (!gz) return HBT_ERR_INVALID;
    gzbuffer(gz, 1 << 20);
    hbt_reader *r = new hbt_reader;
    r->gz = gz;
    r->monval = particle_monval;
    r->buffer_size = event_buffer_size;
    r->rap_shift = rap_shift;
    if (rapidity_cut) {
        r->cut = true;
        r->cut_lo = std::tanh(rapidity_cut->HBTrap_min);
        r->cut_hi = std::tanh(rapidity_cut->HBTrap_max);
    }
    r->buf.resize(4 << 20);
    r->worker = std::thread([r] { r->run(); });
    *out = r;
    return HBT_OK;
}

extern "C" int32_t hbt_reader_next(hbt_reader *r, const double **particles, const int64_t **offsets, int64_t *all_particles) {
    if (!r) return HBT_ERR_INVALID;
    std::unique_lock<std::mutex> lk(r->mu);
    r->cv.wait(lk, [&] { return !r->ready.empty() || r->done; });
    if (r->ready.empty()) {
        r->current.reset();
        if (particles) *particles = nullptr;
        if (offsets) *offsets = nullptr;
        if (all_particles) *all_particles = 0;
        return r->error.empty() ? 0 : HBT_ERR_INVALID;
    }
    r->current = std::move(r->ready.front());
    r->ready.pop_front();
    r->cv.notify_all();
    if (particles) *particles = r->current->p.data();
    if (offsets) *offsets = r->current->off.data();
    if (all_particles) *all_particles = r->current->all_particles;
    return static_cast<int32_t>(r->current->off.size() - 1);
}

extern "C" const char *hbt_reader_error(const hbt_reader *r) { return r ? r->error.c_str() : "null reader"; }

extern "C" uint64_t hbt_reader_bytes(const hbt_reader *r) { return r ? r->bytes_inflated : 0; }

extern "C" void hbt_reader_close(hbt_reader *r) {
    if (!r) return;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        r->stop = true;
    }
    r->cv.notify_all();
    if (r->worker.joinable()) r->worker.join();
    if (r->gz) gzclose(r->gz);
    delete r;
}
