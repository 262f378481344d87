// Fast reader for the particle samples that feed the HBT path.
//
// Replaces, for read_in_mode = 10 ("particle_samples.gz": gzipped iSS text), 2 ("particle_list.dat":
// gzipped UrQMD text), 21 ("particle_list.bin": UrQMD binary), 0 ("OSCAR.DAT": OSCAR1997A text), 1
// ("particle_list.dat": UrQMD file-13 style text; 4: the UrQMD 3.3p header; 3: no header block), 5
// ("particle_list.dat": JAM text), 9 ("particle_list.bin": iSS binary) and 7 ("particle_list.dat": gzipped
// SMASH text), the chain
//   particleSamples::read_in_particle_samples_UrQMD_3p3      src/particleSamples.cpp:1463-1529
//   particleSamples::read_in_particle_samples_Sangwook       src/particleSamples.cpp:1955-2020
//   particleSamples::read_in_particle_samples_JAM            src/particleSamples.cpp:716-756 (+ header :205-210)
//   particleSamples::open_SMASH_binary / read_in_particle_samples_SMASH_binary   :288-323, :1104-1201  (mode 8)
//   particleSamples::read_in_particle_samples_binary         src/particleSamples.cpp:1203-1245
//   particleSamples::read_in_particle_samples_SMASH_gzipped  src/particleSamples.cpp:1061-1102
//   particleSamples::read_in_particle_samples_OSCAR          src/particleSamples.cpp:680-714 (+ header :198-202)
//   particleSamples::read_in_particle_samples_UrQMD          src/particleSamples.cpp:838-908
//   particleSamples::read_in_particle_samples_gzipped        src/particleSamples.cpp:1247-1286
//   particleSamples::read_in_particle_samples_UrQMD_zipped   src/particleSamples.cpp:910-974
//   particleSamples::read_in_particle_samples_UrQMD_binary   src/particleSamples.cpp:976-1059
//   build_map_urqmd_to_pdg_id / get_pdg_id                   src/particleSamples.cpp:325-357,389-400
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
#include "hbt_inflate.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
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

// UrQMD (particle id, 2 x isospin projection) -> Monte-Carlo number, for the species the reference
// knows (src/particleSamples.cpp:325-357); anything else gets 0 and is dropped by the filter, but
// still counts towards event_buffer_size.
struct UrqmdSpecies {
    int id, iso3, pdg;
};
const UrqmdSpecies kUrqmdSpecies[] = {
    // mesons: pi, K, phi, eta, photon
    {101, 2, 211},   {101, 0, 111},    {101, -2, -211}, {106, 1, 321},    {106, -1, 311},  {-106, 1, -311},
    {-106, -1, -321}, {109, 0, 333},   {102, 0, 221},   {100, 0, 22},
    // baryons: N, Sigma, Xi, Lambda, Omega and their antiparticles
    {1, 1, 2212},    {1, -1, 2112},    {-1, -1, -2212}, {-1, 1, -2112},   {40, 2, 3222},   {-40, -2, -3222},
    {40, 0, 3212},   {-40, 0, -3212},  {40, -2, 3112},  {-40, 2, -3112},  {49, 1, 3322},   {-49, -1, -3322},
    {49, -1, 3312},  {-49, 1, -3312},  {27, 0, 3122},   {-27, 0, -3122},  {55, 0, 3334},   {-55, 0, -3334},
};
int urqmd_to_pdg(long long id, long long iso3) {
    for (const UrqmdSpecies &s : kUrqmdSpecies)
        if (s.id == id && s.iso3 == iso3) return s.pdg;
    return 0;
}

}  // namespace

struct hbt_reader {
    HbtGz *gz = nullptr;  // the reader's own gzip decoder (hbt_inflate.cpp: ~1.6x zlib's inflate on this text)
    FILE *bin = nullptr;  // read_in_mode 21, 9, 8
    uint16_t smash_version = 0;  // read_in_mode 8: format version of the file header
    int32_t mode = 10;
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

    // Inflating is the slower half of reading gzipped text (one deflate stream is sequential: ~130-350 MB/s of
    // output with zlib, 1.6x that with hbt_inflate.cpp, against ~250-500 MB/s for parsing it), so it runs on its own
    // thread, two pieces of kChunk bytes ahead of the parser: the reader then moves at the pace of the slower of the two.
    static constexpr size_t kChunk = 4u << 20;
    struct Chunk {
        std::vector<char> data;
        size_t n = 0;
        bool last = false;
    };
    std::thread inflater;
    std::mutex imu;
    std::condition_variable icv;
    std::deque<std::unique_ptr<Chunk>> inflated, spare;
    bool istop = false;
    std::string ierror;

    void inflate_loop() {
        for (;;) {
            std::unique_ptr<Chunk> c;
            {
                std::unique_lock<std::mutex> lk(imu);
                icv.wait(lk, [&] { return inflated.size() < 3 || istop; });
                if (istop) return;
                if (!spare.empty()) {
                    c = std::move(spare.front());
                    spare.pop_front();
                }
            }
            if (!c) {
                c.reset(new Chunk);
                c->data.resize(kChunk);
            }
            const long got = hbt_gz_read(gz, c->data.data(), kChunk);
            std::unique_lock<std::mutex> lk(imu);
            if (got < 0) {
                ierror = hbt_gz_error(gz);
                c->n = 0;
                c->last = true;
            } else {
                c->n = static_cast<size_t>(got);
                c->last = static_cast<size_t>(got) < kChunk;
            }
            const bool last = c->last;
            inflated.push_back(std::move(c));
            icv.notify_all();
            if (last) return;
        }
    }

    bool fill() {
        if (at_eof) return false;
        if (pos > 0 && pos < end) std::memmove(buf.data(), buf.data() + pos, end - pos);
        end -= pos;
        pos = 0;
        std::unique_ptr<Chunk> c;
        {
            std::unique_lock<std::mutex> lk(imu);
            icv.wait(lk, [&] { return !inflated.empty() || istop; });
            if (inflated.empty()) {
                at_eof = true;
                return false;
            }
            c = std::move(inflated.front());
            inflated.pop_front();
            icv.notify_all();
        }
        if (!ierror.empty()) {
            error = ierror;
            at_eof = true;
            return false;
        }
        if (end + c->n > buf.size()) buf.resize(std::max(buf.size() * 2, end + c->n));  // a line longer than the buffer
        std::memcpy(buf.data() + end, c->data.data(), c->n);
        const size_t got = c->n;
        end += got;
        bytes_inflated += static_cast<uint64_t>(got);
        if (c->last) at_eof = true;
        {
            std::lock_guard<std::mutex> lk(imu);
            spare.push_back(std::move(c));
        }
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

    // boostParticles (:447-452, evaluated as written), the single-species filter (:672-676) and,
    // optionally, the HBT gather's rapidity cut (src/HBT_correlation.cpp:261-266)
    void keep(Batch &b, long long mv, double ch, double sh, double t, double x, double y, double z, double E, double px,
              double py, double pz) const {
        if (mv != monval) return;
        const double E_s = E * ch + pz * sh;
        const double pz_s = pz * ch + E * sh;
        if (cut) {
            const double ratio = pz_s / E_s;
            if (!(ratio > cut_lo && ratio < cut_hi)) return;
        }
        const double rec[8] = {px, py, pz_s, E_s, x, y, z, t};
        b.p.insert(b.p.end(), rec, rec + 8);
    }

    std::unique_ptr<Batch> read_batch() {
        switch (mode) {
            case 21: return read_batch_urqmd_binary();
            case 2: return read_batch_urqmd_text();
            case 1: return read_batch_urqmd_f13(16);
            case 4: return read_batch_urqmd_f13(13);
            case 3: return read_batch_urqmd_f13(-1);
            case 5: return read_batch_jam();
            case 8: return read_batch_smash_binary();
            case 0: return read_batch_oscar();
            case 9: return read_batch_iss_binary();
            case 7: return read_batch_smash_text();
            default: return read_batch_iss();
        }
    }

    // read_in_mode 8, src/particleSamples.cpp:1110-1199 (extended SMASH binary; the file header is read when
    // the file is opened, :296-322): blocks 'f' (end of an event: u32, f64 and, from format version 7 on, one
    // more byte: skipped) and 'p' (u32 n, then n records of 128 bytes: t x y z m p0 px py pz as f64, pdg id
    // charge ncoll as i32, two f64, two i32, time_last_coll f64, two i32); anything else ends the batch.  The
    // particle is moved back along its velocity to the time of its last collision (:1177-1181, as written).
    std::unique_ptr<Batch> read_batch_smash_binary() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        std::vector<unsigned char> rec;
        while (num_particles < buffer_size) {
            char block_type;
            if (std::fread(&block_type, 1, 1, bin) != 1) break;  // !SMASH_inputfile (:1113)
            bytes_inflated += 1;
            if (block_type == 'f') {
                unsigned char skip[13];
                const size_t want = smash_version > 6 ? 13 : 12;
                const size_t got = std::fread(skip, 1, want, bin);
                bytes_inflated += got;
                if (got < want) break;
                continue;
            }
            if (block_type != 'p') break;
            uint32_t n_part = 0;
            if (std::fread(&n_part, 4, 1, bin) != 1) {
                error = "particles_binary.bin ends inside a block";
                return b;
            }
            bytes_inflated += 4;
            rec.resize(static_cast<size_t>(n_part) * 128);
            if (n_part && std::fread(rec.data(), 128, n_part, bin) != n_part) {
                error = "particles_binary.bin ends inside a block";
                return b;
            }
            bytes_inflated += rec.size();
            for (uint32_t ip = 0; ip < n_part; ip++) {
                const unsigned char *r = rec.data() + static_cast<size_t>(ip) * 128;
                double v[9], t_last;  // t x y z m p0 px py pz
                int32_t pdg;
                std::memcpy(v, r, 72);
                std::memcpy(&pdg, r + 72, 4);
                std::memcpy(&t_last, r + 112, 8);
                const double vx = v[6] / v[5], vy = v[7] / v[5], vz = v[8] / v[5];
                const double dt = v[0] - t_last;
                const double xm = vx * dt, ym = vy * dt, zm = vz * dt;
                keep(*b, pdg, ch, sh, t_last, v[1] - xm, v[2] - ym, v[3] - zm, v[5], v[6], v[7], v[8]);
            }
            num_particles += n_part;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 9, src/particleSamples.cpp:1208-1243: int n, then per particle int pdg and 9 floats (mass t x
    // y z E px py pz).  The reference never advances its event index in this routine: every particle of the
    // batch lands in the batch's FIRST event and the other events stay empty; reproduced as it is.
    std::unique_ptr<Batch> read_batch_iss_binary() {
        std::unique_ptr<Batch> b(new Batch);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        int nev = 0;
        std::vector<unsigned char> rec;
        while (num_particles < buffer_size) {
            int32_t n_particle = 0;
            const size_t got = std::fread(&n_particle, 1, 4, bin);
            bytes_inflated += got;
            if (got < 4) break;  // inputfile.eof() after the read of n_particle (:1211)
            nev++;
            if (n_particle > 0) {
                rec.resize(static_cast<size_t>(n_particle) * 40);
                if (std::fread(rec.data(), 40, static_cast<size_t>(n_particle), bin) != static_cast<size_t>(n_particle)) {
                    error = "particle_list.bin ends inside an event";
                    b->off.assign(1, 0);
                    return b;
                }
                bytes_inflated += rec.size();
            }
            for (int32_t ip = 0; ip < n_particle; ip++) {
                int32_t pdg;
                float v[9];
                std::memcpy(&pdg, rec.data() + static_cast<size_t>(ip) * 40, 4);
                std::memcpy(v, rec.data() + static_cast<size_t>(ip) * 40 + 4, 36);
                keep(*b, pdg, ch, sh, v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
            }
            num_particles += n_particle;
        }
        b->off.assign(1, 0);
        for (int e = 0; e < nev; e++) b->off.push_back(static_cast<int64_t>(b->p.size() / 8));  // all in event 0
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 7, src/particleSamples.cpp:1072-1100: "<n>" then n lines
    // "pdg charge process mother1 mother2 mass t x y z E px py pz"
    std::unique_ptr<Batch> read_batch_smash_text() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        while (num_particles < buffer_size) {
            const char *lb, *le;
            readline(&lb, &le);
            if (hit_eof) break;  // gzeof() after the header read (:1075)
            long long n_particle = 0;
            parse_int(lb, le, &n_particle);
            for (long long ip = 0; ip < n_particle; ip++) {
                if (!readline(&lb, &le) && hit_eof) {
                    error = "particle_list.dat ends inside an event";
                    return b;
                }
                long long pdg = 0, charge = 0, proc = 0, m1 = 0, m2 = 0;
                double mass, t, x, y, z, E, px, py, pz;
                const char *q = parse_int(lb, le, &pdg);
                q = parse_int(q, le, &charge);
                q = parse_int(q, le, &proc);
                q = parse_int(q, le, &m1);
                q = parse_int(q, le, &m2);
                q = parse_double(q, le, &mass);
                q = parse_double(q, le, &t);
                q = parse_double(q, le, &x);
                q = parse_double(q, le, &y);
                q = parse_double(q, le, &z);
                q = parse_double(q, le, &E);
                q = parse_double(q, le, &px);
                q = parse_double(q, le, &py);
                q = parse_double(q, le, &pz);
                keep(*b, pdg, ch, sh, t, x, y, z, E, px, py, pz);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 0, src/particleSamples.cpp:689-712 (the three header lines of the file are skipped when it
    // is opened, :198-202): "<event id> <n> ..." then n lines "<index> <monval> px py pz E mass x y z t"
    std::unique_ptr<Batch> read_batch_oscar() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        while (num_particles < buffer_size) {
            const char *lb, *le;
            readline(&lb, &le);
            if (hit_eof) break;  // inputfile.eof() after the header read (:694)
            long long event_id = 0, n_particle = 0;
            const char *q = parse_int(lb, le, &event_id);
            parse_int(q, le, &n_particle);
            for (long long ip = 0; ip < n_particle; ip++) {
                if (!readline(&lb, &le) && hit_eof) {
                    error = "OSCAR.DAT ends inside an event";
                    return b;
                }
                long long idx = 0, mv = 0;
                double px, py, pz, E, mass, x, y, z, t;
                q = parse_int(lb, le, &idx);
                q = parse_int(q, le, &mv);
                q = parse_double(q, le, &px);
                q = parse_double(q, le, &py);
                q = parse_double(q, le, &pz);
                q = parse_double(q, le, &E);
                q = parse_double(q, le, &mass);
                q = parse_double(q, le, &x);
                q = parse_double(q, le, &y);
                q = parse_double(q, le, &z);
                q = parse_double(q, le, &t);
                keep(*b, mv, ch, sh, t, x, y, z, E, px, py, pz);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 5, src/particleSamples.cpp:726-754 (the first line of the file is skipped when it is opened,
    // :205-210): "<char> <event id> <n>" then n lines "monval mass px py pz x y z t"; E from the mass shell
    std::unique_ptr<Batch> read_batch_jam() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        while (num_particles < buffer_size) {
            const char *lb, *le;
            readline(&lb, &le);
            if (hit_eof) break;  // inputfile.eof() after the header read (:731)
            long long event_id = 0, n_particle = 0;
            const char *q = skip_ws(lb, le);
            if (q < le) q++;  // operator>>(char): one non-blank character
            q = parse_int(q, le, &event_id);
            parse_int(q, le, &n_particle);
            for (long long ip = 0; ip < n_particle; ip++) {
                if (!readline(&lb, &le) && hit_eof) {
                    error = "particle_list.dat ends inside an event";
                    return b;
                }
                long long mv = 0;
                double mass, px, py, pz, x, y, z, t;
                q = parse_int(lb, le, &mv);
                q = parse_double(q, le, &mass);
                q = parse_double(q, le, &px);
                q = parse_double(q, le, &py);
                q = parse_double(q, le, &pz);
                q = parse_double(q, le, &x);
                q = parse_double(q, le, &y);
                q = parse_double(q, le, &z);
                q = parse_double(q, le, &t);
                const double E = std::sqrt(mass * mass + px * px + py * py + pz * pz);  // :743-747, as written
                keep(*b, mv, ch, sh, t, x, y, z, E, px, py, pz);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 1, src/particleSamples.cpp:849-906: 17 header lines, "<n> ...", one line that is skipped,
    // then n lines "r0 rx ry rz p0 px py pz m ityp 2i3 chg lcl# ncl or t x y z E px py pz" (the last eight are
    // the freeze-out coordinates and momenta that are used).  read_in_mode 4 (:1473-1527) has 14 header lines,
    // read_in_mode 3 (:1968-2018) none: `skip` = header lines after the first one, -1 = the first line is "<n>".
    std::unique_ptr<Batch> read_batch_urqmd_f13(int skip) {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        while (num_particles < buffer_size) {
            const char *lb, *le;
            readline(&lb, &le);
            if (hit_eof) break;  // inputfile.eof() after the first line (:851, :1475, :1970)
            if (skip >= 0) {
                for (int i = 0; i < skip; i++) readline(&lb, &le);
                readline(&lb, &le);
            }
            long long n_particle = 0;
            parse_int(lb, le, &n_particle);
            if (!readline(&lb, &le) && hit_eof && n_particle > 0) {
                error = "particle_list.dat ends inside an event";
                return b;
            }
            for (long long ip = 0; ip < n_particle; ip++) {
                if (!readline(&lb, &le) && hit_eof) {
                    error = "particle_list.dat ends inside an event";
                    return b;
                }
                double d, mass, t, x, y, z, E, px, py, pz;
                long long id = 0, iso3 = 0, charge = 0, proc = 0;
                const char *q = lb;
                for (int i = 0; i < 8; i++) q = parse_double(q, le, &d);
                q = parse_double(q, le, &mass);
                q = parse_int(q, le, &id);
                q = parse_int(q, le, &iso3);
                q = parse_int(q, le, &charge);
                q = parse_double(q, le, &d);
                q = parse_double(q, le, &d);
                q = parse_int(q, le, &proc);
                q = parse_double(q, le, &t);
                q = parse_double(q, le, &x);
                q = parse_double(q, le, &y);
                q = parse_double(q, le, &z);
                q = parse_double(q, le, &E);
                q = parse_double(q, le, &px);
                q = parse_double(q, le, &py);
                q = parse_double(q, le, &pz);
                keep(*b, urqmd_to_pdg(id, iso3), ch, sh, t, x, y, z, E, px, py, pz);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 2, src/particleSamples.cpp:921-972: "<n>", one line that is skipped, then n lines
    // "id iso3 charge <2 numbers> process mass t x y z E px py pz"
    std::unique_ptr<Batch> read_batch_urqmd_text() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        while (num_particles < buffer_size) {
            const char *lb, *le;
            readline(&lb, &le);
            if (hit_eof) break;  // gzeof() after the header read (:924)
            long long n_particle = 0;
            parse_int(lb, le, &n_particle);
            if (!readline(&lb, &le) && hit_eof && n_particle > 0) {
                error = "particle_list.dat ends inside an event";
                return b;
            }
            for (long long ip = 0; ip < n_particle; ip++) {
                if (!readline(&lb, &le) && hit_eof) {
                    error = "particle_list.dat ends inside an event";
                    return b;
                }
                long long id = 0, iso3 = 0, charge = 0, proc = 0;
                double d1, d2, mass, t, x, y, z, E, px, py, pz;
                const char *q = parse_int(lb, le, &id);
                q = parse_int(q, le, &iso3);
                q = parse_int(q, le, &charge);
                q = parse_double(q, le, &d1);
                q = parse_double(q, le, &d2);
                q = parse_int(q, le, &proc);
                q = parse_double(q, le, &mass);
                q = parse_double(q, le, &t);
                q = parse_double(q, le, &x);
                q = parse_double(q, le, &y);
                q = parse_double(q, le, &z);
                q = parse_double(q, le, &E);
                q = parse_double(q, le, &px);
                q = parse_double(q, le, &py);
                q = parse_double(q, le, &pz);
                keep(*b, urqmd_to_pdg(id, iso3), ch, sh, t, x, y, z, E, px, py, pz);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 21, src/particleSamples.cpp:985-1057: int n, 8 ints that are skipped, then per
    // particle 6 ints (id, iso3, ...) and 9 floats (mass t x y z E px py pz), native byte order
    std::unique_ptr<Batch> read_batch_urqmd_binary() {
        std::unique_ptr<Batch> b(new Batch);
        b->off.push_back(0);
        const double ch = std::cosh(rap_shift), sh = std::sinh(rap_shift);
        int64_t num_particles = 0;
        std::vector<unsigned char> rec;
        while (num_particles < buffer_size) {
            int32_t head[9];
            const size_t got = std::fread(head, 1, 4, bin);
            bytes_inflated += got;
            if (got < 4) break;  // inputfile.eof() after the read of n_particle (:986)
            const int32_t n_particle = head[0];
            if (std::fread(head + 1, 4, 8, bin) != 8 && n_particle > 0) {
                error = "particle_list.bin ends inside an event";
                return b;
            }
            bytes_inflated += 32;
            if (n_particle > 0) {
                rec.resize(static_cast<size_t>(n_particle) * 60);
                if (std::fread(rec.data(), 60, static_cast<size_t>(n_particle), bin) != static_cast<size_t>(n_particle)) {
                    error = "particle_list.bin ends inside an event";
                    return b;
                }
                bytes_inflated += rec.size();
            }
            for (int32_t ip = 0; ip < n_particle; ip++) {
                int32_t info[6];
                float v[9];
                std::memcpy(info, rec.data() + static_cast<size_t>(ip) * 60, 24);
                std::memcpy(v, rec.data() + static_cast<size_t>(ip) * 60 + 24, 36);
                keep(*b, urqmd_to_pdg(info[0], info[1]), ch, sh, v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    // read_in_mode 10, src/particleSamples.cpp:1256-1284
    std::unique_ptr<Batch> read_batch_iss() {
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
                keep(*b, mv, ch, sh, t, x, y, z, E, px, py, pz);
            }
            num_particles += n_particle;
            b->off.push_back(static_cast<int64_t>(b->p.size() / 8));
        }
        b->all_particles = num_particles;
        return b;
    }

    void run() {
        if (mode == 0 || mode == 5) {  // the file header of OSCAR1997A (src/particleSamples.cpp:198-202), of JAM (:205-210)
            const char *lb, *le;
            for (int i = 0; i < (mode == 0 ? 3 : 1); i++) readline(&lb, &le);
        }
        for (;;) {
            std::unique_ptr<Batch> b = read_batch();
            const bool last = b->off.size() == 1 || !error.empty();  // no event: end of the file
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready.size() < 2 || stop; });
            if (stop) return;
            if (b->off.size() > 1 && error.empty()) ready.push_back(std::move(b));  // a batch cut short by an error is not delivered
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
    if (read_in_mode != 10 && read_in_mode != 2 && read_in_mode != 21 && read_in_mode != 0 && read_in_mode != 1 &&
        read_in_mode != 9 && read_in_mode != 7 && read_in_mode != 3 && read_in_mode != 4 && read_in_mode != 5 && read_in_mode != 8)
        return HBT_ERR_INVALID;
    // species groups (9999, 9998, ... : all charged, ...) need the particle table; single species only
    const int32_t a = particle_monval < 0 ? -particle_monval : particle_monval;
    if (a >= 9996 && a <= 99999 && (a <= 9999 || a == 99999)) return HBT_ERR_INVALID;
    HbtGz *gz = nullptr;
    FILE *bin = nullptr;
    uint16_t smash_version = 0;
    if (read_in_mode == 21 || read_in_mode == 9 || read_in_mode == 8) {
        bin = std::fopen(path, "rb");
        if (!bin) return HBT_ERR_INVALID;
        std::setvbuf(bin, nullptr, _IOFBF, 1 << 20);
        if (read_in_mode == 8) {  // open_SMASH_binary, src/particleSamples.cpp:296-322
            char magic[4];
            uint16_t variant = 0;
            uint32_t len = 0;
            bool ok = std::fread(magic, 1, 4, bin) == 4 && std::fread(&smash_version, 2, 1, bin) == 1 &&
                      std::fread(&variant, 2, 1, bin) == 1 && std::fread(&len, 4, 1, bin) == 1;
            ok = ok && std::memcmp(magic, "SMSH", 4) == 0 && variant == 1 && len < 4096;  // the extended format only
            if (ok && len) ok = std::fseek(bin, static_cast<long>(len), SEEK_CUR) == 0;
            if (!ok) {
                std::fclose(bin);
                return HBT_ERR_INVALID;
            }
        }
    } else {
        gz = hbt_gz_open(path);
        if (!gz) return HBT_ERR_INVALID;
    }
    hbt_reader *r = new hbt_reader;
    r->gz = gz;
    r->bin = bin;
    r->smash_version = smash_version;
    r->mode = read_in_mode;
    r->monval = particle_monval;
    r->buffer_size = event_buffer_size;
    r->rap_shift = rap_shift;
    if (rapidity_cut) {
        r->cut = true;
        r->cut_lo = std::tanh(rapidity_cut->HBTrap_min);
        r->cut_hi = std::tanh(rapidity_cut->HBTrap_max);
    }
    r->buf.resize(2 * hbt_reader::kChunk);
    if (gz) r->inflater = std::thread([r] { r->inflate_loop(); });
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
    {
        std::lock_guard<std::mutex> lk(r->imu);
        r->istop = true;
    }
    r->icv.notify_all();
    if (r->worker.joinable()) r->worker.join();
    if (r->inflater.joinable()) r->inflater.join();
    if (r->gz) hbt_gz_close(r->gz);
    if (r->bin) std::fclose(r->bin);
    delete r;
}
