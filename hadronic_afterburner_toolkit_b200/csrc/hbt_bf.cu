// Device path of the BalanceFunction pair loops, C ABI in include/hbt_b200.h.
//
//   BalanceFunction::combine_and_bin_particle_pairs         src/BalanceFunction.cpp:120-157
//   BalanceFunction::combine_and_bin_mixed_particle_pairs   src/BalanceFunction.cpp:159-197
//
// The loops only histogram (Delta y, Delta phi) of particle pairs of ONE event (or one event and
// its drawn partner); per pair the reference does IEEE subtract / add / divide / floor / int cast on
// values the host computed per particle.  The two divisions are by grid constants, and
// floor(RN(x / d)) is a monotone step function of the double x: its steps are found ONCE on the host
// (bisection over doubles on the reference's own expression) and a pair's index is a multiply-estimate
// corrected against the exact thresholds — the reference's bin for every pair, without a division in
// the loop (the literal __ddiv_rn chain, ~110 instructions per pair, remains as the fallback for values
// outside the tables).  One thread block = 128 list-a particles of one event against the partner
// event's list b staged through shared memory; the [Bnpts][20] histogram is privatised per block
// in shared memory (u32) and flushed with one u64 atomic per non-empty bin.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hbt_b200.h"

namespace {

constexpr int kTile = 128;

// phi index k = floor((dphi - Bphi_min)/dphi_bin) for |phi_p| <= pi and a rotation in [0, 2 pi): k in [-15, 45]
constexpr int kPhiLo = -24, kPhiN = 80;   // thresholds for k = kPhiLo .. kPhiLo + kPhiN (inclusive upper sentinel)

struct BfGrid {
    int nrap;         // Bnpts
    double rap_min;   // Brap_min = -|Brap_max| - drap/2
    double drap;
    double phi_min;   // -pi/2
    double dphi;      // 2 pi / 20
    double inv_dphi, inv_drap;  // estimates only
    // binary32 decision (see bf_pairs): bin units u_y = (dy - rap_min) / drap, u_phi = (dphi + rotation - phi_min) / dphi
    float rap_min_f, inv_drap_f, inv_dphi_f;
    float ky, k0y;  // |u_y(float) - u_y| <= ky (|y_a| + |y_b|) + k0y
    float kp;       // |u_phi(float) - u_phi| <= kp (|phi_a| + |phi_b|) + k0p, k0p = kp0a |rotation - phi_min| + kp0b
    float kp0a, kp0b;
};

struct BfSeg {  // one event of list a: its particles, the partner event's particles, the rotation
    long long a0, b0;
    int na, nb;
    int block0;  // first thread block of this event
    int pad;
    double rotation;
};

// thr_phi[i]: smallest double x with floor(x / dphi) >= kPhiLo + i (i = 0 .. kPhiN); thr_rap[k]: smallest double x >= 0
// with int(x / drap) >= k (k = 0 .. nrap)
// The reference's bin of one pair, every operation in binary64 as written there; -1 = not binned.
__device__ __forceinline__ int bf_pair_exact(const BfGrid &g, const double2 pa, const double2 pb, const double rotation,
                                             const double *s_tphi, const double *s_trap, const unsigned char *s_pbin) {
    // :141-151 / :182-192
    const double dy = __dsub_rn(pa.y, pb.y);
    if (fabs(dy) < 1e-10) return -1;
    if (dy < g.rap_min) return -1;
    const double xr = __dsub_rn(dy, g.rap_min);  // >= 0
    int y_idx = __double2int_rz(xr * g.inv_drap);
    if (y_idx >= 0 && y_idx <= g.nrap) {  // estimate within one step of int(RN(xr / drap)): settle it on the thresholds
        if (y_idx > 0 && xr < s_trap[y_idx]) y_idx--;
        else if (y_idx < g.nrap && xr >= s_trap[y_idx + 1]) y_idx++;
    } else {
        y_idx = (xr * g.inv_drap > static_cast<double>(g.nrap)) ? g.nrap : __double2int_rz(__ddiv_rn(xr, g.drap));  // far outside / NaN
    }
    if (!(y_idx >= 0 && y_idx < g.nrap)) return -1;
    // :134-139 / :175-180 — (a.phi - b.phi) + rotation, then floor((. - Bphi_min)/dphi) % Bnphi
    const double dphi_local = __dadd_rn(__dsub_rn(pa.x, pb.x), rotation);
    const double xp = __dsub_rn(dphi_local, g.phi_min);
    int phi_idx;
    const int ke = __double2int_rd(xp * g.inv_dphi) - kPhiLo;  // table position of the estimate
    if (ke >= 1 && ke < kPhiN - 1) {
        int kk = ke;
        if (xp < s_tphi[kk]) kk--;
        else if (xp >= s_tphi[kk + 1]) kk++;
        phi_idx = s_pbin[kk];
    } else {  // outside the tables (|phi_p| > pi, NaN, ...): the reference's expression as written
        phi_idx = static_cast<int>(floor(__ddiv_rn(xp, g.dphi))) % HBT_BF_NPHI;
        if (phi_idx < 0) phi_idx += HBT_BF_NPHI;
    }
    return y_idx * HBT_BF_NPHI + phi_idx;
}

// floor(u) and |u - rint(u)| for |u| < 2^22 (u + 1.5 * 2^23 holds rint(u) in its mantissa); NaN / inf give far = false
__device__ __forceinline__ void bf_classify(float u, float band, int &i, bool &far) {
    const float magic = 12582912.f;
    const float t = u + magic;
    const float d = u - (t - magic);
    i = __float_as_int(t) - 0x4B400000 - (d < 0.f ? 1 : 0);
    far = fabsf(d) > band;
}

// Per pair the bin is first evaluated in binary32 from float copies of (phi, rapidity): in bin units the bin edges are
// the integers, and the float value differs from the exact one by at most `band` (every input rounding and every
// operation accounted for, x2; constants in BfGrid, derivation in hbt_bf_create).  A pair farther than its band from
// every integer in both coordinates (and with |dy| safely above the reference's 1e-10 self-pair test) has the
// reference's bin — or is certainly outside the rapidity range — without any binary64 operation; the others
// (~1e-5 of the pairs) take bf_pair_exact.
__global__ void __launch_bounds__(kTile) bf_pairs(const double2 *__restrict__ a, const double2 *__restrict__ b,
                                                   const BfSeg *__restrict__ segs, int nseg, BfGrid g,
                                                   const double *__restrict__ thr_phi, const double *__restrict__ thr_rap,
                                                   unsigned long long *__restrict__ hist) {
    extern __shared__ __align__(16) unsigned char dyn[];
    double *const s_tphi = reinterpret_cast<double *>(dyn);     // [kPhiN + 1]
    double *const s_trap = s_tphi + (kPhiN + 1);                 // [nrap + 1]
    unsigned *const s_hist = reinterpret_cast<unsigned *>(s_trap + g.nrap + 1);  // [nrap * 20]
    unsigned char *const s_pbin = reinterpret_cast<unsigned char *>(s_hist + g.nrap * HBT_BF_NPHI);  // [kPhiN]: k -> k mod 20
    __shared__ double2 sb[kTile];
    __shared__ float2 sbf[kTile];
    const int t = threadIdx.x;
    const int nbins = g.nrap * HBT_BF_NPHI;
    for (int k = t; k < nbins; k += kTile) s_hist[k] = 0u;
    for (int k = t; k <= kPhiN; k += kTile) s_tphi[k] = thr_phi[k];
    for (int k = t; k <= g.nrap; k += kTile) s_trap[k] = thr_rap[k];
    for (int k = t; k < kPhiN; k += kTile) s_pbin[k] = static_cast<unsigned char>(((kPhiLo + k) % HBT_BF_NPHI + HBT_BF_NPHI) % HBT_BF_NPHI);
    int lo = 0, hi = nseg - 1;  // the event that owns this block
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].block0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
    }
    const BfSeg sg = segs[lo];
    const int ia = (static_cast<int>(blockIdx.x) - sg.block0) * kTile + t;
    const bool live = ia < sg.na;
    double2 pa = make_double2(0.0, 0.0);
    if (live) pa = a[sg.a0 + ia];
    const float paf_x = static_cast<float>(pa.x), paf_y = static_cast<float>(pa.y);
    const double rotm = sg.rotation - g.phi_min;
    const float rotm_f = static_cast<float>(rotm);
    const float k0p = fmaf(fabsf(rotm_f), g.kp0a, g.kp0b);
    const unsigned nrap = static_cast<unsigned>(g.nrap);
    for (int j0 = 0; j0 < sg.nb; j0 += kTile) {
        __syncthreads();
        if (j0 + t < sg.nb) {
            const double2 v = b[sg.b0 + j0 + t];
            sb[t] = v;
            sbf[t] = make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
        }
        __syncthreads();
        if (!live) continue;
        const int nj = min(kTile, sg.nb - j0);
        for (int j = 0; j < nj; j++) {
            const float2 pbf = sbf[j];
            const float ysum = fabsf(paf_y) + fabsf(pbf.y);
            const float dyf = paf_y - pbf.y;
            int iy, ip;
            bool far_y, far_p;
            bf_classify((dyf - g.rap_min_f) * g.inv_drap_f, fmaf(ysum, g.ky, g.k0y), iy, far_y);
            bf_classify(((paf_x - pbf.x) + rotm_f) * g.inv_dphi_f, fmaf(fabsf(paf_x) + fabsf(pbf.x), g.kp, k0p), ip, far_p);
            const bool not_self = fabsf(dyf) > fmaf(ysum, 2.4e-7f, 1e-9f);  // |dy| certainly >= 1e-10 (:141 / :182)
            int bin;
            if (far_y && not_self && static_cast<unsigned>(iy) >= nrap) continue;  // certainly outside the rapidity range
            if (far_y && not_self && far_p && static_cast<unsigned>(ip + 160) < 320u) {
                bin = iy * HBT_BF_NPHI + static_cast<int>(static_cast<unsigned>(ip + 160) % HBT_BF_NPHI);
            } else {
                bin = bf_pair_exact(g, pa, sb[j], sg.rotation, s_tphi, s_trap, s_pbin);
                if (bin < 0) continue;
            }
            atomicAdd(&s_hist[bin], 1u);
        }
    }
    __syncthreads();
    for (int k = t; k < nbins; k += kTile)
        if (s_hist[k]) atomicAdd(&hist[k], static_cast<unsigned long long>(s_hist[k]));
}

}  // namespace

struct hbt_bf {
    int device = 0;
    BfGrid grid{};
    cudaStream_t stream = nullptr;
    unsigned long long *d_hist = nullptr;  // [8][nrap][20]
    double *d_thr = nullptr;               // [kPhiN + 1] phi thresholds, then [nrap + 1] rapidity thresholds
    double2 *d_a = nullptr, *d_b = nullptr;
    BfSeg *d_seg = nullptr;
    size_t cap_a = 0, cap_b = 0, cap_seg = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double kernel_ms = 0.0;
    uint64_t pairs = 0;
    std::string err;
};

namespace {
std::string g_bf_error;

int bf_fail(hbt_bf *bf, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    (bf ? bf->err : g_bf_error) = buf;
    return code;
}

#define BFCU(bf, call)                                                                              \
    do {                                                                                            \
        const cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return bf_fail(bf, HBT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int bf_reserve(hbt_bf *bf, T **p, size_t *cap, size_t n) {
    if (n <= *cap) return HBT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = n + n / 4 + 1024;
    BFCU(bf, cudaMalloc(p, *cap * sizeof(T)));
    return HBT_OK;
}
}  // namespace

namespace {
// smallest double x in [lo, hi] with pred(x) true (pred monotone: false ... false true ... true; pred(hi) true)
template <typename F>
double first_true(double lo, double hi, F pred) {
    if (pred(lo)) return lo;
    // bisection on the ordered bit patterns of non-negative / negative doubles
    auto key = [](double v) { int64_t b; std::memcpy(&b, &v, 8); return b < 0 ? static_cast<int64_t>(0x8000000000000000ull) - b : b; };
    auto val = [](int64_t k) { int64_t b = k < 0 ? static_cast<int64_t>(0x8000000000000000ull) - k : k; double v; std::memcpy(&v, &b, 8); return v; };
    int64_t a = key(lo), c = key(hi);  // pred(val(a)) false, pred(val(c)) true
    while (c - a > 1) {
        const int64_t m = a + (c - a) / 2;
        if (pred(val(m))) c = m; else a = m;
    }
    return val(c);
}
}  // namespace

extern "C" const char *hbt_bf_last_error(const hbt_bf *bf) { return bf ? bf->err.c_str() : g_bf_error.c_str(); }

extern "C" int hbt_bf_create(int32_t Bnpts, double Brap_max, int32_t device, hbt_bf **out) {
    if (!out) return HBT_ERR_INVALID;
    *out = nullptr;
    if (Bnpts < 2 || Bnpts > 2048) return bf_fail(nullptr, HBT_ERR_INVALID, "Bnpts = %d out of range", Bnpts);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return bf_fail(nullptr, HBT_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return bf_fail(nullptr, HBT_ERR_INVALID, "device %d of %d", device, ndev);
    hbt_bf *bf = new hbt_bf;
    bf->device = device;
    // the constructor's grid, evaluated the same way (src/BalanceFunction.cpp:27-36)
    bf->grid.nrap = Bnpts;
    bf->grid.drap = 2. * std::abs(Brap_max) / (Bnpts - 1);
    bf->grid.rap_min = -std::abs(Brap_max) - 0.5 * bf->grid.drap;
    bf->grid.dphi = 2. * M_PI / HBT_BF_NPHI;
    bf->grid.phi_min = -M_PI / 2.;
    *out = bf;
    BFCU(bf, cudaSetDevice(device));
    BFCU(bf, cudaStreamCreateWithFlags(&bf->stream, cudaStreamNonBlocking));
    BFCU(bf, cudaEventCreate(&bf->e0));
    BFCU(bf, cudaEventCreate(&bf->e1));
    const size_t n = static_cast<size_t>(HBT_BF_NHIST) * Bnpts * HBT_BF_NPHI;
    BFCU(bf, cudaMalloc(&bf->d_hist, n * 8));
    BFCU(bf, cudaMemset(bf->d_hist, 0, n * 8));
    // steps of the two index expressions (src/BalanceFunction.cpp:135-137, :149-150), exact: the host divides as
    // the reference does (IEEE, round to nearest)
    bf->grid.inv_dphi = 1.0 / bf->grid.dphi;
    bf->grid.inv_drap = 1.0 / bf->grid.drap;
    {   // Error bands of the binary32 decision, u = 2^-24.  With y_a, y_b rounded to float, dy_f = fl(y_a - y_b),
        // t = fl(dy_f - fl(rap_min)), u_y = fl(t * fl(1/drap)):
        //   |u_y(float) - u_y| <= (u / drap) (3 (|y_a| + |y_b|) + 2 |rap_min|) + 2 u |u_y|
        // and the same with (phi_a, phi_b, rotation - phi_min, dphi).  |u_y| <= nrap + 1 and |u_phi| <= 160 wherever a
        // decision inside the grid is taken (beyond that only "certainly outside" is concluded, which the relative
        // term cannot overturn).  Everything x2, plus 1e-9 for the roundings of the reference's own binary64 chain.
        const double u = 5.9604644775390625e-8, S = 2.0;
        BfGrid &g = bf->grid;
        g.rap_min_f = static_cast<float>(g.rap_min);
        g.inv_drap_f = static_cast<float>(g.inv_drap);
        g.inv_dphi_f = static_cast<float>(g.inv_dphi);
        g.ky = static_cast<float>(S * 3.0 * u * g.inv_drap * 1.001);
        g.k0y = static_cast<float>((S * (2.0 * std::fabs(g.rap_min) * g.inv_drap * u + 2.0 * u * (Bnpts + 1)) + 1e-9) * 1.001);
        g.kp = static_cast<float>(S * 3.0 * u * g.inv_dphi * 1.001);
        g.kp0a = static_cast<float>(S * 2.0 * u * g.inv_dphi * 1.001);
        g.kp0b = static_cast<float>((S * 2.0 * u * 160.0 + 1e-9) * 1.001);
    }
    std::vector<double> thr(static_cast<size_t>(kPhiN + 1) + Bnpts + 1);
    const double dphi = bf->grid.dphi, drap = bf->grid.drap;
    for (int i = 0; i <= kPhiN; i++) {
        const int k = kPhiLo + i;
        thr[i] = first_true((k - 2) * dphi, (k + 2) * dphi, [&](double x) { return std::floor(x / dphi) >= k; });
    }
    thr[kPhiN + 1] = 0.0;  // int(x / drap) >= 0 for every x >= 0
    for (int k = 1; k <= Bnpts; k++)
        thr[kPhiN + 1 + k] = first_true(std::max(0.0, (k - 2) * drap), (k + 2) * drap, [&](double x) { return static_cast<int>(x / drap) >= k; });
    BFCU(bf, cudaMalloc(&bf->d_thr, thr.size() * 8));
    BFCU(bf, cudaMemcpy(bf->d_thr, thr.data(), thr.size() * 8, cudaMemcpyHostToDevice));
    return HBT_OK;
}

extern "C" void hbt_bf_destroy(hbt_bf *bf) {
    if (!bf) return;
    cudaSetDevice(bf->device);
    if (bf->stream) cudaStreamSynchronize(bf->stream);
    cudaFree(bf->d_hist);
    cudaFree(bf->d_thr);
    cudaFree(bf->d_a);
    cudaFree(bf->d_b);
    cudaFree(bf->d_seg);
    if (bf->e0) cudaEventDestroy(bf->e0);
    if (bf->e1) cudaEventDestroy(bf->e1);
    if (bf->stream) cudaStreamDestroy(bf->stream);
    delete bf;
}

extern "C" int hbt_bf_accumulate(hbt_bf *bf, int32_t hist, const double *a, const int64_t *off_a, int32_t nev,
                                 const double *b, const int64_t *off_b, int32_t nev_b, const int32_t *partner,
                                 const double *rotation) {
    if (!bf || hist < 0 || hist >= HBT_BF_NHIST || nev < 0 || nev_b < 0) return bf_fail(bf, HBT_ERR_INVALID, "hbt_bf_accumulate: bad argument");
    if (nev == 0) return HBT_OK;
    if (!off_a || !off_b || !partner || !rotation) return bf_fail(bf, HBT_ERR_INVALID, "hbt_bf_accumulate: null argument");
    const int64_t na = off_a[nev], nb = off_b[nev_b];
    if ((na > 0 && !a) || (nb > 0 && !b)) return bf_fail(bf, HBT_ERR_INVALID, "hbt_bf_accumulate: null particles");
    std::vector<BfSeg> segs;
    long long blocks = 0;
    uint64_t pairs = 0;
    for (int iev = 0; iev < nev; iev++) {
        const int id = partner[iev];
        if (id < 0 || id >= nev_b) return bf_fail(bf, HBT_ERR_INVALID, "partner event %d out of range", id);
        BfSeg s;
        s.a0 = off_a[iev];
        s.na = static_cast<int>(off_a[iev + 1] - off_a[iev]);
        s.b0 = off_b[id];
        s.nb = static_cast<int>(off_b[id + 1] - off_b[id]);
        s.rotation = rotation[iev];
        s.pad = 0;
        if (s.na <= 0 || s.nb <= 0) continue;
        s.block0 = static_cast<int>(blocks);
        blocks += (s.na + kTile - 1) / kTile;
        pairs += static_cast<uint64_t>(s.na) * s.nb;
        segs.push_back(s);
    }
    if (segs.empty()) return HBT_OK;
    if (blocks > 0x7fffffffLL) return bf_fail(bf, HBT_ERR_INVALID, "batch too large: %lld blocks", blocks);
    BFCU(bf, cudaSetDevice(bf->device));
    BFCU(bf, cudaStreamSynchronize(bf->stream));  // the staging buffers below are reused (pageable sources)
    int rc = bf_reserve(bf, &bf->d_a, &bf->cap_a, static_cast<size_t>(na));
    if (rc) return rc;
    rc = bf_reserve(bf, &bf->d_b, &bf->cap_b, static_cast<size_t>(nb));
    if (rc) return rc;
    rc = bf_reserve(bf, &bf->d_seg, &bf->cap_seg, segs.size());
    if (rc) return rc;
    BFCU(bf, cudaMemcpyAsync(bf->d_a, a, static_cast<size_t>(na) * 16, cudaMemcpyHostToDevice, bf->stream));
    BFCU(bf, cudaMemcpyAsync(bf->d_b, b, static_cast<size_t>(nb) * 16, cudaMemcpyHostToDevice, bf->stream));
    BFCU(bf, cudaMemcpyAsync(bf->d_seg, segs.data(), segs.size() * sizeof(BfSeg), cudaMemcpyHostToDevice, bf->stream));
    const size_t nbins = static_cast<size_t>(bf->grid.nrap) * HBT_BF_NPHI;
    BFCU(bf, cudaEventRecord(bf->e0, bf->stream));
    const size_t smem = (static_cast<size_t>(kPhiN + 1) + bf->grid.nrap + 1) * 8 + nbins * sizeof(unsigned) + kPhiN;
    bf_pairs<<<static_cast<unsigned>(blocks), kTile, smem, bf->stream>>>(
        bf->d_a, bf->d_b, bf->d_seg, static_cast<int>(segs.size()), bf->grid, bf->d_thr, bf->d_thr + kPhiN + 1,
        bf->d_hist + static_cast<size_t>(hist) * nbins);
    BFCU(bf, cudaGetLastError());
    BFCU(bf, cudaEventRecord(bf->e1, bf->stream));
    BFCU(bf, cudaStreamSynchronize(bf->stream));  // segs / a / b may be freed by the caller on return
    float ms = 0.f;
    BFCU(bf, cudaEventElapsedTime(&ms, bf->e0, bf->e1));
    bf->kernel_ms += ms;
    bf->pairs += pairs;
    return HBT_OK;
}

extern "C" int hbt_bf_read(hbt_bf *bf, uint64_t *hist) {
    if (!bf || !hist) return HBT_ERR_INVALID;
    BFCU(bf, cudaSetDevice(bf->device));
    BFCU(bf, cudaStreamSynchronize(bf->stream));
    const size_t n = static_cast<size_t>(HBT_BF_NHIST) * bf->grid.nrap * HBT_BF_NPHI;
    BFCU(bf, cudaMemcpy(hist, bf->d_hist, n * 8, cudaMemcpyDeviceToHost));
    return HBT_OK;
}

extern "C" int hbt_bf_get_timers(hbt_bf *bf, double *kernel_ms, uint64_t *pairs) {
    if (!bf) return HBT_ERR_INVALID;
    if (kernel_ms) *kernel_ms = bf->kernel_ms;
    if (pairs) *pairs = bf->pairs;
    return HBT_OK;
}
