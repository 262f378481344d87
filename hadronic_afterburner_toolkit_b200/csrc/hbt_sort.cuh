// Momentum-space ordering of a same-event particle list (sm_100a).
//
// The accepted region of a pair is small in transverse momentum: |q_out|,|q_side| <= W means
// |p_T,i - p_T,j| <= sqrt(2) W (0.29 GeV for the benchmark grid) while the particles spread over
// ~2 GeV.  After sorting the list along a Morton (Z-order) curve in (px,py), a tile of 128
// consecutive particles covers a small cell, and most tile pairs can be discarded from their
// bounding boxes alone.  The reference's pair orientation (q = p_i - p_j with i before j in the
// gather order, src/HBT_correlation.cpp:291-301,327-330) is kept through the original index.
#ifndef HBT_SORT_CUH_
#define HBT_SORT_CUH_

#include <cub/cub.cuh>

#include "hbt_common.h"

#define HBT_BBOX_TILE 32

struct HbtBBox {
    double xlo, xhi, ylo, yhi;
};

__device__ __forceinline__ unsigned hbt_spread16(unsigned v) {
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// max(|px|, |py|) over the list, as the bit pattern of a non-negative float (atomicMax-able)
__global__ void hbt_sort_range(const double *__restrict__ p, long long n, unsigned *__restrict__ rmax) {
    float m = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double2 v = *reinterpret_cast<const double2 *>(p + 8 * i);
        const float a = fmaxf(fabsf(static_cast<float>(v.x)), fabsf(static_cast<float>(v.y)));
        if (a == a && a < 3.0e38f) m = fmaxf(m, a);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(rmax, __float_as_uint(m));
}

// rmax: device word written by hbt_sort_range, or null when the host already knows the range (r_host)
__global__ void hbt_sort_keys(const double *__restrict__ p, long long n, const unsigned *__restrict__ rmax, float r_host,
                              unsigned *__restrict__ keys, unsigned *__restrict__ idx) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float R = fmaxf(rmax ? __uint_as_float(*rmax) : r_host, 1e-30f) * 1.0001f;
    const float scale = 32767.5f / R;
    const double2 v = *reinterpret_cast<const double2 *>(p + 8 * i);
    const float fx = (static_cast<float>(v.x) + R) * scale, fy = (static_cast<float>(v.y) + R) * scale;
    const unsigned ux = static_cast<unsigned>(fminf(fmaxf(fx, 0.f), 65535.f));  // NaN -> 0
    const unsigned uy = static_cast<unsigned>(fminf(fmaxf(fy, 0.f), 65535.f));
    keys[i] = hbt_spread16(ux) | (hbt_spread16(uy) << 1);
    idx[i] = static_cast<unsigned>(i);
}

// ---- per-event pT order for the mixed-event loops ---------------------------------------------
// key = (event index << 32) | bits of float(pT^2) (non-negative floats order like their bit
// patterns; NaN sorts last): one radix sort orders every event of the buffer by pT, events stay
// where they are.  evoff[0..nev] are the event boundaries (in particles) of the buffer.
__global__ void hbt_mix_keys(const double *__restrict__ p, long long n, const long long *__restrict__ evoff, int nev,
                             unsigned long long *__restrict__ keys, unsigned *__restrict__ idx) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nev - 1;  // last event with evoff[e] <= i
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (evoff[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const double2 v = *reinterpret_cast<const double2 *>(p + 8 * i);
    const float pt2 = static_cast<float>(v.x * v.x + v.y * v.y);
    keys[i] = (static_cast<unsigned long long>(lo) << 32) | (__float_as_uint(pt2) & 0x7fffffffu);
    idx[i] = static_cast<unsigned>(i);
}

// sorted[k] = p[idx[k]] (64 bytes each, two threads per particle would not pay: L2 resident)
__global__ void hbt_sort_gather(const double *__restrict__ p, const unsigned *__restrict__ idx, long long n,
                                double *__restrict__ sorted) {
    const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long k = t >> 2;
    const int part = static_cast<int>(t & 3);
    if (k >= n) return;
    const double2 *src = reinterpret_cast<const double2 *>(p + 8ll * idx[k]);
    reinterpret_cast<double2 *>(sorted + 8 * k)[part] = src[part];
}

// bounding box in (px,py) of every tile of HBT_BBOX_TILE (= one warp) consecutive sorted particles
// orig (may be null): original index of each sorted slot; HBT_PAD_SLOT marks the padding between the batches of a
// multi-batch list, which takes no part in a box
#define HBT_PAD_SLOT 0xffffffffu
__global__ void __launch_bounds__(HBT_BBOX_TILE) hbt_sort_bbox(const double *__restrict__ sorted, long long n,
                                                              HbtBBox *__restrict__ bbox, const unsigned *__restrict__ orig = nullptr) {
    const long long k = blockIdx.x * static_cast<long long>(HBT_BBOX_TILE) + threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double xlo = inf, xhi = -inf, ylo = inf, yhi = -inf;
    const bool pad = orig && k < n && orig[k] == HBT_PAD_SLOT;
    const bool all_pad = __all_sync(0xffffffffu, pad || k >= n) && orig;
    if (k < n && !pad) {
        const double2 v = *reinterpret_cast<const double2 *>(sorted + 8 * k);
        xlo = xhi = v.x;
        ylo = yhi = v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
        xlo = fmin(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
        xhi = fmax(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
        ylo = fmin(ylo, __shfl_xor_sync(0xffffffffu, ylo, o));
        yhi = fmax(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
    }
    if (threadIdx.x == 0) {
        // a NaN coordinate would poison the box test: widen to everything (never culled)
        if (!all_pad && (!(xlo <= xhi) || !(ylo <= yhi))) { xlo = -inf; xhi = inf; ylo = -inf; yhi = inf; }
        bbox[blockIdx.x] = {xlo, xhi, ylo, yhi};  // (a tile of padding only: the empty box, which merges neutrally)
    }
}

// ---- several small batches in one launch ------------------------------------------------------------
// Oversample groups of 10-20 events (15 000-30 000 particles) hold ~0.3 ms of pair work each: the ~8 helper launches
// and the ramp / tail of a persistent kernel per batch cost as much.  Up to HBT_MULTI_MAX batches are therefore
// submitted as ONE launch: their same-event lists are Morton-sorted with the batch number on top of the key, each
// batch padded with NaN particles to a multiple of 64 slots (a work unit then never holds particles of two
// batches, and NaN rows fail the K_T cut), and the cull kernel keeps only units whose row and tile belong to the
// same batch.  The batches may live in different device buffers (src[b]); cbase = prefix of their particle
// counts ("logical" concatenated index), pbase = prefix of the padded counts (slots of the sorted copy).
#define HBT_MULTI_MAX 64
struct HbtMulti {
    int nb;
    int pad;
    const double *src[HBT_MULTI_MAX];
    long long cbase[HBT_MULTI_MAX + 1];
    long long pbase[HBT_MULTI_MAX + 1];
};

__device__ __forceinline__ int hbt_multi_find(const long long *base, int nb, long long i) {
    int lo = 0, hi = nb - 1;  // last b with base[b] <= i
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (base[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void hbt_multi_range(const HbtMulti M, unsigned *__restrict__ rmax) {
    float m = 0.f;
    const long long n = M.cbase[M.nb];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = hbt_multi_find(M.cbase, M.nb, i);
        const double2 v = *reinterpret_cast<const double2 *>(M.src[b] + 8 * (i - M.cbase[b]));
        const float a = fmaxf(fabsf(static_cast<float>(v.x)), fabsf(static_cast<float>(v.y)));
        if (a == a && a < 3.0e38f) m = fmaxf(m, a);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(rmax, __float_as_uint(m));
}

// one key per SLOT of the padded layout: (batch << 24) | top 24 bits of the Morton key; padding slots sort to the end
// of their batch and carry HBT_PAD_SLOT as index, real ones their logical index
__global__ void hbt_multi_keys(const HbtMulti M, const unsigned *__restrict__ rmax, float r_host,
                               unsigned *__restrict__ keys, unsigned *__restrict__ idx) {
    const long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (s >= M.pbase[M.nb]) return;
    const int b = hbt_multi_find(M.pbase, M.nb, s);
    const long long l = s - M.pbase[b], nb_ = M.cbase[b + 1] - M.cbase[b];
    if (l >= nb_) {
        keys[s] = (static_cast<unsigned>(b) << 24) | 0xffffffu;
        idx[s] = HBT_PAD_SLOT;
        return;
    }
    const float R = fmaxf(rmax ? __uint_as_float(*rmax) : r_host, 1e-30f) * 1.0001f;
    const float scale = 32767.5f / R;
    const double2 v = *reinterpret_cast<const double2 *>(M.src[b] + 8 * l);
    const float fx = (static_cast<float>(v.x) + R) * scale, fy = (static_cast<float>(v.y) + R) * scale;
    const unsigned ux = static_cast<unsigned>(fminf(fmaxf(fx, 0.f), 65535.f));  // NaN -> 0
    const unsigned uy = static_cast<unsigned>(fminf(fmaxf(fy, 0.f), 65535.f));
    keys[s] = (static_cast<unsigned>(b) << 24) | ((hbt_spread16(ux) | (hbt_spread16(uy) << 1)) >> 8);
    idx[s] = static_cast<unsigned>(M.cbase[b] + l);
}

// per-event pT keys over the logical concatenation (evoff: its event boundaries)
__global__ void hbt_multi_mix_keys(const HbtMulti M, const long long *__restrict__ evoff, int nev,
                                   unsigned long long *__restrict__ keys, unsigned *__restrict__ idx) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= M.cbase[M.nb]) return;
    const int e = hbt_multi_find(evoff, nev, i);
    const int b = hbt_multi_find(M.cbase, M.nb, i);
    const double2 v = *reinterpret_cast<const double2 *>(M.src[b] + 8 * (i - M.cbase[b]));
    const float pt2 = static_cast<float>(v.x * v.x + v.y * v.y);
    keys[i] = (static_cast<unsigned long long>(e) << 32) | (__float_as_uint(pt2) & 0x7fffffffu);
    idx[i] = static_cast<unsigned>(i);
}

// sorted[k] = particle with logical index idx[k] (NaN particle for a padding slot)
__global__ void hbt_multi_gather(const HbtMulti M, const unsigned *__restrict__ idx, long long n, double *__restrict__ sorted) {
    const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long k = t >> 2;
    const int part = static_cast<int>(t & 3);
    if (k >= n) return;
    const unsigned i = idx[k];
    double2 v;
    if (i == HBT_PAD_SLOT) {
        v.x = v.y = __longlong_as_double(0x7ff8000000000000ll);
    } else {
        const int b = hbt_multi_find(M.cbase, M.nb, i);
        v = reinterpret_cast<const double2 *>(M.src[b] + 8 * (static_cast<long long>(i) - M.cbase[b]))[part];
    }
    reinterpret_cast<double2 *>(sorted + 8 * k)[part] = v;
}

// true when NO pair between the two boxes can pass both the K_T cut and the q_out/q_side
// windows.  W2 = W^2 of the symmetric window, k2lo/k2hi = 4 KT_min^2 / 4 KT_max^2.
__device__ __forceinline__ bool hbt_boxes_culled(const HbtBBox &a, const HbtBBox &b, double W2, double k2lo, double k2hi) {
    // smallest possible |q_T|^2: q_out^2 + q_side^2 = |q_T|^2, so |q_T|^2 > 2 W^2 rules out the pair
    const double gx = fmax(0.0, fmax(a.xlo - b.xhi, b.xlo - a.xhi));
    const double gy = fmax(0.0, fmax(a.ylo - b.yhi, b.ylo - a.yhi));
    if (gx * gx + gy * gy > 2.0 * W2 * (1.0 + 1e-9)) return true;
    // range of s = p_i + p_j: k2 = |s|^2 = 4 K_T^2
    const double sxl = a.xlo + b.xlo, sxh = a.xhi + b.xhi, syl = a.ylo + b.ylo, syh = a.yhi + b.yhi;
    const double nx = (sxl <= 0.0 && sxh >= 0.0) ? 0.0 : fmin(fabs(sxl), fabs(sxh));
    const double ny = (syl <= 0.0 && syh >= 0.0) ? 0.0 : fmin(fabs(syl), fabs(syh));
    const double mx = fmax(fabs(sxl), fabs(sxh)), my = fmax(fabs(syl), fabs(syh));
    if (mx * mx + my * my < k2lo * (1.0 - 1e-9)) return true;
    if (nx * nx + ny * ny > k2hi * (1.0 + 1e-9)) return true;
    return false;
}

#endif  // HBT_SORT_CUH_
