// libhbt_b200: context, staging, launches and the C ABI (include/hbt_b200.h).
//
// One context = one B200.  Per batch ("oversample group") the host hands in the
// rapidity-cut particle lists and the mixed-event plan; particles go through a ring of
// pinned staging slots to HBM on a copy stream while the previous batch's pair kernels
// run on the compute stream; histograms live in HBM (L2-resident at the benchmark sizes)
// for the whole analysis and are read back once.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "hbt_common.h"
#include "hbt_kernels_v1.cuh"
#ifdef HBT_HAVE_V2
#include "hbt_kernels_v3.cuh"
#include "hbt_kernels_v4.cuh"
#else
#define HBT_V3_SUB_MIXED 128
#define HBT_V3_TJ_MIXED 128
#define HBT_V4_TI 128
#endif

namespace {

thread_local std::string g_create_error;

constexpr int kSlots = 4;
constexpr int kLanes = 2;  // compute streams that take production batches in turn (HBT_B200_LANES / HBT_OPT_LANES)
constexpr unsigned kDeferredCapacity = 1u << 16;
constexpr int kTileV1 = 128;

struct Slot {
    double *h_p = nullptr, *d_p = nullptr;  // pinned staging / device copy of the particle lists
    size_t cap = 0;                         // particles
    HbtMixSeg *h_seg = nullptr, *d_seg = nullptr;
    size_t cap_seg = 0;
    cudaEvent_t uploaded = nullptr, done = nullptr;
    bool in_flight = false;
};

// One compute lane = a stream plus the scratch a production launch needs (unit pop counter, culled
// unit list, Morton-sorted copy of the same-event list), so that consecutive batches can be in
// flight together: the sort / cull helpers of batch k+1 and its first units run while the last
// persistent warps of batch k are still draining.  Every accumulation is a RED into the
// context's histograms, so the batches commute.  Lane 0's stream is ctx->compute, which also
// carries everything ordered (cap replay, reductions, the all-reduce, the stopwatch).
struct Lane {
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;                 // the v4 mixed-event kernel of a whole batch, next to the same-event kernel (co-run)
    cudaEvent_t fork = nullptr, join = nullptr;  // side waits for fork (the batch's preparation), stream waits for join (side's kernel)
    cudaEvent_t tail = nullptr;                  // last submission (joins into ctx->compute)
    // page-locked staging of the small per-launch tables (event offsets, segments, row limits).  cudaMemcpyAsync from
    // PAGEABLE memory first waits for everything queued on its stream — the host would sit out the lane's previous pair
    // kernel at every launch; from here the copy is only enqueued.
    unsigned char *h_meta = nullptr;
    size_t h_meta_cap = 0, h_meta_used = 0;
    cudaEvent_t meta_done = nullptr;             // the copies out of h_meta have been made (it may be overwritten)
    bool meta_pending = false;
    bool busy = false;                           // something was enqueued since the last join
    unsigned *d_work = nullptr;                  // [0] unit pop counter, [1] number of units (culled list)
    unsigned *d_units = nullptr;                 // surviving units of the sorted same-event list
    size_t units_cap = 0;
    unsigned *sort_keys[2] = {nullptr, nullptr}, *sort_idx[2] = {nullptr, nullptr}, *sort_rmax = nullptr;
    double *sort_p = nullptr;
    HbtBBox *sort_bbox = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0, sort_cap = 0;
    // per-event pT-sorted copy of the buffer the mixed-event loops read (production mode)
    unsigned long long *mix_keys[2] = {nullptr, nullptr};
    unsigned *mix_idx[2] = {nullptr, nullptr};
    double *mix_p = nullptr;
    long long *mix_off = nullptr;
    void *mix_tmp = nullptr;
    size_t mix_tmp_bytes = 0, mix_cap = 0, mix_off_cap = 0;
    // several small batches in one launch (flush_pending): tile limit of every row, the launch's segments
    unsigned *row_end = nullptr;
    size_t row_end_cap = 0;
    HbtMixSeg *mseg = nullptr;
    size_t mseg_cap = 0;
};

// Small batches waiting to be launched together (see hbt_sort.cuh, "several small batches in one launch").
struct PendBatch {
    const double *d_src;  // device address of the batch's particles
    int64_t n;
};
struct Pending {
    std::vector<PendBatch> b;
    std::vector<long long> evoff;   // event boundaries of the logical concatenation (evoff[0] = 0)
    std::vector<HbtMixSeg> segs;    // mixed-event segments, offsets in the logical concatenation
    unsigned long long pairs_same = 0, pairs_mixed = 0;
    long long nblocks = 0, units_bound = 0;
    int64_t n_logical = 0, n_padded = 0;
    bool do_mixed = false, host = false;
    bool direct = false;            // host batches in page-locked caller buffers: uploaded one by one as they arrive, no staging copy
    Slot *slot = nullptr;           // host-staged batches: their common staging slot (ONE upload at the flush)
    size_t staged = 0;              // particles staged in slot->h_p so far
    float host_range = 0.f;         // max(|px|, |py|) of the staged particles
    double psi_ref = 0.;
};

struct TimerRec {
    cudaEvent_t start, stop;
    int kind;           // 0 same, 1 mixed, 2 fused (one kernel working through both unit lists)
    double frac_same;   // fused: share of the launch's pairs that are same-event pairs
};

// ---- NCCL, loaded lazily so that the library itself has no link-time dependency ----------
struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return false;
        }
#define HBT_SYM(field, sym)                                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));                 \
    if (!field) {                                                                  \
        error = std::string("libnccl lacks ") + sym;                               \
        return false;                                                              \
    }
        HBT_SYM(GetUniqueId, "ncclGetUniqueId")
        HBT_SYM(CommInitRank, "ncclCommInitRank")
        HBT_SYM(CommInitAll, "ncclCommInitAll")
        HBT_SYM(CommDestroy, "ncclCommDestroy")
        HBT_SYM(AllReduce, "ncclAllReduce")
        HBT_SYM(GroupStart, "ncclGroupStart")
        HBT_SYM(GroupEnd, "ncclGroupEnd")
        HBT_SYM(GetErrorString, "ncclGetErrorString")
#undef HBT_SYM
        return true;
    }
};
NcclApi g_nccl;

}  // namespace

struct hbt_ctx {
    int device = 0;
    hbt_params params{};
    HbtGrid grid{};
    HbtAccum acc{};
    unsigned long long *blob_u64 = nullptr;  // all integer accumulators, contiguous
    double *blob_f64 = nullptr;              // all floating sums, contiguous
    unsigned long long *red_u64 = nullptr;   // all-reduced copies (allocated on first use)
    double *red_f64 = nullptr;
    bool reduced = false;                    // red_* hold the sum over ranks of the current state
    size_t n_u64 = 0, n_f64 = 0;
    cudaStream_t compute = nullptr, copy = nullptr;  // compute == lanes[0].stream
    Lane lanes[kLanes];
    int n_lanes = kLanes, next_lane = 0;
    cudaEvent_t epoch = nullptr;    // time origin of the launch timers, renewed whenever the context is idle
    double covered_ms = 0.;         // end (since epoch) of the latest-ending launch already counted
    // mixed-event loops read a per-event pT-sorted copy: 0 never, 1 when the batch has enough mixed-event pairs to
    // pay for the sort kernels (default), 2 always (HBT_B200_PTSORT / HBT_OPT_PTSORT)
    int ptsort = 1;
    unsigned long long ptsort_min_pairs = 500000000ull;
    std::vector<long long> evoff;  // event boundaries of the buffer being sorted
    bool fuse = true;  // whole batches run the fused same+mixed kernel (HBT_B200_FUSE=0 / HBT_OPT_FUSE: separate kernels)
    // production mixed-event loops run hbt_pairs_v4_mixed (binary32 tiles, 24 resident warps; HBT_B200_MIXED4=0: the v3 kernel)
    bool mixed4 = true;
    size_t coalesce_host = 8;  // host batches per launch (HBT_B200_COALESCE_HOST)
    // a whole batch = the same-event kernel and the v4 mixed-event kernel, each with its own registers / shared memory /
    // resident warps (HBT_B200_SPLIT=0: the fused v3 kernel, one allocation for both)
    bool split = true;
    // ... next to each other (launch_split_pair): the same-event kernel alone is held by its reductions into the L2 (5
    // per accepted pair), the mixed-event kernel by instruction issue, so the mixed-event warps fill the issue slots the
    // same-event warps leave.  corun_same / corun_mixed = resident warps per SM of the first launch of each kernel; a
    // second launch of each, behind the OTHER kernel in stream order, brings it to its full grid once that one is done.
    // HBT_B200_CORUN=0: one after the other; HBT_B200_CORUN_SAME / HBT_B200_CORUN_MIXED
    bool corun = true;
    int corun_same = 12, corun_mixed = 8;
    // small production batches are collected and launched together (HBT_B200_COALESCE=0 / HBT_OPT_COALESCE: one launch each)
    bool coalesce = true;
    Pending pend;
    Slot slots[kSlots];
    int next_slot = 0;
    std::vector<TimerRec> timers;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
    double same_ms = 0., mixed_ms = 0.;
    uint64_t same_launches = 0, mixed_launches = 0, kernel_launches = 0;
    HbtDeferred *h_deferred = nullptr;  // pinned
    unsigned *h_defcount = nullptr;     // pinned [2]
    HbtCorrection *d_corr = nullptr;
    uint64_t deferred_total = 0;
    cudaEvent_t sw0 = nullptr, sw1 = nullptr;  // hbt_timer_start / hbt_timer_stop
#ifdef HBT_HAVE_V2
    V2Const v2c{};
    V2Dev *d_dv = nullptr;  // global-memory copy for the non-inlined device functions
#endif
    int kernel_version = 2;
    bool direct_upload = true;                   // page-locked caller buffers are uploaded without a staging copy (HBT_B200_DIRECT=0: never)
    ncclComm_t comm = nullptr;
    int nranks = 1;
    // needed_number_of_pairs bookkeeping (ordered cap)
    std::vector<uint64_t> exact_num, exact_den;  // per-slab accepted pairs as of the last refresh
    uint64_t pending_num = 0, pending_den = 0;   // pairs submitted since (upper bound of what they add)
    // accepted pairs per cap channel accumulated by the OTHER contexts of a group (hbt_cap_set_foreign): the cap is
    // cumulative over all batches of the analysis, whichever GPU took them
    std::vector<uint64_t> foreign_num, foreign_den;
    std::vector<unsigned char> closed;           // [2*nslab] slabs whose counter exceeds the cap
    unsigned char *d_closed = nullptr;
    bool any_closed = false;
    // production mode of the v2 same-event kernel: Morton-sorted copy of the list + tile boxes
    bool stats = false;                          // exact stage populations B, C, D (no culling)
    int n_sm = 148;
    int occ_same = 12, occ_same_stats = 12, occ_mixed = 12, occ_mixed_stats = 12, occ_fused = 12, occ_mixed4 = 12;  // resident warps per SM
    int occ_same_q = 12, occ_mixed_q = 12;       // q_inv mode
    double *d_qinv_thr = nullptr;                // q_inv mode: exact bin thresholds in s space (V2Const::qinv_thr)
    unsigned long long *d_qrep_u64 = nullptr;    // q_inv mode: replicated accumulators (V2Const::qrep_*)
    double *d_qrep_f64 = nullptr;
    std::vector<int> row_item0;                  // same-event unit prefix per row (instrumented runs)
    int *d_rows = nullptr;
    size_t d_rows_cap = 0;
    unsigned long long *snap_u64 = nullptr;      // rollback copies of the accumulators
    double *snap_f64 = nullptr;
    std::string err;
};

namespace {

int fail(hbt_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define CU(ctx, call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(ctx, HBT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

int ensure_slot(hbt_ctx *ctx, Slot &s, size_t particles, size_t segs) {
    if (particles > s.cap) {
        if (s.h_p) cudaFreeHost(s.h_p);
        if (s.d_p) cudaFree(s.d_p);
        s.h_p = nullptr; s.d_p = nullptr;
        size_t cap = std::max<size_t>(particles + particles / 4, 4096);
        CU(ctx, cudaMallocHost(&s.h_p, cap * 64));
        CU(ctx, cudaMalloc(&s.d_p, cap * 64));
        s.cap = cap;
    }
    if (segs > s.cap_seg) {
        if (s.h_seg) cudaFreeHost(s.h_seg);
        if (s.d_seg) cudaFree(s.d_seg);
        s.h_seg = nullptr; s.d_seg = nullptr;
        size_t cap = std::max<size_t>(segs + segs / 4, 256);
        CU(ctx, cudaMallocHost(&s.h_seg, cap * sizeof(HbtMixSeg)));
        CU(ctx, cudaMalloc(&s.d_seg, cap * sizeof(HbtMixSeg)));
        s.cap_seg = cap;
    }
    return HBT_OK;
}

// take the next staging slot; waits (host side) until its previous batch has finished
int acquire_slot(hbt_ctx *ctx, Slot **out) {
    Slot &s = ctx->slots[ctx->next_slot];
    ctx->next_slot = (ctx->next_slot + 1) % kSlots;
    if (s.in_flight) {
        CU(ctx, cudaEventSynchronize(s.done));
        s.in_flight = false;
    }
    *out = &s;
    return HBT_OK;
}

int get_event_pair(hbt_ctx *ctx, cudaEvent_t *a, cudaEvent_t *b) {
    if (!ctx->event_pool.empty()) {
        *a = ctx->event_pool.back().first;
        *b = ctx->event_pool.back().second;
        ctx->event_pool.pop_back();
        return HBT_OK;
    }
    CU(ctx, cudaEventCreate(a));
    CU(ctx, cudaEventCreate(b));
    return HBT_OK;
}

// fold finished timer records into the totals (all of them when `all`, which requires the
// streams to be idle or blocks until each stop event has fired).  Launches of different lanes
// overlap in time, so a record contributes only the part of [start, stop] that lies after the
// end of everything counted before it: same_ms + mixed_ms is the time during which at least one
// pair launch was running.  Times are taken against an epoch event that hbt_synchronize renews
// whenever the context is idle (float milliseconds: microsecond resolution over minutes).
int drain_timers(hbt_ctx *ctx, bool all) {
    size_t keep = all ? 0 : 64;
    while (ctx->timers.size() > keep) {
        TimerRec r = ctx->timers.front();
        CU(ctx, cudaEventSynchronize(r.stop));
        float t0 = 0.f, t1 = 0.f;
        CU(ctx, cudaEventElapsedTime(&t0, ctx->epoch, r.start));
        CU(ctx, cudaEventElapsedTime(&t1, ctx->epoch, r.stop));
        const double a = std::max<double>(t0, ctx->covered_ms), b = t1;
        const double ms = b > a ? b - a : 0.0;
        ctx->covered_ms = std::max<double>(ctx->covered_ms, b);
        if (r.kind == 2) {
            ctx->same_ms += ms * r.frac_same;
            ctx->mixed_ms += ms * (1.0 - r.frac_same);
        } else {
            (r.kind == 0 ? ctx->same_ms : ctx->mixed_ms) += ms;
        }
        ctx->event_pool.emplace_back(r.start, r.stop);
        ctx->timers.erase(ctx->timers.begin());
    }
    return HBT_OK;
}

// every lane's work so far happens-before whatever is enqueued on ctx->compute next
int join_lanes(hbt_ctx *ctx) {
    for (int l = 1; l < kLanes; l++) {
        Lane &L = ctx->lanes[l];
        if (!L.busy) continue;
        CU(ctx, cudaEventRecord(L.tail, L.stream));
        CU(ctx, cudaStreamWaitEvent(ctx->compute, L.tail, 0));
        L.busy = false;
    }
    return HBT_OK;
}

// the lane that takes the next production batch
Lane &next_lane(hbt_ctx *ctx) {
    Lane &L = ctx->lanes[ctx->next_lane];
    ctx->next_lane = (ctx->next_lane + 1) % std::max(1, ctx->n_lanes);
    L.busy = true;
    return L;
}

// production batches alternate between the lanes; instrumented runs, the literal kernels and
// everything near the pair cap stay on lane 0 (= ctx->compute)
Lane &pick_lane(hbt_ctx *ctx, bool ordered = false) {
    if (ordered || ctx->n_lanes < 2 || ctx->stats || ctx->kernel_version == 1) return ctx->lanes[0];
    return next_lane(ctx);
}

size_t dyn_smem_bytes(const HbtGrid &g) {
    const size_t nslab_pad = (g.nslab + 1) & ~1;
    const size_t nqi = g.qinv ? static_cast<size_t>(g.nKT) * g.nq : 0;
    return nslab_pad * 4 + nqi * (8 + 8 + 4) + (g.qinv ? g.nKT * 4 : 0) + 16;
}

// ---- launches ------------------------------------------------------------------------
// small host tables go to the device through the lane's page-locked staging buffer: meta_begin (once per launch, waits
// for the previous launch's copies out of the buffer — done long ago unless the GPU is a whole launch behind), any
// number of meta_upload, meta_end
int meta_begin(hbt_ctx *ctx, Lane &L, size_t bytes_needed) {
    if (L.meta_pending) {
        CU(ctx, cudaEventSynchronize(L.meta_done));
        L.meta_pending = false;
    }
    if (bytes_needed > L.h_meta_cap) {
        if (L.h_meta) cudaFreeHost(L.h_meta);
        L.h_meta = nullptr;
        L.h_meta_cap = std::max<size_t>(bytes_needed * 2, 1u << 16);
        CU(ctx, cudaHostAlloc(&L.h_meta, L.h_meta_cap, cudaHostAllocDefault));
    }
    L.h_meta_used = 0;
    return HBT_OK;
}
int meta_upload(hbt_ctx *ctx, Lane &L, void *dst_dev, const void *src, size_t bytes) {
    if (!bytes) return HBT_OK;
    if (L.h_meta_used + bytes > L.h_meta_cap) return fail(ctx, HBT_ERR_STATE, "staging buffer of the launch tables too small");
    unsigned char *at = L.h_meta + L.h_meta_used;
    std::memcpy(at, src, bytes);
    L.h_meta_used += (bytes + 15) & ~static_cast<size_t>(15);
    CU(ctx, cudaMemcpyAsync(dst_dev, at, bytes, cudaMemcpyHostToDevice, L.stream));
    return HBT_OK;
}
int meta_end(hbt_ctx *ctx, Lane &L) {
    if (!L.meta_done) CU(ctx, cudaEventCreateWithFlags(&L.meta_done, cudaEventDisableTiming));
    CU(ctx, cudaEventRecord(L.meta_done, L.stream));
    L.meta_pending = true;
    return HBT_OK;
}

#ifdef HBT_HAVE_V2
int ensure_work(hbt_ctx *ctx, Lane &L) {
    if (!L.d_work) CU(ctx, cudaMalloc(&L.d_work, 16));  // [0] pop counter, [1] units kept by the culling, [2] pop counter of the v4 mixed-event kernel
    return HBT_OK;
}

int ensure_units(hbt_ctx *ctx, Lane &L, long long all_units) {
    if (static_cast<size_t>(all_units) > L.units_cap) {
        CU(ctx, cudaStreamSynchronize(L.stream));  // the previous launch of this lane may still read the list
        cudaFree(L.d_units);
        L.d_units = nullptr;
        L.units_cap = static_cast<size_t>(all_units) + static_cast<size_t>(all_units) / 4;
        CU(ctx, cudaMalloc(&L.d_units, L.units_cap * 4));
    }
    return HBT_OK;
}

// Morton-sort the same-event list on the lane's stream (keys, radix sort of (key, index),
// gather, tile boxes): ~4 small kernels, microseconds against the pair kernel's milliseconds
// host_range: max(|px|, |py|) of the list when the caller has the particles on the host (saves the range
// kernel and its memset), negative otherwise
int ensure_sort_buffers(hbt_ctx *ctx, Lane &L, int64_t n) {
    if (static_cast<size_t>(n) > L.sort_cap) {
        CU(ctx, cudaStreamSynchronize(L.stream));  // the previous launch of this lane may still read the buffers
        for (int k = 0; k < 2; k++) { cudaFree(L.sort_keys[k]); cudaFree(L.sort_idx[k]); }
        cudaFree(L.sort_p); cudaFree(L.sort_bbox); cudaFree(L.sort_tmp);
        for (int k = 0; k < 2; k++) L.sort_keys[k] = L.sort_idx[k] = nullptr;
        L.sort_p = nullptr; L.sort_bbox = nullptr; L.sort_tmp = nullptr; L.sort_cap = 0;
        const size_t cap = static_cast<size_t>(n) + static_cast<size_t>(n) / 4 + 1024;
        for (int k = 0; k < 2; k++) {
            CU(ctx, cudaMalloc(&L.sort_keys[k], cap * 4));
            CU(ctx, cudaMalloc(&L.sort_idx[k], cap * 4));
        }
        CU(ctx, cudaMalloc(&L.sort_p, cap * 64));
        CU(ctx, cudaMalloc(&L.sort_bbox, (cap / HBT_BBOX_TILE + 2) * sizeof(HbtBBox)));
        if (!L.sort_rmax) CU(ctx, cudaMalloc(&L.sort_rmax, 4));
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, L.sort_keys[0], L.sort_keys[1], L.sort_idx[0], L.sort_idx[1],
                                        static_cast<int>(cap), 0, 32, L.stream);
        CU(ctx, cudaMalloc(&L.sort_tmp, bytes));
        L.sort_tmp_bytes = bytes;
        L.sort_cap = cap;
    }
    return HBT_OK;
}

int prepare_sorted(hbt_ctx *ctx, Lane &L, const double *d_p, int64_t n, float host_range = -1.f) {
    int rc0 = ensure_sort_buffers(ctx, L, n);
    if (rc0) return rc0;
    const int th = 256;
    const unsigned nb = static_cast<unsigned>((n + th - 1) / th);
    if (host_range >= 0.f) {
        hbt_sort_keys<<<nb, th, 0, L.stream>>>(d_p, n, nullptr, host_range, L.sort_keys[0], L.sort_idx[0]);
    } else {
        CU(ctx, cudaMemsetAsync(L.sort_rmax, 0, 4, L.stream));
        hbt_sort_range<<<std::min(nb, 1184u), th, 0, L.stream>>>(d_p, n, L.sort_rmax);
        hbt_sort_keys<<<nb, th, 0, L.stream>>>(d_p, n, L.sort_rmax, 0.f, L.sort_keys[0], L.sort_idx[0]);
        ctx->kernel_launches++;
    }
    // the top 24 bits of the Morton key (4096 x 4096 cells) order the tiles as well as all 32: one radix pass fewer
    size_t bytes = L.sort_tmp_bytes;
    CU(ctx, cub::DeviceRadixSort::SortPairs(L.sort_tmp, bytes, L.sort_keys[0], L.sort_keys[1], L.sort_idx[0],
                                            L.sort_idx[1], static_cast<int>(n), 8, 32, L.stream));
    hbt_sort_gather<<<static_cast<unsigned>((4 * n + th - 1) / th), th, 0, L.stream>>>(d_p, L.sort_idx[1], n, L.sort_p);
    hbt_sort_bbox<<<static_cast<unsigned>((n + HBT_BBOX_TILE - 1) / HBT_BBOX_TILE), HBT_BBOX_TILE, 0, L.stream>>>(L.sort_p, n, L.sort_bbox);
    ctx->kernel_launches += 4;  // keys, gather, boxes + the radix sort (counted once); the range kernel above
    CU(ctx, cudaGetLastError());
    return HBT_OK;
}
#endif

#ifdef HBT_HAVE_V2
int ensure_mix_buffers(hbt_ctx *ctx, Lane &L, int64_t n, size_t n_evoff) {
    if (static_cast<size_t>(n) > L.mix_cap) {
        CU(ctx, cudaStreamSynchronize(L.stream));
        for (int k = 0; k < 2; k++) { cudaFree(L.mix_keys[k]); cudaFree(L.mix_idx[k]); L.mix_keys[k] = nullptr; L.mix_idx[k] = nullptr; }
        cudaFree(L.mix_p); cudaFree(L.mix_tmp);
        L.mix_p = nullptr; L.mix_tmp = nullptr; L.mix_cap = 0;
        const size_t cap = static_cast<size_t>(n) + static_cast<size_t>(n) / 4 + 1024;
        for (int k = 0; k < 2; k++) {
            CU(ctx, cudaMalloc(&L.mix_keys[k], cap * 8));
            CU(ctx, cudaMalloc(&L.mix_idx[k], cap * 4));
        }
        CU(ctx, cudaMalloc(&L.mix_p, cap * 64));
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, L.mix_keys[0], L.mix_keys[1], L.mix_idx[0], L.mix_idx[1],
                                        static_cast<int>(cap), 0, 64, L.stream);
        CU(ctx, cudaMalloc(&L.mix_tmp, bytes));
        L.mix_tmp_bytes = bytes;
        L.mix_cap = cap;
    }
    if (n_evoff > L.mix_off_cap) {
        CU(ctx, cudaStreamSynchronize(L.stream));
        cudaFree(L.mix_off);
        L.mix_off = nullptr;
        L.mix_off_cap = n_evoff * 2;
        CU(ctx, cudaMalloc(&L.mix_off, L.mix_off_cap * 8));
    }
    return HBT_OK;
}

// Per-event pT order of the buffer the mixed-event loops read (ctx->evoff: its event boundaries):
// keys, one radix sort of (event, pT^2), gather.  pT is invariant under the partner rotation, so one
// sort per batch serves every (event, partner) segment; v3_run_unit then skips the list-2 particles
// whose pT is farther than the q_out window from the sub-tile's pT range.  Returns the sorted copy.
int prepare_mixed_sorted(hbt_ctx *ctx, Lane &L, const double *d_p, int64_t n, const double **out) {
    *out = d_p;
    const int nev = static_cast<int>(ctx->evoff.size()) - 1;
    if (n < 2 || nev < 1) return HBT_OK;
    int rc0 = ensure_mix_buffers(ctx, L, n, ctx->evoff.size());
    if (rc0) return rc0;
    rc0 = meta_begin(ctx, L, ctx->evoff.size() * 8 + 16);
    if (rc0) return rc0;
    rc0 = meta_upload(ctx, L, L.mix_off, ctx->evoff.data(), ctx->evoff.size() * 8);
    if (rc0) return rc0;
    rc0 = meta_end(ctx, L);
    if (rc0) return rc0;
    const int th = 256;
    hbt_mix_keys<<<static_cast<unsigned>((n + th - 1) / th), th, 0, L.stream>>>(d_p, n, L.mix_off, nev, L.mix_keys[0], L.mix_idx[0]);
    int ev_bits = 1;
    while ((1ll << ev_bits) < nev) ev_bits++;
    size_t bytes = L.mix_tmp_bytes;
    CU(ctx, cub::DeviceRadixSort::SortPairs(L.mix_tmp, bytes, L.mix_keys[0], L.mix_keys[1], L.mix_idx[0], L.mix_idx[1],
                                            static_cast<int>(n), 8, 32 + ev_bits, L.stream));  // (pT^2 to 2^-15: the order need not be exact)
    hbt_sort_gather<<<static_cast<unsigned>((4 * n + th - 1) / th), th, 0, L.stream>>>(d_p, L.mix_idx[1], n, L.mix_p);
    ctx->kernel_launches += 3;  // keys, gather + the radix sort (counted once)
    CU(ctx, cudaGetLastError());
    *out = L.mix_p;
    return HBT_OK;
}
#endif

// event boundaries of a buffer that holds list 1 (off1) followed, unless it aliases list 1, by list 2
void set_evoff(hbt_ctx *ctx, const int64_t *off1, int32_t nev1, const int64_t *off2, int32_t nev2, int64_t base2, bool alias) {
    ctx->evoff.clear();
    for (int e = 0; e <= nev1; e++) ctx->evoff.push_back(off1[e]);
    if (!alias)
        for (int e = 1; e <= nev2; e++) ctx->evoff.push_back(base2 + off2[e]);
}

bool production_mixed(const hbt_ctx *ctx, unsigned long long npairs) {
    return (ctx->ptsort == 2 || (ctx->ptsort == 1 && npairs >= ctx->ptsort_min_pairs)) && !ctx->stats && ctx->kernel_version != 1 &&
           !ctx->grid.qinv;  // (the pT range restriction rests on the q_out window, which q_inv does not respect)
}

const unsigned char *closed_ptr(const hbt_ctx *ctx) { return ctx->any_closed ? ctx->d_closed : nullptr; }

// is this host buffer page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)?
bool is_page_locked(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// the literal kernels: when asked for, when the grid needs them, for the two ordered-cap passes, and for
// instrumented runs in q_inv mode (the tuned q_inv kernels keep no stage counters)
bool use_literal(const hbt_ctx *ctx, int mode) {
    return ctx->kernel_version == 1 || mode != 0 || (ctx->grid.qinv && ctx->stats);
}

int tile_i(const hbt_ctx *ctx);  // list-1 particles per mixed-event unit of the kernel this context launches

// production mixed-event loops on the v4 kernel (binary32 tiles): not instrumented runs, not q_inv mode (both need the
// binary64 values), not when the FP32 decision is switched off
bool use_mixed4(const hbt_ctx *ctx) {
#ifdef HBT_HAVE_V2
    return ctx->mixed4 && ctx->kernel_version != 1 && !ctx->stats && !ctx->grid.qinv && ctx->v2c.f32_mixed;
#else
    return false;
#endif
}

#ifdef HBT_HAVE_V2
// q_inv mode: the replicas of the q_inv accumulators go into the histograms after every launch (same stream)
int fold_qinv(hbt_ctx *ctx, Lane &L) {
    hbt_qinv_fold<<<static_cast<unsigned>(ctx->grid.nKT * ctx->grid.nq), 128, 0, L.stream>>>(ctx->v2c, ctx->acc, ctx->grid.nKT, ctx->grid.nq);
    ctx->kernel_launches++;
    CU(ctx, cudaGetLastError());
    return HBT_OK;
}
#endif

// mode 0: the production kernels (v2 unless the grid needs v1); mode 1 / 2: the two ordered-cap
// passes, always on the literal v1 kernels
// `L`: the lane the launch goes to (the ordered-cap passes and the literal kernels always get lane 0)
int launch_same(hbt_ctx *ctx, Lane &L, const double *d_p, int64_t n, double psi_ref, int mode = 0, const HbtCap *capin = nullptr,
                float host_range = -1.f) {
    if (n < 2) return HBT_OK;
    ctx->reduced = false;
    cudaEvent_t e0, e1;
    int rc = get_event_pair(ctx, &e0, &e1);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(e0, L.stream));
    const unsigned long long npairs = static_cast<unsigned long long>(n) * (n - 1) / 2;
    HbtCap cap{};
    if (capin) cap = *capin;
    cap.closed = closed_ptr(ctx);
    if (use_literal(ctx, mode)) {
        const long long T = (n + kTileV1 - 1) / kTileV1;
        const long long blocks = T * (T + 1) / 2;
        if (blocks > 0x7fffffffLL) return fail(ctx, HBT_ERR_INVALID, "batch too large: %lld tiles", blocks);
        const unsigned nb = static_cast<unsigned>(blocks);
        const size_t sm = dyn_smem_bytes(ctx->grid);
        if (mode != 1) {
            hbt_add_stage_a<<<1, 1, 0, L.stream>>>(ctx->acc, 0, npairs);
            ctx->kernel_launches++;
        }
        if (mode == 0) hbt_pairs_v1<kTileV1, false, 0><<<nb, kTileV1, sm, L.stream>>>(d_p, d_p, n, nullptr, ctx->grid, ctx->acc, psi_ref, cap);
        else if (mode == 1) hbt_pairs_v1<kTileV1, false, 1><<<nb, kTileV1, sm, L.stream>>>(d_p, d_p, n, nullptr, ctx->grid, ctx->acc, psi_ref, cap);
        else hbt_pairs_v1<kTileV1, false, 2><<<nb, kTileV1, sm, L.stream>>>(d_p, d_p, n, nullptr, ctx->grid, ctx->acc, psi_ref, cap);
        ctx->kernel_launches++;
    } else {
#ifdef HBT_HAVE_V2
        rc = ensure_work(ctx, L);
        if (rc) return rc;
        CU(ctx, cudaMemsetAsync(L.d_work, 0, 16, L.stream));
        const bool qinv = ctx->grid.qinv != 0;
        const bool sorted = !ctx->stats && !qinv;
        const unsigned grid = static_cast<unsigned>(ctx->n_sm * (qinv ? ctx->occ_same_q : sorted ? ctx->occ_same : ctx->occ_same_stats));
        const long long all_units = hbt_v3_same_units(n, ctx->row_item0);
        if (all_units > 0x7fffffffLL || n > HBT_V3_MAX_SORTED)  // 8.8e12 pairs in one same-event list
            return fail(ctx, HBT_ERR_INVALID, "batch too large: %lld work units", all_units);
        if (sorted) {
            // production: Morton-sorted copy + tile boxes, units that can hold an accepted pair
            rc = prepare_sorted(ctx, L, d_p, n, host_range);
            if (rc) return rc;
            rc = ensure_units(ctx, L, all_units);
            if (rc) return rc;
            const long long n_rows = (n + HBT_V3_SUB_SAME - 1) / HBT_V3_SUB_SAME, ntj = (n + HBT_V3_TJ_SAME - 1) / HBT_V3_TJ_SAME;
            hbt_cull_units<<<dim3(static_cast<unsigned>((ntj + 127) / 128), static_cast<unsigned>(n_rows)), 128, 0, L.stream>>>(
                L.sort_bbox, n, ctx->v2c.W2, ctx->v2c.k2lo, ctx->v2c.k2hi, L.d_units, L.d_work);
            ctx->kernel_launches++;
            hbt_pairs_v3<false, false><<<grid, 32, 0, L.stream>>>(
                L.sort_p, L.sort_p, n, nullptr, nullptr, 0, L.d_units, L.d_work, 0, ctx->grid, ctx->v2c, ctx->d_dv,
                ctx->acc, psi_ref, npairs, cap.closed, L.sort_idx[1]);
        } else {
            // instrumented: every unit, reference order
            const size_t rb = ctx->row_item0.size() * sizeof(int);
            if (rb > ctx->d_rows_cap) {
                cudaFree(ctx->d_rows);
                CU(ctx, cudaMalloc(&ctx->d_rows, rb * 2));
                ctx->d_rows_cap = rb * 2;
            }
            // (pageable source: the copy is staged before the call returns, the vector can be reused)
            CU(ctx, cudaMemcpyAsync(ctx->d_rows, ctx->row_item0.data(), rb, cudaMemcpyHostToDevice, L.stream));
            const int n_rows = static_cast<int>(ctx->row_item0.size()) - 1;
            if (qinv) {
                // q_inv mode: every unit of the upper triangle (the transverse window does not bound q_inv: no culling)
                hbt_pairs_v3<false, false, true><<<grid, 32, 0, L.stream>>>(
                    d_p, d_p, n, nullptr, ctx->d_rows, n_rows, nullptr, L.d_work, static_cast<unsigned>(all_units), ctx->grid,
                    ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, npairs, cap.closed, nullptr);
                rc = fold_qinv(ctx, L);
                if (rc) return rc;
            } else {
                hbt_pairs_v3<false, true><<<grid, 32, 0, L.stream>>>(
                    d_p, d_p, n, nullptr, ctx->d_rows, n_rows, nullptr, L.d_work, static_cast<unsigned>(all_units), ctx->grid,
                    ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, npairs, cap.closed, nullptr);
            }
        }
        ctx->kernel_launches++;
        if (rc) return fail(ctx, rc, "v2 same-event launch failed");
#endif
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaEventRecord(e1, L.stream));
    ctx->timers.push_back({e0, e1, 0, 0.0});
    ctx->same_launches++;
    ctx->pending_num += npairs;
    return drain_timers(ctx, false);
}

// builds the (event, partner) segments of one batch into seg[]; returns their number
size_t build_segments(const int64_t *off1, int32_t nev1, const int64_t *off2, int64_t base2,
                      const int32_t *ids, const double *cs, int32_t nmix, int tile_i, int tile_j,
                      HbtMixSeg *seg, unsigned long long *npairs, long long *nblocks) {
    size_t ns = 0;
    long long block0 = 0;
    unsigned long long pairs = 0;
    for (int iev = 0; iev < nev1; iev++) {
        const int64_t i0 = off1[iev], ni = off1[iev + 1] - off1[iev];
        int64_t pos0 = 0;  // position inside the row's concatenated partner list (src :493-552)
        for (int c = 0; c < nmix; c++) {
            const size_t k = static_cast<size_t>(iev) * nmix + c;
            const int id = ids[k];
            const int64_t j0 = off2[id], nj = off2[id + 1] - off2[id];
            if (ni <= 0 || nj <= 0) continue;
            HbtMixSeg &s = seg[ns++];
            s.i0 = i0; s.ni = static_cast<int32_t>(ni);
            s.j0 = base2 + j0; s.nj = static_cast<int32_t>(nj);
            s.c = cs[2 * k]; s.s = cs[2 * k + 1];
            s.tiles_j = static_cast<int32_t>((nj + tile_j - 1) / tile_j);
            const int32_t tiles_i = static_cast<int32_t>((ni + tile_i - 1) / tile_i);
            s.block0 = block0;
            s.pos0 = pos0;
            s.pad = 0;
            block0 += static_cast<long long>(tiles_i) * s.tiles_j;
            pairs += static_cast<unsigned long long>(ni) * nj;
            pos0 += nj;
        }
    }
    *npairs = pairs;
    *nblocks = block0;
    return ns;
}

int launch_mixed(hbt_ctx *ctx, Lane &L, const double *d_p1, const double *d_p2, const HbtMixSeg *d_seg, size_t nseg,
                 long long nblocks, unsigned long long npairs, double psi_ref, int mode = 0, const HbtCap *capin = nullptr) {
    if (nseg == 0 || nblocks == 0) return HBT_OK;
    ctx->reduced = false;
    cudaEvent_t e0, e1;
    int rc = get_event_pair(ctx, &e0, &e1);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(e0, L.stream));
    if (nblocks > 0x7fffffffLL) return fail(ctx, HBT_ERR_INVALID, "batch too large: %lld tiles", nblocks);
    HbtCap cap{};
    if (capin) cap = *capin;
    cap.closed = closed_ptr(ctx);
    if (use_literal(ctx, mode)) {
        const unsigned nb = static_cast<unsigned>(nblocks);
        const size_t sm = dyn_smem_bytes(ctx->grid);
        const long long ns = static_cast<long long>(nseg);
        if (mode != 1) {
            hbt_add_stage_a<<<1, 1, 0, L.stream>>>(ctx->acc, 6, npairs);
            ctx->kernel_launches++;
        }
        if (mode == 0) hbt_pairs_v1<kTileV1, true, 0><<<nb, kTileV1, sm, L.stream>>>(d_p1, d_p2, ns, d_seg, ctx->grid, ctx->acc, psi_ref, cap);
        else if (mode == 1) hbt_pairs_v1<kTileV1, true, 1><<<nb, kTileV1, sm, L.stream>>>(d_p1, d_p2, ns, d_seg, ctx->grid, ctx->acc, psi_ref, cap);
        else hbt_pairs_v1<kTileV1, true, 2><<<nb, kTileV1, sm, L.stream>>>(d_p1, d_p2, ns, d_seg, ctx->grid, ctx->acc, psi_ref, cap);
        ctx->kernel_launches++;
    } else {
#ifdef HBT_HAVE_V2
        rc = ensure_work(ctx, L);
        if (rc) return rc;
        CU(ctx, cudaMemsetAsync(L.d_work, 0, 16, L.stream));
        const unsigned grid = static_cast<unsigned>(std::min<long long>(
            nblocks, static_cast<long long>(ctx->n_sm) * (ctx->grid.qinv ? ctx->occ_mixed_q : ctx->stats ? ctx->occ_mixed_stats : ctx->occ_mixed)));
        if (ctx->grid.qinv) {
            hbt_pairs_v3<true, false, true><<<grid, 32, 0, L.stream>>>(
                d_p1, d_p2, static_cast<long long>(nseg), d_seg, nullptr, 0, nullptr, L.d_work, static_cast<unsigned>(nblocks), ctx->grid,
                ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, npairs, cap.closed, nullptr);
            rc = fold_qinv(ctx, L);
            if (rc) return rc;
        } else if (ctx->stats)
            hbt_pairs_v3<true, true><<<grid, 32, 0, L.stream>>>(
                d_p1, d_p2, static_cast<long long>(nseg), d_seg, nullptr, 0, nullptr, L.d_work, static_cast<unsigned>(nblocks), ctx->grid,
                ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, npairs, cap.closed, nullptr);
        else if (use_mixed4(ctx))
            hbt_pairs_v4_mixed<<<static_cast<unsigned>(std::min<long long>(nblocks, static_cast<long long>(ctx->n_sm) * ctx->occ_mixed4)), 32, 0, L.stream>>>(
                d_p1, d_p2, static_cast<int>(nseg), d_seg, L.d_work + 2, static_cast<unsigned>(nblocks), ctx->grid, ctx->v2c, ctx->d_dv,
                ctx->acc, psi_ref, npairs, cap.closed);
        else
            hbt_pairs_v3<true, false><<<grid, 32, 0, L.stream>>>(
                d_p1, d_p2, static_cast<long long>(nseg), d_seg, nullptr, 0, nullptr, L.d_work, static_cast<unsigned>(nblocks), ctx->grid,
                ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, npairs, cap.closed, nullptr);
        ctx->kernel_launches++;
        if (rc) return fail(ctx, rc, "v2 mixed-event launch failed");
#endif
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaEventRecord(e1, L.stream));
    ctx->timers.push_back({e0, e1, 1, 0.0});
    ctx->mixed_launches++;
    ctx->pending_den += npairs;
    return drain_timers(ctx, false);
}

#ifdef HBT_HAVE_V2
// the two kernels of a split batch (everything they read has been enqueued on L.stream).
// Co-run: the same-event kernel with corun_same warps per SM on the lane's stream and the mixed-event kernel with
// corun_mixed warps per SM on its side stream share every SM.  Both are persistent (a warp leaves when its unit list is
// empty), so whichever list is finished first frees its share at once: a second launch of the OTHER kernel, queued behind
// it in stream order with the warps that make up that kernel's full grid, then joins the work that is left (it pops the
// same unit counter; if nothing is left its warps leave immediately).
int launch_split_pair(hbt_ctx *ctx, Lane &L, const double *d_same, long long n_same, const double *d_p1, const double *d_p2,
                      const HbtMixSeg *d_seg, size_t nseg, long long nblocks, double psi_ref, unsigned long long npairs_same,
                      unsigned long long npairs_mixed) {
    const bool corun = ctx->corun && ctx->corun_same < ctx->occ_same && ctx->corun_mixed < ctx->occ_mixed4;
    auto same = [&](cudaStream_t st, int warps, unsigned long long np) {
        hbt_pairs_v3<false, false><<<static_cast<unsigned>(ctx->n_sm * warps), 32, 0, st>>>(
            d_same, d_same, n_same, nullptr, nullptr, 0, L.d_units, L.d_work, 0, ctx->grid, ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, np,
            closed_ptr(ctx), L.sort_idx[1]);
        ctx->kernel_launches++;
    };
    auto mixed = [&](cudaStream_t st, int warps, unsigned long long np) {
        hbt_pairs_v4_mixed<<<static_cast<unsigned>(std::min<long long>(nblocks, static_cast<long long>(ctx->n_sm) * warps)), 32, 0, st>>>(
            d_p1, d_p2, static_cast<int>(nseg), d_seg, L.d_work + 2, static_cast<unsigned>(nblocks), ctx->grid, ctx->v2c, ctx->d_dv,
            ctx->acc, psi_ref, np, closed_ptr(ctx));
        ctx->kernel_launches++;
    };
    if (!corun) {
        same(L.stream, ctx->occ_same, npairs_same);
        mixed(L.stream, ctx->occ_mixed4, npairs_mixed);
        ctx->kernel_launches--;  // (the caller counts one of the two)
        return HBT_OK;
    }
    static const bool trace = getenv("HBT_B200_CORUN_TRACE") != nullptr;  // experiments: when does each of the four launches end
    cudaEvent_t te[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (trace) for (cudaEvent_t &e : te) CU(ctx, cudaEventCreate(&e));
    CU(ctx, cudaEventRecord(L.fork, L.stream));
    CU(ctx, cudaStreamWaitEvent(L.side, L.fork, 0));
    if (trace) CU(ctx, cudaEventRecord(te[0], L.stream));
    same(L.stream, ctx->corun_same, npairs_same);
    if (trace) CU(ctx, cudaEventRecord(te[1], L.stream));
    mixed(L.side, ctx->corun_mixed, npairs_mixed);
    if (trace) CU(ctx, cudaEventRecord(te[2], L.side));
    same(L.side, ctx->occ_same - ctx->corun_same, 0);        // after the mixed-event list is empty
    if (trace) CU(ctx, cudaEventRecord(te[3], L.side));
    mixed(L.stream, ctx->occ_mixed4 - ctx->corun_mixed, 0);  // after the same-event list is empty
    if (trace) CU(ctx, cudaEventRecord(te[4], L.stream));
    CU(ctx, cudaEventRecord(L.join, L.side));
    CU(ctx, cudaStreamWaitEvent(L.stream, L.join, 0));
    if (trace) {
        CU(ctx, cudaEventSynchronize(te[3]));
        CU(ctx, cudaEventSynchronize(te[4]));
        float t[5] = {0, 0, 0, 0, 0};
        for (int k = 1; k < 5; k++) cudaEventElapsedTime(&t[k], te[0], te[k]);
        fprintf(stderr, "[corun %d:%d] same(%d) ends %.2f ms, mixed(%d) ends %.2f, late same(%d) ends %.2f, late mixed(%d) ends %.2f\n",
                ctx->corun_same, ctx->corun_mixed, ctx->corun_same, t[1], ctx->corun_mixed, t[2], ctx->occ_same - ctx->corun_same, t[3],
                ctx->occ_mixed4 - ctx->corun_mixed, t[4]);
        for (cudaEvent_t e : te) cudaEventDestroy(e);
    }
    ctx->kernel_launches--;
    return HBT_OK;
}

// production launch of a whole batch: sort + cull of the same-event list, then one kernel that
// works through the same-event and the mixed-event units interleaved (hbt_pairs_v3_fused)
int launch_fused(hbt_ctx *ctx, Lane &L, const double *d_p, int64_t n, const double *d_p1, const double *d_p2, const HbtMixSeg *d_seg,
                 size_t nseg, long long nblocks, unsigned long long npairs_mixed, double psi_ref, float host_range = -1.f) {
    ctx->reduced = false;
    cudaEvent_t e0, e1;
    int rc = get_event_pair(ctx, &e0, &e1);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(e0, L.stream));
    const unsigned long long npairs_same = static_cast<unsigned long long>(n) * (n - 1) / 2;
    rc = ensure_work(ctx, L);
    if (rc) return rc;
    CU(ctx, cudaMemsetAsync(L.d_work, 0, 16, L.stream));
    const long long all_units = hbt_v3_same_units(n, ctx->row_item0);
    if (all_units > 0x7fffffffLL || n > HBT_V3_MAX_SORTED || nblocks + all_units > 0x7fffffffLL)
        return fail(ctx, HBT_ERR_INVALID, "batch too large: %lld work units", all_units + nblocks);
    rc = prepare_sorted(ctx, L, d_p, n, host_range);
    if (rc) return rc;
    rc = ensure_units(ctx, L, all_units);
    if (rc) return rc;
    const long long n_rows = (n + HBT_V3_SUB_SAME - 1) / HBT_V3_SUB_SAME, ntj = (n + HBT_V3_TJ_SAME - 1) / HBT_V3_TJ_SAME;
    hbt_cull_units<<<dim3(static_cast<unsigned>((ntj + 127) / 128), static_cast<unsigned>(n_rows)), 128, 0, L.stream>>>(
        L.sort_bbox, n, ctx->v2c.W2, ctx->v2c.k2lo, ctx->v2c.k2hi, L.d_units, L.d_work);
    if (ctx->split && use_mixed4(ctx)) {
        rc = launch_split_pair(ctx, L, L.sort_p, n, d_p1, d_p2, d_seg, nseg, nblocks, psi_ref, npairs_same, npairs_mixed);
        if (rc) return rc;
    } else {
        const unsigned grid = static_cast<unsigned>(ctx->n_sm * ctx->occ_fused);
        hbt_pairs_v3_fused<<<grid, 32, 0, L.stream>>>(L.sort_p, n, L.d_units, L.sort_idx[1], d_p1, d_p2,
                                                          static_cast<long long>(nseg), d_seg, static_cast<unsigned>(nblocks), L.d_work,
                                                          ctx->grid, ctx->v2c, ctx->d_dv, ctx->acc, psi_ref, npairs_same, npairs_mixed,
                                                          closed_ptr(ctx));
    }
    ctx->kernel_launches += 2;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaEventRecord(e1, L.stream));
    ctx->timers.push_back({e0, e1, 2, static_cast<double>(npairs_same) / static_cast<double>(npairs_same + npairs_mixed)});
    ctx->same_launches++;
    ctx->mixed_launches++;
    ctx->pending_num += npairs_same;
    ctx->pending_den += npairs_mixed;
    return drain_timers(ctx, false);
}
#endif

#ifdef HBT_HAVE_V2
// ---- several small batches in one launch -----------------------------------------------------------
constexpr size_t kCoalesceBatches = 32;                     // batches per launch (device-resident batches)
// host batches are staged one after the other by the calling thread: a launch every 8 keeps the GPU working on
// the previous ones meanwhile
constexpr size_t kCoalesceBatchesHost = 8;
constexpr unsigned long long kCoalescePairs = 1500000000ull;  // a batch with fewer pairs than this is "small"
constexpr unsigned long long kCoalesceFlushPairs = 8000000000ull;
constexpr size_t kCoalesceSlotParticles = 1u << 17;         // smallest staging slot of a group of host batches (8 MB)

// host-side time of the small-batch path by section (HBT_B200_HOSTPROF=1: printed at hbt_destroy)
struct HostProf {
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool on = getenv("HBT_B200_HOSTPROF") != nullptr;
    ~HostProf() {
        if (!on) return;
        static const char *name[8] = {"append: upload enqueue", "append: key range", "append: wait for the upload", "append: segments + bookkeeping",
                                      "flush: until the sort", "flush: sorts + helper launches", "flush: pair kernels + events", "append: page-locked query"};
        for (int k = 0; k < 8; k++) if (n[k]) fprintf(stderr, "[hostprof] %-34s %8.1f us per call, %llu calls\n", name[k], 1e6 * t[k] / n[k], n[k]);
    }
};
HostProf g_hostprof;
struct HostProfScope {
    int k;
    std::chrono::steady_clock::time_point t0;
    explicit HostProfScope(int k_) : k(k_), t0(std::chrono::steady_clock::now()) {}
    ~HostProfScope() {
        if (!g_hostprof.on) return;
        g_hostprof.t[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        g_hostprof.n[k]++;
    }
};

// launches everything that waits in ctx->pend as ONE production launch
int flush_pending(hbt_ctx *ctx) {
    Pending &P = ctx->pend;
    if (P.b.empty()) return HBT_OK;
    auto hp_t = std::chrono::steady_clock::now();
    auto hp_mark = [&](int k) {
        if (!g_hostprof.on) return;
        const auto now = std::chrono::steady_clock::now();
        g_hostprof.t[k] += std::chrono::duration<double>(now - hp_t).count();
        g_hostprof.n[k]++;
        hp_t = now;
    };
    Lane &L = next_lane(ctx);
    const int nb = static_cast<int>(P.b.size());
    if (P.host) {
        if (!P.direct) CU(ctx, cudaMemcpyAsync(P.slot->d_p, P.slot->h_p, P.staged * 64, cudaMemcpyHostToDevice, ctx->copy));
        CU(ctx, cudaEventRecord(P.slot->uploaded, ctx->copy));
        CU(ctx, cudaStreamWaitEvent(L.stream, P.slot->uploaded, 0));
    }
    HbtMulti M;
    M.nb = nb;
    M.pad = 0;
    long long c = 0, q = 0;
    for (int b = 0; b < nb; b++) {
        M.src[b] = P.b[b].d_src;
        M.cbase[b] = c;
        M.pbase[b] = q;
        c += P.b[b].n;
        q += (P.b[b].n + 63) & ~63ll;
    }
    for (int b = nb; b <= HBT_MULTI_MAX; b++) { M.cbase[b] = c; M.pbase[b] = q; }
    for (int b = nb; b < HBT_MULTI_MAX; b++) M.src[b] = nullptr;
    const long long n_log = c, n_pad = q;
    ctx->reduced = false;
    cudaEvent_t e0, e1;
    int rc = get_event_pair(ctx, &e0, &e1);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(e0, L.stream));
    rc = ensure_work(ctx, L);
    if (rc) return rc;
    CU(ctx, cudaMemsetAsync(L.d_work, 0, 16, L.stream));
    rc = ensure_sort_buffers(ctx, L, n_pad);
    if (rc) return rc;
    rc = ensure_units(ctx, L, P.units_bound);
    if (rc) return rc;
    // tile limit of every row: a unit pairs a row only with tiles of its own batch
    const long long n_rows = n_pad / HBT_V3_SUB_SAME;
    std::vector<unsigned> row_end(static_cast<size_t>(n_rows));
    for (int b = 0; b < nb; b++)
        for (long long a = M.pbase[b] / HBT_V3_SUB_SAME; a < M.pbase[b + 1] / HBT_V3_SUB_SAME; a++)
            row_end[static_cast<size_t>(a)] = static_cast<unsigned>(M.pbase[b + 1] / HBT_V3_TJ_SAME);
    if (row_end.size() > L.row_end_cap) {
        CU(ctx, cudaStreamSynchronize(L.stream));
        cudaFree(L.row_end);
        L.row_end = nullptr;
        L.row_end_cap = row_end.size() * 2;
        CU(ctx, cudaMalloc(&L.row_end, L.row_end_cap * 4));
    }
    rc = meta_begin(ctx, L, row_end.size() * 4 + P.evoff.size() * 8 + P.segs.size() * sizeof(HbtMixSeg) + 64);
    if (rc) return rc;
    rc = meta_upload(ctx, L, L.row_end, row_end.data(), row_end.size() * 4);
    if (rc) return rc;
    const int th = 256;
    hp_mark(4);
    // ---- same-event list: (batch, Morton) order of the padded slots, boxes, surviving units
    if (P.host && !P.direct) {
        hbt_multi_keys<<<static_cast<unsigned>((n_pad + th - 1) / th), th, 0, L.stream>>>(M, nullptr, P.host_range, L.sort_keys[0], L.sort_idx[0]);
    } else {
        CU(ctx, cudaMemsetAsync(L.sort_rmax, 0, 4, L.stream));
        hbt_multi_range<<<static_cast<unsigned>(std::min<long long>((n_log + th - 1) / th, 1184)), th, 0, L.stream>>>(M, L.sort_rmax);
        hbt_multi_keys<<<static_cast<unsigned>((n_pad + th - 1) / th), th, 0, L.stream>>>(M, L.sort_rmax, 0.f, L.sort_keys[0], L.sort_idx[0]);
        ctx->kernel_launches++;
    }
    int b_bits = 1;
    while ((1 << b_bits) < nb) b_bits++;
    size_t bytes = L.sort_tmp_bytes;
    CU(ctx, cub::DeviceRadixSort::SortPairs(L.sort_tmp, bytes, L.sort_keys[0], L.sort_keys[1], L.sort_idx[0], L.sort_idx[1],
                                            static_cast<int>(n_pad), 0, 24 + b_bits, L.stream));
    hbt_multi_gather<<<static_cast<unsigned>((4 * n_pad + th - 1) / th), th, 0, L.stream>>>(M, L.sort_idx[1], n_pad, L.sort_p);
    hbt_sort_bbox<<<static_cast<unsigned>((n_pad + HBT_BBOX_TILE - 1) / HBT_BBOX_TILE), HBT_BBOX_TILE, 0, L.stream>>>(L.sort_p, n_pad, L.sort_bbox, L.sort_idx[1]);
    const long long ntj = n_pad / HBT_V3_TJ_SAME;
    hbt_cull_units<<<dim3(static_cast<unsigned>((ntj + 127) / 128), static_cast<unsigned>(n_rows)), 128, 0, L.stream>>>(
        L.sort_bbox, n_pad, ctx->v2c.W2, ctx->v2c.k2lo, ctx->v2c.k2hi, L.d_units, L.d_work, L.row_end);
    ctx->kernel_launches += 5;
    CU(ctx, cudaGetLastError());
    if (P.do_mixed) {
        // ---- mixed-event lists: per-event pT order of the logical concatenation, one contiguous copy
        const int nev = static_cast<int>(P.evoff.size()) - 1;
        rc = ensure_mix_buffers(ctx, L, n_log, P.evoff.size());
        if (rc) return rc;
        rc = meta_upload(ctx, L, L.mix_off, P.evoff.data(), P.evoff.size() * 8);
        if (rc) return rc;
        hbt_multi_mix_keys<<<static_cast<unsigned>((n_log + th - 1) / th), th, 0, L.stream>>>(M, L.mix_off, nev, L.mix_keys[0], L.mix_idx[0]);
        int ev_bits = 1;
        while ((1ll << ev_bits) < nev) ev_bits++;
        bytes = L.mix_tmp_bytes;
        CU(ctx, cub::DeviceRadixSort::SortPairs(L.mix_tmp, bytes, L.mix_keys[0], L.mix_keys[1], L.mix_idx[0], L.mix_idx[1],
                                                static_cast<int>(n_log), 8, 32 + ev_bits, L.stream));
        hbt_multi_gather<<<static_cast<unsigned>((4 * n_log + th - 1) / th), th, 0, L.stream>>>(M, L.mix_idx[1], n_log, L.mix_p);
        if (P.segs.size() > L.mseg_cap) {
            CU(ctx, cudaStreamSynchronize(L.stream));
            cudaFree(L.mseg);
            L.mseg = nullptr;
            L.mseg_cap = P.segs.size() * 2;
            CU(ctx, cudaMalloc(&L.mseg, L.mseg_cap * sizeof(HbtMixSeg)));
        }
        rc = meta_upload(ctx, L, L.mseg, P.segs.data(), P.segs.size() * sizeof(HbtMixSeg));
        if (rc) return rc;
        ctx->kernel_launches += 3;
        rc = meta_end(ctx, L);
        if (rc) return rc;
        hp_mark(5);
        if (ctx->split && use_mixed4(ctx)) {
            rc = launch_split_pair(ctx, L, L.sort_p, n_pad, L.mix_p, L.mix_p, L.mseg, P.segs.size(), P.nblocks, P.psi_ref, P.pairs_same,
                                   P.pairs_mixed);
            if (rc) return rc;
        } else {
            const unsigned grid = static_cast<unsigned>(ctx->n_sm * ctx->occ_fused);
            hbt_pairs_v3_fused<<<grid, 32, 0, L.stream>>>(L.sort_p, n_pad, L.d_units, L.sort_idx[1], L.mix_p, L.mix_p,
                                                              static_cast<long long>(P.segs.size()), L.mseg, static_cast<unsigned>(P.nblocks),
                                                              L.d_work, ctx->grid, ctx->v2c, ctx->d_dv, ctx->acc, P.psi_ref, P.pairs_same,
                                                              P.pairs_mixed, closed_ptr(ctx));
        }
        ctx->timers.push_back({e0, e1, 2, static_cast<double>(P.pairs_same) / static_cast<double>(P.pairs_same + P.pairs_mixed)});
        ctx->mixed_launches++;
    } else {
        rc = meta_end(ctx, L);
        if (rc) return rc;
        hp_mark(5);
        const unsigned grid = static_cast<unsigned>(ctx->n_sm * ctx->occ_same);
        hbt_pairs_v3<false, false><<<grid, 32, 0, L.stream>>>(
            L.sort_p, L.sort_p, n_pad, nullptr, nullptr, 0, L.d_units, L.d_work, 0, ctx->grid, ctx->v2c, ctx->d_dv,
            ctx->acc, P.psi_ref, P.pairs_same, closed_ptr(ctx), L.sort_idx[1]);
        ctx->timers.push_back({e0, e1, 0, 0.0});
    }
    ctx->kernel_launches++;
    ctx->same_launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaEventRecord(e1, L.stream));
    if (P.slot) {
        CU(ctx, cudaEventRecord(P.slot->done, L.stream));
        P.slot->in_flight = true;
    }
    P = Pending();
    rc = drain_timers(ctx, false);
    hp_mark(6);
    return rc;
}

// Takes a small production batch into the pending launch (flushing first when it does not fit).  p_host: the
// batch's particles in host memory (they are staged now), or null with d_src = their device address.
int append_pending(hbt_ctx *ctx, const double *p_host, const double *d_src, const int64_t *off, int32_t nev,
                   const int32_t *ids, const double *cs, int32_t nmix, bool do_mixed, double psi_ref,
                   unsigned long long sp, unsigned long long mp) {
    Pending &P = ctx->pend;
    const int64_t n = off[nev];
    const bool host = p_host != nullptr;
    bool direct;
    {
        HostProfScope hp(7);
        direct = host && ctx->direct_upload && is_page_locked(p_host);
    }
    if (!P.b.empty() && (P.do_mixed != do_mixed || P.host != host || P.direct != direct || P.b.size() >= HBT_MULTI_MAX ||
                         P.n_padded + ((n + 63) & ~63ll) > HBT_V3_MAX_SORTED ||
                         (host && P.staged + static_cast<size_t>(n) > P.slot->cap))) {
        int rc = flush_pending(ctx);
        if (rc) return rc;
    }
    if (P.b.empty()) {
        P.do_mixed = do_mixed;
        P.host = host;
        P.direct = direct;
        P.psi_ref = psi_ref;  // (only grids without K_phi bins are coalesced: the angle is not read)
        P.evoff.assign(1, 0);
        if (host) {
            int rc = acquire_slot(ctx, &P.slot);
            if (rc) return rc;
            // (8 batches of 15 000-30 000 particles; grows with the batches, once)
            rc = ensure_slot(ctx, *P.slot, std::max<size_t>(static_cast<size_t>(n) * ctx->coalesce_host, kCoalesceSlotParticles), 0);
            if (rc) return rc;
        }
    }
    if (host) {
        const double *dst = p_host;
        {
            HostProfScope hp(0);
            if (direct) {  // straight from the caller's page-locked buffer; waited for below, the caller may reuse it on return
                CU(ctx, cudaMemcpyAsync(P.slot->d_p + 8 * P.staged, p_host, static_cast<size_t>(n) * 64, cudaMemcpyHostToDevice, ctx->copy));
            } else {
                std::memcpy(P.slot->h_p + 8 * P.staged, p_host, static_cast<size_t>(n) * 64);
                dst = P.slot->h_p + 8 * P.staged;
            }
        }
        if (!direct) {  // key range of the staged copy (in cache); page-locked batches are not touched: the device finds it
            HostProfScope hp(1);
            float m = P.host_range;
            for (int64_t i = 0; i < n; i++) {
                const float a = std::max(std::fabs(static_cast<float>(dst[8 * i])), std::fabs(static_cast<float>(dst[8 * i + 1])));
                if (a == a && a < 3.0e38f) m = std::max(m, a);
            }
            P.host_range = m;
        }
        d_src = P.slot->d_p + 8 * P.staged;
        P.staged += static_cast<size_t>(n);
        {
            HostProfScope hp(2);
            if (direct) CU(ctx, cudaStreamSynchronize(ctx->copy));  // (the range loop above ran beside the copy)
        }
    }
    HostProfScope hp3(3);
    if (do_mixed) {
        const size_t s0 = P.segs.size();
        P.segs.resize(s0 + static_cast<size_t>(nev) * nmix);
        unsigned long long np = 0;
        long long nblk = 0;
        const size_t ns = build_segments(off, nev, off, 0, ids, cs, nmix, tile_i(ctx), HBT_V3_TJ_MIXED, P.segs.data() + s0, &np, &nblk);
        P.segs.resize(s0 + ns);
        for (size_t k = s0; k < s0 + ns; k++) {
            P.segs[k].i0 += P.n_logical;
            P.segs[k].j0 += P.n_logical;
            P.segs[k].block0 += P.nblocks;
        }
        P.nblocks += nblk;
    }
    for (int e = 1; e <= nev; e++) P.evoff.push_back(P.n_logical + off[e]);
    P.b.push_back({d_src, n});
    P.n_logical += n;
    P.n_padded += (n + 63) & ~63ll;
    {
        const long long rows = (n + HBT_V3_SUB_SAME - 1) / HBT_V3_SUB_SAME;
        P.units_bound += rows * (rows + 1) / 2;
    }
    P.pairs_same += sp;
    P.pairs_mixed += mp;
    ctx->pending_num += sp;
    ctx->pending_den += mp;
    if (P.b.size() >= (host ? ctx->coalesce_host : kCoalesceBatches) || P.pairs_same + P.pairs_mixed >= kCoalesceFlushPairs)
        return flush_pending(ctx);
    return HBT_OK;
}

unsigned long long mixed_pairs(const int64_t *off1, int32_t nev1, const int64_t *off2, const int32_t *ids, int32_t nmix) {
    unsigned long long mp = 0;
    for (int iev = 0; iev < nev1; iev++)
        for (int k = 0; k < nmix; k++) {
            const int id = ids[static_cast<size_t>(iev) * nmix + k];
            mp += static_cast<unsigned long long>(off1[iev + 1] - off1[iev]) * static_cast<unsigned long long>(off2[id + 1] - off2[id]);
        }
    return mp;
}

// may this batch wait for others?  Production mode, list 2 = list 1, no K_phi bins (one psi_ref per launch), far
// from the pair cap, few pairs.
bool can_coalesce(const hbt_ctx *ctx, bool do_same, bool do_mixed, bool alias, int64_t n1, unsigned long long sp,
                  unsigned long long mp, bool near_cap) {
    return ctx->coalesce && do_same && alias && !ctx->stats && ctx->kernel_version != 1 && !ctx->grid.az && !ctx->grid.qinv && !near_cap &&
           n1 > 1 && sp + mp < kCoalescePairs && (!do_mixed || (ctx->fuse && ctx->ptsort != 0 && mp > 0));
}
#else
int flush_pending(hbt_ctx *) { return HBT_OK; }
#endif

int tile_i(const hbt_ctx *ctx) { return ctx->kernel_version == 1 ? kTileV1 : use_mixed4(ctx) ? HBT_V4_TI : HBT_V3_SUB_MIXED; }
int tile_j(const hbt_ctx *ctx) { return ctx->kernel_version == 1 ? kTileV1 : HBT_V3_TJ_MIXED; }

// host literal evaluation of the pairs the device deferred; leaves the stream idle
// position filter of the ordered cap for host-evaluated pairs
struct CapFilter {
    const std::vector<int64_t> *cut_row = nullptr, *cut_pos = nullptr;  // per slab; null: no cut
};

// copies the device's deferred-pair list to the host (pinned buffer) and clears it
int fetch_deferred(hbt_ctx *ctx, unsigned *n_out) {
    CU(ctx, cudaMemcpyAsync(ctx->h_defcount, ctx->acc.deferred_count, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    if (ctx->h_defcount[1]) return fail(ctx, HBT_ERR_OVERFLOW, "deferred-pair list overflowed (%u entries)", kDeferredCapacity);
    const unsigned n = ctx->h_defcount[0];
    *n_out = n;
    if (n == 0) return HBT_OK;
    CU(ctx, cudaMemcpyAsync(ctx->h_deferred, ctx->acc.deferred, static_cast<size_t>(n) * sizeof(HbtDeferred),
                            cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaMemsetAsync(ctx->acc.deferred_count, 0, 8, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return HBT_OK;
}

// host literal evaluation of the pairs the device deferred; leaves the stream idle
int resolve_deferred(hbt_ctx *ctx, const CapFilter *f = nullptr) {
    unsigned n = 0;
    int rc = fetch_deferred(ctx, &n);
    if (rc || n == 0) return rc;
    std::vector<HbtCorrection> corr;
    HbtStageDelta delta{};
    const int ns = ctx->grid.nslab;
    for (unsigned k = 0; k < n; k++) {
        const HbtDeferred &d = ctx->h_deferred[k];
        HbtCorrection c;
        uint64_t st[6] = {0, 0, 0, 0, 0, 0};
        const int ok = hbt_host_pair_literal(&ctx->grid, d.a, d.b, d.mixed, d.psi_ref, &c, st);
        for (int s = 1; s < 6; s++) delta.v[(d.mixed ? 6 : 0) + s] += st[s];
        if (!ok) continue;
        if (!ctx->closed.empty() && ctx->closed[c.slab + (d.mixed ? ns : 0)]) continue;
        if (f && f->cut_row) {
            const int64_t cr = (*f->cut_row)[c.slab];
            if (d.row > cr || (d.row == cr && d.pos > (*f->cut_pos)[c.slab])) continue;
        }
        corr.push_back(c);
    }
    ctx->deferred_total += n;
    if (!corr.empty())
        CU(ctx, cudaMemcpyAsync(ctx->d_corr, corr.data(), corr.size() * sizeof(HbtCorrection), cudaMemcpyHostToDevice, ctx->compute));
    hbt_apply_corrections<<<(static_cast<int>(corr.size()) + 127) / 128 + 1, 128, 0, ctx->compute>>>(
        ctx->d_corr, static_cast<int>(corr.size()), ctx->acc, delta);
    ctx->kernel_launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return HBT_OK;
}

// Cap channels: c < nslab is the 3-D histogram of slab c, which takes the first needed+1 pairs
// (:402-406, :651-655); in q_inv mode c = nslab + iK is the q_inv histogram of K_T bin iK, which
// takes the first 50*needed (:340-341, :600-602).  The two kinds are independent of each other.
int n_channels(const hbt_ctx *ctx) { return ctx->grid.nslab + (ctx->grid.qinv ? ctx->grid.nKT : 0); }
uint64_t channel_quota(const hbt_ctx *ctx, int c) {
    return c < ctx->grid.nslab ? ctx->grid.needed + 1 : 50 * ctx->grid.needed;
}
size_t closed_index(const hbt_ctx *ctx, int c, bool mixed) {
    const int ns = ctx->grid.nslab, nk = ctx->grid.nKT;
    return c < ns ? static_cast<size_t>(c + (mixed ? ns : 0)) : static_cast<size_t>(2 * ns + (c - ns) + (mixed ? nk : 0));
}

// per-slab accepted-pair counters = sums of the bin counts; refreshes the host copies
int refresh_counts(hbt_ctx *ctx) {
    const long long q3 = static_cast<long long>(ctx->grid.nq) * ctx->grid.nq * ctx->grid.nq;
    hbt_reduce_slabs<<<2 * ctx->grid.nslab, 256, 0, ctx->compute>>>(ctx->acc, ctx->grid.nslab, q3);
    hbt_finish_stage<<<1, 32, 0, ctx->compute>>>(ctx->acc, ctx->grid.nslab);
    ctx->kernel_launches += 2;
    CU(ctx, cudaGetLastError());
    const size_t ns = ctx->grid.nslab, nch = static_cast<size_t>(n_channels(ctx));
    ctx->exact_num.resize(nch);  // [0, ns): slabs; [ns, nch): q_inv K_T bins
    ctx->exact_den.resize(nch);
    CU(ctx, cudaMemcpyAsync(ctx->exact_num.data(), ctx->acc.npairs_num, ns * 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaMemcpyAsync(ctx->exact_den.data(), ctx->acc.npairs_den, ns * 8, cudaMemcpyDeviceToHost, ctx->compute));
    if (nch > ns) {
        CU(ctx, cudaMemcpyAsync(ctx->exact_num.data() + ns, ctx->acc.npairs_num_qinv, (nch - ns) * 8, cudaMemcpyDeviceToHost, ctx->compute));
        CU(ctx, cudaMemcpyAsync(ctx->exact_den.data() + ns, ctx->acc.npairs_den_qinv, (nch - ns) * 8, cudaMemcpyDeviceToHost, ctx->compute));
    }
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    ctx->pending_num = ctx->pending_den = 0;
    return HBT_OK;
}


// ==== needed_number_of_pairs: the ordered cap ==============================================
// The reference accepts a pair into slab K only while the slab's counter is <= needed
// (src/HBT_correlation.cpp:402-406, :651-655), in loop order, cumulatively over the batches: a
// slab takes exactly the first needed+1 pairs that reach it.  Far from the cap nothing special
// happens (cap_may_engage is false, batches stay asynchronous).  Near it each phase (the
// same-event loop, then the mixed-event loops) runs optimistically with a rollback copy; if no
// open slab crossed the cap the result stands, otherwise the phase is replayed in order:
// pass 1 counts the accepted pairs per row (list-1 particle) of the crossing slabs, the host
// finds the row where the quota runs out and evaluates that row literally to get the exact
// position, pass 2 accumulates only pairs at or before it.
struct PhaseInput {
    bool mixed = false;
    const double *h1 = nullptr, *h2 = nullptr;  // host copies of list 1 / list 2
    const double *d1 = nullptr, *d2 = nullptr;  // device copies
    int64_t n1 = 0;
    int64_t base2 = 0;  // offset of list 2 inside the staged device buffer (0 when it aliases list 1)
    const int64_t *off1 = nullptr, *off2 = nullptr;
    int32_t nev1 = 0, nmix = 0;
    const int32_t *ids = nullptr;
    const double *cs = nullptr;
    double psi_ref = 0.;
};

// pairs of channel c the cap has seen so far: this context's own (as of the last refresh) + the other contexts'
uint64_t foreign_count(const hbt_ctx *ctx, int c, bool mixed) {
    const std::vector<uint64_t> &f = mixed ? ctx->foreign_den : ctx->foreign_num;
    return f.empty() ? 0 : f[c];
}
uint64_t have_count(const hbt_ctx *ctx, int c, bool mixed) {
    const std::vector<uint64_t> &ex = mixed ? ctx->exact_den : ctx->exact_num;
    return (ex.empty() ? 0 : ex[c]) + foreign_count(ctx, c, mixed);
}

bool cap_may_engage(const hbt_ctx *ctx, bool mixed, unsigned long long pairs) {
    if (ctx->grid.needed >= (1ull << 56)) return false;
    const uint64_t pend = (mixed ? ctx->pending_den : ctx->pending_num) + pairs;
    const int nch = n_channels(ctx);
    for (int c = 0; c < nch; c++) {
        if (ctx->closed[closed_index(ctx, c, mixed)]) continue;
        if (have_count(ctx, c, mixed) + pend > channel_quota(ctx, c)) return true;
    }
    return false;
}

int sync_closed(hbt_ctx *ctx, bool mixed) {
    const int nch = n_channels(ctx);
    bool changed = false;
    for (int c = 0; c < nch; c++) {
        const size_t ci = closed_index(ctx, c, mixed);
        if (have_count(ctx, c, mixed) >= channel_quota(ctx, c) && !ctx->closed[ci]) { ctx->closed[ci] = 1; changed = true; }
    }
    if (changed) {
        ctx->any_closed = true;
        CU(ctx, cudaMemcpyAsync(ctx->d_closed, ctx->closed.data(), ctx->closed.size(), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->compute));
    }
    return HBT_OK;
}

// does the reference accept the pair (a, b) into cap channel K (a slab, or a q_inv K_T bin)?
bool pair_hits_channel(const hbt_ctx *ctx, const double *a, const double *b, int mixed, double psi_ref, int K) {
    const int ns = ctx->grid.nslab;
    if (K >= ns) {
        int iK = -1, iq = -1;
        return hbt_host_pair_qinv(&ctx->grid, a, b, &iK, &iq) && iK == K - ns;
    }
    HbtCorrection c;
    uint64_t st[6];
    return hbt_host_pair_literal(&ctx->grid, a, b, mixed, psi_ref, &c, st) && c.slab == K;
}

// position of the m-th pair (m >= 1) of row `row` that the reference accepts into channel K
int64_t resolve_row(const hbt_ctx *ctx, const PhaseInput &in, int K, int64_t row, uint64_t m) {
    const double *a = in.h1 + 8 * row;
    if (!in.mixed) {
        for (int64_t j = row + 1; j < in.n1; j++)
            if (pair_hits_channel(ctx, a, in.h1 + 8 * j, 0, in.psi_ref, K) && --m == 0) return j;
        return -1;
    }
    const int iev = static_cast<int>(std::upper_bound(in.off1, in.off1 + in.nev1 + 1, row) - in.off1) - 1;
    int64_t pos = 0;
    for (int s = 0; s < in.nmix; s++) {
        const size_t k = static_cast<size_t>(iev) * in.nmix + s;
        const int id = in.ids[k];
        const double cphi = in.cs[2 * k], sphi = in.cs[2 * k + 1];
        for (int64_t j = in.off2[id]; j < in.off2[id + 1]; j++, pos++) {
            const double *q = in.h2 + 8 * j;
            double b[8] = {0., 0., q[2], q[3], 0., 0., 0., 0.};
            const double t1 = q[0] * cphi, t2 = q[1] * sphi, t3 = q[0] * sphi, t4 = q[1] * cphi;  // :522-523
            b[0] = t1 - t2;
            b[1] = t3 + t4;
            if (pair_hits_channel(ctx, a, b, 1, in.psi_ref, K) && --m == 0) return pos;
        }
    }
    return -1;
}

int capped_phase(hbt_ctx *ctx, const PhaseInput &in, const HbtMixSeg *d_seg, size_t nseg, long long nblocks,
                 unsigned long long npairs) {
    const int nch = n_channels(ctx);
    // exact counters before the phase
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    if (!ctx->snap_u64) {
        CU(ctx, cudaMalloc(&ctx->snap_u64, ctx->n_u64 * 8));
        CU(ctx, cudaMalloc(&ctx->snap_f64, ctx->n_f64 * 8));
    }
    CU(ctx, cudaMemcpyAsync(ctx->snap_u64, ctx->blob_u64, ctx->n_u64 * 8, cudaMemcpyDeviceToDevice, ctx->compute));
    CU(ctx, cudaMemcpyAsync(ctx->snap_f64, ctx->blob_f64, ctx->n_f64 * 8, cudaMemcpyDeviceToDevice, ctx->compute));
    const std::vector<uint64_t> before = in.mixed ? ctx->exact_den : ctx->exact_num;
    // optimistic pass
    Lane &L0 = ctx->lanes[0];
    rc = in.mixed ? launch_mixed(ctx, L0, in.d1, in.d2, d_seg, nseg, nblocks, npairs, in.psi_ref)
                  : launch_same(ctx, L0, in.d1, in.n1, in.psi_ref);
    if (rc) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    // deferred pairs of this pass (K_phi edge cases) count too: fold them in before looking
    rc = resolve_deferred(ctx);
    if (rc) return rc;
    rc = refresh_counts(ctx);
    if (rc) return rc;
    const std::vector<uint64_t> &after = in.mixed ? ctx->exact_den : ctx->exact_num;
    std::vector<int32_t> xidx(nch, -1);
    std::vector<int> crossing;
    for (int k = 0; k < nch; k++)
        if (!ctx->closed[closed_index(ctx, k, in.mixed)] && after[k] + foreign_count(ctx, k, in.mixed) > channel_quota(ctx, k)) {
            xidx[k] = static_cast<int32_t>(crossing.size());
            crossing.push_back(k);
        }
    if (crossing.empty()) return sync_closed(ctx, in.mixed);

    // ---- roll back and replay in order ----------------------------------------------------
    CU(ctx, cudaMemcpyAsync(ctx->blob_u64, ctx->snap_u64, ctx->n_u64 * 8, cudaMemcpyDeviceToDevice, ctx->compute));
    CU(ctx, cudaMemcpyAsync(ctx->blob_f64, ctx->snap_f64, ctx->n_f64 * 8, cudaMemcpyDeviceToDevice, ctx->compute));
    (in.mixed ? ctx->exact_den : ctx->exact_num) = before;
    const size_t nx = crossing.size();
    const int64_t nrows = in.n1;
    // scratch of one replay, released on every exit path
    struct Scratch {
        int32_t *xidx = nullptr;
        unsigned *rowcnt = nullptr;
        int64_t *cut = nullptr;
        HbtMixSeg *seg1 = nullptr;
        ~Scratch() { cudaFree(xidx); cudaFree(rowcnt); cudaFree(cut); cudaFree(seg1); }
    } scratch;
    int32_t *&d_xidx = scratch.xidx;
    unsigned *&d_rowcnt = scratch.rowcnt;
    int64_t *&d_cut = scratch.cut;
    HbtMixSeg *&d_seg1 = scratch.seg1;
    std::vector<HbtMixSeg> seg1;
    unsigned long long np1 = npairs;
    long long nb1 = 0;
    size_t nseg1 = 0;
    CU(ctx, cudaMalloc(&d_xidx, nch * sizeof(int32_t)));
    CU(ctx, cudaMalloc(&d_rowcnt, nx * nrows * sizeof(unsigned)));
    CU(ctx, cudaMalloc(&d_cut, 2 * nch * sizeof(int64_t)));
    CU(ctx, cudaMemcpyAsync(d_xidx, xidx.data(), nch * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->compute));
    CU(ctx, cudaMemsetAsync(d_rowcnt, 0, nx * nrows * sizeof(unsigned), ctx->compute));
    if (in.mixed) {  // the literal kernels tile differently: rebuild the segments for them
        seg1.resize(static_cast<size_t>(in.nev1) * in.nmix);
        nseg1 = build_segments(in.off1, in.nev1, in.off2, in.base2, in.ids, in.cs, in.nmix, kTileV1, kTileV1,
                               seg1.data(), &np1, &nb1);
        CU(ctx, cudaMalloc(&d_seg1, std::max<size_t>(nseg1, 1) * sizeof(HbtMixSeg)));
        CU(ctx, cudaMemcpyAsync(d_seg1, seg1.data(), nseg1 * sizeof(HbtMixSeg), cudaMemcpyHostToDevice, ctx->compute));
    }
    HbtCap cap{};
    cap.xidx = d_xidx;
    cap.rowcnt = d_rowcnt;
    cap.nrows = nrows;
    cap.cut_row = d_cut;
    cap.cut_pos = d_cut + nch;
    // pass 1: per-row counts of the crossing channels
    rc = in.mixed ? launch_mixed(ctx, L0, in.d1, in.d2, d_seg1, nseg1, nb1, np1, in.psi_ref, 1, &cap)
                  : launch_same(ctx, L0, in.d1, in.n1, in.psi_ref, 1, &cap);
    if (rc) return rc;
    std::vector<unsigned> rowcnt(nx * nrows);
    CU(ctx, cudaMemcpyAsync(rowcnt.data(), d_rowcnt, rowcnt.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    {   // pairs the device could not decide (K_phi edge): the host's literal verdict joins the counts
        unsigned nd = 0;
        rc = fetch_deferred(ctx, &nd);
        if (rc) return rc;
        for (unsigned k = 0; k < nd; k++) {
            const HbtDeferred &d = ctx->h_deferred[k];
            HbtCorrection c;
            uint64_t st[6];
            if (hbt_host_pair_literal(&ctx->grid, d.a, d.b, d.mixed, d.psi_ref, &c, st) && xidx[c.slab] >= 0 && d.row >= 0)
                rowcnt[static_cast<size_t>(xidx[c.slab]) * nrows + d.row]++;
        }
    }
    std::vector<int64_t> cut_row(nch, INT64_MAX), cut_pos(nch, INT64_MAX);
    for (size_t x = 0; x < nx; x++) {
        const int K = crossing[x];
        // >= 1 for an open slab; 0 for q_inv with needed = 0
        uint64_t quota = channel_quota(ctx, K) - std::min<uint64_t>(channel_quota(ctx, K), before[K] + foreign_count(ctx, K, in.mixed));
        if (quota == 0) { cut_row[K] = -1; cut_pos[K] = -1; continue; }
        int64_t row = 0;
        const unsigned *rc_x = rowcnt.data() + x * nrows;
        while (row < nrows && rc_x[row] < quota) { quota -= rc_x[row]; row++; }
        if (row >= nrows) return fail(ctx, HBT_ERR_STATE, "ordered cap: quota not reached in the replay (channel %d)", K);
        const int64_t pos = resolve_row(ctx, in, K, row, quota);
        if (pos < 0) return fail(ctx, HBT_ERR_STATE, "ordered cap: host and device disagree on row %lld of channel %d", static_cast<long long>(row), K);
        cut_row[K] = row;
        cut_pos[K] = pos;
    }
    CU(ctx, cudaMemcpyAsync(d_cut, cut_row.data(), nch * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->compute));
    CU(ctx, cudaMemcpyAsync(d_cut + nch, cut_pos.data(), nch * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->compute));
    // pass 2: accumulate up to the cuts
    rc = in.mixed ? launch_mixed(ctx, L0, in.d1, in.d2, d_seg1, nseg1, nb1, np1, in.psi_ref, 2, &cap)
                  : launch_same(ctx, L0, in.d1, in.n1, in.psi_ref, 2, &cap);
    if (rc) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    CapFilter f;
    f.cut_row = &cut_row;
    f.cut_pos = &cut_pos;
    rc = resolve_deferred(ctx, &f);
    if (rc) return rc;
    rc = refresh_counts(ctx);
    if (rc) return rc;
    return sync_closed(ctx, in.mixed);
}

// the accumulators a reader sees: the all-reduced copies while they are current
HbtAccum view(const hbt_ctx *ctx) {
    HbtAccum a = ctx->acc;
    if (!ctx->reduced) return a;
    const ptrdiff_t du = ctx->red_u64 - ctx->blob_u64;
    const ptrdiff_t df = ctx->red_f64 - ctx->blob_f64;
    a.num_count += du; a.den_count += du; a.npairs_num += du; a.npairs_den += du; a.stage += du;
    a.qinv_count += du; a.qinv_den += du; a.npairs_num_qinv += du; a.npairs_den_qinv += du;
    a.num_cos += df; a.sum_qo += df; a.sum_qs += df; a.sum_ql += df; a.qinv_sum += df; a.qinv_cos += df;
    return a;
}

int check_cap(hbt_ctx *ctx, const uint64_t *num, const uint64_t *den, const uint64_t *qn, const uint64_t *qd) {
    // The reference stops accepting pairs into a (K_T[,K_phi]) slab once its counter
    // exceeds needed_number_of_pairs (src/HBT_correlation.cpp:402-406, :651-655), in pair
    // order.  A slab counter <= needed+1 proves the cap never engaged.
    const uint64_t lim = ctx->grid.needed + 1;
    for (int k = 0; k < ctx->grid.nslab; k++)
        if ((num && num[k] > lim) || (den && den[k] > lim))
            return fail(ctx, HBT_ERR_CAP,
                        "needed_number_of_pairs=%llu was exceeded in slab %d (internal error: the ordered pair cap "
                        "should have closed the slab at needed+1 pairs)",
                        static_cast<unsigned long long>(ctx->grid.needed), k);
    if (ctx->grid.qinv)
        for (int k = 0; k < ctx->grid.nKT; k++)
            if (ctx->grid.needed < (1ull << 56) && ((qn && qn[k] > 50 * ctx->grid.needed) || (qd && qd[k] > 50 * ctx->grid.needed)))
                return fail(ctx, HBT_ERR_CAP, "50*needed_number_of_pairs was exceeded in the q_inv histogram, K_T bin %d", k);
    return HBT_OK;
}

}  // namespace

// =========================================================================================
extern "C" int32_t hbt_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" const char *hbt_version(void) { return "hbt_b200 0.1 sm_100a"; }

extern "C" const char *hbt_last_error(const hbt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int hbt_create(const hbt_params *params, int32_t device, hbt_ctx **out) {
    if (!params || !out) return fail(nullptr, HBT_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, HBT_ERR_NO_DEVICE, "no CUDA device available (%s); libhbt_b200 has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, HBT_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    hbt_ctx *ctx = new hbt_ctx;
    ctx->device = device;
    ctx->params = *params;
    char msg[256];
    int rc = hbt_host_derive_grid(params, &ctx->grid, msg, sizeof(msg));
    if (rc) {
        delete ctx;
        return fail(nullptr, rc, "%s", msg);
    }
    if (const char *v = getenv("HBT_B200_KERNEL")) ctx->kernel_version = atoi(v) == 1 ? 1 : 2;
    if (const char *v = getenv("HBT_B200_STATS")) ctx->stats = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_FUSE")) ctx->fuse = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_MIXED4")) ctx->mixed4 = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_COALESCE_HOST")) ctx->coalesce_host = static_cast<size_t>(std::min(32, std::max(1, atoi(v))));
    if (const char *v = getenv("HBT_B200_SPLIT")) ctx->split = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_CORUN")) ctx->corun = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_CORUN_SAME")) ctx->corun_same = std::max(1, atoi(v));
    if (const char *v = getenv("HBT_B200_CORUN_MIXED")) ctx->corun_mixed = std::max(1, atoi(v));
    if (const char *v = getenv("HBT_B200_DIRECT")) ctx->direct_upload = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_COALESCE")) ctx->coalesce = atoi(v) != 0;
    if (const char *v = getenv("HBT_B200_PTSORT")) ctx->ptsort = std::min(2, std::max(0, atoi(v)));
#define CUC(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fail(nullptr, HBT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));       \
            hbt_destroy(ctx);                                                                  \
            return HBT_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        fail(nullptr, HBT_ERR_NO_DEVICE, "device %d is sm_%d%d; libhbt_b200 is built for sm_100a only", device, prop.major, prop.minor);
        hbt_destroy(ctx);
        return HBT_ERR_NO_DEVICE;
    }
    ctx->n_sm = prop.multiProcessorCount;
    for (Lane &L : ctx->lanes) {
        CUC(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CUC(cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking));
        CUC(cudaEventCreateWithFlags(&L.tail, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming));
    }
    ctx->compute = ctx->lanes[0].stream;
    if (const char *v = getenv("HBT_B200_LANES")) ctx->n_lanes = std::min(kLanes, std::max(1, atoi(v)));
    CUC(cudaEventCreate(&ctx->epoch));
    CUC(cudaEventRecord(ctx->epoch, ctx->compute));
    CUC(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
    const HbtGrid &g = ctx->grid;
    const size_t nb = static_cast<size_t>(g.nbins), ns = g.nslab, nqi = static_cast<size_t>(g.nKT) * g.nq;
    ctx->n_u64 = 2 * nb + 2 * ns + 12 + 2 * nqi + 2 * g.nKT;
    ctx->n_f64 = 4 * nb + 2 * nqi;
    CUC(cudaMalloc(&ctx->blob_u64, ctx->n_u64 * 8));
    CUC(cudaMalloc(&ctx->blob_f64, ctx->n_f64 * 8));
    CUC(cudaMemset(ctx->blob_u64, 0, ctx->n_u64 * 8));
    CUC(cudaMemset(ctx->blob_f64, 0, ctx->n_f64 * 8));
    HbtAccum &a = ctx->acc;
    unsigned long long *u = ctx->blob_u64;
    a.num_count = u; u += nb;
    a.den_count = u; u += nb;
    a.npairs_num = u; u += ns;
    a.npairs_den = u; u += ns;
    a.stage = u; u += 12;
    a.qinv_count = u; u += nqi;
    a.qinv_den = u; u += nqi;
    a.npairs_num_qinv = u; u += g.nKT;
    a.npairs_den_qinv = u; u += g.nKT;
    double *f = ctx->blob_f64;
    a.num_cos = f; f += nb;
    a.sum_qo = f; f += nb;
    a.sum_qs = f; f += nb;
    a.sum_ql = f; f += nb;
    a.qinv_sum = f; f += nqi;
    a.qinv_cos = f; f += nqi;
    CUC(cudaMalloc(&a.deferred, sizeof(HbtDeferred) * kDeferredCapacity));
    CUC(cudaMalloc(&a.deferred_count, 8));
    CUC(cudaMemset(a.deferred_count, 0, 8));
    a.deferred_capacity = kDeferredCapacity;
    CUC(cudaMallocHost(&ctx->h_deferred, sizeof(HbtDeferred) * kDeferredCapacity));
    CUC(cudaMallocHost(&ctx->h_defcount, 8));
    CUC(cudaMalloc(&ctx->d_corr, sizeof(HbtCorrection) * kDeferredCapacity));
    const size_t nclosed = 2 * static_cast<size_t>(ns) + 2 * static_cast<size_t>(g.nKT);  // slabs (same, mixed), q_inv K_T bins (same, mixed)
    ctx->closed.assign(nclosed, 0);
    CUC(cudaMalloc(&ctx->d_closed, nclosed));
    CUC(cudaMemset(ctx->d_closed, 0, nclosed));
    for (Slot &s : ctx->slots) {
        CUC(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
#ifdef HBT_HAVE_V2
    ctx->v2c = hbt_v2_consts(g);
    if (const char *v = getenv("HBT_B200_F32MIX")) ctx->v2c.f32_mixed = atoi(v) != 0;  // 0: every survivor through the FP64 path
    // persistent kernels: one grid-full of resident warps
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_same, hbt_pairs_v3<false, false>, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_same_stats, hbt_pairs_v3<false, true>, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_mixed, hbt_pairs_v3<true, false>, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_fused, hbt_pairs_v3_fused, 32, 0));
    // the same-event kernel and the v4 mixed-event kernel share the SMs (launch_split_pair): two kernels are resident
    // on one SM only under the same shared-memory carveout, and left to itself the driver sizes each kernel's to its own
    // full occupancy (228 KB for one, 196 KB for the other)
    CUC(cudaFuncSetAttribute(hbt_pairs_v3<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUC(cudaFuncSetAttribute(hbt_pairs_v4_mixed, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_same, hbt_pairs_v3<false, false>, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_mixed4, hbt_pairs_v4_mixed, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_mixed_stats, hbt_pairs_v3<true, true>, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_same_q, hbt_pairs_v3<false, false, true>, 32, 0));
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_mixed_q, hbt_pairs_v3<true, false, true>, 32, 0));
    if (!hbt_v2_supported(g)) ctx->kernel_version = 1;  // grids the fast path's guards do not cover: literal kernels
    if (g.qinv) {  // thresholds of the q_inv tests and the replicated q_inv accumulators (hbt_kernels_v3.cuh: v3_qinv_pair)
        std::vector<double> thr(static_cast<size_t>(g.nq) + 1);
        hbt_qinv_thresholds(&g, &ctx->v2c.qinv_s_lo, &ctx->v2c.qinv_s_hi, thr.data());
        CUC(cudaMalloc(&ctx->d_qinv_thr, thr.size() * 8));
        CUC(cudaMemcpy(ctx->d_qinv_thr, thr.data(), thr.size() * 8, cudaMemcpyHostToDevice));
        const int R = 256;
        const size_t nrep = static_cast<size_t>(R) * 2 * nqi;
        CUC(cudaMalloc(&ctx->d_qrep_u64, nrep * 8));
        CUC(cudaMalloc(&ctx->d_qrep_f64, nrep * 8));
        CUC(cudaMemset(ctx->d_qrep_u64, 0, nrep * 8));
        CUC(cudaMemset(ctx->d_qrep_f64, 0, nrep * 8));
        ctx->v2c.qinv_thr = ctx->d_qinv_thr;
        ctx->v2c.qrep_u64 = ctx->d_qrep_u64;
        ctx->v2c.qrep_f64 = ctx->d_qrep_f64;
        ctx->v2c.qrep_n = R;
    }
    if (const char *v = getenv("HBT_B200_OCC")) {  // experiments: fewer resident warps per SM than the kernels allow
        const int cap = std::max(1, atoi(v));
        ctx->occ_same = std::min(ctx->occ_same, cap); ctx->occ_mixed = std::min(ctx->occ_mixed, cap);
        ctx->occ_fused = std::min(ctx->occ_fused, cap);
    }
    if (const char *v = getenv("HBT_B200_OCC4")) ctx->occ_mixed4 = std::min(ctx->occ_mixed4, std::max(1, atoi(v)));

    {
        V2Dev dv;
        dv.g = ctx->grid;
        dv.c = ctx->v2c;
        dv.acc = ctx->acc;
        dv.closed = ctx->d_closed;
        CUC(cudaMalloc(&ctx->d_dv, sizeof(V2Dev)));
        CUC(cudaMemcpy(ctx->d_dv, &dv, sizeof(V2Dev), cudaMemcpyHostToDevice));
    }
#else
    ctx->kernel_version = 1;
#endif
    CUC(cudaDeviceSynchronize());
#undef CUC
    *out = ctx;
    return HBT_OK;
}

extern "C" void hbt_destroy(hbt_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (Lane &L : ctx->lanes) if (L.stream) cudaStreamSynchronize(L.stream);
    if (ctx->copy) cudaStreamSynchronize(ctx->copy);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (Slot &s : ctx->slots) {
        if (s.h_p) cudaFreeHost(s.h_p);
        if (s.d_p) cudaFree(s.d_p);
        if (s.h_seg) cudaFreeHost(s.h_seg);
        if (s.d_seg) cudaFree(s.d_seg);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.done) cudaEventDestroy(s.done);
    }
    for (auto &r : ctx->timers) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    for (auto &r : ctx->event_pool) { cudaEventDestroy(r.first); cudaEventDestroy(r.second); }
    if (ctx->blob_u64) cudaFree(ctx->blob_u64);
    if (ctx->blob_f64) cudaFree(ctx->blob_f64);
    if (ctx->red_u64) cudaFree(ctx->red_u64);
    if (ctx->red_f64) cudaFree(ctx->red_f64);
    if (ctx->acc.deferred) cudaFree(ctx->acc.deferred);
    if (ctx->acc.deferred_count) cudaFree(ctx->acc.deferred_count);
    if (ctx->h_deferred) cudaFreeHost(ctx->h_deferred);
    if (ctx->h_defcount) cudaFreeHost(ctx->h_defcount);
    if (ctx->d_corr) cudaFree(ctx->d_corr);
    if (ctx->d_closed) cudaFree(ctx->d_closed);
    if (ctx->d_qinv_thr) cudaFree(ctx->d_qinv_thr);
    if (ctx->d_qrep_u64) cudaFree(ctx->d_qrep_u64);
    if (ctx->d_qrep_f64) cudaFree(ctx->d_qrep_f64);
    for (Lane &L : ctx->lanes) {
        for (int k = 0; k < 2; k++) { if (L.sort_keys[k]) cudaFree(L.sort_keys[k]); if (L.sort_idx[k]) cudaFree(L.sort_idx[k]); }
        if (L.sort_p) cudaFree(L.sort_p);
        if (L.sort_bbox) cudaFree(L.sort_bbox);
        if (L.sort_tmp) cudaFree(L.sort_tmp);
        if (L.sort_rmax) cudaFree(L.sort_rmax);
        if (L.d_work) cudaFree(L.d_work);
        if (L.d_units) cudaFree(L.d_units);
        for (int k = 0; k < 2; k++) { if (L.mix_keys[k]) cudaFree(L.mix_keys[k]); if (L.mix_idx[k]) cudaFree(L.mix_idx[k]); }
        if (L.mix_p) cudaFree(L.mix_p);
        if (L.mix_off) cudaFree(L.mix_off);
        if (L.mix_tmp) cudaFree(L.mix_tmp);
        if (L.row_end) cudaFree(L.row_end);
        if (L.mseg) cudaFree(L.mseg);
        if (L.tail) cudaEventDestroy(L.tail);
    }
    if (ctx->epoch) cudaEventDestroy(ctx->epoch);
    if (ctx->d_rows) cudaFree(ctx->d_rows);
    if (ctx->snap_u64) cudaFree(ctx->snap_u64);
    if (ctx->snap_f64) cudaFree(ctx->snap_f64);
#ifdef HBT_HAVE_V2
    if (ctx->d_dv) cudaFree(ctx->d_dv);
#endif
    if (ctx->sw0) cudaEventDestroy(ctx->sw0);
    if (ctx->sw1) cudaEventDestroy(ctx->sw1);
    for (Lane &L : ctx->lanes) {
        if (L.stream) cudaStreamDestroy(L.stream);
        if (L.side) cudaStreamDestroy(L.side);
        if (L.h_meta) cudaFreeHost(L.h_meta);
        if (L.meta_done) cudaEventDestroy(L.meta_done);
        if (L.fork) cudaEventDestroy(L.fork);
        if (L.join) cudaEventDestroy(L.join);
    }
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    delete ctx;
}

extern "C" int64_t hbt_num_bins(const hbt_ctx *ctx) { return ctx ? ctx->grid.nbins : 0; }
extern "C" int32_t hbt_num_slabs(const hbt_ctx *ctx) { return ctx ? ctx->grid.nslab : 0; }

extern "C" int hbt_reset(hbt_ctx *ctx) {
    if (!ctx) return HBT_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    CU(ctx, cudaMemsetAsync(ctx->blob_u64, 0, ctx->n_u64 * 8, ctx->compute));
    CU(ctx, cudaMemsetAsync(ctx->blob_f64, 0, ctx->n_f64 * 8, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    ctx->same_ms = ctx->mixed_ms = 0.;
    ctx->same_launches = ctx->mixed_launches = 0;
    ctx->deferred_total = 0;
    ctx->reduced = false;
    ctx->exact_num.clear();
    ctx->exact_den.clear();
    ctx->pending_num = ctx->pending_den = 0;
    std::fill(ctx->closed.begin(), ctx->closed.end(), 0);
    ctx->any_closed = false;
    CU(ctx, cudaMemset(ctx->d_closed, 0, ctx->closed.size()));
    return HBT_OK;
}

extern "C" int hbt_synchronize(hbt_ctx *ctx) {
    if (!ctx) return HBT_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    {
        int rc = flush_pending(ctx);  // small batches still waiting for company
        if (rc) return rc;
    }
    CU(ctx, cudaStreamSynchronize(ctx->copy));
    for (Lane &L : ctx->lanes) {
        CU(ctx, cudaStreamSynchronize(L.stream));
        L.busy = false;
    }
    for (Slot &s : ctx->slots) s.in_flight = false;
    int rc = drain_timers(ctx, true);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->epoch, ctx->compute));  // idle: new time origin for the launch timers
    ctx->covered_ms = 0.;
    rc = resolve_deferred(ctx);
    if (rc) return rc;
    // accepted-pair counters per slab = sum of the slab's bin counts (every accepted pair
    // increments exactly one bin, src/HBT_correlation.cpp:406/437 and :655/682)
    return refresh_counts(ctx);
}

// ---- device-resident entry points --------------------------------------------------------
namespace {
// the device-resident loops on a given lane (both loops of one batch go to the same lane)
int same_dev_on(hbt_ctx *ctx, Lane &L, const double *d_p, int64_t n, double psi_ref) {
    if (!ctx || (n > 0 && !d_p) || n < 0) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_same_dev: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (cap_may_engage(ctx, false, n > 1 ? static_cast<unsigned long long>(n) * (n - 1) / 2 : 0))
        return fail(ctx, HBT_ERR_CAP, "needed_number_of_pairs may be reached: use the host-buffer entry points, which replay the cap in order");
    return launch_same(ctx, L, d_p, n, psi_ref);
}
int mixed_dev_on(hbt_ctx *ctx, Lane &L, const double *d_p1, const int64_t *off1, int32_t nev1, const double *d_p2,
                 const int64_t *off2, int32_t nev2, const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                 double psi_ref);
}  // namespace

extern "C" int hbt_accumulate_same_dev(hbt_ctx *ctx, const double *d_p, int64_t n, double psi_ref) {
    if (!ctx) return HBT_ERR_INVALID;
#ifdef HBT_HAVE_V2
    if (d_p && n > 1) {
        CU(ctx, cudaSetDevice(ctx->device));
        const unsigned long long sp0 = static_cast<unsigned long long>(n) * (n - 1) / 2;
        if (can_coalesce(ctx, true, false, true, n, sp0, 0, cap_may_engage(ctx, false, sp0))) {
            const int64_t off[2] = {0, n};
            return append_pending(ctx, nullptr, d_p, off, 1, nullptr, nullptr, 0, false, psi_ref, sp0, 0);
        }
        int rc = flush_pending(ctx);
        if (rc) return rc;
    }
#endif
    return same_dev_on(ctx, pick_lane(ctx), d_p, n, psi_ref);
}

extern "C" int hbt_accumulate_mixed_dev(hbt_ctx *ctx, const double *d_p1, const int64_t *off1, int32_t nev1,
                                        const double *d_p2, const int64_t *off2, int32_t nev2,
                                        const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                                        double psi_ref) {
    if (!ctx) return HBT_ERR_INVALID;
    {
        int rc = flush_pending(ctx);
        if (rc) return rc;
    }
    return mixed_dev_on(ctx, pick_lane(ctx), d_p1, off1, nev1, d_p2, off2, nev2, partner_ids, cos_sin, nmix, psi_ref);
}

namespace {
int mixed_dev_on(hbt_ctx *ctx, Lane &L, const double *d_p1, const int64_t *off1, int32_t nev1, const double *d_p2,
                 const int64_t *off2, int32_t nev2, const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                 double psi_ref) {
    if (!ctx || nev1 < 0 || nmix < 0) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_mixed_dev: bad argument");
    if (nev1 == 0 || nmix == 0) return HBT_OK;
    if (!d_p1 || !off1 || !partner_ids || !cos_sin) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_mixed_dev: null argument");
    if (!d_p2) { d_p2 = d_p1; off2 = off1; nev2 = nev1; }
    for (size_t k = 0; k < static_cast<size_t>(nev1) * nmix; k++)
        if (partner_ids[k] < 0 || partner_ids[k] >= nev2) return fail(ctx, HBT_ERR_INVALID, "partner id %d out of range", partner_ids[k]);
    CU(ctx, cudaSetDevice(ctx->device));
    Slot *s;
    int rc = acquire_slot(ctx, &s);
    if (rc) return rc;
    rc = ensure_slot(ctx, *s, 0, static_cast<size_t>(nev1) * nmix);
    if (rc) return rc;
    unsigned long long npairs;
    long long nblocks;
    const size_t nseg = build_segments(off1, nev1, off2, 0, partner_ids, cos_sin, nmix, tile_i(ctx), tile_j(ctx), s->h_seg, &npairs, &nblocks);
    if (cap_may_engage(ctx, true, npairs))
        return fail(ctx, HBT_ERR_CAP, "needed_number_of_pairs may be reached: use the host-buffer entry points, which replay the cap in order");
    if (nseg) CU(ctx, cudaMemcpyAsync(s->d_seg, s->h_seg, nseg * sizeof(HbtMixSeg), cudaMemcpyHostToDevice, L.stream));
#ifdef HBT_HAVE_V2
    if (d_p2 == d_p1 && nseg && production_mixed(ctx, npairs)) {  // (two separate device lists are read as given)
        set_evoff(ctx, off1, nev1, nullptr, 0, 0, true);
        rc = prepare_mixed_sorted(ctx, L, d_p1, off1[nev1], &d_p1);
        if (rc) return rc;
        d_p2 = d_p1;
    }
#endif
    rc = launch_mixed(ctx, L, d_p1, d_p2, s->d_seg, nseg, nblocks, npairs, psi_ref);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(s->done, L.stream));
    s->in_flight = true;
    return HBT_OK;
}
}  // namespace

// One whole batch, device resident: same-event loop over d_p[0, off[nev]) and the mixed-event
// loops of the events of the same list (list 2 = list 1), as hbt_accumulate_batch does for host
// buffers.  Runs the fused kernel when both halves are there.
extern "C" int hbt_accumulate_batch_dev(hbt_ctx *ctx, const double *d_p, const int64_t *off, int32_t nev,
                                        const int32_t *partner_ids, const double *cos_sin, int32_t nmix, double psi_ref) {
    if (!ctx || nev < 0 || nmix < 0) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_batch_dev: bad argument");
    if (nev == 0) return HBT_OK;
    if (!d_p || !off) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_batch_dev: null argument");
    const int64_t n = off[nev];
#ifdef HBT_HAVE_V2
    const bool do_mixed = nmix > 0 && partner_ids && cos_sin;
    if (do_mixed)
        for (size_t k = 0; k < static_cast<size_t>(nev) * nmix; k++)
            if (partner_ids[k] < 0 || partner_ids[k] >= nev) return fail(ctx, HBT_ERR_INVALID, "partner id %d out of range", partner_ids[k]);
    {
        CU(ctx, cudaSetDevice(ctx->device));
        const unsigned long long sp0 = n > 1 ? static_cast<unsigned long long>(n) * (n - 1) / 2 : 0;
        const unsigned long long mp0 = do_mixed ? mixed_pairs(off, nev, off, partner_ids, nmix) : 0;
        const bool near0 = cap_may_engage(ctx, false, sp0) || (do_mixed && cap_may_engage(ctx, true, mp0));
        if (can_coalesce(ctx, true, do_mixed, true, n, sp0, mp0, near0))
            return append_pending(ctx, nullptr, d_p, off, nev, partner_ids, cos_sin, nmix, do_mixed, psi_ref, sp0, mp0);
        int rc = flush_pending(ctx);
        if (rc) return rc;
    }
    if (do_mixed && ctx->fuse && !ctx->stats && ctx->kernel_version != 1 && !ctx->grid.qinv && n > 1) {
        CU(ctx, cudaSetDevice(ctx->device));
        Slot *s;
        int rc = acquire_slot(ctx, &s);
        if (rc) return rc;
        rc = ensure_slot(ctx, *s, 0, static_cast<size_t>(nev) * nmix);
        if (rc) return rc;
        unsigned long long npairs;
        long long nblocks;
        const size_t nseg = build_segments(off, nev, off, 0, partner_ids, cos_sin, nmix, tile_i(ctx), tile_j(ctx), s->h_seg, &npairs, &nblocks);
        const unsigned long long sp = static_cast<unsigned long long>(n) * (n - 1) / 2;
        if (cap_may_engage(ctx, false, sp) || cap_may_engage(ctx, true, npairs))
            return fail(ctx, HBT_ERR_CAP, "needed_number_of_pairs may be reached: use the host-buffer entry points, which replay the cap in order");
        if (nseg && nblocks) {
            Lane &L = pick_lane(ctx);
            CU(ctx, cudaMemcpyAsync(s->d_seg, s->h_seg, nseg * sizeof(HbtMixSeg), cudaMemcpyHostToDevice, L.stream));
            const double *d_mix = d_p;
            if (production_mixed(ctx, npairs)) {
                set_evoff(ctx, off, nev, nullptr, 0, 0, true);
                rc = prepare_mixed_sorted(ctx, L, d_p, n, &d_mix);
                if (rc) return rc;
            }
            rc = launch_fused(ctx, L, d_p, n, d_mix, d_mix, s->d_seg, nseg, nblocks, npairs, psi_ref);
            if (rc) return rc;
            CU(ctx, cudaEventRecord(s->done, L.stream));
            s->in_flight = true;
            return HBT_OK;
        }
    }
#endif
    Lane &L = pick_lane(ctx);
    int rc = same_dev_on(ctx, L, d_p, n, psi_ref);
    if (rc) return rc;
    if (nmix > 0) rc = mixed_dev_on(ctx, L, d_p, off, nev, nullptr, nullptr, 0, partner_ids, cos_sin, nmix, psi_ref);
    return rc;
}

// ---- host-buffer entry points ------------------------------------------------------------
extern "C" int hbt_accumulate_batch(hbt_ctx *ctx, const double *p1, const int64_t *off1, int32_t nev1,
                                    const double *p2, const int64_t *off2, int32_t nev2,
                                    const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                                    double psi_ref, int32_t do_same, int32_t do_mixed) {
    if (!ctx || nev1 < 0) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_batch: bad argument");
    if (nev1 == 0) return HBT_OK;  // the reader's trailing empty batch (src/Analysis.cpp:821)
    if (!off1) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_batch: null offsets");
    const int64_t n1 = off1[nev1];
    if (n1 > 0 && !p1) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_batch: null particles");
    const bool alias = (p2 == nullptr);
    if (alias) { off2 = off1; nev2 = nev1; }
    const int64_t n2 = alias ? 0 : off2[nev2];
    do_mixed = do_mixed && nmix > 0;
    if (do_mixed) {
        if (!partner_ids || !cos_sin) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_batch: null mixed-event plan");
        for (size_t k = 0; k < static_cast<size_t>(nev1) * nmix; k++)
            if (partner_ids[k] < 0 || partner_ids[k] >= nev2) return fail(ctx, HBT_ERR_INVALID, "partner id %d out of range", partner_ids[k]);
    }
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
#ifdef HBT_HAVE_V2
    {   // small production batches wait for each other and go out as one launch
        const unsigned long long sp0 = (do_same && n1 > 1) ? static_cast<unsigned long long>(n1) * (n1 - 1) / 2 : 0;
        const unsigned long long mp0 = do_mixed ? mixed_pairs(off1, nev1, off2, partner_ids, nmix) : 0;
        const bool near0 = (do_same && cap_may_engage(ctx, false, sp0)) || (do_mixed && cap_may_engage(ctx, true, mp0));
        if (can_coalesce(ctx, do_same, do_mixed, alias, n1, sp0, mp0, near0))
            return append_pending(ctx, p1, nullptr, off1, nev1, partner_ids, cos_sin, nmix, do_mixed, psi_ref, sp0, mp0);
        rc = flush_pending(ctx);
        if (rc) return rc;
    }
#endif
    Slot *s;
    rc = acquire_slot(ctx, &s);
    if (rc) return rc;
    rc = ensure_slot(ctx, *s, static_cast<size_t>(n1 + n2), do_mixed ? static_cast<size_t>(nev1) * nmix : 0);
    if (rc) return rc;
    // the device buffers of this slot may still be read by the compute stream's previous
    // use; acquire_slot already waited on `done`.  Stage and upload on the copy stream.
    // Page-locked caller buffers (what a host that cares about throughput hands over) go to the device directly:
    // the DMA engine reads them at PCIe speed, ~0.2 ms for a config-5 group against ~1.5 ms for a staging memcpy by
    // the calling thread; the call waits for that copy, since the caller may reuse its buffers on return.  Not near
    // the pair cap: the ordered replay reads the staged host copy.
    const double *h1 = s->h_p, *h2 = s->h_p + 8 * n1;
    bool direct = false;
    {
        const unsigned long long sp0 = (do_same && n1 > 1) ? static_cast<unsigned long long>(n1) * (n1 - 1) / 2 : 0;
        const unsigned long long mp0 = do_mixed ? mixed_pairs(off1, nev1, off2, partner_ids, nmix) : 0;
        const bool near0 = (do_same && cap_may_engage(ctx, false, sp0)) || (do_mixed && cap_may_engage(ctx, true, mp0));
        direct = ctx->direct_upload && !near0 && n1 > 0 && is_page_locked(p1) && (n2 == 0 || is_page_locked(p2));
    }
    if (direct) {
        CU(ctx, cudaMemcpyAsync(s->d_p, p1, static_cast<size_t>(n1) * 64, cudaMemcpyHostToDevice, ctx->copy));
        if (n2) CU(ctx, cudaMemcpyAsync(s->d_p + 8 * n1, p2, static_cast<size_t>(n2) * 64, cudaMemcpyHostToDevice, ctx->copy));
        h1 = p1;
        h2 = alias ? p1 : p2;
    } else {
        if (n1) std::memcpy(s->h_p, p1, static_cast<size_t>(n1) * 64);
        if (n2) std::memcpy(s->h_p + 8 * n1, p2, static_cast<size_t>(n2) * 64);
        if (n1 + n2) CU(ctx, cudaMemcpyAsync(s->d_p, s->h_p, static_cast<size_t>(n1 + n2) * 64, cudaMemcpyHostToDevice, ctx->copy));
    }
    size_t nseg = 0;
    unsigned long long npairs = 0;
    long long nblocks = 0;
    if (do_mixed) {
        nseg = build_segments(off1, nev1, off2, alias ? 0 : n1, partner_ids, cos_sin, nmix, tile_i(ctx), tile_j(ctx), s->h_seg, &npairs, &nblocks);
        if (nseg) CU(ctx, cudaMemcpyAsync(s->d_seg, s->h_seg, nseg * sizeof(HbtMixSeg), cudaMemcpyHostToDevice, ctx->copy));
    }
    CU(ctx, cudaEventRecord(s->uploaded, ctx->copy));
    // max(|px|, |py|) of list 1 for the Morton keys of the same-event sort, from the host copy (as
    // hbt_sort_range computes it: finite values only)
    // (only of a copy this thread has just made: a page-locked caller buffer is left to the DMA engine, and the device
    // finds the range — one small kernel — instead of 1 ms of the calling thread per 150 000 particles before the launch)
    float host_range = -1.f;
    if (do_same && !ctx->stats && ctx->kernel_version != 1 && n1 > 1 && !direct) {
        float m = 0.f;
        for (int64_t i = 0; i < n1; i++) {
            const float a = std::max(std::fabs(static_cast<float>(h1[8 * i])), std::fabs(static_cast<float>(h1[8 * i + 1])));
            if (a == a && a < 3.0e38f) m = std::max(m, a);
        }
        host_range = m;
    }
    const unsigned long long sp_all = n1 > 1 ? static_cast<unsigned long long>(n1) * (n1 - 1) / 2 : 0;
    const bool near_cap = (do_same && cap_may_engage(ctx, false, sp_all)) || (do_mixed && cap_may_engage(ctx, true, npairs));
    Lane &L = pick_lane(ctx, near_cap);
    CU(ctx, cudaStreamWaitEvent(L.stream, s->uploaded, 0));
    // the caller's buffers must be free again when the call returns — but not before: the launches below are enqueued
    // while the DMA engine still reads them (the kernels themselves are not waited for)
    struct UploadWait {
        cudaEvent_t e;
        bool on;
        ~UploadWait() { if (on) cudaEventSynchronize(e); }
    } upload_wait{s->uploaded, direct};
    PhaseInput in;
    in.h1 = h1;
    in.h2 = alias ? h1 : h2;
    in.d1 = s->d_p;
    in.d2 = s->d_p;  // segments carry the list-2 base offset
    in.n1 = n1;
    in.base2 = alias ? 0 : n1;
    in.off1 = off1;
    in.off2 = off2;
    in.nev1 = nev1;
    in.nmix = nmix;
    in.ids = partner_ids;
    in.cs = cos_sin;
    in.psi_ref = psi_ref;
#ifdef HBT_HAVE_V2
    if (do_same && do_mixed && ctx->fuse && !ctx->stats && ctx->kernel_version != 1 && !ctx->grid.qinv && n1 > 1 && nseg > 0 && nblocks > 0 &&
        !near_cap) {
        const double *d_mix = s->d_p;
        if (production_mixed(ctx, npairs)) {
            set_evoff(ctx, off1, nev1, off2, nev2, n1, alias);
            rc = prepare_mixed_sorted(ctx, L, s->d_p, n1 + n2, &d_mix);
            if (rc) return rc;
        }
        rc = launch_fused(ctx, L, s->d_p, n1, d_mix, d_mix, s->d_seg, nseg, nblocks, npairs, psi_ref, host_range);
        if (rc) return rc;
        do_same = do_mixed = 0;
    }
#endif
    if (do_same) {
        const unsigned long long sp = sp_all;
        in.mixed = false;
        rc = cap_may_engage(ctx, false, sp) ? capped_phase(ctx, in, nullptr, 0, 0, sp)
                                            : launch_same(ctx, L, s->d_p, n1, psi_ref, 0, nullptr, host_range);
        if (rc) return rc;
    }
    if (do_mixed) {
        in.mixed = true;
        if (cap_may_engage(ctx, true, npairs)) {
            rc = capped_phase(ctx, in, s->d_seg, nseg, nblocks, npairs);
        } else {
            const double *d_mix = s->d_p;
#ifdef HBT_HAVE_V2
            if (nseg && production_mixed(ctx, npairs)) {
                set_evoff(ctx, off1, nev1, off2, nev2, n1, alias);
                rc = prepare_mixed_sorted(ctx, L, s->d_p, n1 + n2, &d_mix);
                if (rc) return rc;
            }
#endif
            rc = launch_mixed(ctx, L, d_mix, d_mix, s->d_seg, nseg, nblocks, npairs, psi_ref);
        }
        if (rc) return rc;
    }
    CU(ctx, cudaEventRecord(s->done, L.stream));
    s->in_flight = true;
    return HBT_OK;
}

extern "C" int hbt_accumulate_same(hbt_ctx *ctx, const double *p, int64_t n, double psi_ref) {
    if (!ctx || n < 0) return fail(ctx, HBT_ERR_INVALID, "hbt_accumulate_same: bad argument");
    if (n == 0) return HBT_OK;
    const int64_t off[2] = {0, n};
    return hbt_accumulate_batch(ctx, p, off, 1, nullptr, nullptr, 0, nullptr, nullptr, 0, psi_ref, 1, 0);
}

extern "C" int hbt_accumulate_mixed(hbt_ctx *ctx, const double *p1, const int64_t *off1, int32_t nev1,
                                    const double *p2, const int64_t *off2, int32_t nev2,
                                    const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                                    double psi_ref) {
    return hbt_accumulate_batch(ctx, p1, off1, nev1, p2, off2, nev2, partner_ids, cos_sin, nmix, psi_ref, 0, 1);
}

// ---- results -----------------------------------------------------------------------------
extern "C" int hbt_read(hbt_ctx *ctx, uint64_t *num_count, double *num_cos, double *sum_qo, double *sum_qs,
                        double *sum_ql, uint64_t *den_count, uint64_t *npairs_num, uint64_t *npairs_den) {
    if (!ctx) return HBT_ERR_INVALID;
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    const size_t nb = static_cast<size_t>(ctx->grid.nbins) * 8, ns = static_cast<size_t>(ctx->grid.nslab) * 8;
    const HbtAccum a = view(ctx);
    if (num_count) CU(ctx, cudaMemcpy(num_count, a.num_count, nb, cudaMemcpyDeviceToHost));
    if (num_cos) CU(ctx, cudaMemcpy(num_cos, a.num_cos, nb, cudaMemcpyDeviceToHost));
    if (sum_qo) CU(ctx, cudaMemcpy(sum_qo, a.sum_qo, nb, cudaMemcpyDeviceToHost));
    if (sum_qs) CU(ctx, cudaMemcpy(sum_qs, a.sum_qs, nb, cudaMemcpyDeviceToHost));
    if (sum_ql) CU(ctx, cudaMemcpy(sum_ql, a.sum_ql, nb, cudaMemcpyDeviceToHost));
    if (den_count) CU(ctx, cudaMemcpy(den_count, a.den_count, nb, cudaMemcpyDeviceToHost));
    std::vector<uint64_t> kn(ctx->grid.nslab), kd(ctx->grid.nslab), qn(ctx->grid.nKT), qd(ctx->grid.nKT);
    CU(ctx, cudaMemcpy(kn.data(), a.npairs_num, ns, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(kd.data(), a.npairs_den, ns, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(qn.data(), a.npairs_num_qinv, ctx->grid.nKT * 8, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(qd.data(), a.npairs_den_qinv, ctx->grid.nKT * 8, cudaMemcpyDeviceToHost));
    if (npairs_num) std::memcpy(npairs_num, kn.data(), ns);
    if (npairs_den) std::memcpy(npairs_den, kd.data(), ns);
    return check_cap(ctx, kn.data(), kd.data(), qn.data(), qd.data());
}

extern "C" int hbt_read_qinv(hbt_ctx *ctx, uint64_t *count, double *sum_qinv, double *sum_cos, uint64_t *den,
                             uint64_t *npairs_num, uint64_t *npairs_den) {
    if (!ctx) return HBT_ERR_INVALID;
    if (!ctx->grid.qinv) return fail(ctx, HBT_ERR_STATE, "invariant_radius_flag is 0");
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    const size_t n = static_cast<size_t>(ctx->grid.nKT) * ctx->grid.nq * 8, nk = ctx->grid.nKT * 8;
    const HbtAccum a = view(ctx);
    if (count) CU(ctx, cudaMemcpy(count, a.qinv_count, n, cudaMemcpyDeviceToHost));
    if (sum_qinv) CU(ctx, cudaMemcpy(sum_qinv, a.qinv_sum, n, cudaMemcpyDeviceToHost));
    if (sum_cos) CU(ctx, cudaMemcpy(sum_cos, a.qinv_cos, n, cudaMemcpyDeviceToHost));
    if (den) CU(ctx, cudaMemcpy(den, a.qinv_den, n, cudaMemcpyDeviceToHost));
    if (npairs_num) CU(ctx, cudaMemcpy(npairs_num, a.npairs_num_qinv, nk, cudaMemcpyDeviceToHost));
    if (npairs_den) CU(ctx, cudaMemcpy(npairs_den, a.npairs_den_qinv, nk, cudaMemcpyDeviceToHost));
    return HBT_OK;
}

extern "C" int hbt_get_stage_counters(hbt_ctx *ctx, uint64_t same[6], uint64_t mixed[6]) {
    if (!ctx) return HBT_ERR_INVALID;
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    uint64_t st[12];
    CU(ctx, cudaMemcpy(st, view(ctx).stage, sizeof(st), cudaMemcpyDeviceToHost));
    if (same) std::memcpy(same, st, 48);
    if (mixed) std::memcpy(mixed, st + 6, 48);
    return HBT_OK;
}

extern "C" int hbt_get_timers(hbt_ctx *ctx, double *same_ms, double *mixed_ms, uint64_t *same_launches,
                              uint64_t *mixed_launches) {
    if (!ctx) return HBT_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = drain_timers(ctx, true);
    if (rc) return rc;
    if (same_ms) *same_ms = ctx->same_ms;
    if (mixed_ms) *mixed_ms = ctx->mixed_ms;
    if (same_launches) *same_launches = ctx->same_launches;
    if (mixed_launches) *mixed_launches = ctx->mixed_launches;
    return HBT_OK;
}

extern "C" int hbt_timer_start(hbt_ctx *ctx) {
    if (!ctx) return HBT_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!ctx->sw0) {
        CU(ctx, cudaEventCreate(&ctx->sw0));
        CU(ctx, cudaEventCreate(&ctx->sw1));
    }
    {
        int rc = flush_pending(ctx);  // what was submitted before the stopwatch starts before it
        if (rc) return rc;
    }
    CU(ctx, cudaStreamSynchronize(ctx->copy));
    int rc = join_lanes(ctx);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->sw0, ctx->compute));
    for (int l = 1; l < kLanes; l++) CU(ctx, cudaStreamWaitEvent(ctx->lanes[l].stream, ctx->sw0, 0));  // nothing starts before the stopwatch
    return HBT_OK;
}

extern "C" int hbt_timer_stop(hbt_ctx *ctx, double *ms) {
    if (!ctx || !ms || !ctx->sw0) return HBT_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    {
        int rc = flush_pending(ctx);
        if (rc) return rc;
    }
    CU(ctx, cudaStreamSynchronize(ctx->copy));  // everything uploaded has been handed to a compute lane
    int rc = join_lanes(ctx);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->sw1, ctx->compute));
    CU(ctx, cudaEventSynchronize(ctx->sw1));
    float t = 0.f;
    CU(ctx, cudaEventElapsedTime(&t, ctx->sw0, ctx->sw1));
    *ms = t;
    return HBT_OK;
}

extern "C" int hbt_set_option(hbt_ctx *ctx, int32_t option, int32_t value) {
    if (!ctx) return HBT_ERR_INVALID;
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    switch (option) {
        case HBT_OPT_STAGE_COUNTERS:
            ctx->stats = value != 0;
            return HBT_OK;
        case HBT_OPT_KERNEL:
            if (value != 1 && value != 2) return fail(ctx, HBT_ERR_INVALID, "kernel version must be 1 or 2");
#ifdef HBT_HAVE_V2
            ctx->kernel_version = (value == 2 && hbt_v2_supported(ctx->grid)) ? 2 : 1;
#endif
            return HBT_OK;
        case HBT_OPT_FUSE:
            ctx->fuse = value != 0;
            return HBT_OK;
        case HBT_OPT_COALESCE:
            ctx->coalesce = value != 0;
            return HBT_OK;
        case HBT_OPT_PTSORT:
            if (value < 0 || value > 2) return fail(ctx, HBT_ERR_INVALID, "ptsort must be 0, 1 or 2");
            ctx->ptsort = value;
            return HBT_OK;
        case HBT_OPT_LANES:
            if (value < 1 || value > kLanes) return fail(ctx, HBT_ERR_INVALID, "lanes must be in [1, %d]", kLanes);
            ctx->n_lanes = value;
            ctx->next_lane = 0;
            return HBT_OK;
        default:
            return fail(ctx, HBT_ERR_INVALID, "unknown option %d", option);
    }
}

extern "C" int hbt_get_launch_count(hbt_ctx *ctx, uint64_t *n) {
    if (!ctx || !n) return HBT_ERR_INVALID;
    *n = ctx->kernel_launches;
    return HBT_OK;
}

extern "C" int hbt_get_deferred_pairs(hbt_ctx *ctx, uint64_t *n) {
    if (!ctx || !n) return HBT_ERR_INVALID;
    *n = ctx->deferred_total;
    return HBT_OK;
}

// ---- FP64 roofline denominator ---------------------------------------------------------
namespace {
// 8 independent DFMA chains per thread; every result feeds the final store so nothing is
// eliminated.  2 flops per DFMA.
__global__ void __launch_bounds__(256) hbt_dfma_chain(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int hbt_measure_fp64_peak(int32_t device, double ms, double *tflops) {
    if (!tflops) return HBT_ERR_INVALID;
    hbt_ctx *ctx = nullptr;
    CU(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(ctx, cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double *out = nullptr;
    CU(ctx, cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CU(ctx, cudaEventCreate(&e0));
    CU(ctx, cudaEventCreate(&e1));
    int iters = 1 << 14;
    double best = 0.0, spent = 0.0;
    for (int rep = 0; rep < 64 && (spent < ms || rep < 3); rep++) {
        CU(ctx, cudaEventRecord(e0, 0));
        hbt_dfma_chain<<<blocks, threads>>>(out, iters, 0.9999999, 1e-9);
        CU(ctx, cudaEventRecord(e1, 0));
        CU(ctx, cudaEventSynchronize(e1));
        float t = 0.f;
        CU(ctx, cudaEventElapsedTime(&t, e0, e1));
        const double flops = 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads;
        if (rep > 0) best = std::max(best, flops / (t * 1e-3) / 1e12);  // rep 0 = warm-up
        spent += t;
        if (t < 5.f && iters < (1 << 20)) iters *= 2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return HBT_OK;
}

// ---- multi-GPU -------------------------------------------------------------------------
#define NC(ctx, call)                                                                          \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess)                                                                 \
            return fail(ctx, HBT_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

extern "C" int hbt_comm_unique_id(char id[128]) {
    if (!g_nccl.load()) return fail(nullptr, HBT_ERR_NCCL, "%s", g_nccl.error.c_str());
    ncclUniqueId uid;
    NC(nullptr, g_nccl.GetUniqueId(&uid));
    std::memcpy(id, uid.internal, 128);
    return HBT_OK;
}

extern "C" int hbt_comm_init_rank(hbt_ctx *ctx, int32_t nranks, int32_t rank, const char id[128]) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, HBT_ERR_INVALID, "hbt_comm_init_rank: bad argument");
    if (!g_nccl.load()) return fail(ctx, HBT_ERR_NCCL, "%s", g_nccl.error.c_str());
    CU(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    NC(ctx, g_nccl.CommInitRank(&ctx->comm, nranks, uid, rank));
    ctx->nranks = nranks;
    return HBT_OK;
}

extern "C" int hbt_comm_init_all(hbt_ctx **ctxs, int32_t n) {
    if (!ctxs || n < 1) return HBT_ERR_INVALID;
    if (!g_nccl.load()) return fail(ctxs[0], HBT_ERR_NCCL, "%s", g_nccl.error.c_str());
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    NC(ctxs[0], g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->nranks = n; }
    return HBT_OK;
}

extern "C" int hbt_allreduce_all(hbt_ctx **ctxs, int32_t n) {
    if (!ctxs || n < 1) return HBT_ERR_INVALID;
    for (int i = 0; i < n; i++) {
        int rc = hbt_synchronize(ctxs[i]);  // deferred pairs are folded in before the sum
        if (rc) return rc;
    }
    if (n == 1 && !ctxs[0]->comm) return HBT_OK;  // a single GPU is its own sum
    for (int i = 0; i < n; i++)
        if (!ctxs[i]->comm) return fail(ctxs[i], HBT_ERR_STATE, "hbt_allreduce: no communicator (call hbt_comm_init_*)");
    for (int i = 0; i < n; i++) {
        hbt_ctx *c = ctxs[i];
        CU(c, cudaSetDevice(c->device));
        if (!c->red_u64) {
            CU(c, cudaMalloc(&c->red_u64, c->n_u64 * 8));
            CU(c, cudaMalloc(&c->red_f64, c->n_f64 * 8));
        }
    }
    // the local accumulators stay local (more batches may follow); the sums over ranks land
    // in red_*, which hbt_read / hbt_get_stage_counters return until the next accumulate
    NC(ctxs[0], g_nccl.GroupStart());
    for (int i = 0; i < n; i++) {
        hbt_ctx *c = ctxs[i];
        NC(c, g_nccl.AllReduce(c->blob_u64, c->red_u64, c->n_u64, ncclUint64, ncclSum, c->comm, c->compute));
        NC(c, g_nccl.AllReduce(c->blob_f64, c->red_f64, c->n_f64, ncclFloat64, ncclSum, c->comm, c->compute));
    }
    NC(ctxs[0], g_nccl.GroupEnd());
    for (int i = 0; i < n; i++) {
        CU(ctxs[i], cudaSetDevice(ctxs[i]->device));
        CU(ctxs[i], cudaStreamSynchronize(ctxs[i]->compute));
        ctxs[i]->reduced = true;
    }
    return HBT_OK;
}

extern "C" int hbt_allreduce(hbt_ctx *ctx) { return hbt_allreduce_all(&ctx, 1); }


// ---- cap bookkeeping across contexts ---------------------------------------------------------------
extern "C" int32_t hbt_cap_channels(const hbt_ctx *ctx) { return ctx ? n_channels(ctx) : 0; }

extern "C" int hbt_cap_get_counts(hbt_ctx *ctx, uint64_t *num, uint64_t *den) {
    if (!ctx) return HBT_ERR_INVALID;
    int rc = hbt_synchronize(ctx);  // refreshes the exact per-channel counters
    if (rc) return rc;
    const size_t nch = static_cast<size_t>(n_channels(ctx));
    if (num) std::memcpy(num, ctx->exact_num.data(), nch * 8);
    if (den) std::memcpy(den, ctx->exact_den.data(), nch * 8);
    return HBT_OK;
}

extern "C" int hbt_cap_set_foreign(hbt_ctx *ctx, const uint64_t *num, const uint64_t *den) {
    if (!ctx || !num || !den) return HBT_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = hbt_synchronize(ctx);
    if (rc) return rc;
    const size_t nch = static_cast<size_t>(n_channels(ctx));
    ctx->foreign_num.assign(num, num + nch);
    ctx->foreign_den.assign(den, den + nch);
    rc = sync_closed(ctx, false);  // channels another context filled are closed here too
    if (rc) return rc;
    return sync_closed(ctx, true);
}

// ---- group of contexts: one analysis over several GPUs of one process -----------------------------
namespace {
__global__ void hbt_add_blob_u64(unsigned long long *__restrict__ dst, const unsigned long long *__restrict__ src, size_t n) {
    const size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (k < n) dst[k] += src[k];
}
__global__ void hbt_add_blob_f64(double *__restrict__ dst, const double *__restrict__ src, size_t n) {
    const size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (k < n) dst[k] += src[k];
}
}  // namespace

struct hbt_group {
    std::vector<hbt_ctx *> ctx;
    int next = 0;
    bool distinct = true;      // all contexts on different devices: the sum is one NCCL all-reduce
    bool cap_possible = false;
    int nch = 0, nslab = 0;
    uint64_t needed = 0;
    std::vector<uint64_t> g_num, g_den;                     // accepted pairs per channel, all contexts, as of the last refresh
    std::vector<std::vector<uint64_t>> own_num, own_den;    // per context
    uint64_t pend_num = 0, pend_den = 0;                    // pairs submitted to any context since
    uint64_t ordered_batches = 0;                           // batches that had to run in sequence (near the cap)
    std::string err;
};

namespace {
int gfail(hbt_group *g, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (g) g->err = buf; else g_create_error = buf;
    return code;
}
uint64_t gquota(const hbt_group *g, int c) { return c < g->nslab ? g->needed + 1 : 50 * g->needed; }

// synchronize every context, sum the exact per-channel counters and tell each context what the others hold
int group_refresh(hbt_group *g) {
    const size_t nch = static_cast<size_t>(g->nch);
    std::fill(g->g_num.begin(), g->g_num.end(), 0);
    std::fill(g->g_den.begin(), g->g_den.end(), 0);
    for (size_t i = 0; i < g->ctx.size(); i++) {
        int rc = hbt_cap_get_counts(g->ctx[i], g->own_num[i].data(), g->own_den[i].data());
        if (rc) return gfail(g, rc, "%s", hbt_last_error(g->ctx[i]));
        for (size_t c = 0; c < nch; c++) { g->g_num[c] += g->own_num[i][c]; g->g_den[c] += g->own_den[i][c]; }
    }
    std::vector<uint64_t> fn(nch), fd(nch);
    for (size_t i = 0; i < g->ctx.size(); i++) {
        for (size_t c = 0; c < nch; c++) { fn[c] = g->g_num[c] - g->own_num[i][c]; fd[c] = g->g_den[c] - g->own_den[i][c]; }
        int rc = hbt_cap_set_foreign(g->ctx[i], fn.data(), fd.data());
        if (rc) return gfail(g, rc, "%s", hbt_last_error(g->ctx[i]));
    }
    g->pend_num = g->pend_den = 0;
    return HBT_OK;
}

bool group_far(const hbt_group *g, uint64_t sp, uint64_t mp) {
    for (int c = 0; c < g->nch; c++) {
        const uint64_t q = gquota(g, c);
        if (sp && g->g_num[c] < q && g->g_num[c] + g->pend_num + sp > q) return false;
        if (mp && g->g_den[c] < q && g->g_den[c] + g->pend_den + mp > q) return false;
    }
    return true;
}
}  // namespace

extern "C" int hbt_group_create(const hbt_params *params, int32_t n_devices, const int32_t *devices, hbt_group **out) {
    if (!params || !out || n_devices < 1) return gfail(nullptr, HBT_ERR_INVALID, "hbt_group_create: bad argument");
    *out = nullptr;
    hbt_group *g = new hbt_group;
    for (int i = 0; i < n_devices; i++) {
        hbt_ctx *c = nullptr;
        const int dev = devices ? devices[i] : i;
        const int rc = hbt_create(params, dev, &c);
        if (rc) {
            for (hbt_ctx *x : g->ctx) hbt_destroy(x);
            delete g;
            return rc;  // message in hbt_last_error(NULL)
        }
        for (hbt_ctx *x : g->ctx) if (x->device == dev) g->distinct = false;
        g->ctx.push_back(c);
    }
    hbt_ctx *c0 = g->ctx[0];
    g->nch = n_channels(c0);
    g->nslab = c0->grid.nslab;
    g->needed = c0->grid.needed;
    g->cap_possible = g->needed < (1ull << 56);
    g->g_num.assign(g->nch, 0);
    g->g_den.assign(g->nch, 0);
    g->own_num.assign(g->ctx.size(), std::vector<uint64_t>(g->nch, 0));
    g->own_den.assign(g->ctx.size(), std::vector<uint64_t>(g->nch, 0));
    if (g->ctx.size() > 1 && g->distinct) {
        const int rc = hbt_comm_init_all(g->ctx.data(), static_cast<int>(g->ctx.size()));
        if (rc) {
            gfail(nullptr, rc, "%s", hbt_last_error(c0));
            for (hbt_ctx *x : g->ctx) hbt_destroy(x);
            delete g;
            return rc;
        }
    }
    *out = g;
    return HBT_OK;
}

extern "C" void hbt_group_destroy(hbt_group *g) {
    if (!g) return;
    for (hbt_ctx *c : g->ctx) hbt_destroy(c);
    delete g;
}

extern "C" const char *hbt_group_last_error(const hbt_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }
extern "C" int32_t hbt_group_size(const hbt_group *g) { return g ? static_cast<int32_t>(g->ctx.size()) : 0; }
extern "C" hbt_ctx *hbt_group_ctx(hbt_group *g, int32_t i) {
    return (g && i >= 0 && i < static_cast<int32_t>(g->ctx.size())) ? g->ctx[i] : nullptr;
}
extern "C" int hbt_group_ordered_batches(const hbt_group *g, uint64_t *n) {
    if (!g || !n) return HBT_ERR_INVALID;
    *n = g->ordered_batches;
    return HBT_OK;
}

extern "C" int hbt_group_accumulate_batch(hbt_group *g, const double *p1, const int64_t *off1, int32_t nev1,
                                          const double *p2, const int64_t *off2, int32_t nev2,
                                          const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                                          double psi_ref, int32_t do_same, int32_t do_mixed) {
    if (!g || nev1 < 0) return gfail(g, HBT_ERR_INVALID, "hbt_group_accumulate_batch: bad argument");
    if (nev1 == 0) return HBT_OK;
    if (!off1) return gfail(g, HBT_ERR_INVALID, "hbt_group_accumulate_batch: null offsets");
    hbt_ctx *c = g->ctx[g->next];
    g->next = (g->next + 1) % static_cast<int>(g->ctx.size());
    auto submit = [&]() {
        const int rc = hbt_accumulate_batch(c, p1, off1, nev1, p2, off2, nev2, partner_ids, cos_sin, nmix, psi_ref, do_same, do_mixed);
        return rc ? gfail(g, rc, "%s", hbt_last_error(c)) : HBT_OK;
    };
    if (!g->cap_possible || g->ctx.size() == 1) return submit();
    // pairs of this batch (what the loops will visit: an upper bound of what they accept)
    const uint64_t n1 = static_cast<uint64_t>(off1[nev1]);
    const uint64_t sp = (do_same && n1 > 1) ? n1 * (n1 - 1) / 2 : 0;
    uint64_t mp = 0;
    if (do_mixed && nmix > 0 && partner_ids) {
        const int64_t *o2 = p2 ? off2 : off1;
        const int32_t ne2 = p2 ? nev2 : nev1;
        for (int iev = 0; iev < nev1; iev++)
            for (int k = 0; k < nmix; k++) {
                const int id = partner_ids[static_cast<size_t>(iev) * nmix + k];
                if (id < 0 || id >= ne2) return gfail(g, HBT_ERR_INVALID, "partner id %d out of range", id);
                mp += static_cast<uint64_t>(off1[iev + 1] - off1[iev]) * static_cast<uint64_t>(o2[id + 1] - o2[id]);
            }
    }
    // Far from the cap the batches stay asynchronous on their GPUs.  The needed_number_of_pairs cap is cumulative
    // over the batches IN ORDER (src/HBT_correlation.cpp:402-406, :651-655): when this batch could close a channel,
    // every context is brought up to date first, then the batch runs alone, with the other contexts' counters as
    // its starting point (hbt_cap_set_foreign), so its own ordered replay cuts at the reference's pair.
    if (!group_far(g, sp, mp)) {
        int rc = group_refresh(g);
        if (rc) return rc;
        if (!group_far(g, sp, mp)) {
            rc = submit();
            if (rc) return rc;
            g->ordered_batches++;
            return group_refresh(g);
        }
    }
    g->pend_num += sp;
    g->pend_den += mp;
    return submit();
}

// sum of all contexts' accumulators, readable from context 0 afterwards
extern "C" int hbt_group_reduce(hbt_group *g) {
    if (!g) return HBT_ERR_INVALID;
    const int n = static_cast<int>(g->ctx.size());
    if (n == 1) return hbt_synchronize(g->ctx[0]) ? gfail(g, HBT_ERR_CUDA, "%s", hbt_last_error(g->ctx[0])) : HBT_OK;
    if (g->distinct) {  // one NCCL all-reduce over NVLink
        const int rc = hbt_allreduce_all(g->ctx.data(), n);
        return rc ? gfail(g, rc, "%s", hbt_last_error(g->ctx[0])) : HBT_OK;
    }
    // several contexts share a device (test configurations): NCCL refuses duplicate devices in one communicator;
    // context 0 adds the others' blobs itself (peer copies where the device differs)
    for (hbt_ctx *c : g->ctx) {
        const int rc = hbt_synchronize(c);
        if (rc) return gfail(g, rc, "%s", hbt_last_error(c));
    }
    hbt_ctx *c0 = g->ctx[0];
    CU(c0, cudaSetDevice(c0->device));
    if (!c0->red_u64) {
        CU(c0, cudaMalloc(&c0->red_u64, c0->n_u64 * 8));
        CU(c0, cudaMalloc(&c0->red_f64, c0->n_f64 * 8));
    }
    CU(c0, cudaMemcpyAsync(c0->red_u64, c0->blob_u64, c0->n_u64 * 8, cudaMemcpyDeviceToDevice, c0->compute));
    CU(c0, cudaMemcpyAsync(c0->red_f64, c0->blob_f64, c0->n_f64 * 8, cudaMemcpyDeviceToDevice, c0->compute));
    unsigned long long *tmp_u = nullptr;
    double *tmp_f = nullptr;
    for (int i = 1; i < n; i++) {
        hbt_ctx *ci = g->ctx[i];
        const unsigned long long *su = ci->blob_u64;
        const double *sf = ci->blob_f64;
        if (ci->device != c0->device) {
            if (!tmp_u) {
                CU(c0, cudaMalloc(&tmp_u, c0->n_u64 * 8));
                CU(c0, cudaMalloc(&tmp_f, c0->n_f64 * 8));
            }
            CU(c0, cudaMemcpyPeerAsync(tmp_u, c0->device, ci->blob_u64, ci->device, c0->n_u64 * 8, c0->compute));
            CU(c0, cudaMemcpyPeerAsync(tmp_f, c0->device, ci->blob_f64, ci->device, c0->n_f64 * 8, c0->compute));
            su = tmp_u;
            sf = tmp_f;
        }
        hbt_add_blob_u64<<<static_cast<unsigned>((c0->n_u64 + 255) / 256), 256, 0, c0->compute>>>(c0->red_u64, su, c0->n_u64);
        hbt_add_blob_f64<<<static_cast<unsigned>((c0->n_f64 + 255) / 256), 256, 0, c0->compute>>>(c0->red_f64, sf, c0->n_f64);
        c0->kernel_launches += 2;
        CU(c0, cudaGetLastError());
        CU(c0, cudaStreamSynchronize(c0->compute));
    }
    cudaFree(tmp_u);
    cudaFree(tmp_f);
    c0->reduced = true;
    return HBT_OK;
}
