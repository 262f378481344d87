// v1 pair kernels: the literal chain for every pair, one thread per list-1 particle, list-2
// tile staged through shared memory, global atomics into the L2-resident histograms.
// This is the straightforward, obviously-correct kernel; the tuned kernels (hbt_kernels_v2)
// keep it as their reference and slow path.
#ifndef HBT_KERNELS_V1_CUH_
#define HBT_KERNELS_V1_CUH_

#include "hbt_pair.cuh"

// decode a linear block index into an upper-triangular tile pair (ti <= tj), row-major
__device__ __forceinline__ void tri_decode(long long b, long long T, int &ti, int &tj) {
    // row r starts at S(r) = r*T - r*(r-1)/2
    double Td = static_cast<double>(T);
    long long r = static_cast<long long>(floor((2.0 * Td + 1.0 - sqrt((2.0 * Td + 1.0) * (2.0 * Td + 1.0) - 8.0 * static_cast<double>(b))) * 0.5));
    if (r < 0) r = 0;
    if (r > T - 1) r = T - 1;
    while (r > 0 && r * T - r * (r - 1) / 2 > b) r--;
    while ((r + 1) * T - (r + 1) * r / 2 <= b) r++;
    ti = static_cast<int>(r);
    tj = static_cast<int>(r + (b - (r * T - r * (r - 1) / 2)));
}

// index of the segment that owns thread block b: the last one with block0 <= b
__device__ __forceinline__ int find_segment(const HbtMixSeg *__restrict__ segs, int nseg, long long b) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].block0 <= b) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

// MODE 0: accumulate (closed slabs skipped); MODE 1: ordered-cap pass 1 (count accepted pairs
// per row for the crossing slabs, nothing else is touched); MODE 2: ordered-cap pass 2
// (accumulate only up to the per-slab cut position).
template <int TILE, bool MIXED, int MODE>
__global__ void __launch_bounds__(TILE)
hbt_pairs_v1(const double *__restrict__ p1, const double *__restrict__ p2, long long n_same,
             const HbtMixSeg *__restrict__ segs, const HbtGrid g, const HbtAccum acc,
             const double psi_ref, const HbtCap cap) {
    // same-event: n_same = particles in the merged list; mixed: n_same = number of segments
    constexpr int NC = MIXED ? 4 : 8;
    __shared__ double sj[NC][TILE];
    __shared__ unsigned s_stage[6];
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned *s_slab = reinterpret_cast<unsigned *>(dyn);
    const int nslab_pad = (g.nslab + 1) & ~1;
    const int nqi = g.qinv ? g.nKT * g.nq : 0;
    double *s_qsum = reinterpret_cast<double *>(s_slab + nslab_pad);
    double *s_qcos = s_qsum + nqi;
    unsigned *s_qcnt = reinterpret_cast<unsigned *>(s_qcos + nqi);
    unsigned *s_qpairs = s_qcnt + nqi;  // [nKT] accepted q_inv pairs

    const int t = threadIdx.x;
    long long i0, j0;
    int ni, nj;
    bool diag = false;
    double rc = 1.0, rs = 0.0;
    long long pos_base;  // loop-order position of the tile's first list-2 particle inside a row
    if (MIXED) {
        const HbtMixSeg sg = segs[find_segment(segs, static_cast<int>(n_same), blockIdx.x)];
        const int local = static_cast<int>(blockIdx.x - sg.block0);
        const int ti = local / sg.tiles_j, tj = local - ti * sg.tiles_j;
        i0 = sg.i0 + static_cast<long long>(ti) * TILE;
        j0 = sg.j0 + static_cast<long long>(tj) * TILE;
        ni = min(TILE, sg.ni - ti * TILE);
        nj = min(TILE, sg.nj - tj * TILE);
        rc = sg.c; rs = sg.s;
        pos_base = sg.pos0 + static_cast<long long>(tj) * TILE;
    } else {
        const long long T = (n_same + TILE - 1) / TILE;
        int ti, tj;
        tri_decode(blockIdx.x, T, ti, tj);
        i0 = static_cast<long long>(ti) * TILE;
        j0 = static_cast<long long>(tj) * TILE;
        ni = static_cast<int>(min(static_cast<long long>(TILE), n_same - i0));
        nj = static_cast<int>(min(static_cast<long long>(TILE), n_same - j0));
        diag = (ti == tj);
        pos_base = j0;
    }

    for (int k = t; k < nqi; k += TILE) { s_qsum[k] = 0.0; s_qcos[k] = 0.0; s_qcnt[k] = 0; }
    if (g.qinv) for (int k = t; k < g.nKT; k += TILE) s_qpairs[k] = 0;
    if (t < 6) s_stage[t] = 0;

    // stage the list-2 tile (rotated for mixed events, :522-523)
    for (int k = t; k < nj; k += TILE) {
        const double2 *src = reinterpret_cast<const double2 *>(p2 + 8 * (j0 + k));
        const double2 a = src[0], b = src[1];
        if (MIXED) {
            sj[0][k] = __dsub_rn(__dmul_rn(a.x, rc), __dmul_rn(a.y, rs));
            sj[1][k] = __dadd_rn(__dmul_rn(a.x, rs), __dmul_rn(a.y, rc));
        } else {
            sj[0][k] = a.x;
            sj[1][k] = a.y;
        }
        sj[2][k] = b.x;
        sj[3][k] = b.y;
        if (!MIXED) {
            const double2 c = src[2], d = src[3];
            sj[4][k] = c.x; sj[5][k] = c.y; sj[6][k] = d.x; sj[7][k] = d.y;
        }
    }
    __syncthreads();

    unsigned nB = 0, nC = 0, nD = 0, nE = 0;
    if (t < ni) {
        double a[8];
        {
            const double2 *src = reinterpret_cast<const double2 *>(p1 + 8 * (i0 + t));
            const double2 v0 = src[0], v1 = src[1];
            a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                a[4] = v2.x; a[5] = v2.y; a[6] = v3.x; a[7] = v3.y;
            } else {
                a[4] = a[5] = a[6] = a[7] = 0.0;
            }
        }
        const int jstart = diag ? t + 1 : 0;  // same-event: j > i only (:301)
        for (int j = jstart; j < nj; j++) {
            const double bx = sj[0][j], by = sj[1][j], bz = sj[2][j], bE = sj[3][j];
            if (g.qinv) {
                // q_inv branch (:339-356 / :595-607); needs the K_T bin, so the cut first
                const double Kx = __dmul_rn(0.5, __dadd_rn(a[0], bx));
                const double Ky = __dmul_rn(0.5, __dadd_rn(a[1], by));
                const double K2 = __dadd_rn(__dmul_rn(Kx, Kx), __dmul_rn(Ky, Ky));
                if (K2 >= g.KT_min_sq && K2 <= g.KT_max_sq) {
                    const int iK = __double2int_rz(__ddiv_rn(__dsub_rn(__dsqrt_rn(K2), g.KT_min), g.dKT));
                    const double qx = __dsub_rn(a[0], bx), qy = __dsub_rn(a[1], by);
                    const double qz = __dsub_rn(a[2], bz), qE = __dsub_rn(a[3], bE);
                    const double m2 = __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(qE, qE), __dmul_rn(qx, qx)),
                                                          __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
                    const double qinv = __dsqrt_rn(-m2);
                    if (qinv > g.q_lo && qinv < g.q_hi) {
                        const int iq = __double2int_rz(__ddiv_rn(__dsub_rn(qinv, g.q_base), g.dq));
                        // the q_inv histogram of a K_T bin is its own cap channel (nslab + iK): first 50*needed pairs
                        bool take = iq < g.nq && !(cap.closed && cap.closed[2 * g.nslab + iK + (MIXED ? g.nKT : 0)]);
                        if (MODE == 1) {
                            if (take) {
                                const int x = cap.xidx[g.nslab + iK];
                                if (x >= 0) atomicAdd(&cap.rowcnt[x * cap.nrows + (i0 + t)], 1u);
                            }
                            take = false;
                        }
                        if (MODE == 2 && take) {
                            const long long row = i0 + t, cr = cap.cut_row[g.nslab + iK];
                            if (row > cr || (row == cr && pos_base + j > cap.cut_pos[g.nslab + iK])) take = false;
                        }
                        if (take) {
                            atomicAdd(&s_qpairs[iK], 1u);
                            atomicAdd(&s_qcnt[iK * g.nq + iq], 1u);
                            if (!MIXED) {
                                const double cq = pair_cos(g, qx, qy, qz, qE, a[4] - sj[4][j], a[5] - sj[5][j],
                                                           a[6] - sj[6][j], a[7] - sj[7][j]);
                                atomicAdd(&s_qsum[iK * g.nq + iq], qinv);
                                atomicAdd(&s_qcos[iK * g.nq + iq], cq);
                            }
                        }
                    }
                }
            }
            PairBin pb;
            const int st = pair_literal(g, a[0], a[1], a[2], a[3], bx, by, bz, bE, MIXED, psi_ref, pb);
            if (st == PAIR_REJ_KT) continue;
            nB++;
            if (st == PAIR_REJ_QO) continue;
            nC++;
            if (st == PAIR_REJ_QS) continue;
            nD++;
            if (st == PAIR_REJ_QL) continue;
            nE++;
            if (st == PAIR_REJ_PHI) continue;
            if (st == PAIR_DEFER) {
                // undo the stage counts: the host's literal evaluation recounts the pair
                nB--; nC--; nD--; nE--;
                double b8[8] = {bx, by, bz, bE, 0., 0., 0., 0.};
                if (!MIXED) { b8[4] = sj[4][j]; b8[5] = sj[5][j]; b8[6] = sj[6][j]; b8[7] = sj[7][j]; }
                defer_pair(acc, a, b8, psi_ref, MIXED ? 1 : 0, i0 + t, pos_base + j);
                continue;
            }
            if (cap.closed && cap.closed[pb.slab + (MIXED ? g.nslab : 0)]) continue;  // cap reached earlier
            if (MODE == 1) {
                const int x = cap.xidx[pb.slab];
                if (x >= 0) atomicAdd(&cap.rowcnt[x * cap.nrows + (i0 + t)], 1u);
                continue;
            }
            if (MODE == 2) {
                const long long row = i0 + t, cr = cap.cut_row[pb.slab];
                if (row > cr || (row == cr && pos_base + j > cap.cut_pos[pb.slab])) continue;
            }
            const long long bin = bin_index(g, pb);
            if (MIXED) {
                atomicAdd(&acc.den_count[bin], 1ull);
            } else {
                const double cv = pair_cos(g, a[0] - bx, a[1] - by, a[2] - bz, a[3] - bE, a[4] - sj[4][j],
                                           a[5] - sj[5][j], a[6] - sj[6][j], a[7] - sj[7][j]);
                atomicAdd(&acc.num_count[bin], 1ull);
                atomicAdd(&acc.sum_qo[bin], pb.qo);
                atomicAdd(&acc.sum_qs[bin], pb.qs);
                atomicAdd(&acc.sum_ql[bin], pb.ql);
                atomicAdd(&acc.num_cos[bin], cv);
            }
        }
    }
    // block-level merge of the counters, then one global atomic per non-zero entry
    // (accepted pairs per slab and in total are derived from the bin counts at synchronize)
    if (MODE == 1) return;  // pass 1 only counts rows
    nB = warp_sum(nB); nC = warp_sum(nC); nD = warp_sum(nD); nE = warp_sum(nE);
    if ((t & 31) == 0) {
        atomicAdd(&s_stage[1], nB); atomicAdd(&s_stage[2], nC); atomicAdd(&s_stage[3], nD);
        atomicAdd(&s_stage[4], nE);
    }
    __syncthreads();
    unsigned long long *stage = acc.stage + (MIXED ? 6 : 0);
    if (t >= 1 && t < 5 && s_stage[t]) atomicAdd(&stage[t], static_cast<unsigned long long>(s_stage[t]));
    if (g.qinv) {
        for (int k = t; k < nqi; k += TILE) {
            if (!s_qcnt[k]) continue;
            if (MIXED) {
                atomicAdd(&acc.qinv_den[k], static_cast<unsigned long long>(s_qcnt[k]));
            } else {
                atomicAdd(&acc.qinv_count[k], static_cast<unsigned long long>(s_qcnt[k]));
                atomicAdd(&acc.qinv_sum[k], s_qsum[k]);
                atomicAdd(&acc.qinv_cos[k], s_qcos[k]);
            }
        }
        unsigned long long *nq = MIXED ? acc.npairs_den_qinv : acc.npairs_num_qinv;
        for (int k = t; k < g.nKT; k += TILE)
            if (s_qpairs[k]) atomicAdd(&nq[k], static_cast<unsigned long long>(s_qpairs[k]));
    }
}

// npairs_num[slab] / npairs_den[slab] = sum over the slab's q^3 bins of num_count / den_count.
// grid = 2*nslab blocks: block b < nslab reduces the numerator slab b, the others the denominator.
__global__ void __launch_bounds__(256) hbt_reduce_slabs(const HbtAccum acc, int nslab, long long q3) {
    __shared__ unsigned long long part[256];
    const bool den = static_cast<int>(blockIdx.x) >= nslab;
    const int slab = den ? blockIdx.x - nslab : blockIdx.x;
    const unsigned long long *src = (den ? acc.den_count : acc.num_count) + slab * q3;
    unsigned long long s = 0;
    for (long long k = threadIdx.x; k < q3; k += 256) s += src[k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) (den ? acc.npairs_den : acc.npairs_num)[slab] = part[0];
}

// stage[5] / stage[11] (accepted pairs) = sum of the per-slab counters
__global__ void hbt_finish_stage(const HbtAccum acc, int nslab) {
    if (threadIdx.x >= 2) return;
    const unsigned long long *src = threadIdx.x ? acc.npairs_den : acc.npairs_num;
    unsigned long long s = 0;
    for (int k = 0; k < nslab; k++) s += src[k];
    acc.stage[threadIdx.x ? 11 : 5] = s;
}

// total pair count of a launch (stage A is known analytically on the host)
__global__ void hbt_add_stage_a(const HbtAccum acc, int slot, unsigned long long npairs) {
    atomicAdd(&acc.stage[slot], npairs);
}

// adds host-evaluated deferred pairs back into the device accumulators
__global__ void hbt_apply_corrections(const HbtCorrection *__restrict__ c, int n, const HbtAccum acc,
                                      const HbtStageDelta delta) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 12 && k != 5 && k != 11 && delta.v[k]) atomicAdd(&acc.stage[k], delta.v[k]);
    if (k >= n) return;
    const HbtCorrection x = c[k];
    if (x.mixed) {
        atomicAdd(&acc.den_count[x.bin], 1ull);
    } else {
        atomicAdd(&acc.num_count[x.bin], 1ull);
        atomicAdd(&acc.sum_qo[x.bin], x.qo);
        atomicAdd(&acc.sum_qs[x.bin], x.qs);
        atomicAdd(&acc.sum_ql[x.bin], x.ql);
        atomicAdd(&acc.num_cos[x.bin], x.cosv);
    }
}

#endif  // HBT_KERNELS_V1_CUH_
