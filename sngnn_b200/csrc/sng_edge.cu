// Edge-restricted similarity / selection / aggregation kernels (K0, K2, K2b, K3, K4, SDDMM).
//
// All of these are HBM/L2-gather bound: one warp owns one target row; a group of
// G = C/4 lanes owns one in-edge at a time and fetches its source row with ONE 128-bit load per lane, so
// a warp step touches 32/G full rows (coalesced 16..512-byte segments).  Scores, selection and the
// weighted reduction never leave registers / shared memory; nothing of size E is written.
//
// Reference call sites replaced: R: models/models.py:121-137,139-158 (++), :233-263 (+), :322-334 (base).
#include "sng_common.cuh"
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

namespace sng {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

template <int G> struct Unroll { static constexpr int value = (G >= 4) ? 4 : G; };
// forward scoring: all loads of a 32-edge chunk in flight at once while that fits in registers (G <= 8)
template <int G> struct UnrollFwd { static constexpr int value = (G <= 8) ? G : 8; };

// ------------------------------------------------------------------------------------------ K0
__global__ void __launch_bounds__(kThreads) rownorm_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ldx,
                                                          float* __restrict__ xf, int ldf, __half* __restrict__ xh, int ldh,
                                                          float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * kWarpsPerBlock;
    const int kmax = max(xf ? ldf : 0, xh ? ldh : 0);
    for (; row < n; row += stride) {
        const float* xr = x + row * ldx;
        float ss = 0.f;
        for (int k = lane; k < d; k += 32) { float v = __ldg(xr + k); ss = fmaf(v, v, ss); }
        ss = group_sum<32>(ss);
        const float r = fmaxf(sqrtf(ss), kNormEps);
        if (inv && lane == 0) inv[row] = 1.0f / r;
        for (int k = lane; k < kmax; k += 32) {
            const float v = k < d ? __ldg(xr + k) / r : 0.f;     // true division, as F.normalize does; zero padding beyond d
            if (xf && k < ldf) xf[row * ldf + k] = v;
            if (xh && k < ldh) xh[row * ldh + k] = __float2half_rn(v);
        }
    }
}

// ------------------------------------------------------------------------------------------ K2 forward
// Sorted (score desc, arrival order on ties) candidate list of one warp, in shared memory.
struct TopList {
    float* s;
    int* j;
    int cnt;
    float kth;   // score of the last entry once the list is full, else -inf
    __device__ __forceinline__ void insert(float sc, int src, int top_k, int lane) {
        // number of entries that stay in front: those with score >= sc (earlier position wins ties)
        int pos = 0;
        for (int t0 = 0; t0 < cnt; t0 += 32) {
            const int t = t0 + lane;
            pos += __popc(__ballot_sync(0xffffffffu, t < cnt && s[t] >= sc));
        }
        const int ncnt = min(cnt + 1, top_k);
        // shift [pos, ncnt-1) one slot down; top_k <= 64 => two slots per lane
        float a0 = 0.f, a1 = 0.f; int b0 = 0, b1 = 0;
        const int t0 = lane, t1 = lane + 32;
        const bool m0 = t0 > pos && t0 < ncnt, m1 = t1 > pos && t1 < ncnt;
        if (m0) { a0 = s[t0 - 1]; b0 = j[t0 - 1]; }
        if (m1) { a1 = s[t1 - 1]; b1 = j[t1 - 1]; }
        __syncwarp();
        if (m0) { s[t0] = a0; j[t0] = b0; }
        if (m1) { s[t1] = a1; j[t1] = b1; }
        if (lane == 0) { s[pos] = sc; j[pos] = src; }
        __syncwarp();
        cnt = ncnt;
        kth = (cnt == top_k) ? s[top_k - 1] : -CUDART_INF_F;
    }
};

// inv_r[i] = 1 / max(||h_i||, 1e-12): one warp per row.  The edge kernels gather this scalar per edge (the array is
// L2 resident) instead of recomputing every source row's norm from its gathered features.
__global__ void __launch_bounds__(kThreads) row_inv_norm_kernel(const float* __restrict__ h, int64_t n, int c, int64_t ld, float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * kWarpsPerBlock) {
        float ss = 0.f;
        for (int k = lane * 4; k < c; k += 128) { const float4 v = ldg4(h + row * ld + k); ss += dot4(v, v); }
        ss = group_sum<32>(ss);
        if (lane == 0) inv[row] = inv_norm_of(ss);
    }
}

// K2 forward.  One warp per target row; lane e of the warp owns in-edge e of the current 32-edge chunk (source id,
// source 1/norm, score); a group of G lanes fetches one source row per step with one 128-bit load per lane.
// Loads are never predicated: missing edges are redirected to the target row itself and masked afterwards.
template <int G, bool SELECT_ALL>
__global__ void __launch_bounds__(kThreads) edge_topk_agg_fwd_kernel(
    const float* __restrict__ h, const float* __restrict__ inv_r, int n, int row_offset, int c, int ldh, const int* __restrict__ rowptr,
    const int* __restrict__ col, int top_k, float thr, float* __restrict__ out, int ldo,
    int* __restrict__ sel_src, float* __restrict__ sel_w, int* __restrict__ sel_cnt) {
    constexpr int EPW = 32 / G;                 // edges per warp step
    constexpr int U = UnrollFwd<G>::value;      // steps whose loads are issued together
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G;
    const bool ch_ok = q * 4 < c;
    const int c4 = ch_ok ? q * 4 : 0;           // lanes beyond the channel count read channel 0 and contribute with weight 0
    const float* hb = h + c4;
    TopList L;
    L.s = smem + (size_t)warp * 2 * max(top_k, 1);
    L.j = reinterpret_cast<int*>(L.s + max(top_k, 1));
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int row = blockIdx.x * kWarpsPerBlock + warp; row < n; row += gridDim.x * kWarpsPerBlock) {
        const int grow = row_offset + row;
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 ni = scale4(ldg4(hb + (int64_t)grow * ldh), __ldg(inv_r + grow));     // target row, normalised
        if (!ch_ok) ni = z4;
        float4 acc = z4;
        L.cnt = 0; L.kth = -CUDART_INF_F;

        for (int base = beg; base < end; base += 32) {
            const int nchunk = min(32, end - base);
            const bool has = lane < nchunk;
            const int jl = has ? __ldg(col + base + lane) : grow;
            const float irl = __ldg(inv_r + jl);
            float my_d = 0.f;                                    // <n_i, h_j> of edge base+lane
            for (int st = 0; st < nchunk; st += EPW * U) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = __shfl_sync(0xffffffffu, jl, (st + u * EPW + grp) & 31);
                    v[u] = ldg4(hb + (int64_t)j * ldh);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float d = group_sum<G>(dot4(ni, v[u]));
                    if (SELECT_ALL) {
                        // weight of this step's edge = its score, available in the owning lane only after the assembly below;
                        // recompute it here from the edge's own 1/norm instead
                        const int e = st + u * EPW + grp;
                        const float w = __shfl_sync(0xffffffffu, irl, e & 31);
                        if (e < nchunk && ch_ok) fma4(acc, d * w, v[u]);
                    } else {
                        const float t = __shfl_sync(0xffffffffu, d, (lane % EPW) * G);      // hand (step, group) to the edge's lane
                        if (lane / EPW == st / EPW + u) my_d = t;
                    }
                }
            }
            if (!SELECT_ALL) {
                const float my_s = my_d * irl + 0.0f;            // + 0: -0 becomes +0, so equal scores are equal as integers too
                const bool valid = has && my_s >= thr;
                if (L.cnt == 0) {
                    // empty list (normally the row's only chunk): top_k rounds of a one-instruction warp max over an
                    // order-preserving integer image of the score; the lowest lane among the maxima = lowest edge position
                    const unsigned ub = __float_as_uint(my_s);
                    unsigned key = valid ? ((ub & 0x80000000u) ? ~ub : (ub | 0x80000000u)) : 0u;     // 0 = not a candidate
                    int t = 0;
                    for (; t < top_k; ++t) {
                        const unsigned mx = __reduce_max_sync(0xffffffffu, key);
                        if (mx == 0u) break;
                        const int w = __ffs(__ballot_sync(0xffffffffu, key == mx)) - 1;
                        if (lane == w) { L.s[t] = my_s; L.j[t] = jl; key = 0u; }
                    }
                    __syncwarp();
                    L.cnt = t;
                    L.kth = (t == top_k) ? L.s[top_k - 1] : -CUDART_INF_F;
                } else {
                    unsigned m = __ballot_sync(0xffffffffu, valid && (L.cnt < top_k || my_s > L.kth));
                    while (m) {                                              // ascending lane == ascending edge position
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const float sg = __shfl_sync(0xffffffffu, my_s, l);
                        const int jg = __shfl_sync(0xffffffffu, jl, l);
                        if (L.cnt < top_k || sg > L.kth) L.insert(sg, jg, top_k, lane);
                    }
                }
            }
        }
        if (!SELECT_ALL) {
            const int cnt = L.cnt;
            for (int st = 0; st < cnt; st += EPW) {
                const int t = min(st + grp, cnt - 1);                        // clamp: the duplicate gets weight 0
                const float w = st + grp < cnt ? L.s[t] : 0.f;
                fma4(acc, w, ldg4(hb + (int64_t)L.j[t] * ldh));
            }
            if (lane < top_k) {
                sel_src[(int64_t)row * top_k + lane] = lane < cnt ? L.j[lane] : -1;
                sel_w[(int64_t)row * top_k + lane] = lane < cnt ? L.s[lane] : 0.f;
            }
            if (lane + 32 < top_k) {
                sel_src[(int64_t)row * top_k + lane + 32] = lane + 32 < cnt ? L.j[lane + 32] : -1;
                sel_w[(int64_t)row * top_k + lane + 32] = lane + 32 < cnt ? L.s[lane + 32] : 0.f;
            }
            if (lane == 0) sel_cnt[row] = cnt;
            __syncwarp();
        }
        acc.x = cross_group_sum<G>(acc.x); acc.y = cross_group_sum<G>(acc.y);
        acc.z = cross_group_sum<G>(acc.z); acc.w = cross_group_sum<G>(acc.w);
        if (grp == 0 && ch_ok) {
            const float invd = 1.0f / (float)max(end - beg, 1);
            *reinterpret_cast<float4*>(out + (int64_t)row * ldo + q * 4) = scale4(acc, invd);
        }
    }
}

// K2 forward, selection variant (top_k <= 32).  One warp per target row, 32 in-edges per chunk.  A group of G = C/4 lanes
// fetches one source row per step with one 128-bit load per lane (32/G full rows per warp step, G steps per chunk), and
// the G partial dot products a lane then holds are reduced ACROSS its group with a transposing butterfly (G-1 shuffles
// instead of G log2 G), which leaves every lane with the finished score of exactly one edge -- the edge whose source id
// and 1/norm it loaded in the first place.  The running top-k lives in registers (lane t = rank t).  The first chunk is
// ranked by top_k rounds of a one-instruction warp max; later chunks only insert the candidates that beat the k-th score.
// The <= top_k winners are re-gathered at the end (L1 hits), so nothing of a chunk has to stay live across chunks.
template <int G, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) edge_topk_sel_fwd_kernel(
    const float* __restrict__ h, const float* __restrict__ inv_r, int n, int row_offset, int c, int ldh, const int* __restrict__ rowptr,
    const int* __restrict__ col, int top_k, float thr, float* __restrict__ out, int ldo,
    int* __restrict__ sel_src, float* __restrict__ sel_w, int* __restrict__ sel_cnt) {
    constexpr int EPW = 32 / G;                 // edges per warp step
    constexpr int UB = G < 8 ? G : 8;           // steps whose loads are issued together
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G;
    const int my_e = q * EPW + grp;             // the edge of a chunk this lane owns: fetched at step q by group grp
    const bool ch_ok = q * 4 < c;
    const float* hb = h + (ch_ok ? q * 4 : 0);  // lanes beyond the channel count read channel 0 and contribute zeros
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // The next row's metadata (rowptr pair, first chunk of source ids) is loaded at the END of a row, ahead of its use.
    // (Fetching it a whole row ahead costs 16 registers -- one resident block per SM -- and measured no gain.)
    const int stride = gridDim.x * kWarpsPerBlock;
    int row = blockIdx.x * kWarpsPerBlock + warp;
    int beg = 0, end = 0, jl0 = 0;
    if (row < n) {
        beg = __ldg(rowptr + row); end = __ldg(rowptr + row + 1);
        jl0 = my_e < end - beg ? __ldg(col + beg + my_e) : row_offset + row;
    }
    for (; row < n; row += stride) {
        const int grow = row_offset + row;
        const int nrow = row + stride;
        int nbeg = 0, nend = 0, njl = 0;
        float4 ni = scale4(ldg4(hb + (int64_t)grow * ldh), __ldg(inv_r + grow));     // target row, normalised
        if (!ch_ok) ni = z4;
        float ls = 0.f; int lj = -1;            // rank `lane` of the running top-k
        int cnt = 0; float kth = -CUDART_INF_F; // entries in the list; score of rank top_k-1 once full

        for (int base = beg; base < end; base += 32) {
            const int nchunk = min(32, end - base);
            const bool has = my_e < nchunk;
            const int jl = base == beg ? jl0 : (has ? __ldg(col + base + my_e) : grow);   // missing edges read the target row itself
            const float irl = __ldg(inv_r + jl);
            float d[G];
#pragma unroll
            for (int u0 = 0; u0 < G; u0 += UB) {
                float4 v[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int j = __shfl_sync(0xffffffffu, jl, grp * G + u0 + u);     // owner of edge (u0+u)*EPW + grp
                    v[u] = (u0 + u) * EPW < nchunk ? ldg4(hb + (int64_t)j * ldh) : z4;
                }
#pragma unroll
                for (int u = 0; u < UB; ++u) d[u0 + u] = dot4(ni, v[u]);
            }
            // transposing reduction over the group: lane q ends up with sum over the group of d[q]
#pragma unroll
            for (int o = G / 2; o >= 1; o >>= 1) {
                const bool up = (q & o) != 0;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                    const float send = up ? d[i] : d[i + o];
                    const float keep = up ? d[i + o] : d[i];
                    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            const float my_s = d[0] * irl + 0.0f;                                    // + 0: -0 becomes +0, equal scores get equal keys
            const bool cand = has && my_s >= thr && (cnt < top_k || my_s > kth);
            const unsigned ub = __float_as_uint(my_s);
            unsigned key = cand ? ((ub & 0x80000000u) ? ~ub : (ub | 0x80000000u)) : 0u;   // order-preserving image; 0 = not a candidate
            if (cnt == 0) {
                // empty list (the row's first, usually only, chunk): candidates come out in rank order, so rank t goes to lane t
                int t = 0;
                for (; t < top_k; ++t) {
                    const unsigned mx = __reduce_max_sync(0xffffffffu, key);
                    if (mx == 0u) break;
                    const unsigned m = __ballot_sync(0xffffffffu, key == mx);
                    int w = __ffs(m) - 1;
                    if (m & (m - 1)) {                                               // exact tie: the lowest edge position wins
                        const unsigned emin = __reduce_min_sync(0xffffffffu, key == mx ? (unsigned)my_e : 64u);
                        w = __ffs(__ballot_sync(0xffffffffu, key == mx && (unsigned)my_e == emin)) - 1;
                    }
                    const float sw = __shfl_sync(0xffffffffu, my_s, w);
                    const int jw = __shfl_sync(0xffffffffu, jl, w);
                    if (lane == w) key = 0u;
                    if (lane == t) { ls = sw; lj = jw; }
                }
                cnt = t;
                if (cnt == top_k) kth = __shfl_sync(0xffffffffu, ls, top_k - 1);
                continue;
            }
            while (true) {
                const unsigned mx = __reduce_max_sync(0xffffffffu, key);
                if (mx == 0u) break;
                unsigned m = __ballot_sync(0xffffffffu, key == mx);
                int w = __ffs(m) - 1;
                if (m & (m - 1)) {                                                   // exact tie: the lowest edge position wins
                    const unsigned emin = __reduce_min_sync(0xffffffffu, key == mx ? (unsigned)my_e : 64u);
                    w = __ffs(__ballot_sync(0xffffffffu, key == mx && (unsigned)my_e == emin)) - 1;
                }
                const float sw = __shfl_sync(0xffffffffu, my_s, w);
                const int jw = __shfl_sync(0xffffffffu, jl, w);
                if (lane == w) key = 0u;
                // rank of the newcomer: list entries with score >= sw stay in front (earlier positions win ties)
                const int pos = __popc(__ballot_sync(0xffffffffu, lane < cnt && ls >= sw));
                if (pos >= top_k) break;                                            // nothing that remains can enter either
                const float us = __shfl_up_sync(0xffffffffu, ls, 1);
                const int uj = __shfl_up_sync(0xffffffffu, lj, 1);
                if (lane > pos) { ls = us; lj = uj; }
                if (lane == pos) { ls = sw; lj = jw; }
                cnt = min(cnt + 1, top_k);
                if (cnt == top_k) kth = __shfl_sync(0xffffffffu, ls, top_k - 1);
                if (cnt == top_k) key = (my_s > kth) ? key : 0u;                     // candidates the new k-th score rules out
            }
        }
        // weighted sum of the winners' rows (re-gathered: they were loaded moments ago)
        float4 acc = z4;
        for (int st = 0; st < cnt; st += EPW) {
            const int t = min(st + grp, cnt - 1);                                    // clamp: the duplicate gets weight 0
            const float w = __shfl_sync(0xffffffffu, ls, t);
            const int j = __shfl_sync(0xffffffffu, lj, t);
            fma4(acc, st + grp < cnt ? w : 0.f, ldg4(hb + (int64_t)j * ldh));
        }
        if (lane < top_k) {
            sel_src[(int64_t)row * top_k + lane] = lane < cnt ? lj : -1;
            sel_w[(int64_t)row * top_k + lane] = lane < cnt ? ls : 0.f;
        }
        if (lane == 0) sel_cnt[row] = cnt;
        acc.x = cross_group_sum<G>(acc.x); acc.y = cross_group_sum<G>(acc.y);
        acc.z = cross_group_sum<G>(acc.z); acc.w = cross_group_sum<G>(acc.w);
        if (grp == 0 && ch_ok) {
            const float invd = 1.0f / (float)max(end - beg, 1);
            *reinterpret_cast<float4*>(out + (int64_t)row * ldo + q * 4) = scale4(acc, invd);
        }
        if (nrow < n) {
            nbeg = __ldg(rowptr + nrow); nend = __ldg(rowptr + nrow + 1);
            njl = my_e < nend - nbeg ? __ldg(col + nbeg + my_e) : row_offset + nrow;
        }
        beg = nbeg; end = nend; jl0 = njl;
    }
}

// list-only aggregation (all-pairs mode)
template <int G>
__global__ void __launch_bounds__(kThreads) list_agg_fwd_kernel(
    const float* __restrict__ h, int n_rows, int c, int64_t ldh, int list_k, const int* __restrict__ sel_src,
    const float* __restrict__ sel_w, const int* __restrict__ sel_cnt, const float* __restrict__ inv_denom,
    float* __restrict__ out, int64_t ldo) {
    constexpr int EPW = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G, c4 = q * 4;
    const bool ch_ok = c4 < c;
    for (int row = blockIdx.x * kWarpsPerBlock + warp; row < n_rows; row += gridDim.x * kWarpsPerBlock) {
        const int cnt = min(__ldg(sel_cnt + row), list_k);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int st = 0; st < cnt; st += EPW) {
            const int t = st + grp;
            if (t < cnt && ch_ok) {
                const int j = __ldg(sel_src + (int64_t)row * list_k + t);
                fma4(acc, __ldg(sel_w + (int64_t)row * list_k + t), ldg4(h + (int64_t)j * ldh + c4));
            }
        }
        acc.x = cross_group_sum<G>(acc.x); acc.y = cross_group_sum<G>(acc.y);
        acc.z = cross_group_sum<G>(acc.z); acc.w = cross_group_sum<G>(acc.w);
        if (grp == 0 && ch_ok)
            *reinterpret_cast<float4*>(out + (int64_t)row * ldo + c4) = scale4(acc, inv_denom ? __ldg(inv_denom + row) : 1.f);
    }
}

// ------------------------------------------------------------------------------------------ K2b backward
template <int G, bool SELECT_ALL>
__global__ void __launch_bounds__(kThreads) edge_agg_bwd_scatter_kernel(
    const float* __restrict__ h, const float* __restrict__ inv_r, const float* __restrict__ g, int n, int row_offset, int c, int64_t ld, const int* __restrict__ rowptr,
    const int* __restrict__ col, int top_k, const int* __restrict__ sel_src, const float* __restrict__ sel_w,
    const int* __restrict__ sel_cnt, const float* __restrict__ inv_denom, float* __restrict__ dval, float* __restrict__ dnrm) {
    constexpr int EPW = 32 / G;
    constexpr int U = Unroll<G>::value;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G, c4 = q * 4;
    const bool ch_ok = c4 < c;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int row = blockIdx.x * kWarpsPerBlock + warp; row < n; row += gridDim.x * kWarpsPerBlock) {
        int beg, cnt;
        float invd;
        if (SELECT_ALL) {
            beg = __ldg(rowptr + row);
            cnt = __ldg(rowptr + row + 1) - beg;
            invd = 1.0f / (float)max(cnt, 1);
        } else {
            beg = 0;
            cnt = __ldg(sel_cnt + row);
            invd = __ldg(inv_denom + row);
        }
        if (cnt == 0) continue;
        const int grow = row_offset + row;                       // h / inv_r / dval / dnrm hold all nodes, g and the lists are shard-local
        const float4 hi = ch_ok ? ldg4(h + (int64_t)grow * ld + c4) : z4;
        const float4 ni = scale4(hi, __ldg(inv_r + grow));
        const float4 gs = ch_ok ? scale4(ldg4(g + (int64_t)row * ld + c4), invd) : z4;     // g_i / deg_i
        float4 dni = z4;
        for (int st = 0; st < cnt; st += EPW * U) {
            float4 v[U]; int j[U]; float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = st + u * EPW + grp;
                j[u] = -1; w[u] = 0.f;
                if (t < cnt) {
                    if (SELECT_ALL) j[u] = __ldg(col + beg + t);
                    else { j[u] = __ldg(sel_src + (int64_t)row * top_k + t); w[u] = __ldg(sel_w + (int64_t)row * top_k + t); }
                }
                v[u] = (j[u] >= 0 && ch_ok) ? ldg4(h + (int64_t)j[u] * ld + c4) : z4;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (st + u * EPW >= cnt) break;
                const float inv_rj = j[u] >= 0 ? __ldg(inv_r + j[u]) : 0.f;
                const float s = SELECT_ALL ? group_sum<G>(dot4(ni, v[u])) * inv_rj + 0.0f : w[u];
                const float ds = group_sum<G>(dot4(v[u], gs));                 // dL/ds_e = (h_j . g_i)/deg_i
                if (j[u] >= 0 && ch_ok) {
                    atomicAdd(reinterpret_cast<float4*>(dval + (int64_t)j[u] * ld + c4), scale4(gs, s));
                    atomicAdd(reinterpret_cast<float4*>(dnrm + (int64_t)j[u] * ld + c4), scale4(ni, ds));
                    fma4(dni, ds * inv_rj, v[u]);                              // ds * n_j
                }
            }
        }
        dni.x = cross_group_sum<G>(dni.x); dni.y = cross_group_sum<G>(dni.y);
        dni.z = cross_group_sum<G>(dni.z); dni.w = cross_group_sum<G>(dni.w);
        if (grp == 0 && ch_ok) atomicAdd(reinterpret_cast<float4*>(dnrm + (int64_t)grow * ld + c4), dni);
    }
}

// pass 2: dh = dval + (dnrm - n (n . dnrm)) / r      (one lane-group per row)
template <int G>
__global__ void __launch_bounds__(kThreads) norm_bwd_finish_kernel(const float* __restrict__ h, const float* __restrict__ inv_rv, int n, int c, int64_t ld,
                                                                  const float* __restrict__ dval, const float* __restrict__ dnrm,
                                                                  float* __restrict__ dh) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G, c4 = q * 4;
    const bool ch_ok = c4 < c;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = (blockIdx.x * kWarpsPerBlock + warp) * RPW; r0 < n; r0 += gridDim.x * kWarpsPerBlock * RPW) {
        const int row = r0 + grp;
        const bool ok = row < n && ch_ok;
        const float4 hi = ok ? ldg4(h + (int64_t)row * ld + c4) : z4;
        const float inv_r = row < n ? __ldg(inv_rv + row) : 0.f;
        const float4 ni = scale4(hi, inv_r);
        const float4 dn = ok ? ldg4(dnrm + (int64_t)row * ld + c4) : z4;
        const float proj = group_sum<G>(dot4(ni, dn));
        if (ok) {
            const float4 dv = ldg4(dval + (int64_t)row * ld + c4);
            float4 o;
            o.x = dv.x + (dn.x - ni.x * proj) * inv_r; o.y = dv.y + (dn.y - ni.y * proj) * inv_r;
            o.z = dv.z + (dn.z - ni.z * proj) * inv_r; o.w = dv.w + (dn.w - ni.w * proj) * inv_r;
            *reinterpret_cast<float4*>(dh + (int64_t)row * ld + c4) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------ K3 / K4
template <int G>
__device__ __forceinline__ float4 gather_sum(const float* __restrict__ x, int64_t ldx, const int* __restrict__ col,
                                             const float* __restrict__ val, int beg, int end, int c4, bool ch_ok, int grp, int lane) {
    constexpr int EPW = 32 / G;
    constexpr int U = Unroll<G>::value;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = z4;
    for (int base = beg; base < end; base += 32) {
        const int nchunk = min(32, end - base);
        const int jl = lane < nchunk ? __ldg(col + base + lane) : -1;
        const float wl = (val && lane < nchunk) ? __ldg(val + base + lane) : 1.f;
        for (int st = 0; st < nchunk; st += EPW * U) {
            float4 v[U]; float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = st + u * EPW + grp;
                int j = __shfl_sync(0xffffffffu, jl, e & 31);
                w[u] = __shfl_sync(0xffffffffu, wl, e & 31);
                if (e >= nchunk) j = -1;
                v[u] = (j >= 0 && ch_ok) ? ldg4(x + (int64_t)j * ldx + c4) : z4;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) fma4(acc, w[u], v[u]);
        }
    }
    acc.x = cross_group_sum<G>(acc.x); acc.y = cross_group_sum<G>(acc.y);
    acc.z = cross_group_sum<G>(acc.z); acc.w = cross_group_sum<G>(acc.w);
    return acc;
}

template <int G>
__global__ void __launch_bounds__(kThreads) spmm_fwd_kernel(const float* __restrict__ x, int n_rows, int c, int64_t ldx,
                                                           const int* __restrict__ rowptr, const int* __restrict__ col,
                                                           const float* __restrict__ val, const float* __restrict__ rowscale,
                                                           const float* __restrict__ bias, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G, c4 = q * 4;
    const bool ch_ok = c4 < c;
    for (int row = blockIdx.x * kWarpsPerBlock + warp; row < n_rows; row += gridDim.x * kWarpsPerBlock) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 acc = gather_sum<G>(x, ldx, col, val, beg, end, c4, ch_ok, grp, lane);
        if (grp == 0 && ch_ok) {
            if (rowscale) acc = scale4(acc, __ldg(rowscale + row));
            if (bias) { const float4 b = ldg4(bias + c4); acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w; }
            *reinterpret_cast<float4*>(out + (int64_t)row * ldo + c4) = acc;
        }
    }
}

template <int G>
__global__ void __launch_bounds__(kThreads) pp_fuse_fwd_kernel(const float* __restrict__ wt, int n, int c, int64_t ld,
                                                              const int* __restrict__ rowptr, const int* __restrict__ col,
                                                              const float* __restrict__ b_w, const float* __restrict__ beta_p,
                                                              const float* __restrict__ out1, const float* __restrict__ bias,
                                                              float* __restrict__ out0, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane % G, grp = lane / G, c4 = q * 4;
    const bool ch_ok = c4 < c;
    const float beta = __ldg(beta_p);
    for (int row = blockIdx.x * kWarpsPerBlock + warp; row < n; row += gridDim.x * kWarpsPerBlock) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 a = gather_sum<G>(wt, ld, col, nullptr, beg, end, c4, ch_ok, grp, lane);
        if (grp == 0 && ch_ok) {
            const float4 b = ldg4(b_w + c4);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            const int64_t o = (int64_t)row * ld + c4;
            *reinterpret_cast<float4*>(out0 + o) = a;
            const float4 o1 = ldg4(out1 + o);
            float4 r;
            r.x = beta * a.x + (1.f - beta) * o1.x; r.y = beta * a.y + (1.f - beta) * o1.y;
            r.z = beta * a.z + (1.f - beta) * o1.z; r.w = beta * a.w + (1.f - beta) * o1.w;
            if (bias) { const float4 bb = ldg4(bias + c4); r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w; }
            *reinterpret_cast<float4*>(out + o) = r;
        }
    }
}

__global__ void __launch_bounds__(256) pp_beta_grad_kernel(const float* __restrict__ out0, const float* __restrict__ out1,
                                                          const float* __restrict__ g, int64_t numel, float* __restrict__ dbeta) {
    float acc = 0.f;
    const int64_t n4 = numel / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = ldg4(out0 + 4 * i), b = ldg4(out1 + 4 * i), gg = ldg4(g + 4 * i);
        acc = fmaf(a.x - b.x, gg.x, fmaf(a.y - b.y, gg.y, fmaf(a.z - b.z, gg.z, fmaf(a.w - b.w, gg.w, acc))));
    }
    acc = group_sum<32>(acc);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
        v = group_sum<32>(v);
        if (threadIdx.x == 0) atomicAdd(dbeta, v);
    }
}

// ------------------------------------------------------------------------------------------ SDDMM
__global__ void __launch_bounds__(kThreads) sddmm_dot_kernel(const float* __restrict__ xh, int d, int64_t ld,
                                                            const int* __restrict__ a, const int* __restrict__ b,
                                                            int64_t ne, float* __restrict__ s) {
    const int lane = threadIdx.x & 31;
    int64_t e = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const bool vec = (ld % 4 == 0) && (d % 4 == 0);
    for (; e < ne; e += (int64_t)gridDim.x * kWarpsPerBlock) {
        const float* pa = xh + (int64_t)__ldg(a + e) * ld;
        const float* pb = xh + (int64_t)__ldg(b + e) * ld;
        float acc = 0.f;
        if (vec) for (int k = lane * 4; k < d; k += 128) acc += dot4(ldg4(pa + k), ldg4(pb + k));
        else for (int k = lane; k < d; k += 32) acc = fmaf(__ldg(pa + k), __ldg(pb + k), acc);
        acc = group_sum<32>(acc);
        if (lane == 0) s[e] = acc;
    }
}

static int grid_for_rows(int64_t rows, int rows_per_block) {
    int64_t need = (rows + rows_per_block - 1) / rows_per_block;
    int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 1) * 8;   // 8 resident 256-thread CTAs per SM, grid-stride
    int64_t g = need < cap ? need : cap;
    return (int)(g < 1 ? 1 : g);
}

// Grid of a grid-stride row kernel = exactly the number of CTAs that are resident at once (SMs x occupancy of THIS kernel):
// every CTA then gets the same share of rows and there is no partial second wave (a fixed 8 CTAs per SM left the gather
// kernels, which fit 5-6, with a 1.3-1.6 wave launch whose tail ran on a half-empty GPU).
template <typename Kernel>
static int grid_resident(Kernel kernel, int64_t rows, int rows_per_block, size_t smem = 0, int threads = kThreads) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const int64_t need = (rows + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 1) * occ;
    const int64_t g = need < cap ? need : cap;
    return (int)(g < 1 ? 1 : g);
}

#define SNG_DISPATCH_G(cexpr, ...)                                              \
    switch (group_lanes(cexpr)) {                                               \
        case 1: { constexpr int G = 1; __VA_ARGS__; } break;                    \
        case 2: { constexpr int G = 2; __VA_ARGS__; } break;                    \
        case 4: { constexpr int G = 4; __VA_ARGS__; } break;                    \
        case 8: { constexpr int G = 8; __VA_ARGS__; } break;                    \
        case 16: { constexpr int G = 16; __VA_ARGS__; } break;                  \
        default: { constexpr int G = 32; __VA_ARGS__; } break;                  \
    }

static int check_rows(const char* fn, int64_t n, int64_t c, int64_t ld) {
    if (n < 0 || n >= (1ll << 31)) { set_error("%s: n=%lld out of range", fn, (long long)n); return SNG_ERR_ARG; }
    if (c <= 0 || c % 4 != 0 || ld % 4 != 0 || ld < c) { set_error("%s: c=%lld ld=%lld must be multiples of 4, ld>=c", fn, (long long)c, (long long)ld); return SNG_ERR_ARG; }
    if (c > 128) { set_error("%s: c=%lld > 128 channels not supported by the edge kernels", fn, (long long)c); return SNG_ERR_UNSUPPORTED; }
    return SNG_OK;
}

}  // namespace sng

using namespace sng;

extern "C" int sng_rownorm_f32(const float* x, int64_t n, int64_t d, int64_t ldx, float* xhat_f32, int64_t ld_f32,
                               uint16_t* xhat_f16, int64_t ld_f16, float* inv_norm, void* stream) {
    SNG_REQUIRE(x && n >= 0 && d > 0 && ldx >= d && d < (1ll << 30), "sng_rownorm_f32: bad x/n/d/ldx");
    SNG_REQUIRE(!xhat_f32 || (ld_f32 >= d && ld_f32 < (1ll << 30)), "sng_rownorm_f32: ld_f32 < d");
    SNG_REQUIRE(!xhat_f16 || (ld_f16 >= d && ld_f16 < (1ll << 30)), "sng_rownorm_f32: ld_f16 < d");
    if (n == 0) return SNG_OK;
    rownorm_kernel<<<grid_resident(rownorm_kernel, n, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(
        x, n, (int)d, ldx, xhat_f32, (int)ld_f32, reinterpret_cast<__half*>(xhat_f16), (int)ld_f16, inv_norm);
    return check_launch("sng_rownorm_f32");
}

extern "C" int sng_edge_topk_agg_fwd(const float* h, int64_t n_total, int64_t n, int64_t row_offset, int64_t c, int64_t ldh,
                                     const int32_t* rowptr, const int32_t* col, int top_k, float thr, float* out, int64_t ldo,
                                     int32_t* sel_src, float* sel_w, int32_t* sel_cnt, float* inv_norm, void* stream) {
    if (int rc = check_rows("sng_edge_topk_agg_fwd", n, c, ldh)) return rc;
    SNG_REQUIRE(h && rowptr && col && out && ldo % 4 == 0 && ldo >= c, "sng_edge_topk_agg_fwd: null pointer or bad ldo");
    SNG_REQUIRE(row_offset >= 0 && row_offset + n <= n_total && n_total < (1ll << 31) && ldh < (1ll << 31) && ldo < (1ll << 31),
                "sng_edge_topk_agg_fwd: bad row_offset / n_total");
    SNG_REQUIRE(inv_norm, "sng_edge_topk_agg_fwd: inv_norm [n_total] is required");
    SNG_REQUIRE(top_k <= SNG_MAX_TOPK, "sng_edge_topk_agg_fwd: top_k=%d > %d", top_k, SNG_MAX_TOPK);
    SNG_REQUIRE(top_k <= 0 || (sel_src && sel_w && sel_cnt), "sng_edge_topk_agg_fwd: selection outputs required when top_k>0");
    SNG_REQUIRE(top_k <= 0 || thr > -1.1f, "sng_edge_topk_agg_fwd: thr must be > -1.1 (knock-out sentinel of R models.py:153)");
    if (n == 0) return SNG_OK;
    const size_t smem = (size_t)kWarpsPerBlock * 2 * (top_k > 0 ? top_k : 1) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    row_inv_norm_kernel<<<grid_resident(row_inv_norm_kernel, n_total, kWarpsPerBlock), kThreads, 0, st>>>(h, n_total, (int)c, ldh, inv_norm);
    SNG_DISPATCH_G(c,
        if (top_k > 0 && top_k <= 32 && !getenv("SNG_K2_OLD")) {
            constexpr int MB = G >= 16 ? 2 : (G == 8 ? 4 : 6);
            edge_topk_sel_fwd_kernel<G, MB><<<grid_resident(edge_topk_sel_fwd_kernel<G, MB>, n, kWarpsPerBlock), kThreads, 0, st>>>(h, inv_norm, (int)n, (int)row_offset, (int)c, (int)ldh, rowptr, col, top_k, thr, out, (int)ldo, sel_src, sel_w, sel_cnt);
        }
        else if (top_k > 0) edge_topk_agg_fwd_kernel<G, false><<<grid_resident(edge_topk_agg_fwd_kernel<G, false>, n, kWarpsPerBlock, smem), kThreads, smem, st>>>(h, inv_norm, (int)n, (int)row_offset, (int)c, (int)ldh, rowptr, col, top_k, thr, out, (int)ldo, sel_src, sel_w, sel_cnt);
        else edge_topk_agg_fwd_kernel<G, true><<<grid_resident(edge_topk_agg_fwd_kernel<G, true>, n, kWarpsPerBlock, smem), kThreads, smem, st>>>(h, inv_norm, (int)n, (int)row_offset, (int)c, (int)ldh, rowptr, col, 0, thr, out, (int)ldo, nullptr, nullptr, nullptr));
    return check_launch("sng_edge_topk_agg_fwd");
}

extern "C" int sng_list_agg_fwd(const float* h, int64_t n_rows, int64_t c, int64_t ldh, int list_k, const int32_t* sel_src,
                                const float* sel_w, const int32_t* sel_cnt, const float* inv_denom, float* out, int64_t ldo,
                                void* stream) {
    if (int rc = check_rows("sng_list_agg_fwd", n_rows, c, ldh)) return rc;
    SNG_REQUIRE(h && sel_src && sel_w && sel_cnt && out && list_k > 0 && ldo % 4 == 0 && ldo >= c, "sng_list_agg_fwd: bad arguments");
    if (n_rows == 0) return SNG_OK;
    SNG_DISPATCH_G(c, list_agg_fwd_kernel<G><<<grid_resident(list_agg_fwd_kernel<G>, n_rows, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(h, (int)n_rows, (int)c, ldh, list_k, sel_src, sel_w, sel_cnt, inv_denom, out, ldo));
    return check_launch("sng_list_agg_fwd");
}

extern "C" int sng_edge_agg_bwd(const float* h, const float* inv_norm, const float* g, int64_t n_total, int64_t n, int64_t row_offset, int64_t c, int64_t ld,
                                const int32_t* rowptr, const int32_t* col, int top_k, const int32_t* sel_src, const float* sel_w, const int32_t* sel_cnt,
                                const float* inv_denom, float* dval, float* dnrm, float* dh, void* stream) {
    if (int rc = check_rows("sng_edge_agg_bwd", n_total, c, ld)) return rc;
    SNG_REQUIRE(h && inv_norm && g && dval && dnrm && dh, "sng_edge_agg_bwd: null pointer");
    SNG_REQUIRE(n >= 0 && row_offset >= 0 && row_offset + n <= n_total, "sng_edge_agg_bwd: bad n / row_offset / n_total");
    SNG_REQUIRE(top_k > 0 ? (sel_src && sel_w && sel_cnt && inv_denom) : (rowptr && col), "sng_edge_agg_bwd: missing selection list / CSR");
    if (n_total == 0) return SNG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SNG_DISPATCH_G(c,
        if (n > 0 && top_k > 0) edge_agg_bwd_scatter_kernel<G, false><<<grid_resident(edge_agg_bwd_scatter_kernel<G, false>, n, kWarpsPerBlock), kThreads, 0, st>>>(h, inv_norm, g, (int)n, (int)row_offset, (int)c, ld, rowptr, col, top_k, sel_src, sel_w, sel_cnt, inv_denom, dval, dnrm);
        else if (n > 0) edge_agg_bwd_scatter_kernel<G, true><<<grid_resident(edge_agg_bwd_scatter_kernel<G, true>, n, kWarpsPerBlock), kThreads, 0, st>>>(h, inv_norm, g, (int)n, (int)row_offset, (int)c, ld, rowptr, col, 0, nullptr, nullptr, nullptr, nullptr, dval, dnrm);
        norm_bwd_finish_kernel<G><<<grid_resident(norm_bwd_finish_kernel<G>, n_total, kWarpsPerBlock * (32 / G)), kThreads, 0, st>>>(h, inv_norm, (int)n_total, (int)c, ld, dval, dnrm, dh));
    return check_launch("sng_edge_agg_bwd");
}

extern "C" int sng_spmm_fwd(const float* x, int64_t n_rows, int64_t c, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                            const float* val, const float* rowscale, const float* bias, float* out, int64_t ldo, void* stream) {
    if (int rc = check_rows("sng_spmm_fwd", n_rows, c, ldx)) return rc;
    SNG_REQUIRE(x && rowptr && col && out && ldo % 4 == 0 && ldo >= c, "sng_spmm_fwd: null pointer or bad ldo");
    if (n_rows == 0) return SNG_OK;
    SNG_DISPATCH_G(c, spmm_fwd_kernel<G><<<grid_resident(spmm_fwd_kernel<G>, n_rows, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(x, (int)n_rows, (int)c, ldx, rowptr, col, val, rowscale, bias, out, ldo));
    return check_launch("sng_spmm_fwd");
}

extern "C" int sng_pp_fuse_fwd(const float* wt, int64_t n, int64_t c, int64_t ld, const int32_t* rowptr_out, const int32_t* col_out,
                               const float* b_w, const float* beta, const float* out1, const float* bias, float* out0, float* out,
                               void* stream) {
    if (int rc = check_rows("sng_pp_fuse_fwd", n, c, ld)) return rc;
    SNG_REQUIRE(wt && rowptr_out && col_out && b_w && beta && out1 && out0 && out, "sng_pp_fuse_fwd: null pointer");
    if (n == 0) return SNG_OK;
    SNG_DISPATCH_G(c, pp_fuse_fwd_kernel<G><<<grid_resident(pp_fuse_fwd_kernel<G>, n, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(wt, (int)n, (int)c, ld, rowptr_out, col_out, b_w, beta, out1, bias, out0, out));
    return check_launch("sng_pp_fuse_fwd");
}

extern "C" int sng_pp_beta_grad(const float* out0, const float* out1, const float* g, int64_t numel, float* dbeta, void* stream) {
    SNG_REQUIRE(out0 && out1 && g && dbeta && numel >= 0 && numel % 4 == 0, "sng_pp_beta_grad: bad arguments (numel must be a multiple of 4)");
    if (numel == 0) return SNG_OK;
    const int grid = grid_for_rows(numel / 4, 256);
    pp_beta_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out0, out1, g, numel, dbeta);
    return check_launch("sng_pp_beta_grad");
}

extern "C" int sng_sddmm_dot(const float* xhat, int64_t n, int64_t d, int64_t ld, const int32_t* a, const int32_t* b,
                             int64_t num_edges, float* s, void* stream) {
    SNG_REQUIRE(xhat && a && b && s && n >= 0 && d > 0 && ld >= d && num_edges >= 0, "sng_sddmm_dot: bad arguments");
    if (num_edges == 0) return SNG_OK;
    sddmm_dot_kernel<<<grid_resident(sddmm_dot_kernel, num_edges, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(xhat, (int)d, ld, a, b, num_edges, s);
    return check_launch("sng_sddmm_dot");
}

// ------------------------------------------------------------------------------------------ toolbox helpers
// Dense all-pairs cosine S = Xhat Xhat^T in FP32 for the toolbox's *_small variants, which RETURN the N x N values
// (R: SimGFAToolbox/dense.py:138-149).  64x64 output tile per CTA, 16-wide K slices through shared memory, 4x4 micro-tile
// per thread.  Only meant for matrices that are materialised anyway; the kNN builder never forms S.
namespace sng {
__global__ void __launch_bounds__(256) allpairs_dense_kernel(const float* __restrict__ xh, int n, int d, int64_t ld, float* __restrict__ out) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d; k0 += 16) {
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            const int r = t >> 4, k = t & 15;
            As[k][r] = (i0 + r < n && k0 + k < d) ? __ldg(xh + (int64_t)(i0 + r) * ld + k0 + k) : 0.f;
            Bs[k][r] = (j0 + r < n && k0 + k < d) ? __ldg(xh + (int64_t)(j0 + r) * ld + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[k][ty * 4 + u]; b[u] = Bs[k][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
            if (i < n && j < n) out[(int64_t)i * n + j] = acc[u][v];
        }
}

// Per-class sums of unit rows, S[c] = sum_{i: y_i = c} xhat_i  (FP64 accumulation): the closed form behind every
// "sum of all pairwise similarities" metric, sum_{i in a, j in b} <xhat_i, xhat_j> = <S_a, S_b>
// (R: SimGFAToolbox/dense.py:9-30, 104-130, 167-179 compute the same sums by materialising N x N blocks).
__global__ void __launch_bounds__(256) class_sum_kernel(const float* __restrict__ xh, const int* __restrict__ y, int64_t n, int d, int64_t ld,
                                                       int num_classes, double* __restrict__ sums, double* __restrict__ counts) {
    // grid.x tiles rows, grid.y tiles columns (256 per block): each thread owns one column for a slab of rows
    const int col = blockIdx.y * 256 + threadIdx.x;
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
    const bool counter = counts && blockIdx.y == 0 && threadIdx.x == 0;
    int cur = -1;
    double acc = 0.0, run = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        int c = y ? __ldg(y + r) : 0;
        if (c < 0 || c >= num_classes) c = -1;                         // rows with an out-of-range label are ignored
        if (c != cur) {
            if (cur >= 0 && col < d) atomicAdd(sums + (int64_t)cur * d + col, acc);
            if (cur >= 0 && counter) atomicAdd(counts + cur, run);
            cur = c; acc = 0.0; run = 0.0;
        }
        if (c >= 0 && col < d) acc += (double)__ldg(xh + r * ld + col);
        run += 1.0;
    }
    if (cur >= 0 && col < d) atomicAdd(sums + (int64_t)cur * d + col, acc);
    if (cur >= 0 && counter) atomicAdd(counts + cur, run);
}
}  // namespace sng

extern "C" int sng_allpairs_dense_f32(const float* xhat, int64_t n, int64_t d, int64_t ld, float* out, void* stream) {
    SNG_REQUIRE(xhat && out && n > 0 && d > 0 && ld >= d && n < 65536 * 64ll, "sng_allpairs_dense_f32: bad arguments");
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((n + 63) / 64));
    sng::allpairs_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xhat, (int)n, (int)d, ld, out);
    return check_launch("sng_allpairs_dense_f32");
}

extern "C" int sng_class_sums_f64(const float* xhat, const int32_t* y, int64_t n, int64_t d, int64_t ld, int num_classes, double* sums,
                                  double* counts, void* stream) {
    SNG_REQUIRE(xhat && sums && n >= 0 && d > 0 && ld >= d && num_classes >= 1 && (y || num_classes == 1), "sng_class_sums_f64: bad arguments");
    if (n == 0) return SNG_OK;
    const int gx = (int)(n < 4096 ? (n + 63) / 64 : 2048);
    dim3 grid((unsigned)(gx > 0 ? gx : 1), (unsigned)((d + 255) / 256));
    sng::class_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xhat, y, n, (int)d, ld, num_classes, sums, counts);
    return check_launch("sng_class_sums_f64");
}
