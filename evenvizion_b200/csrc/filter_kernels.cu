// Ingest (descriptor pack, norms/ckey, coordinate canonical ids) and K1b/K2
// (Lowe ratio test, many-to-one filter, coordinate de-dup, ordered compaction, gather).
// Byte/integer work bound by memory traffic: one CTA per frame / per pair, coalesced row
// accesses, per-pair tables in shared memory.
#include "evz_common.cuh"
#include <climits>

namespace evz {

constexpr int kMaxKpSmem = 12288;   // per-frame keypoint limit of the shared-memory tables

__device__ __forceinline__ unsigned long long coord_key(float x, float y) {
    // Python dict keys (x, y): 0.0 and -0.0 are the same key
    const unsigned int bx = x == 0.f ? 0u : __float_as_uint(x);
    const unsigned int by = y == 0.f ? 0u : __float_as_uint(y);
    return (static_cast<unsigned long long>(bx) << 32) | by;
}
__device__ __forceinline__ unsigned int hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return static_cast<unsigned int>(k);
}

// one CTA per frame: pack descriptor rows to 128 zero-padded bytes, ckey, coords; zero the padding rows
__global__ void __launch_bounds__(256)
ingest_pack_kernel(const void* __restrict__ raw_desc, int raw_is_f32, int d,
                   const float* __restrict__ raw_coords, const int64_t* __restrict__ raw_off,
                   const int32_t* __restrict__ row_off,
                   uint8_t* __restrict__ desc, int32_t* __restrict__ ckey, float* __restrict__ coords,
                   int32_t* __restrict__ bad_count) {
    const int f = blockIdx.x;
    const int64_t r0 = raw_off[f];
    const int n = static_cast<int>(raw_off[f + 1] - r0);
    const int p0 = row_off[f], p1 = row_off[f + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int bad = 0;
    for (int i = warp; i < p1 - p0; i += nwarps) {
        const int64_t prow = static_cast<int64_t>(p0) + i;
        unsigned int packed = 0;
        if (i < n && lane * 4 < d) {
            if (raw_is_f32) {
                const float* src = static_cast<const float*>(raw_desc) + (r0 + i) * d + lane * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (lane * 4 + j < d) {
                        const float v = src[j];
                        const float r = rintf(v);
                        if (!(v == r) || v < 0.f || v > 255.f) ++bad;
                        packed |= (static_cast<unsigned int>(fminf(fmaxf(r, 0.f), 255.f)) & 255u) << (8 * j);
                    }
                }
            } else {
                const uint8_t* src = static_cast<const uint8_t*>(raw_desc) + (r0 + i) * d + lane * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (lane * 4 + j < d) packed |= static_cast<unsigned int>(src[j]) << (8 * j);
            }
        }
        reinterpret_cast<unsigned int*>(desc + prow * EVZ_DESC_BYTES)[lane] = packed;
        unsigned int s = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const unsigned int b = (packed >> (8 * j)) & 255u; s += b * b; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
        if (lane == 0) ckey[prow] = i < n ? static_cast<int32_t>((s << 8) | static_cast<unsigned int>(prow & 255)) : INT_MAX;
        if (lane < 2) coords[prow * 2 + lane] = i < n ? raw_coords[(r0 + i) * 2 + lane] : 0.f;
    }
    if (bad) atomicAdd(bad_count, bad);
}

// one CTA per frame: canon[i] = smallest keypoint index of the frame with the same (x, y)
__global__ void __launch_bounds__(256)
ingest_canon_kernel(const float* __restrict__ coords, const int64_t* __restrict__ raw_off,
                    const int32_t* __restrict__ row_off, int32_t* __restrict__ canon, int table_size) {
    extern __shared__ int32_t table[];
    const int f = blockIdx.x;
    const int n = static_cast<int>(raw_off[f + 1] - raw_off[f]);
    const int p0 = row_off[f], p1 = row_off[f + 1];
    const float2* c = reinterpret_cast<const float2*>(coords) + p0;
    int cap = 64;
    while (cap < 2 * n) cap <<= 1;
    cap = min(cap, table_size);
    const unsigned int mask = cap - 1;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) table[i] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float2 ci = c[i];
        const unsigned long long key = coord_key(ci.x, ci.y);
        unsigned int s = hash64(key) & mask;
        while (true) {
            int cur = table[s];
            if (cur < 0) {
                cur = atomicCAS(&table[s], -1, i);
                if (cur < 0) break;
            }
            const float2 cc = c[cur];
            if (coord_key(cc.x, cc.y) == key) { atomicMin(&table[s], i); break; }
            s = (s + 1) & mask;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p1 - p0; i += blockDim.x) {
        int out = -1;
        if (i < n) {
            const float2 ci = c[i];
            const unsigned long long key = coord_key(ci.x, ci.y);
            unsigned int s = hash64(key) & mask;
            while (true) {
                const int cur = table[s];
                const float2 cc = c[cur];
                if (coord_key(cc.x, cc.y) == key) { out = cur; break; }
                s = (s + 1) & mask;
            }
        }
        canon[p0 + i] = out;
    }
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024); returns the exclusive
// prefix and the block total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffff, incl, d); if (lane >= d) incl += t; }
    __syncthreads();
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarps ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffff, w, d); if (lane >= d) w += t; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    total = warp_sums[nwarps - 1];
    return incl - v + (warp > 0 ? warp_sums[warp - 1] : 0);
}

struct FilterArgs {
    const int32_t* top2_idx; const int32_t* top2_d2;
    const float* coords; const int32_t* canon;
    const int32_t* row_off; const int32_t* n_kp;
    const int32_t* pair_q; const int32_t* pair_t; const int32_t* out_off;
    double ratio; int min_matching_pts;
    uint8_t* surv; int32_t* m_idx; float* m_pts; int32_t* m_cnt; int32_t* n_filtered; int32_t* status;
};

// one CTA per pair.  Every thread owns FOUR consecutive queries of a sweep (128-bit loads of the top-2 records, the
// canonical indices and the coordinates; one block scan per 2048 queries): the kernel is a chain of dependent
// round trips to L2, so the fewer sweeps and the more loads in flight per trip the better.
__global__ void __launch_bounds__(512)
filter_matches_kernel(const FilterArgs a) {
    extern __shared__ __align__(16) int32_t fsm[];
    __shared__ int warp_sums[32];
    __shared__ int kept_total;
    const int p = blockIdx.x;
    const int qf = a.pair_q[p], tf = a.pair_t[p];
    const int nq = a.n_kp[qf], nt = a.n_kp[tf];
    const int q0 = a.row_off[qf], t0 = a.row_off[tf];
    const int64_t o0 = a.out_off[p];
    const int nq4 = (nq + 3) & ~3;    // the arrays below are padded to a multiple of 4 (the per-row arrays of a frame to 256)
    int32_t* tsel = fsm;              // [nq4] claimed train index of a surviving query, else -1
    int32_t* first_c = tsel + nq4;    // [nq4] smallest kept query of a coordinate group
    int32_t* last_c = first_c + nq4;  // [nq4] largest kept query of a coordinate group
    int32_t* cnt_t = last_c + nq4;    // [nt] surviving queries per train index
    for (int i = threadIdx.x; i < nq; i += blockDim.x) { first_c[i] = INT_MAX; last_c[i] = -1; }
    for (int i = threadIdx.x; i < nt; i += blockDim.x) cnt_t[i] = 0;
    if (threadIdx.x == 0) kept_total = 0;
    __syncthreads();
    const int sweep = blockDim.x * 4;
    // Lowe ratio test: matches[0].distance < matches[1].distance * ratio, f32 sqrt, double compare
    for (int qb = 0; qb < nq; qb += sweep) {
        const int q = qb + 4 * threadIdx.x;
        if (q >= nq) continue;
        // (64-bit loads: out_off[p] is only known to be a row number, and a pair owns exactly n_kp rows)
        const int2* pi = reinterpret_cast<const int2*>(a.top2_idx) + o0 + q;
        const int2* pd = reinterpret_cast<const int2*>(a.top2_d2) + o0 + q;
        const int2 none = make_int2(-1, -1);
        const int2 i0 = pi[0], i1 = q + 1 < nq ? pi[1] : none, i2 = q + 2 < nq ? pi[2] : none, i3 = q + 3 < nq ? pi[3] : none;
        const int2 d0v = pd[0], d1v = q + 1 < nq ? pd[1] : none, d2v = q + 2 < nq ? pd[2] : none, d3v = q + 3 < nq ? pd[3] : none;
        const int id0[4] = {i0.x, i1.x, i2.x, i3.x}, id1[4] = {i0.y, i1.y, i2.y, i3.y};
        const int dd0[4] = {d0v.x, d1v.x, d2v.x, d3v.x}, dd1[4] = {d0v.y, d1v.y, d2v.y, d3v.y};
        int sel[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool s = false;
            if (q + k < nq && id1[k] >= 0) {
                const double d0 = static_cast<double>(__fsqrt_rn(static_cast<float>(dd0[k])));
                const double d1 = static_cast<double>(__fsqrt_rn(static_cast<float>(dd1[k])));
                s = d0 < __dmul_rn(d1, a.ratio);
            }
            sel[k] = s ? id0[k] : -1;
            if (a.surv && q + k < nq) a.surv[o0 + q + k] = s ? 1 : 0;
            if (s) atomicAdd(&cnt_t[id0[k]], 1);
        }
        *reinterpret_cast<int4*>(tsel + q) = make_int4(sel[0], sel[1], sel[2], sel[3]);
    }
    __syncthreads();
    // many-to-one filter, then group the kept queries by coordinate
    int kept = 0;
    for (int qb = 0; qb < nq; qb += sweep) {
        const int q = qb + 4 * threadIdx.x;
        if (q >= nq) continue;
        const int4 cn = *reinterpret_cast<const int4*>(a.canon + q0 + q);
        int4 ts = *reinterpret_cast<const int4*>(tsel + q);
        int sel[4] = {ts.x, ts.y, ts.z, ts.w};
        const int cc[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (sel[k] >= 0 && cnt_t[sel[k]] != 1) sel[k] = -1;
            if (sel[k] >= 0) {
                ++kept;
                atomicMin(&first_c[cc[k]], q + k);
                atomicMax(&last_c[cc[k]], q + k);
            }
        }
        *reinterpret_cast<int4*>(tsel + q) = make_int4(sel[0], sel[1], sel[2], sel[3]);
    }
    if (kept) atomicAdd(&kept_total, kept);
    __syncthreads();
    const int n_filtered = kept_total;
    const bool ok = n_filtered >= a.min_matching_pts;
    // ordered compaction: first position of every coordinate group, value of its last member
    int base = 0;
    const float2* cq = reinterpret_cast<const float2*>(a.coords) + q0;
    const float2* ct = reinterpret_cast<const float2*>(a.coords) + t0;
    for (int qb = 0; qb < nq; qb += sweep) {
        const int q = qb + 4 * threadIdx.x;
        int emit[4] = {0, 0, 0, 0}, tr[4] = {0, 0, 0, 0};
        float2 pa[4], pb[4];
        int n_emit = 0;
        if (ok && q < nq) {
            const int4 ts = *reinterpret_cast<const int4*>(tsel + q);
            const int4 cn = *reinterpret_cast<const int4*>(a.canon + q0 + q);
            const int sel[4] = {ts.x, ts.y, ts.z, ts.w}, cc[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (q + k < nq && sel[k] >= 0 && first_c[cc[k]] == q + k) {
                    emit[k] = 1;
                    tr[k] = tsel[last_c[cc[k]]];
                    pa[k] = cq[q + k];             // the gathers leave before the scan and land behind it
                    pb[k] = ct[tr[k]];
                    ++n_emit;
                }
            }
        }
        int total;
        int pos = base + block_excl_scan(n_emit, warp_sums, total);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (emit[k]) {
                reinterpret_cast<int2*>(a.m_idx)[o0 + pos] = make_int2(q + k, tr[k]);
                reinterpret_cast<float4*>(a.m_pts)[o0 + pos] = make_float4(pa[k].x, pa[k].y, pb[k].x, pb[k].y);
                ++pos;
            }
        }
        base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.m_cnt[p] = ok ? base : 0;
        a.n_filtered[p] = n_filtered;
        a.status[p] = ok ? EVZ_ST_OK : EVZ_ST_FEW_MATCHES;
    }
}

// K5b: concatenate + de-duplicate the static points of several feature types, one CTA per pair.  Shared-memory
// open-addressing table over the a-point coordinates: first[s] / last[s] = smallest / largest element index with the
// key that owns slot s (element = position in the concatenation over the types).
struct ConcatArgs {
    const float* pts[EVZ_MAX_TYPES]; const int32_t* off[EVZ_MAX_TYPES]; const int32_t* cnt[EVZ_MAX_TYPES];
    int n_types;
    const int32_t* status; const int32_t* out_off; float* out_pts; int32_t* out_cnt;
};
constexpr int kConcatTable = 16384;          // slots: load factor <= 0.75 at EVZ_MAX_KP elements

__global__ void __launch_bounds__(256)
concat_dedup_kernel(const ConcatArgs a) {
    extern __shared__ int32_t csm[];
    __shared__ int warp_sums[32];
    int32_t* first = csm;
    int32_t* last = csm + kConcatTable;
    const int p = blockIdx.x;
    int base_t[EVZ_MAX_TYPES + 1];
    base_t[0] = 0;
#pragma unroll
    for (int t = 0; t < EVZ_MAX_TYPES; ++t) base_t[t + 1] = base_t[t] + (t < a.n_types ? a.cnt[t][p] : 0);
    const int total = base_t[EVZ_MAX_TYPES];
    if (a.status[p] != EVZ_ST_OK || total <= 0 || total > EVZ_MAX_KP) { if (threadIdx.x == 0) a.out_cnt[p] = 0; return; }
    auto elem = [&](int e) -> float4 {
        int t = 0;
#pragma unroll
        for (int u = 1; u < EVZ_MAX_TYPES; ++u) t += (u < a.n_types && e >= base_t[u]) ? 1 : 0;
        return reinterpret_cast<const float4*>(a.pts[t])[static_cast<int64_t>(a.off[t][p]) + (e - base_t[t])];
    };
    int cap = 64;
    while (cap < total + total / 3 + 1 && cap < kConcatTable) cap <<= 1;
    const unsigned int mask = cap - 1;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) { first[i] = -1; last[i] = -1; }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const float4 v = elem(e);
        const unsigned long long key = coord_key(v.x, v.y);
        unsigned int s = hash64(key) & mask;
        while (true) {
            int cur = first[s];
            if (cur < 0) {
                cur = atomicCAS(&first[s], -1, e);
                if (cur < 0) { atomicMax(&last[s], e); break; }
            }
            const float4 c = elem(cur);
            if (coord_key(c.x, c.y) == key) { atomicMin(&first[s], e); atomicMax(&last[s], e); break; }
            s = (s + 1) & mask;
        }
    }
    __syncthreads();
    const int64_t o0 = a.out_off[p];
    int base = 0;
    for (int eb = 0; eb < total; eb += blockDim.x) {
        const int e = eb + threadIdx.x;
        int emit = 0, l = 0;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < total) {
            v = elem(e);
            const unsigned long long key = coord_key(v.x, v.y);
            unsigned int s = hash64(key) & mask;
            while (true) {
                const float4 c = elem(first[s]);
                if (coord_key(c.x, c.y) == key) break;
                s = (s + 1) & mask;
            }
            emit = first[s] == e;
            l = last[s];
        }
        int tot;
        const int pos = base + block_excl_scan(emit, warp_sums, tot);
        if (emit) {
            const float4 w = elem(l);
            reinterpret_cast<float4*>(a.out_pts)[o0 + pos] = make_float4(v.x, v.y, w.z, w.w);
        }
        base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.out_cnt[p] = base;
}

}  // namespace evz

extern "C" int evz_ingest(evz_handle* h, const void* raw_desc, int raw_is_f32, int d,
                          const float* raw_coords, const int64_t* raw_off, const int32_t* row_off, int n_frames,
                          uint8_t* desc, int32_t* ckey, float* coords, int32_t* canon, int32_t* bad_count,
                          void* stream) {
    if (!h) return EVZ_E_ARG;
    EVZ_ENTER(h);
    EVZ_REQUIRE(h, raw_desc && raw_coords && raw_off && row_off && desc && ckey && coords && canon && bad_count, "null pointer");
    EVZ_REQUIRE(h, d > 0 && d <= EVZ_DESC_BYTES && d % 4 == 0, "descriptor width must be a multiple of 4, at most 128");
    if (n_frames <= 0) return EVZ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    evz::ingest_pack_kernel<<<n_frames, 256, 0, st>>>(raw_desc, raw_is_f32, d, raw_coords, raw_off, row_off, desc, ckey, coords, bad_count);
    EVZ_LAUNCH_CHECK(h);
    const int table = 32768;    // >= 2 * kMaxKpSmem
    if (!h->attr_canon) {
        EVZ_CUDA_CHECK(h, cudaFuncSetAttribute(evz::ingest_canon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, table * 4));
        h->attr_canon = true;
    }
    evz::ingest_canon_kernel<<<n_frames, 256, table * 4, st>>>(coords, raw_off, row_off, canon, table);
    EVZ_LAUNCH_CHECK(h);
    return EVZ_OK;
}

extern "C" int evz_filter_matches(evz_handle* h, const int32_t* top2_idx, const int32_t* top2_d2,
                                  const float* coords, const int32_t* canon,
                                  const int32_t* row_off, const int32_t* n_kp,
                                  const int32_t* pair_q, const int32_t* pair_t, const int32_t* out_off, int n_pairs,
                                  int max_kp, double ratio, int min_matching_pts,
                                  uint8_t* surv, int32_t* m_idx, float* m_pts, int32_t* m_cnt,
                                  int32_t* n_filtered, int32_t* status, void* stream) {
    if (!h) return EVZ_E_ARG;
    EVZ_ENTER(h);
    EVZ_REQUIRE(h, top2_idx && top2_d2 && coords && canon && row_off && n_kp && pair_q && pair_t && out_off &&
                   m_idx && m_pts && m_cnt && n_filtered && status, "null pointer");
    if (max_kp > evz::kMaxKpSmem) {
        EVZ_SET_ERR(h, "evz_filter_matches: max_kp %d exceeds the supported %d keypoints per frame", max_kp, evz::kMaxKpSmem);
        return EVZ_E_UNSUPPORTED;
    }
    if (n_pairs <= 0) return EVZ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int smem = 16 * (max_kp > 0 ? max_kp : 1) + 64;      // three tables padded to a multiple of 4 entries
    if (smem > h->attr_filter) {
        EVZ_CUDA_CHECK(h, cudaFuncSetAttribute(evz::filter_matches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        h->attr_filter = smem;
    }
    evz::FilterArgs a{top2_idx, top2_d2, coords, canon, row_off, n_kp, pair_q, pair_t, out_off, ratio, min_matching_pts,
                      surv, m_idx, m_pts, m_cnt, n_filtered, status};
    evz::filter_matches_kernel<<<n_pairs, 512, smem, st>>>(a);
    EVZ_LAUNCH_CHECK(h);
    return EVZ_OK;
}

extern "C" int evz_concat_dedup(evz_handle* h, int n_types, const float* const* pts, const int32_t* const* off, const int32_t* const* cnt,
                                int n_pairs, const int32_t* status, int max_total, const int32_t* out_off,
                                float* out_pts, int32_t* out_cnt, void* stream) {
    if (!h) return EVZ_E_ARG;
    EVZ_ENTER(h);
    EVZ_REQUIRE(h, n_types >= 1 && n_types <= EVZ_MAX_TYPES, "n_types must be in [1, EVZ_MAX_TYPES]");
    EVZ_REQUIRE(h, pts && off && cnt && status && out_off && out_pts && out_cnt, "null pointer");
    if (max_total > EVZ_MAX_KP) {
        EVZ_SET_ERR(h, "evz_concat_dedup: max_total %d exceeds the supported %d points per pair", max_total, EVZ_MAX_KP);
        return EVZ_E_UNSUPPORTED;
    }
    if (n_pairs <= 0) return EVZ_OK;
    evz::ConcatArgs a{};
    for (int t = 0; t < n_types; ++t) {
        EVZ_REQUIRE(h, pts[t] && off[t] && cnt[t], "null pointer");
        a.pts[t] = pts[t]; a.off[t] = off[t]; a.cnt[t] = cnt[t];
    }
    a.n_types = n_types; a.status = status; a.out_off = out_off; a.out_pts = out_pts; a.out_cnt = out_cnt;
    const int smem = evz::kConcatTable * 8;
    if (!h->attr_concat) {
        EVZ_CUDA_CHECK(h, cudaFuncSetAttribute(evz::concat_dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        h->attr_concat = true;
    }
    evz::concat_dedup_kernel<<<n_pairs, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    EVZ_LAUNCH_CHECK(h);
    return EVZ_OK;
}
