// K1: exact brute-force 2-NN of u8 descriptors as a batched u8 x u8 -> s32 tcgen05 contraction
// with a fused distance / top-2 epilogue.  Replaces cv2 BFMatcher.knnMatch(q, t, 2)
// (reference evenvizion/processing/matching.py:102-108).
//
// Work item = (pair, block of 256 query rows).  For every 256-row train tile the CTA issues
//   acc[sub] (128 x 256, s32, TMEM) = Q[sub] (128 x 128 u8) . T (256 x 128 u8)^T     sub = 0, 1
// as 4 x tcgen05.mma.kind::i8 (K = 32 each).  One train tile in shared memory feeds both query
// sub-tiles, and the two 256-column accumulators double-buffer TMEM: the epilogue drains acc[0]
// while the tensor core fills acc[1].
//
// Epilogue, per accumulator element (row r = query, column c = train):
//   key = ckey[c] - 512 * acc = ((||t_c||^2 - 2 q.t_c) << 8) | c      (one IMAD on the FMA pipe; the
//   multiplier is a kernel argument so that ptxas cannot strength-reduce it onto the ALU pipe)
// ckey[c] = (||t_c||^2 << 8) | (c & 255) comes from the frame store, INT32_MAX for padding rows;
// it is staged next to the train tile by a 1 KB bulk copy.  ||t||^2 - 2 q.t >= -||q||^2 >= -8323200,
// so the key fits in int32, and a signed min over keys is the lexicographic (distance, index)
// minimum: ties go to the lowest train index exactly as OpenCV's batchDistance does.  ||q||^2 is
// added once per row after the reduction.  A running top-2 per row is merged across tiles on
// (value, frame-local index).
//
// The ALU pipe (min/max) bounds the epilogue, so the default variant does not track an exact top-2
// per element.  It reduces every aligned chunk of 8 columns to its minimum with 3-input mins
// (4 ops / 8 elements) and keeps the two smallest CHUNK minima (3 ops / 8 elements).  The smallest
// is the exact nearest neighbour; the second neighbour is either the other chunk minimum or one of
// the 7 remaining columns of the winning chunk, which a per-row fix-up at the end of the item
// recomputes exactly (u8 dp4a, query row from the swizzled smem tile, 1 KB of train rows from L2).
// Variant 1 (EVZ_OPT_MATCH_VARIANT) is the straightforward exact top-2 per element, kept for A/B.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over items):
//   warp 0 : TMA producer        warp 1 : MMA issuer       warp 2 : TMEM allocator
//   warps 4-11 : epilogue; warp w reads TMEM lanes 32*(w%4).. and columns 128*((w-4)/4)..
#include "evz_common.cuh"
#include "evz_ptx.cuh"
#include <climits>

namespace evz {

constexpr int kBlockQ      = 256;   // query rows per item (two 128-row MMA sub-tiles)
constexpr int kBlockT      = 256;   // train rows per tile (MMA N)
constexpr int kRowBytes    = 128;   // descriptor bytes = one SWIZZLE_128B row
constexpr int kStages      = 4;     // train-tile ring depth
constexpr int kEpiWarps    = 8;
constexpr int kThreads     = 128 + kEpiWarps * 32;
constexpr int kTileBytes   = kBlockT * kRowBytes;   // 32 KB
constexpr int kQBytes      = kBlockQ * kRowBytes;   // 32 KB
constexpr int kCkeyBytes   = kBlockT * 4;           // 1 KB
constexpr int kAbsent      = 0x7FFFFF;              // INT32_MAX >> 8: "no neighbour"

struct MatchSmem {
    // offsets into dynamic shared memory (base aligned to 1024)
    static constexpr int q_off     = 0;                               // 2 x 32 KB
    static constexpr int t_off     = q_off + 2 * kQBytes;             // kStages x 32 KB
    static constexpr int ckey_off  = t_off + kStages * kTileBytes;    // kStages x 1 KB
    static constexpr int merge_off = ckey_off + kStages * kCkeyBytes; // 2 x 256 rows x int4
    static constexpr int bar_off   = merge_off + 2 * 256 * 16;
    static constexpr int n_bars    = 2 * kStages + 2 + 2 + 2 + 2;
    static constexpr int tmem_off  = bar_off + n_bars * 8;
    static constexpr int total     = tmem_off + 16;
};
constexpr int kMatchSmemBytes = MatchSmem::total + 1024;   // + alignment slack

struct MatchArgs {
    const int32_t* ckey;
    const int32_t* row_off;
    const int32_t* n_kp;
    const int32_t* pair_q;
    const int32_t* pair_t;
    const int32_t* out_off;
    const int32_t* items;      // [n_items][2] = (pair, query block)
    const int32_t* n_items;    // device scalar
    int32_t* top2_idx;
    int32_t* top2_d2;
    const uint8_t* desc;       // frame store descriptors (fix-up reads train rows through L2)
    int neg512;                // -512, passed at run time (see header)
};

struct Item {
    int q_row0, nq_left, t_row0, nt, n_tiles, n_sub, out_row0;
};

__device__ __forceinline__ Item load_item(const MatchArgs& a, int it) {
    Item r;
    const int p = a.items[2 * it], blk = a.items[2 * it + 1];
    const int qf = a.pair_q[p], tf = a.pair_t[p];
    const int nq = a.n_kp[qf];
    r.q_row0 = a.row_off[qf] + blk * kBlockQ;
    r.nq_left = nq - blk * kBlockQ;                 // valid query rows in this block (may exceed 256)
    r.n_sub = r.nq_left > 128 ? 2 : 1;
    r.t_row0 = a.row_off[tf];
    r.nt = a.n_kp[tf];
    r.n_tiles = (r.nt + kBlockT - 1) / kBlockT;
    r.out_row0 = a.out_off[p] + blk * kBlockQ;
    return r;
}

// running top-2 of packed keys, two elements at a time (5 ALU ops / 2 elements)
__device__ __forceinline__ void top2_pair(int k0, int k1, int& m1, int& m2) {
    const int lo = min(k0, k1), hi = max(k0, k1);
    m2 = __vimin3_s32(m2, hi, max(m1, lo));
    m1 = min(m1, lo);
}
// lexicographic (value, index) insertion into a running top-2
__device__ __forceinline__ void top2_insert(int v, int i, int& V1, int& I1, int& V2, int& I2) {
    const bool lt1 = (v < V1) || (v == V1 && i < I1);
    const bool lt2 = (v < V2) || (v == V2 && i < I2);
    if (lt1)      { V2 = V1; I2 = I1; V1 = v; I1 = i; }
    else if (lt2) { V2 = v;  I2 = i; }
}

// drain one 128-column half of a 128 x 256 accumulator into a tile-local top-2 of packed keys
__device__ __forceinline__ void drain_half(uint32_t taddr, const int32_t* ck, int& m1, int& m2) {
    int a1 = INT_MAX, a2 = INT_MAX, b1 = INT_MAX, b2 = INT_MAX;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + c * 64, r0);
        tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
        tmem_ld_wait();
        const int4* ck4 = reinterpret_cast<const int4*>(ck + c * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int4 k = ck4[j];
            top2_pair(k.x - 512 * static_cast<int>(r0[4 * j + 0]), k.y - 512 * static_cast<int>(r0[4 * j + 1]), a1, a2);
            top2_pair(k.z - 512 * static_cast<int>(r0[4 * j + 2]), k.w - 512 * static_cast<int>(r0[4 * j + 3]), b1, b2);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int4 k = ck4[8 + j];
            top2_pair(k.x - 512 * static_cast<int>(r1[4 * j + 0]), k.y - 512 * static_cast<int>(r1[4 * j + 1]), a1, a2);
            top2_pair(k.z - 512 * static_cast<int>(r1[4 * j + 2]), k.w - 512 * static_cast<int>(r1[4 * j + 3]), b1, b2);
        }
    }
    m1 = min(a1, b1);
    m2 = __vimin3_s32(max(a1, b1), a2, b2);
}

__device__ __forceinline__ int mad_key(uint32_t acc, int mul, int ck) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(static_cast<int>(acc)), "r"(mul), "r"(ck));
    return r;
}
// minimum packed key of 8 consecutive columns: 3 x min3 + 1 x min
__device__ __forceinline__ int chunk_min8(const uint32_t* r, const int4 ka, const int4 kb, int mul) {
    const int m0 = __vimin3_s32(mad_key(r[0], mul, ka.x), mad_key(r[1], mul, ka.y), mad_key(r[2], mul, ka.z));
    const int m1 = __vimin3_s32(mad_key(r[3], mul, ka.w), mad_key(r[4], mul, kb.x), mad_key(r[5], mul, kb.y));
    return __vimin3_s32(m0, m1, min(mad_key(r[6], mul, kb.z), mad_key(r[7], mul, kb.w)));
}
__device__ __forceinline__ void top2_one(int k, int& m1, int& m2) {
    m2 = min(m2, max(m1, k));
    m1 = min(m1, k);
}
// drain one 128-column half into the two smallest CHUNK minima (chunks of 8 columns)
__device__ __forceinline__ void drain_half_chunked(uint32_t taddr, const int32_t* ck, int mul, int& m1, int& m2) {
    int a1 = INT_MAX, a2 = INT_MAX, b1 = INT_MAX, b2 = INT_MAX;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + c * 64, r0);
        tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
        tmem_ld_wait();
        const int4* ck4 = reinterpret_cast<const int4*>(ck + c * 64);
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
            top2_one(chunk_min8(r0 + 8 * q, ck4[2 * q], ck4[2 * q + 1], mul), a1, a2);
            top2_one(chunk_min8(r0 + 8 * q + 8, ck4[2 * q + 2], ck4[2 * q + 3], mul), b1, b2);
        }
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
            top2_one(chunk_min8(r1 + 8 * q, ck4[8 + 2 * q], ck4[8 + 2 * q + 1], mul), a1, a2);
            top2_one(chunk_min8(r1 + 8 * q + 8, ck4[8 + 2 * q + 2], ck4[8 + 2 * q + 3], mul), b1, b2);
        }
    }
    m1 = min(a1, b1);
    m2 = __vimin3_s32(max(a1, b1), a2, b2);
}

// exact second neighbour: the 7 other columns of the winning chunk against (V2, I2)
__device__ __forceinline__ void fixup_second(const uint8_t* q_tile, int r, const uint8_t* desc, const int32_t* ckey,
                                             int t_row0, int I1, int& V2, int& I2) {
    const int cb = I1 & ~7;
    const uint4* trow = reinterpret_cast<const uint4*>(desc + (static_cast<size_t>(t_row0) + cb) * kRowBytes);
    unsigned int dot[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dot[j] = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 qv = *reinterpret_cast<const uint4*>(q_tile + r * kRowBytes + ((c ^ (r & 7)) << 4));   // SWIZZLE_128B
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 tv = __ldg(trow + j * 8 + c);
            dot[j] = __dp4a(qv.x, tv.x, dot[j]); dot[j] = __dp4a(qv.y, tv.y, dot[j]);
            dot[j] = __dp4a(qv.z, tv.z, dot[j]); dot[j] = __dp4a(qv.w, tv.w, dot[j]);
        }
    }
    const int4 k0 = __ldg(reinterpret_cast<const int4*>(ckey + t_row0 + cb));
    const int4 k1 = __ldg(reinterpret_cast<const int4*>(ckey + t_row0 + cb) + 1);
    const int ck[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (cb + j != I1 && ck[j] != INT_MAX) {
            const int v = (ck[j] >> 8) - 2 * static_cast<int>(dot[j]);
            const int idx = cb + j;
            if (v < V2 || (v == V2 && idx < I2) || I2 < 0) { V2 = v; I2 = idx; }
        }
    }
}

template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1)
match_top2_kernel(const __grid_constant__ CUtensorMap tmap, const MatchArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_s = smem + MatchSmem::q_off;
    uint8_t* t_s = smem + MatchSmem::t_off;
    int32_t* ckey_s = reinterpret_cast<int32_t*>(smem + MatchSmem::ckey_off);
    int4* merge_s = reinterpret_cast<int4*>(smem + MatchSmem::merge_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MatchSmem::bar_off);
    uint64_t* full = bars;                       // [kStages] train tile + ckey landed (TMA tx)
    uint64_t* empty = full + kStages;            // [kStages] 1 MMA commit + kEpiWarps arrivals
    uint64_t* q_full = empty + kStages;          // [2]
    uint64_t* q_empty = q_full + 2;              // [2]
    uint64_t* acc_full = q_empty + 2;            // [2] accumulator ready (MMA commit)
    uint64_t* acc_empty = acc_full + 2;          // [2] accumulator drained (kEpiWarps arrivals)
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + MatchSmem::tmem_off);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + kEpiWarps); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1 + kEpiWarps);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_s, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    const int n_items = *args.n_items;

    if (warp == 0) {
        // ------------------------------------------------------------- TMA producer
        if (lane == 0) {
            uint32_t stage = 0, sphase = 0, qi = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const Item im = load_item(args, it);
                if (im.n_tiles == 0) continue;            // empty train frame: nothing to multiply
                const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
                ++qi;
                mbar_wait(&q_empty[qb], qph ^ 1);
                mbar_arrive_expect_tx(&q_full[qb], kQBytes);
                tma_load_2d(q_s + qb * kQBytes, &tmap, 0, im.q_row0, &q_full[qb]);
                for (int n = 0; n < im.n_tiles; ++n) {
                    mbar_wait(&empty[stage], sphase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kTileBytes + kCkeyBytes);
                    tma_load_2d(t_s + stage * kTileBytes, &tmap, 0, im.t_row0 + n * kBlockT, &full[stage]);
                    bulk_load_1d(ckey_s + stage * kBlockT, args.ckey + im.t_row0 + n * kBlockT, kCkeyBytes, &full[stage]);
                    if (++stage == kStages) { stage = 0; sphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_u8(128, kBlockT);
            uint32_t stage = 0, sphase = 0, qi = 0, g = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const Item im = load_item(args, it);
                if (im.n_tiles == 0) continue;            // empty train frame: nothing to multiply
                const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
                ++qi;
                mbar_wait(&q_full[qb], qph);
                const uint32_t q_addr = smem_u32(q_s + qb * kQBytes);
                for (int n = 0; n < im.n_tiles; ++n) {
                    mbar_wait(&full[stage], sphase);
                    const uint32_t t_addr = smem_u32(t_s + stage * kTileBytes);
                    for (int sub = 0; sub < im.n_sub; ++sub, ++g) {
                        const uint32_t acc = g & 1, aph = (g >> 1) & 1;
                        mbar_wait(&acc_empty[acc], aph ^ 1);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; ++k) {
                            const uint64_t da = umma_desc_sw128(q_addr + sub * (128 * kRowBytes) + k * 32);
                            const uint64_t db = umma_desc_sw128(t_addr + k * 32);
                            umma_i8(tmem_base + acc * kBlockT, da, db, idesc, k > 0 ? 1u : 0u);
                        }
                        umma_commit(&acc_full[acc]);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == kStages) { stage = 0; sphase ^= 1; }
                }
                umma_commit(&q_empty[qb]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------- epilogue
        const int e = warp - 4;
        const int quarter = warp & 3;           // TMEM lane quarter this warp may read
        const int half = e >> 2;                // column half of the accumulator
        const int row_in_sub = quarter * 32 + lane;
        uint32_t stage = 0, sphase = 0, g = 0, qi = 0;
        int4* merge2_s = merge_s + 256;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const Item im = load_item(args, it);
            const uint32_t qb = qi & 1;
            if (im.n_tiles > 0) ++qi;
            int V1[2] = {INT_MAX, INT_MAX}, I1[2] = {-1, -1}, V2[2] = {INT_MAX, INT_MAX}, I2[2] = {-1, -1};
            for (int n = 0; n < im.n_tiles; ++n) {
                mbar_wait(&full[stage], sphase);          // ckey tile visible to this thread
                const int32_t* ck = ckey_s + stage * kBlockT + half * 128;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    if (sub < im.n_sub) {
                        const uint32_t acc = g & 1, aph = (g >> 1) & 1;
                        mbar_wait(&acc_full[acc], aph);
                        tc_fence_after();
                        int m1, m2;
                        const uint32_t taddr = tmem_base + acc * kBlockT + half * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
                        if (kVariant == 0) drain_half_chunked(taddr, ck, args.neg512, m1, m2);
                        else               drain_half(taddr, ck, m1, m2);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);
                        if (m1 != INT_MAX) top2_insert(m1 >> 8, (m1 & 255) + n * kBlockT, V1[sub], I1[sub], V2[sub], I2[sub]);
                        if (m2 != INT_MAX) top2_insert(m2 >> 8, (m2 & 255) + n * kBlockT, V1[sub], I1[sub], V2[sub], I2[sub]);
                        ++g;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == kStages) { stage = 0; sphase ^= 1; }
            }
            // merge the two column halves through shared memory
            if (half == 1) {
                merge_s[row_in_sub] = make_int4(V1[0], I1[0], V2[0], I2[0]);
                merge_s[128 + row_in_sub] = make_int4(V1[1], I1[1], V2[1], I2[1]);
            }
            named_bar_sync(1, kEpiWarps * 32);
            if (half == 0) {
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int4 o = merge_s[sub * 128 + row_in_sub];
                    if (o.y >= 0) top2_insert(o.x, o.y, V1[sub], I1[sub], V2[sub], I2[sub]);
                    if (o.w >= 0) top2_insert(o.z, o.w, V1[sub], I1[sub], V2[sub], I2[sub]);
                    if (kVariant == 0) merge2_s[sub * 128 + row_in_sub] = make_int4(V1[sub], I1[sub], V2[sub], I2[sub]);
                }
            }
            if (kVariant == 0) {
                // second barrier: the merged candidates are visible; warps of column half h finish the rows of sub-tile h
                named_bar_sync(1, kEpiWarps * 32);
                const int r = half * 128 + row_in_sub;
                if (r < im.nq_left) {
                    const int4 o = merge2_s[r];
                    int v1 = o.x, i1 = o.y, v2 = o.z, i2 = o.w;
                    if (i1 >= 0) fixup_second(q_s + qb * kQBytes, r, args.desc, args.ckey, im.t_row0, i1, v2, i2);
                    const int qn = args.ckey[im.q_row0 + r] >> 8;
                    const int64_t o_row = static_cast<int64_t>(im.out_row0) + r;
                    int2 oi, od;
                    oi.x = i1; od.x = i1 >= 0 ? v1 + qn : -1;
                    oi.y = i2; od.y = i2 >= 0 ? v2 + qn : -1;
                    reinterpret_cast<int2*>(args.top2_idx)[o_row] = oi;
                    reinterpret_cast<int2*>(args.top2_d2)[o_row] = od;
                }
            } else {
                if (half == 0) {
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int r = sub * 128 + row_in_sub;
                        if (r < im.nq_left) {
                            const int qn = args.ckey[im.q_row0 + r] >> 8;
                            const int64_t o_row = static_cast<int64_t>(im.out_row0) + r;
                            int2 oi, od;
                            oi.x = I1[sub]; od.x = I1[sub] >= 0 ? V1[sub] + qn : -1;
                            oi.y = I2[sub]; od.y = I2[sub] >= 0 ? V2[sub] + qn : -1;
                            reinterpret_cast<int2*>(args.top2_idx)[o_row] = oi;
                            reinterpret_cast<int2*>(args.top2_d2)[o_row] = od;
                        }
                    }
                }
                named_bar_sync(1, kEpiWarps * 32);
            }
            // the query tile in shared memory may now be overwritten (the fix-up read it)
            if (im.n_tiles > 0) { __syncwarp(); if (lane == 0) mbar_arrive(&q_empty[qb]); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// items[i] = (pair, query block): one entry per 256-row block of every pair's query frame.
// Single CTA; pairs are scanned in chunks of blockDim.x.
__global__ void build_items_kernel(const int32_t* n_kp, const int32_t* pair_q, int n_pairs,
                                   int32_t* items, int32_t* n_items, int capacity) {
    __shared__ int warp_sums[32];
    __shared__ int base_s;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int p0 = 0; p0 < n_pairs; p0 += blockDim.x) {
        const int p = p0 + threadIdx.x;
        const int nb = p < n_pairs ? (n_kp[pair_q[p]] + kBlockQ - 1) / kBlockQ : 0;
        int incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffff, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nwarps ? warp_sums[lane] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffff, w, d); if (lane >= d) w += t; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int base = base_s;
        int start = base + incl - nb + (warp > 0 ? warp_sums[warp - 1] : 0);
        for (int b = 0; b < nb; ++b) {
            if (start + b < capacity) { items[2 * (start + b)] = p; items[2 * (start + b) + 1] = b; }
        }
        __syncthreads();
        if (threadIdx.x == 0) base_s = base + warp_sums[nwarps - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_items = min(base_s, capacity);
}

}  // namespace evz

static int ensure_tmap(evz_handle* h, const uint8_t* desc, int64_t total_rows) {
    if (h->tmap_ptr == desc && h->tmap_rows == total_rows) return EVZ_OK;
    if (!h->encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        EVZ_CUDA_CHECK(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            EVZ_SET_ERR(h, "cuTensorMapEncodeTiled not available from the driver");
            return EVZ_E_CUDA;
        }
        h->encode = reinterpret_cast<evz_encode_tiled_fn>(fn);
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(evz::kRowBytes), static_cast<cuuint64_t>(total_rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(evz::kRowBytes)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(evz::kRowBytes), 256u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = h->encode(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(desc), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        EVZ_SET_ERR(h, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld)", static_cast<int>(r), static_cast<long long>(total_rows));
        return EVZ_E_CUDA;
    }
    h->tmap_ptr = desc;
    h->tmap_rows = total_rows;
    return EVZ_OK;
}

extern "C" int evz_match_top2(evz_handle* h, const uint8_t* desc, const int32_t* ckey, int64_t total_rows,
                              const int32_t* row_off, const int32_t* n_kp,
                              const int32_t* pair_q, const int32_t* pair_t, const int32_t* out_off, int n_pairs,
                              int32_t* top2_idx, int32_t* top2_d2, void* stream) {
    if (!h) return EVZ_E_ARG;
    EVZ_REQUIRE(h, desc && ckey && row_off && n_kp && pair_q && pair_t && out_off && top2_idx && top2_d2, "null pointer");
    EVZ_REQUIRE(h, total_rows > 0 && total_rows % EVZ_ROW_ALIGN == 0 && total_rows < (int64_t(1) << 31), "total_rows must be a positive multiple of 256 below 2^31");
    EVZ_REQUIRE(h, (reinterpret_cast<uintptr_t>(desc) & 127) == 0 && (reinterpret_cast<uintptr_t>(ckey) & 15) == 0, "desc must be 128-byte and ckey 16-byte aligned");
    if (n_pairs <= 0) return EVZ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = ensure_tmap(h, desc, total_rows);
    if (rc) return rc;
    // every pair owns at most ceil(n_kp/256) <= rows/256 + 1 items
    const size_t capacity = static_cast<size_t>(total_rows / evz::kBlockQ) + static_cast<size_t>(n_pairs);
    EVZ_REQUIRE(h, capacity < (size_t(1) << 30), "too many work items");
    void* scr = nullptr;
    rc = evz_scratch(h, capacity * 8 + 256, &scr);
    if (rc) return rc;
    int32_t* n_items = static_cast<int32_t*>(scr);
    int32_t* items = n_items + 64;
    evz::build_items_kernel<<<1, 1024, 0, st>>>(n_kp, pair_q, n_pairs, items, n_items, static_cast<int>(capacity));
    EVZ_LAUNCH_CHECK(h);
    if (!h->match_attr_set) {
        EVZ_CUDA_CHECK(h, cudaFuncSetAttribute(evz::match_top2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, evz::kMatchSmemBytes));
        EVZ_CUDA_CHECK(h, cudaFuncSetAttribute(evz::match_top2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, evz::kMatchSmemBytes));
        h->match_attr_set = true;
    }
    evz::MatchArgs a{ckey, row_off, n_kp, pair_q, pair_t, out_off, items, n_items, top2_idx, top2_d2, desc, -512};
    if (h->opt_match_variant == 1)
        evz::match_top2_kernel<1><<<h->sm_count, evz::kThreads, evz::kMatchSmemBytes, st>>>(h->tmap, a);
    else
        evz::match_top2_kernel<0><<<h->sm_count, evz::kThreads, evz::kMatchSmemBytes, st>>>(h->tmap, a);
    EVZ_LAUNCH_CHECK(h);
    return EVZ_OK;
}
