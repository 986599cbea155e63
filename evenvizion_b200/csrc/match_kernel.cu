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
// it is staged by 1 KB bulk copies into its own ring.  ||t||^2 - 2 q.t >= -||q||^2 >= -8323200,
// so the key fits in int32, and a signed min over keys is the lexicographic (distance, index)
// minimum: ties go to the lowest train index exactly as OpenCV's batchDistance does.  ||q||^2 is
// added once per row after the reduction.
//
// The K dimension is only 128, so the kernel is bound by the epilogue's instruction issue, not by
// the tensor pipe: an exact running top-2 costs 2.5 min/max per element on the ALU pipe (kChunk = 0,
// kept for A/B).  The default epilogue spends ~1 ALU op per element instead:
//   * every aligned chunk of kChunk (8 or 16) columns is reduced to its minimum key with 3-input
//     mins, and a running top-2 of CHUNK minima is kept per thread;
//   * whenever a chunk minimum beats the thread's running best, the chunk's raw keys are saved to
//     a per-thread shared-memory slot by two (four) predicated 128-bit stores, so at the end of the
//     item the slot holds the keys of the chunk that contains the nearest neighbour;
//   * the second neighbour is the smaller of the second-best chunk minimum and the best of the
//     other keys in that slot -- both exact, no recomputation.
// The running best crosses tile boundaries as a sentinel key (V1 << 8, made distinct from every
// real key of the thread's column range), so a later tile only wins with a strictly smaller distance.
//
// Warp roles (128 + 32 kEW threads, 1 CTA / SM, persistent over items):
//   warp 0 : TMA producer (query block, train tiles, item descriptor ring)   warp 1 : MMA issuer
//   warp 2 : TMEM allocator        warp 3 : ckey producer
//   warps 4.. : kEW epilogue warps; warp w reads TMEM lanes 32*(w%4).. and the column group (w-4)/4
#include "evz_common.cuh"
#include "evz_ptx.cuh"
#include <climits>

namespace evz {

constexpr int kBlockQ      = 256;   // query rows per item (two 128-row MMA sub-tiles)
constexpr int kBlockT      = 256;   // train rows per tile (MMA N)
constexpr int kRowBytes    = 128;   // descriptor bytes = one SWIZZLE_128B row
constexpr int kStages      = 3;     // train-tile ring depth (released by the MMA commit alone)
constexpr int kCkStages    = 8;     // ckey-tile ring depth (released by the epilogue)
constexpr int kTileBytes   = kBlockT * kRowBytes;   // 32 KB
constexpr int kQBytes      = kBlockQ * kRowBytes;   // 32 KB
constexpr int kCkeyBytes   = kBlockT * 4;           // 1 KB
constexpr int kAbsent      = 0x7FFFFF;              // INT32_MAX >> 8: "no neighbour"

struct Item {
    int q_row0, nq_left, t_row0, nt, n_tiles, n_sub, out_row0, pad;
};

// kEW epilogue warps (8 or 16); kChunk keys per save slot (8 or 16; 0 = exact per-element top-2, no slots)
template <int kChunk, int kEW>
struct MatchCfg {
    static constexpr int epi_threads = kEW * 32;
    static constexpr int threads     = 128 + epi_threads;
    static constexpr int groups      = kEW / 4;                   // column groups of the accumulator
    static constexpr int cols        = kBlockT / groups;          // accumulator columns per epilogue thread and tile
    static constexpr int part_stride = epi_threads * 16;          // save slots are laid out [sub][part][thread] x 16 B
    static constexpr int parts       = kChunk / 4;
    // offsets into dynamic shared memory (base aligned to 1024)
    static constexpr int q_off     = 0;                               // 2 x 32 KB
    static constexpr int t_off     = q_off + 2 * kQBytes;             // kStages x 32 KB
    static constexpr int ckey_off  = t_off + kStages * kTileBytes;    // kCkStages x 1 KB
    static constexpr int merge_rows = (groups - 1) * 256;             // one merge buffer: (groups-1) x 256 rows x int4
    static constexpr int merge_off = ckey_off + kCkStages * kCkeyBytes; // two buffers, alternating per item
    static constexpr int slot_off  = merge_off + 2 * merge_rows * 16;
    static constexpr int item_off  = slot_off + 2 * parts * part_stride;
    static constexpr int bar_off   = item_off + 2 * static_cast<int>(sizeof(Item));
    static constexpr int n_bars    = 2 * kStages + 2 * kCkStages + 2 + 2 + 2 + 2;
    static constexpr int tmem_off  = bar_off + n_bars * 8;
    static constexpr int total     = tmem_off + 16;
    static constexpr int smem_bytes = total + 1024;               // + alignment slack
    static_assert(smem_bytes <= 227 * 1024, "match kernel shared memory exceeds 227 KB");
    static_assert(kEW == 8 || kEW == 16, "epilogue warps come in groups of four (TMEM lane quarters)");
};

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
    int neg512;                // -512, passed at run time (see header)
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
    r.pad = 0;
    return r;
}

// lexicographic (value, index) insertion into a running top-2
__device__ __forceinline__ void top2_insert(int v, int i, int& V1, int& I1, int& V2, int& I2) {
    const bool lt1 = (v < V1) || (v == V1 && i < I1);
    const bool lt2 = (v < V2) || (v == V2 && i < I2);
    if (lt1)      { V2 = V1; I2 = I1; V1 = v; I1 = i; }
    else if (lt2) { V2 = v;  I2 = i; }
}

// 128-bit shared-memory load through a 32-bit shared address (the dynamic smem base is realigned by
// integer arithmetic, which hides the address space from nvcc: plain loads would become generic LD)
__device__ __forceinline__ int4 lds128(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int mad_key(uint32_t acc, int mul, int ck) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(static_cast<int>(acc)), "r"(mul), "r"(ck));
    return r;
}

// ---- kChunk = 0: exact running top-2 of packed keys, two elements at a time (5 ALU ops / 2 elements)
__device__ __forceinline__ void top2_pair(int k0, int k1, int& m1, int& m2) {
    const int lo = min(k0, k1), hi = max(k0, k1);
    m2 = __vimin3_s32(m2, hi, max(m1, lo));
    m1 = min(m1, lo);
}

// ---- kChunk = 8 / 16: chunk minima + predicated save of the best chunk's keys
// returns min(cm, m1); when cm < m1 the chunk's keys go to the slot (setp + selp fuse into one
// VIMNMX with a predicate output)
template <int kStride>
__device__ __forceinline__ int save_if_less8(int cm, int m1, uint32_t slot, const int* k) {
    int mn;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.lt.s32 p, %1, %2;\n\t"
                 "selp.s32 %0, %1, %2, p;\n\t"
                 "@p st.shared.v4.b32 [%3], {%4, %5, %6, %7};\n\t"
                 "@p st.shared.v4.b32 [%3+%12], {%8, %9, %10, %11};\n\t}"
                 : "=r"(mn) : "r"(cm), "r"(m1), "r"(slot), "r"(k[0]), "r"(k[1]), "r"(k[2]), "r"(k[3]),
                    "r"(k[4]), "r"(k[5]), "r"(k[6]), "r"(k[7]), "n"(kStride));
    return mn;
}
template <int kStride>
__device__ __forceinline__ int save_if_less16(int cm, int m1, uint32_t slot, const int* k) {
    int mn;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.lt.s32 p, %1, %2;\n\t"
                 "selp.s32 %0, %1, %2, p;\n\t"
                 "@p st.shared.v4.b32 [%3], {%4, %5, %6, %7};\n\t"
                 "@p st.shared.v4.b32 [%3+%20], {%8, %9, %10, %11};\n\t"
                 "@p st.shared.v4.b32 [%3+2*%20], {%12, %13, %14, %15};\n\t"
                 "@p st.shared.v4.b32 [%3+3*%20], {%16, %17, %18, %19};\n\t}"
                 : "=r"(mn) : "r"(cm), "r"(m1), "r"(slot), "r"(k[0]), "r"(k[1]), "r"(k[2]), "r"(k[3]),
                    "r"(k[4]), "r"(k[5]), "r"(k[6]), "r"(k[7]), "r"(k[8]), "r"(k[9]), "r"(k[10]), "r"(k[11]),
                    "r"(k[12]), "r"(k[13]), "r"(k[14]), "r"(k[15]), "n"(kStride));
    return mn;
}

// 32 accumulator columns (r) against their ckey values (ck): update the thread's running (m1, m2)
template <int kChunk, int kEW>
__device__ __forceinline__ void pass32(const uint32_t (&r)[32], const int4 (&ck)[8], int mul, int& m1, int& m2,
                                       int& b1, int& b2, uint32_t slot) {
    if (kChunk == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            top2_pair(mad_key(r[4 * j + 0], mul, ck[j].x), mad_key(r[4 * j + 1], mul, ck[j].y), m1, m2);
            top2_pair(mad_key(r[4 * j + 2], mul, ck[j].z), mad_key(r[4 * j + 3], mul, ck[j].w), b1, b2);
        }
    } else {
        constexpr int kC = kChunk == 0 ? 8 : kChunk;
#pragma unroll
        for (int q = 0; q < 32 / kC; ++q) {
            int k[16];
#pragma unroll
            for (int j = 0; j < kC / 4; ++j) {
                const int4 c = ck[q * (kC / 4) + j];
                k[4 * j + 0] = mad_key(r[q * kC + 4 * j + 0], mul, c.x);
                k[4 * j + 1] = mad_key(r[q * kC + 4 * j + 1], mul, c.y);
                k[4 * j + 2] = mad_key(r[q * kC + 4 * j + 2], mul, c.z);
                k[4 * j + 3] = mad_key(r[q * kC + 4 * j + 3], mul, c.w);
            }
            int cm, mn;
            if (kC == 8) {
                cm = __vimin3_s32(__vimin3_s32(k[0], k[1], k[2]), __vimin3_s32(k[3], k[4], k[5]), min(k[6], k[7]));
                mn = save_if_less8<MatchCfg<kC, kEW>::part_stride>(cm, m1, slot, k);
            } else {
                const int a = __vimin3_s32(k[0], k[1], k[2]), b = __vimin3_s32(k[3], k[4], k[5]), c = __vimin3_s32(k[6], k[7], k[8]);
                const int d = __vimin3_s32(k[9], k[10], k[11]), e = __vimin3_s32(k[12], k[13], k[14]);
                cm = min(__vimin3_s32(a, b, c), __vimin3_s32(d, e, k[15]));
                mn = save_if_less16<MatchCfg<kC, kEW>::part_stride>(cm, m1, slot, k);
            }
            m2 = min(m2, max(m1, cm));
            m1 = mn;
        }
    }
}

// Drain this thread's columns of one accumulator in batches of 32: the TMEM load and the eight
// broadcast ckey loads of a batch are issued together (one exposed latency per batch, hidden by the
// other epilogue warps of the scheduler), and the accumulator is handed back to the MMA warp as soon
// as its last column is in registers, before the last batch is processed.
template <int kChunk, int kEW>
__device__ __forceinline__ void drain_acc(uint32_t taddr, uint32_t ck, int mul, int& m1, int& m2, uint32_t slot,
                                          uint64_t* acc_empty, int lane) {
    constexpr int kCols = MatchCfg<kChunk, kEW>::cols;
    constexpr int kBW = (kEW == 8 && kChunk != 0) ? 64 : 32;   // columns per batch: fewer, longer batches expose fewer load latencies
    constexpr int kBatches = kCols / kBW;
    int b1 = INT_MAX, b2 = INT_MAX;            // kChunk = 0: second interleaved chain
#pragma unroll
    for (int b = 0; b < kBatches; ++b) {
        uint32_t r[kBW / 32][32];
        int4 ckv[kBW / 32][8];
#pragma unroll
        for (int h = 0; h < kBW / 32; ++h) tmem_ld_32x32b_x32(taddr + b * kBW + h * 32, r[h]);
#pragma unroll
        for (int h = 0; h < kBW / 32; ++h)
#pragma unroll
            for (int j = 0; j < 8; ++j) ckv[h][j] = lds128(ck + (b * kBW + h * 32) * 4 + j * 16);
#pragma unroll
        for (int h = 0; h < kBW / 32; ++h) tmem_ld_wait_dep(r[h]);
        if (b == kBatches - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
#pragma unroll
        for (int h = 0; h < kBW / 32; ++h) pass32<kChunk, kEW>(r[h], ckv[h], mul, m1, m2, b1, b2, slot);
    }
    if (kChunk == 0) {
        const int lo = min(m1, b1);
        m2 = __vimin3_s32(max(m1, b1), m2, b2);
        m1 = lo;
    }
}

template <int kChunk, int kEW>
__global__ void __launch_bounds__(MatchCfg<kChunk, kEW>::threads, 1)
match_top2_kernel(const __grid_constant__ CUtensorMap tmap, const MatchArgs args) {
    using Cfg = MatchCfg<kChunk, kEW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_s = smem + Cfg::q_off;
    uint8_t* t_s = smem + Cfg::t_off;
    int32_t* ckey_s = reinterpret_cast<int32_t*>(smem + Cfg::ckey_off);
    int4* merge_s = reinterpret_cast<int4*>(smem + Cfg::merge_off);
    Item* item_s = reinterpret_cast<Item*>(smem + Cfg::item_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::bar_off);
    uint64_t* full = bars;                       // [kStages] train tile landed (TMA tx)
    uint64_t* empty = full + kStages;            // [kStages] MMA commit
    uint64_t* ck_full = empty + kStages;         // [kCkStages] ckey tile landed
    uint64_t* ck_empty = ck_full + kCkStages;    // [kCkStages] kEW arrivals
    uint64_t* q_full = ck_empty + kCkStages;     // [2] query block landed + item descriptor published
    uint64_t* q_empty = q_full + 2;              // [2] MMA commit + kEW arrivals
    uint64_t* acc_full = q_empty + 2;            // [2] accumulator ready (MMA commit)
    uint64_t* acc_empty = acc_full + 2;          // [2] accumulator drained (kEW arrivals)
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + Cfg::tmem_off);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kCkStages; ++i) { mbar_init(&ck_full[i], 1); mbar_init(&ck_empty[i], kEW); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1 + kEW);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEW);
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
                const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
                ++qi;
                mbar_wait(&q_empty[qb], qph ^ 1);
                item_s[qb] = im;                           // published by the arrive below (release)
                if (im.n_tiles == 0) { mbar_arrive(&q_full[qb]); continue; }   // empty train frame: nothing to multiply
                mbar_arrive_expect_tx(&q_full[qb], kQBytes);
                tma_load_2d(q_s + qb * kQBytes, &tmap, 0, im.q_row0, &q_full[qb]);
                for (int n = 0; n < im.n_tiles; ++n) {
                    mbar_wait(&empty[stage], sphase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kTileBytes);
                    tma_load_2d(t_s + stage * kTileBytes, &tmap, 0, im.t_row0 + n * kBlockT, &full[stage]);
                    if (++stage == kStages) { stage = 0; sphase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------- ckey producer
        if (lane == 0) {
            uint32_t cs = 0, cph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const Item im = load_item(args, it);
                for (int n = 0; n < im.n_tiles; ++n) {
                    mbar_wait(&ck_empty[cs], cph ^ 1);
                    mbar_arrive_expect_tx(&ck_full[cs], kCkeyBytes);
                    bulk_load_1d(ckey_s + cs * kBlockT, args.ckey + im.t_row0 + n * kBlockT, kCkeyBytes, &ck_full[cs]);
                    if (++cs == kCkStages) { cs = 0; cph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_u8(128, kBlockT);
            uint32_t stage = 0, sphase = 0, qi = 0, g = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
                ++qi;
                mbar_wait(&q_full[qb], qph);
                const int n_tiles = item_s[qb].n_tiles, n_sub = item_s[qb].n_sub;
                const uint32_t q_addr = smem_u32(q_s + qb * kQBytes);
                for (int n = 0; n < n_tiles; ++n) {
                    mbar_wait(&full[stage], sphase);
                    const uint32_t t_addr = smem_u32(t_s + stage * kTileBytes);
                    for (int sub = 0; sub < n_sub; ++sub, ++g) {
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
        const int et = threadIdx.x - 128;       // epilogue thread id
        const int quarter = warp & 3;           // TMEM lane quarter this warp may read
        const int cgrp = (warp - 4) >> 2;       // column group of the accumulator
        const int colbase = cgrp * Cfg::cols;
        const int row_in_sub = quarter * 32 + lane;
        const int sent_adj = colbase == 0 ? 1 : 0;   // sentinel distinct from every key of this column range
        const uint32_t slot0 = smem_u32(smem + Cfg::slot_off) + et * 16;
        const uint32_t ckey_base = smem_u32(ckey_s);
        uint32_t cs = 0, cph = 0, g = 0, qi = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
            ++qi;
            mbar_wait(&q_full[qb], qph);
            const Item im = item_s[qb];
            // ||q||^2 of this thread's two query rows: loaded now, needed when the item is written out
            int qkey[2] = {0, 0};
            if (cgrp == 0) {
                qkey[0] = __ldg(args.ckey + im.q_row0 + row_in_sub);
                qkey[1] = __ldg(args.ckey + im.q_row0 + 128 + row_in_sub);
            }
            // running best / second best of this thread's column range as packed keys (distance << 8 | column
            // within the tile) plus the tile they came from; decoded once per item
            int B1[2] = {kAbsent * 256, kAbsent * 256}, T1[2] = {-1, -1}, B2[2] = {INT_MAX, INT_MAX}, T2[2] = {-1, -1};
            for (int n = 0; n < im.n_tiles; ++n) {
                mbar_wait(&ck_full[cs], cph);
                const uint32_t ck = ckey_base + (cs * kBlockT + colbase) * 4;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    if (sub < im.n_sub) {
                        const uint32_t acc = g & 1, aph = (g >> 1) & 1;
                        mbar_wait(&acc_full[acc], aph);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + acc * kBlockT + colbase + (static_cast<uint32_t>(quarter * 32) << 16);
                        if (kChunk == 0) {
                            int m1 = INT_MAX, m2 = INT_MAX;
                            drain_acc<kChunk, kEW>(taddr, ck, args.neg512, m1, m2, 0u, &acc_empty[acc], lane);
                            // a later tile only displaces an earlier one with a strictly smaller distance
#pragma unroll
                            for (int w = 0; w < 2; ++w) {
                                const int k = w ? m2 : m1;
                                if ((k >> 8) < (B1[sub] >> 8))      { B2[sub] = B1[sub]; T2[sub] = T1[sub]; B1[sub] = k; T1[sub] = n; }
                                else if ((k >> 8) < (B2[sub] >> 8)) { B2[sub] = k; T2[sub] = n; }
                            }
                        } else {
                            // the running best enters the tile as a sentinel: a chunk only wins (and is saved)
                            // with a strictly smaller distance
                            const int sentinel = (B1[sub] & ~255) - sent_adj;
                            int m1 = sentinel, m2 = INT_MAX;
                            drain_acc<kChunk, kEW>(taddr, ck, args.neg512, m1, m2, slot0 + sub * (Cfg::parts * Cfg::part_stride),
                                                   &acc_empty[acc], lane);
                            if (m1 != sentinel) {
                                const bool demote = m2 == sentinel;      // the old best is the new second best
                                B2[sub] = demote ? B1[sub] : m2; T2[sub] = demote ? T1[sub] : n;
                                B1[sub] = m1; T1[sub] = n;
                            } else if (m2 < (B2[sub] & ~255)) {
                                B2[sub] = m2; T2[sub] = n;
                            }
                        }
                        ++g;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ck_empty[cs]);
                if (++cs == kCkStages) { cs = 0; cph ^= 1; }
            }
            int V1[2], I1[2], V2[2], I2[2];
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                V1[sub] = T1[sub] >= 0 ? B1[sub] >> 8 : kAbsent; I1[sub] = T1[sub] >= 0 ? T1[sub] * kBlockT + (B1[sub] & 255) : -1;
                V2[sub] = T2[sub] >= 0 ? B2[sub] >> 8 : kAbsent; I2[sub] = T2[sub] >= 0 ? T2[sub] * kBlockT + (B2[sub] & 255) : -1;
            }
            if (kChunk != 0) {
                // exact second neighbour: the other keys of the chunk that holds the nearest one
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    if (I1[sub] >= 0) {
                        const int wkey = B1[sub];
                        int cand = INT_MAX;
#pragma unroll
                        for (int part = 0; part < Cfg::parts; ++part) {
                            const int4 k = lds128(slot0 + sub * (Cfg::parts * Cfg::part_stride) + part * Cfg::part_stride);
                            cand = min(cand, k.x == wkey ? INT_MAX : k.x);
                            cand = min(cand, k.y == wkey ? INT_MAX : k.y);
                            cand = min(cand, k.z == wkey ? INT_MAX : k.z);
                            cand = min(cand, k.w == wkey ? INT_MAX : k.w);
                        }
                        const int v = cand >> 8, i = (I1[sub] & ~255) + (cand & 255);
                        if (v < V2[sub] || (v == V2[sub] && i < I2[sub])) { V2[sub] = v; I2[sub] = i; }
                    }
                }
            }
            // merge the column groups through shared memory: groups 1.. hand their candidates to group 0 and
            // move on (bar.arrive); the merge buffer alternates per item, and a writer can be at most two
            // accumulators ahead of group 0, so a buffer is never rewritten before it has been read
            int4* mbuf = merge_s + (qi & 1) * Cfg::merge_rows;
            if (cgrp > 0) {
                mbuf[(cgrp - 1) * 256 + row_in_sub] = make_int4(V1[0], I1[0], V2[0], I2[0]);
                mbuf[(cgrp - 1) * 256 + 128 + row_in_sub] = make_int4(V1[1], I1[1], V2[1], I2[1]);
                named_bar_arrive(1, Cfg::epi_threads);
            } else {
                named_bar_sync(1, Cfg::epi_threads);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                    for (int gq = 0; gq < Cfg::groups - 1; ++gq) {
                        const int4 o = mbuf[gq * 256 + sub * 128 + row_in_sub];
                        if (o.y >= 0) top2_insert(o.x, o.y, V1[sub], I1[sub], V2[sub], I2[sub]);
                        if (o.w >= 0) top2_insert(o.z, o.w, V1[sub], I1[sub], V2[sub], I2[sub]);
                    }
                    const int r = sub * 128 + row_in_sub;
                    if (r < im.nq_left) {
                        const int qn = qkey[sub] >> 8;
                        const int64_t o_row = static_cast<int64_t>(im.out_row0) + r;
                        int2 oi, od;
                        oi.x = I1[sub]; od.x = I1[sub] >= 0 ? V1[sub] + qn : -1;
                        oi.y = I2[sub]; od.y = I2[sub] >= 0 ? V2[sub] + qn : -1;
                        reinterpret_cast<int2*>(args.top2_idx)[o_row] = oi;
                        reinterpret_cast<int2*>(args.top2_d2)[o_row] = od;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_empty[qb]);
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
    evz::MatchArgs a{ckey, row_off, n_kp, pair_q, pair_t, out_off, items, n_items, top2_idx, top2_d2, -512};
    // EVZ_OPT_MATCH_VARIANT: 0 = chunk 8 / 8 epilogue warps (default), 1 = exact per-element top-2 / 8 warps,
    // 2 = chunk 16 / 8 warps, 3 = chunk 8 / 16 warps, 4 = exact per-element / 16 warps
#define EVZ_MATCH_LAUNCH(CH, EW)                                                                                         \
    do {                                                                                                                 \
        using Cfg = evz::MatchCfg<CH, EW>;                                                                               \
        static bool attr_set = false;                                                                                    \
        if (!attr_set) {                                                                                                 \
            EVZ_CUDA_CHECK(h, cudaFuncSetAttribute(evz::match_top2_kernel<CH, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                   Cfg::smem_bytes));                                                    \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        evz::match_top2_kernel<CH, EW><<<h->sm_count, Cfg::threads, Cfg::smem_bytes, st>>>(h->tmap, a);                  \
    } while (0)
    switch (h->opt_match_variant) {
        case 1:  EVZ_MATCH_LAUNCH(0, 8); break;
        case 2:  EVZ_MATCH_LAUNCH(16, 8); break;
        case 3:  EVZ_MATCH_LAUNCH(8, 16); break;
        case 4:  EVZ_MATCH_LAUNCH(0, 16); break;
        default: EVZ_MATCH_LAUNCH(8, 8); break;
    }
#undef EVZ_MATCH_LAUNCH
    EVZ_LAUNCH_CHECK(h);
    return EVZ_OK;
}
