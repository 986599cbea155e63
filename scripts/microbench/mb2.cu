// Round-2 microbenchmarks behind the match-kernel redesign (sm_100a).  One binary, several sections:
//   sts    : cost of (predicated) STS.128 per SM by number of active lanes / warps
//   alu    : VIMNMX3 rate, vote + branch cost
//   mma    : tcgen05.mma.kind::i8 issue rate -- A from shared memory (SS) vs A from tensor memory (TS), N = 256 / 128
//   layout : correctness / layout of the A operand in tensor memory (tcgen05.st and tcgen05.cp)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mb2 mb2.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ------------------------------------------------------------------------------------------------ sts
// mode 0: all lanes store; 1: predicate false on all lanes; 2: lane 0 only; 3: lanes 0-7; 4: uniform branch around
// the stores (not taken); 5: two random lanes per warp (changing per iteration)
__global__ void __launch_bounds__(1024, 1) sts_kernel(int mode, int iters, int zero, long long* cycles) {
    extern __shared__ uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const uint32_t addr = smem_u32(smem) + threadIdx.x * 16;
    const uint32_t stride = blockDim.x * 16;
    uint32_t a = threadIdx.x, b = a * 3, c = a * 5, d = a * 7;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        int on;
        if (mode == 0) on = 1;
        else if (mode == 1) on = zero;
        else if (mode == 2) on = (lane == 0) | zero;
        else if (mode == 3) on = (lane < 8) | zero;
        else if (mode == 5) on = (lane == (i & 31)) | (lane == ((i * 7 + 3) & 31)) | zero;
        else on = zero;
        if (mode == 4) {
            if (__any_sync(0xffffffff, on)) {
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr + stride), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr + 2 * stride), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr + 3 * stride), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
            }
            a += i;
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
                         "@p st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n\t"
                         "@p st.shared.v4.b32 [%0+%6], {%1,%2,%3,%4};\n\t"
                         "@p st.shared.v4.b32 [%0+2*%6], {%1,%2,%3,%4};\n\t"
                         "@p st.shared.v4.b32 [%0+3*%6], {%1,%2,%3,%4};\n\t}"
                         :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(on), "n"(16384) : "memory");
            a += i;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ------------------------------------------------------------------------------------------------ alu
// mode 0: 8 independent VIMNMX3 chains; 1: VIMNMX3 + vote.any + (not taken) branch per 8 ops; 2: IMAD chains (fma pipe)
__global__ void __launch_bounds__(1024, 1) alu_kernel(int mode, int iters, int zero, long long* cycles, uint32_t* sink) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * (j + 1);
    uint32_t x = threadIdx.x ^ 0x55, y = threadIdx.x * 9;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(x), "r"(y));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("max.u32 %0, %0, %1;\n\tmax.u32 %0, %0, %2;" : "+r"(v[j]) : "r"(x), "r"(y));
            if (mode == 1) {
                if (__any_sync(0xffffffff, (v[0] == 0x12345u) | zero)) { v[1] += clock(); }
            }
        }
        x += 1; y += 3;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= v[j];
    if (s == 0x12345678u) sink[0] = s;
}
__global__ void __launch_bounds__(1024, 1) alu3_kernel(int iters, long long* cycles, uint32_t* sink) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * (j + 1);
    uint32_t x = threadIdx.x ^ 0x55, y = threadIdx.x * 9;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __vimax3_u32(v[j], x, y);
        x += 1; y += 3;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= v[j];
    if (s == 0x12345678u) sink[0] = s;
}

// ------------------------------------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo >> 4) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_u8(uint32_t m, uint32_t n) {
    return (2u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
#define TMEM_ST32(addr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
    :: "r"(addr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]), \
       "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory")
#define TMEM_LD32(addr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]), \
      "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(addr) : "memory")

// ------------------------------------------------------------------------------------------------ mma rate
// One thread issues `groups` groups of `per_group` MMAs (K blocks) into rotating accumulators, one commit per group.
// ts: A from tensor memory.  n: MMA N.  nacc accumulators of n columns each from column 0; A (TS) at column 448.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int ts, int n, int nacc, int groups, int per_group, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    uint8_t* a_s = smem;                  // 128 rows x 128 B (SW128 layout, content arbitrary)
    uint8_t* b_s = smem + 16384;          // 256 rows x 128 B
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    {   // A in tensor memory: 32 columns at 448 (arbitrary content)
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = threadIdx.x * 2654435761u + j * 40503u;
        const uint32_t addr = tmem + 448 + (static_cast<uint32_t>(warp * 32) << 16);
        TMEM_ST32(addr, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_u8(128, n);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tmem + (g % nacc) * n;
            for (int k = 0; k < per_group; ++k) {
                const uint64_t db = desc_sw128(smem_u32(b_s) + (k & 3) * 32);
                if (ts) mma_ts(d, tmem + 448 + (k & 3) * 8, db, idesc, k > 0);
                else    mma_ss(d, desc_sw128(smem_u32(a_s) + (k & 3) * 32), db, idesc, k > 0);
            }
        }
        mma_commit(&bar);
        while (!mbar_test(&bar, 0)) { }
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ layout
// D[128][128] = A[128][128 B] . B[128][128 B]^T with A in tensor memory.
// how = 0: A written with tcgen05.st (thread = row, register j = bytes 4j..4j+3 of the row);
// how = 1: tcgen05.cp.128x256b from a no-swizzle K-block panel ([m/8][k half][m%8] x 16 B), one copy per K block;
// how = 2: tcgen05.cp.128x256b from the SWIZZLE_128B row-major tile (descriptor start + 32 * k block);
// how = 3: reference: A from shared memory (SS), no-swizzle panels
// dump: raw tensor-memory contents of the 32 A columns after the copy (how = 1, 2)
__global__ void __launch_bounds__(128, 1) layout_kernel(int how, const uint8_t* A, const uint8_t* B, int32_t* D, uint32_t* dump) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    uint8_t* a_pan = smem;                 // 4 panels x 4 KB  (no swizzle)
    uint8_t* b_pan = smem + 16384;         // 4 panels x 4 KB
    uint8_t* a_sw  = smem + 32768;         // 128 rows x 128 B, SWIZZLE_128B
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) {
        const int m = i >> 7, kb = i & 127, blk = kb >> 5, k = kb & 31;
        const int off = blk * 4096 + (m >> 3) * 256 + (k >> 4) * 128 + (m & 7) * 16 + (k & 15);
        a_pan[off] = A[i];
        b_pan[off] = B[i];
        // SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
        a_sw[m * 128 + ((((kb >> 4) ^ (m & 7)) << 4) | (kb & 15))] = A[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    const uint32_t a_col = 256;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    if (how == 0) {
        uint32_t r[32];
        const uint32_t* row = reinterpret_cast<const uint32_t*>(A + threadIdx.x * 128);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = row[j];
        const uint32_t addr = tmem + a_col + lane_addr;
        TMEM_ST32(addr, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_u8(128, 128);
        if (how == 1 || how == 2) {
            for (int k = 0; k < 4; ++k) {
                const uint64_t sd = how == 1 ? desc_nosw(smem_u32(a_pan) + k * 4096, 128, 256) : desc_sw128(smem_u32(a_sw) + k * 32);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(tmem + a_col + k * 8), "l"(sd) : "memory");
            }
        }
        for (int k = 0; k < 4; ++k) {
            const uint64_t db = desc_nosw(smem_u32(b_pan) + k * 4096, 128, 256);
            if (how == 3) mma_ss(tmem, desc_nosw(smem_u32(a_pan) + k * 4096, 128, 256), db, idesc, k > 0);
            else          mma_ts(tmem, tmem + a_col + k * 8, db, idesc, k > 0);
        }
        mma_commit(&bar);
        while (!mbar_test(&bar, 0)) { }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int b = 0; b < 4; ++b) {
        uint32_t r[32];
        const uint32_t addr = tmem + b * 32 + lane_addr;
        TMEM_LD32(addr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) D[threadIdx.x * 128 + b * 32 + j] = static_cast<int32_t>(r[j]);
    }
    if (dump) {
        uint32_t r[32];
        const uint32_t addr = tmem + a_col + lane_addr;
        TMEM_LD32(addr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) dump[threadIdx.x * 32 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}


// ------------------------------------------------------------------------------------------------ mma2: issue patterns
// pattern 0: groups of `per` MMAs on one accumulator, accumulators rotate per group (= mma section)
// pattern 1: `nacc` accumulators interleaved: k block 0 of every accumulator, then k block 1, ...
// pattern 2: every MMA independent (accumulate = 0), rotating over nacc accumulators
__global__ void __launch_bounds__(128, 1) mma_pattern_kernel(int ts, int n, int nacc, int pattern, int rounds, int per, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + 16384;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = threadIdx.x * 2654435761u + j * 40503u;
        const uint32_t addr = tmem + 448 + (static_cast<uint32_t>(warp * 32) << 16);
        TMEM_ST32(addr, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_u8(128, n);
        const long long t0 = clock64();
        for (int g = 0; g < rounds; ++g) {
            if (pattern == 0) {
                for (int a = 0; a < nacc; ++a)
                    for (int k = 0; k < per; ++k) {
                        const uint64_t db = desc_sw128(smem_u32(b_s) + (k & 3) * 32);
                        if (ts) mma_ts(tmem + a * n, tmem + 448 + (k & 3) * 8, db, idesc, k > 0);
                        else    mma_ss(tmem + a * n, desc_sw128(smem_u32(a_s) + (k & 3) * 32), db, idesc, k > 0);
                    }
            } else {
                for (int k = 0; k < per; ++k)
                    for (int a = 0; a < nacc; ++a) {
                        const uint64_t db = desc_sw128(smem_u32(b_s) + (k & 3) * 32 + (n <= 128 ? (a & 1) * 16384 : 0));
                        const uint32_t accum = pattern == 2 ? 0u : (k > 0);
                        if (ts) mma_ts(tmem + a * n, tmem + 448 + (k & 3) * 8, db, idesc, accum);
                        else    mma_ss(tmem + a * n, desc_sw128(smem_u32(a_s) + (k & 3) * 32), db, idesc, accum);
                    }
            }
        }
        mma_commit(&bar);
        while (!mbar_test(&bar, 0)) { }
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ sts2: clean loops
// kWidth: 4 / 8 / 16 bytes per lane; kMode 0: all lanes, 1: predicate false, 2: lane 0 only, 3: no store at all (loop overhead)
template <int kWidth, int kMode>
__global__ void __launch_bounds__(1024, 1) sts2_kernel(int iters, int zero, long long* cycles) {
    extern __shared__ uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const uint32_t addr = smem_u32(smem) + threadIdx.x * 16;
    uint32_t a = threadIdx.x, b = a * 3, c = a * 5, d = a * 7;
    const int on = kMode == 0 ? 1 : kMode == 1 ? zero : ((lane == 0) | zero);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (kMode != 3) {
                if (kWidth == 16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n\t}"
                                 :: "r"(addr + (u & 3) * 16384), "r"(a), "r"(b), "r"(c), "r"(d), "r"(on) : "memory");
                else if (kWidth == 8)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1,%2};\n\t}"
                                 :: "r"(addr + (u & 3) * 16384), "r"(a), "r"(b), "r"(on) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.b32 [%0], %1;\n\t}"
                                 :: "r"(addr + (u & 3) * 16384), "r"(a), "r"(on) : "memory");
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int kWidth, int kMode>
static void run_sts2(int sms, int nw, long long* d_cycles, const char* name) {
    const int iters = 2000;
    cudaFuncSetAttribute(sts2_kernel<kWidth, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 4 * 16384);
    sts2_kernel<kWidth, kMode><<<sms, nw * 32, 65536 + 4 * 16384>>>(iters, 0, d_cycles);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < sms; ++i) s += h[i];
    s /= sms;
    printf("sts2 warps/SM %2d  %2d B/lane %-12s: %.2f clk per store instruction per SM, %.1f clk per warp per store\n", nw, kWidth, name,
           s / (iters * 8.0 * nw), s / (iters * 8.0));
}


// ------------------------------------------------------------------------------------------------ stg: predicated global stores
// kMode 0: all lanes, 1: predicate false, 2: lane 0 only, 3: two moving lanes.  Every CTA writes its own 64 KB region (L2 resident).
template <int kMode>
__global__ void __launch_bounds__(1024, 1) stg_kernel(int iters, int zero, uint8_t* buf, long long* cycles) {
    const int lane = threadIdx.x & 31;
    uint8_t* addr = buf + static_cast<size_t>(blockIdx.x) * 65536 + threadIdx.x * 16;
    uint32_t a = threadIdx.x, b = a * 3, c = a * 5, d = a * 7;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        const int on = kMode == 0 ? 1 : kMode == 1 ? zero : kMode == 2 ? ((lane == 0) | zero) : ((lane == (i & 31)) | (lane == ((i * 7 + 3) & 31)) | zero);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.global.v4.b32 [%0], {%1,%2,%3,%4};\n\t}"
                         :: "l"(addr + u * 16384), "r"(a), "r"(b), "r"(c), "r"(d), "r"(on) : "memory");
        a += i;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int kMode>
static void run_stg(int sms, int nw, uint8_t* buf, long long* d_cycles, const char* name) {
    const int iters = 2000;
    stg_kernel<kMode><<<sms, nw * 32>>>(iters, 0, buf, d_cycles);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < sms; ++i) s += h[i];
    s /= sms;
    printf("stg  warps/SM %2d  %-16s: %.2f clk per STG.128 per SM, %.1f clk per warp per store\n", nw, name, s / (iters * 4.0 * nw), s / (iters * 4.0));
}


// ------------------------------------------------------------------------------------------------ ld16: register layout of tcgen05.ld.16x256b.x2
// TMEM is filled with value = lane * 1000 + column (tcgen05.st.32x32b); every warp then reads columns [32, 48) of its lane
// quarter with two 16x256b.x2 loads (lanes +0 and +16) and dumps the eight registers of every thread.
__global__ void __launch_bounds__(128, 1) ld16_kernel(uint32_t* out) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_ptr)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    for (int b = 0; b < 2; ++b) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = threadIdx.x * 1000 + b * 32 + j;
        const uint32_t addr = tmem + b * 32 + lane_addr;
        TMEM_ST32(addr, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int h = 0; h < 2; ++h) {
        uint32_t r[8];
        const uint32_t addr = tmem + 32 + (static_cast<uint32_t>(warp * 32 + 16 * h) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) out[(threadIdx.x * 2 + h) * 8 + j] = r[j];
    }
    (void)lane;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64) : "memory");
}

// ------------------------------------------------------------------------------------------------ sts3: quad-shaped predicates, vote / redux cost
// kMode 0: lanes 0-3 store; 1: lanes 0-3 and 20-23; 2: lanes 0-7; 3: no store, one vote.ballot + one redux.or per iteration
template <int kMode>
__global__ void __launch_bounds__(1024, 1) sts3_kernel(int iters, int zero, long long* cycles, uint32_t* sink) {
    extern __shared__ uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const uint32_t addr = smem_u32(smem) + threadIdx.x * 16;
    uint32_t a = threadIdx.x, b = a * 3, c = a * 5, d = a * 7, acc = 0;
    const int on = kMode == 0 ? ((lane < 4) | zero) : kMode == 1 ? ((lane < 4) | (lane >= 20 && lane < 24) | zero) : kMode == 2 ? ((lane < 8) | zero) : zero;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (kMode == 3) {
                const unsigned bal = __ballot_sync(0xffffffff, (a + u) & 1);
                unsigned red;
                asm volatile("redux.sync.or.b32 %0, %1, 0xffffffff;" : "=r"(red) : "r"(a ^ bal));
                acc += red; a += red & 3;
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n\t}"
                             :: "r"(addr + (u & 3) * 16384), "r"(a), "r"(b), "r"(c), "r"(d), "r"(on) : "memory");
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}
template <int kMode>
static void run_sts3(int sms, int nw, long long* d_cycles, uint32_t* d_sink, const char* name) {
    const int iters = 2000;
    cudaFuncSetAttribute(sts3_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 4 * 16384);
    sts3_kernel<kMode><<<sms, nw * 32, 65536 + 4 * 16384>>>(iters, 0, d_cycles, d_sink);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < sms; ++i) s += h[i];
    s /= sms;
    printf("sts3 warps/SM %2d  %-34s: %.2f clk per instruction per SM, %.1f clk per warp\n", nw, name, s / (iters * 8.0 * nw), s / (iters * 8.0));
}

// ================================================================================================ host
static double avg_cycles(long long* d_cycles, int n) {
    std::vector<long long> h(n);
    CK(cudaMemcpy(h.data(), d_cycles, n * sizeof(long long), cudaMemcpyDeviceToHost));
    double s = 0; for (int i = 0; i < n; ++i) s += h[i];
    return s / n;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "all";
    const bool all = !strcmp(what, "all");
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* d_cycles; CK(cudaMalloc(&d_cycles, 1024 * sizeof(long long)));
    uint32_t* d_sink; CK(cudaMalloc(&d_sink, 64));
    printf("SMs %d\n", sms);

    if (all || !strcmp(what, "sts")) {
        CK(cudaFuncSetAttribute(sts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 4 * 16384));
        const char* names[] = {"all lanes", "pred false", "lane 0", "lanes 0-7", "uniform branch, not taken", "2 moving lanes"};
        for (int nw : {4, 8, 16}) for (int mode = 0; mode < 6; ++mode) {
            const int iters = 4000;
            sts_kernel<<<sms, nw * 32, 65536 + 4 * 16384>>>(mode, iters, 0, d_cycles);
            CK(cudaDeviceSynchronize());
            const double c = avg_cycles(d_cycles, sms);
            printf("sts  warps/SM %2d  %-28s: %.2f clk per STS.128 per SM (%.1f clk per iteration of 4 per warp)\n", nw, names[mode],
                   c / (iters * 4.0 * nw), c / iters);
        }
    }
    if (all || !strcmp(what, "alu")) {
        for (int nw : {4, 8, 16}) {
            const int iters = 4000;
            for (int mode = 0; mode < 3; ++mode) {
                alu_kernel<<<sms, nw * 32>>>(mode, iters, 0, d_cycles, d_sink);
                CK(cudaDeviceSynchronize());
                const double c = avg_cycles(d_cycles, sms);
                const int per_iter = mode == 2 ? 8 : 16;
                printf("alu  warps/SM %2d  mode %d (%s): %.2f clk per warp-instruction per SMSP\n", nw, mode,
                       mode == 0 ? "2 x VIMNMX x 8 chains" : mode == 1 ? "same + vote.any + branch" : "IMAD x 8 chains",
                       c / (iters * per_iter * (nw / 4.0)));
            }
            alu3_kernel<<<sms, nw * 32>>>(iters, d_cycles, d_sink);
            CK(cudaDeviceSynchronize());
            printf("alu  warps/SM %2d  VIMNMX3 x 8 chains: %.2f clk per warp-instruction per SMSP\n", nw, avg_cycles(d_cycles, sms) / (iters * 8 * (nw / 4.0)));
        }
    }
    if (all || !strcmp(what, "mma")) {
        CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
        struct V { int ts, n, nacc, per; const char* name; };
        const V vs[] = {{0, 256, 2, 4, "SS N=256 K=4x32"}, {0, 256, 2, 5, "SS N=256 K=5x32"}, {0, 128, 3, 5, "SS N=128 K=5x32"},
                        {1, 256, 1, 5, "TS N=256 K=5x32"}, {1, 128, 3, 5, "TS N=128 K=5x32"}, {1, 128, 3, 4, "TS N=128 K=4x32"},
                        {1, 64, 3, 5, "TS N=64 K=5x32"}, {0, 64, 3, 5, "SS N=64 K=5x32"}};
        for (int grid : {1, sms}) for (const V& v : vs) {
            const int groups = 4000;
            mma_rate_kernel<<<grid, 128, 16384 + 32768 + 1024>>>(v.ts, v.n, v.nacc, groups, v.per, d_cycles);
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            mma_rate_kernel<<<grid, 128, 16384 + 32768 + 1024>>>(v.ts, v.n, v.nacc, groups, v.per, d_cycles);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double c = avg_cycles(d_cycles, grid);
            const double ops = 2.0 * 128 * v.n * 32 * v.per * groups * grid;
            printf("mma  grid %3d  %-16s: %.1f clk per MMA, %.1f clk per group; kernel %.3f ms = %.0f TOP/s\n", grid, v.name,
                   c / (groups * v.per), c / groups, ms, ops / (ms * 1e-3) / 1e12);
        }
    }

    if (all || !strcmp(what, "mma2")) {
        CK(cudaFuncSetAttribute(mma_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
        struct V { int ts, n, nacc, pattern; const char* name; };
        const V vs[] = {{0, 256, 2, 0, "SS N=256 2 acc grouped"}, {0, 256, 2, 1, "SS N=256 2 acc interleaved"}, {0, 256, 2, 2, "SS N=256 independent"},
                        {1, 256, 1, 0, "TS N=256 1 acc grouped"}, {1, 256, 1, 2, "TS N=256 independent"},
                        {0, 128, 4, 0, "SS N=128 4 acc grouped"}, {0, 128, 2, 1, "SS N=128 2 acc interleaved"}, {0, 128, 4, 1, "SS N=128 4 acc interleaved"}, {0, 128, 4, 2, "SS N=128 independent"},
                        {1, 128, 3, 0, "TS N=128 3 acc grouped"}, {1, 128, 2, 1, "TS N=128 2 acc interleaved"}, {1, 128, 3, 1, "TS N=128 3 acc interleaved"}, {1, 128, 3, 2, "TS N=128 independent"},
                        {1, 64, 4, 1, "TS N=64 4 acc interleaved"}, {1, 64, 4, 2, "TS N=64 independent"}};
        for (const V& v : vs) {
            const int rounds = 2000, per = 5;
            mma_pattern_kernel<<<sms, 128, 16384 + 32768 + 1024>>>(v.ts, v.n, v.nacc, v.pattern, rounds, per, d_cycles);
            CK(cudaDeviceSynchronize());
            const double c = avg_cycles(d_cycles, sms);
            const double n_mma = double(rounds) * per * v.nacc;
            printf("mma2 %-28s: %.1f clk per MMA (%.0f MAC/clk/SM)\n", v.name, c / n_mma, 128.0 * v.n * 32 * n_mma / c);
        }
    }
    if (all || !strcmp(what, "sts2")) {
        for (int nw : {1, 4, 16}) {
            run_sts2<16, 3>(sms, nw, d_cycles, "no store");
            run_sts2<16, 0>(sms, nw, d_cycles, "all lanes");
            run_sts2<16, 1>(sms, nw, d_cycles, "pred false");
            run_sts2<16, 2>(sms, nw, d_cycles, "lane 0");
            run_sts2<8, 0>(sms, nw, d_cycles, "all lanes");
            run_sts2<8, 2>(sms, nw, d_cycles, "lane 0");
            run_sts2<4, 0>(sms, nw, d_cycles, "all lanes");
            run_sts2<4, 2>(sms, nw, d_cycles, "lane 0");
        }
    }

    if (all || !strcmp(what, "stg")) {
        uint8_t* buf; CK(cudaMalloc(&buf, static_cast<size_t>(sms) * 65536));
        for (int nw : {8, 16}) {
            run_stg<0>(sms, nw, buf, d_cycles, "all lanes");
            run_stg<1>(sms, nw, buf, d_cycles, "pred false");
            run_stg<2>(sms, nw, buf, d_cycles, "lane 0");
            run_stg<3>(sms, nw, buf, d_cycles, "2 moving lanes");
        }
    }

    if (all || !strcmp(what, "ld16")) {
        uint32_t* d_out; CK(cudaMalloc(&d_out, 128 * 2 * 8 * 4));
        ld16_kernel<<<1, 128>>>(d_out);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> o(128 * 16);
        CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
        for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 33, 100}) for (int h = 0; h < 2; ++h) {
            printf("ld16 thread %3d half %d:", t, h);
            for (int j = 0; j < 8; ++j) printf("  (lane %3u col %2u)", o[(t * 2 + h) * 8 + j] / 1000, o[(t * 2 + h) * 8 + j] % 1000);
            printf("\n");
        }
    }
    if (all || !strcmp(what, "sts3")) {
        for (int nw : {8, 16}) {
            run_sts3<0>(sms, nw, d_cycles, d_sink, "STS.128 lanes 0-3");
            run_sts3<1>(sms, nw, d_cycles, d_sink, "STS.128 lanes 0-3 + 20-23");
            run_sts3<2>(sms, nw, d_cycles, d_sink, "STS.128 lanes 0-7");
            run_sts3<3>(sms, nw, d_cycles, d_sink, "vote.ballot + redux.or (pair)");
        }
    }
    if (all || !strcmp(what, "layout")) {
        CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
        std::vector<uint8_t> A(128 * 128), B(128 * 128);
        srand(1);
        for (auto& x : A) x = rand() & 255;
        for (auto& x : B) x = rand() & 255;
        uint8_t *dA, *dB; int32_t* dD; uint32_t* dDump;
        CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&dDump, 128 * 32 * 4));
        CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
        std::vector<int32_t> ref(128 * 128);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
            int s = 0; for (int k = 0; k < 128; ++k) s += int(A[m * 128 + k]) * int(B[n * 128 + k]);
            ref[m * 128 + n] = s;
        }
        const char* names[] = {"tcgen05.st rows", "tcgen05.cp no-swizzle panels", "tcgen05.cp SWIZZLE_128B tile", "SS reference"};
        for (int how = 0; how < 4; ++how) {
            CK(cudaMemset(dD, 0xff, 128 * 128 * 4)); CK(cudaMemset(dDump, 0, 128 * 32 * 4));
            layout_kernel<<<1, 128, 49152 + 1024>>>(how, dA, dB, dD, dDump);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("layout how %d (%s): CUDA error %s\n", how, names[how], cudaGetErrorString(e)); return 1; }
            std::vector<int32_t> D(128 * 128); std::vector<uint32_t> dump(128 * 32);
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0; for (size_t i = 0; i < D.size(); ++i) bad += D[i] != ref[i];
            int dump_bad = 0;
            for (int m = 0; m < 128; ++m) for (int j = 0; j < 32; ++j) dump_bad += dump[m * 32 + j] != reinterpret_cast<const uint32_t*>(A.data())[m * 32 + j];
            printf("layout %-32s: %d of %zu outputs differ; A image in tensor memory differs from row-major words in %d of 4096\n", names[how], bad,
                   D.size(), dump_bad);
            if (dump_bad && how != 3) {
                printf("   row 0 words 0-7 in TMEM: "); for (int j = 0; j < 8; ++j) printf("%08x ", dump[j]);
                printf("\n   row 0 words 0-7 of A   : "); for (int j = 0; j < 8; ++j) printf("%08x ", reinterpret_cast<const uint32_t*>(A.data())[j]);
                printf("\n   row 1 words 0-7 in TMEM: "); for (int j = 0; j < 8; ++j) printf("%08x ", dump[32 + j]);
                printf("\n   row 1 words 0-7 of A   : "); for (int j = 0; j < 8; ++j) printf("%08x ", reinterpret_cast<const uint32_t*>(A.data())[32 + j]);
                printf("\n");
            }
        }
    }
    return 0;
}
