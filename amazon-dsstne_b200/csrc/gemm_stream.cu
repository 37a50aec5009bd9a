// gemm_stream.cu -- tcgen05 / TMEM / TMA kernels for the OUTPUT layer of a sparse-target network (hot-path rows a11 + a7-a9),
// sm_100a only.  Second generation of the dense path: gemm_tc.cu stays the general kernel (any shape), this file holds the
// kernels for the shape that dominates the training step -- one dimension is the hidden width (<= a few hundred), the
// other two are the batch and the 10^4..10^6 output units:
//
//   fwd  (E/NNLayer.cpp:1073 + kActivation.cu:46-64 + kLoss.cu:595-2352 + kDelta.cu:2193-7305)
//        delta[b][n] = f(sigmoid(X[b][:] . W[:][n] + bias[n]), target[b][n])        out_fwd_kernel      (Z and A never exist)
//   dW   (E/NNLayer.cpp:2223)   G[k][n] = alpha * sum_b X[b][k] * delta[b][n]          stream_kernel<A_MN, EPI_T>
//   dX   (E/NNLayer.cpp:2274)   Dp[b][k] = sum_n delta[b][n] * W[k][n]                 stream_kernel<A_K,  EPI_N>  (split-K)
//
// What round 1's ncu captures said (profiles/r1c_gemm_smem_pipe.md): the 3xTF32 kernels are bound by bytes through the
// L1 / shared-memory data array (st.shared at half rate, operand re-reads by the tensor core), not by HBM or the tensor
// pipe.  The design rule here follows from that:
//   * the BIG operand (delta, or W in the forward pass) is the A operand and never touches shared memory:
//     global -> registers (coalesced) -> hi / lo split -> tcgen05.st -> TENSOR MEMORY -> tcgen05.mma [d], [a_tmem], b_desc;
//   * the SMALL operand (X, X^T or W: <= 14 MB, L2 resident) is the B operand and is moved by the TMA engine only:
//     cp.async.bulk.tensor.2d with a 128-byte-swizzle tensor map drops 128 x 32 panels straight into the UMMA K-major
//     layout -- no loader warps, no st.shared, no proxy fence.  Its "lo" half (x - trunc_tf32(x), 3xTF32) is a second
//     array written once per call by a streaming kernel, in a pitch-padded copy, so the 8-byte-aligned rows of the
//     27,278-wide weight matrix are no obstacle (tensor-map pitches must be multiples of 16 bytes);
//   * "swapped" orientation: the output unit index is the M dimension (TMEM lane = thread), so every global store of the
//     epilogue is a fully coalesced 128-byte row segment straight from the tcgen05.ld registers -- no staging tile;
//   * forward pass: the W slice of a work unit (128 outputs x K <= 128) is split ONCE and stays resident in tensor memory
//     (256 columns) for all batch tiles of the unit; eight epilogue warps finish the element (bias, sigmoid, loss, delta,
//     column sums of delta for the bias gradient) because the epilogue, not the MMA, is the long pole of that kernel.
#include <cuda.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "launch.h"

namespace dsb {
namespace gs {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int PANEL = BN * BK * 4;             // one 128 x 32 fp32 operand panel: 16 KB
constexpr int SLOT = 2 * PANEL;                // hi panel | lo panel
constexpr uint32_t TMEM_COLS = 512;

// ---------------------------------------------------------------- tcgen05 / TMA wrappers
__device__ __forceinline__ void tmem_alloc(uint32_t* smemResult, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smemResult)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], tf32 inputs, fp32 accumulation
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" :: "r"(tmemD), "r"(tmemA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor (version 1), K-major operand in the standard 128-byte swizzle: rows of 32 floats, groups of
// 8 rows 1,024 bytes apart; one K = 8 step = 32 bytes along the row.  Exactly what a SWIZZLE_128B tensor map writes.
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t panelAddr, int kStep)
{
    const uint32_t saddr = panelAddr + kStep * 32;
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, A from tensor memory, B K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 16 consecutive columns: thread t writes TMEM lane (quadrant base + t)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// 16 lanes x 256 bits, twice (columns c..c+7 and c+8..c+15): register s of thread t -> lane t / 4 + 8 * ((s >> 1) & 1),
// column 2 * (t % 4) + 8 * (s >> 2) + (s & 1)   (probed on the B200: tools/umma_st16x256_probe.cu)
__device__ __forceinline__ void tmem_st16x256_x2(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one box of a 2-D tensor map -> shared memory, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(uint32_t smemDst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smemDst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

__device__ __forceinline__ uint32_t ldg_nc_u32(const float* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_nc_u32x2(const float* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t hi_of(uint32_t x) { return x & 0xFFFFE000u; }                 // the 19 bits a tf32 operand keeps
__device__ __forceinline__ uint32_t lo_of(uint32_t x) { return __float_as_uint(__uint_as_float(x) - __uint_as_float(x & 0xFFFFE000u)); }

// =====================================================================================================================
// stream_kernel: D[M][N] = sum_k A(m, k) * B(k, n),  A streamed through registers into a tensor-memory ring, B by TMA.
//   AMODE 0  A(m, k) = A[k * lda + m]   (m contiguous: lanes = consecutive m, 32-bit loads, any alignment)       -- dW
//   AMODE 1  A(m, k) = A[m * lda + k]   (k contiguous: the 16x256b store shape, 4 threads = one 32-byte sector)  -- dX
//   EPI 0    C[n * ldc + m] = alpha * D + beta * C     (transposed store, coalesced over the lanes)
//   EPI 1    C[m * ldc + n] = alpha * D + beta * C, or the raw partial tile [split][m][n] when K is split
// B is given as two tensor maps over K-major arrays Bhi / Blo [N rows][K] (box 32 x 128, 128-byte swizzle, zero fill
// outside), so ragged N and K need no code.
// Warps: 0-7 A loaders (quadrant = warp % 4, k-half = warp / 4), 8-11 epilogue, 12 TMA producer, 13 MMA issuer.
// Tensor memory: 2 x 128 accumulator columns + 4 ring slots of (32 hi | 32 lo) A columns.
// =====================================================================================================================
constexpr int S_LOAD_WARPS = 8, S_EPI_WARP0 = 8, S_EPI_WARPS = 4, S_TMA_WARP = 12, S_MMA_WARP = 13;
constexpr int S_THREADS = 14 * 32;
// The tensor core's fp32 accumulator truncates, so the error of a product grows with the number of MMAs chained into one accumulator
// (measured in round 1: 2e-6 at K <= 128, 1.8e-5 at K = 1,024; an engine test on the exact config-2 network saw 4e-5 on a hidden
// weight after two steps, all of one sign).  Chains are therefore cut every S_CHUNK k-iterations (96 MMAs): the MMA warp switches to
// the other accumulator, and the epilogue warps add the finished chunk into an fp32 tile in SHARED MEMORY (round to nearest; 64 KB,
// layout [column / 4][row][4] so that the 32 rows of a warp are 32 consecutive 16-byte words) while the next chunk runs.
constexpr int S_CHUNK = 8;
constexpr int S_SUM_BYTES = BM * BN * 4;
constexpr int S_SLOTS = 4, S_DEPTH = 3;        // ring depth / register panels in flight per loader thread
constexpr uint32_t S_ACC_COLS = 2 * BN, S_A_COLS = 2 * BK;
constexpr int S_SMEM_BYTES = S_SLOTS * SLOT + S_SUM_BYTES + 1024;
static_assert(S_ACC_COLS + S_SLOTS * S_A_COLS <= TMEM_COLS, "tensor memory budget");

struct SArgs {
    const float* A; uint32_t lda;
    uint32_t M, N, K;
    uint32_t tilesM, tilesN, splits, kPerSplit;    // kPerSplit: multiple of BK
    float* C; uint32_t ldc;
    float* partial;                                // [splits][M][N] when splits > 1
    float alpha, beta;
    int passes;                                    // 3 = 3xTF32, 1 = TF32
    int debug;                                     // option "gemm_debug" (bring-up, wrong results): 512 = no global loads of A, 1024 = no MMAs
};

struct STile { uint32_t m0, n0, split, kBegin, kEnd, numK; };
__device__ __forceinline__ STile stile_of(const SArgs& a, uint32_t t)
{
    STile x;
    const uint32_t mn = a.tilesM * a.tilesN;
    x.split = t / mn;
    const uint32_t r = t - x.split * mn;
    x.n0 = (r / a.tilesM) * BN;
    x.m0 = (r % a.tilesM) * BM;
    x.kBegin = x.split * a.kPerSplit;
    x.kEnd = min(a.K, x.kBegin + a.kPerSplit);
    x.numK = (x.kEnd - x.kBegin + BK - 1) / BK;
    return x;
}

// A loader plans: 16 values of one thread per 128 x 32 panel
struct PlanMN {                                    // thread = row m, 16 consecutive k, stride lda
    const float* ptr; size_t lda; uint32_t kOff; bool rowIn;
    __device__ __forceinline__ void init(const float* A, uint32_t ld, uint32_t m0, uint32_t kBegin, uint32_t M, uint32_t q, uint32_t khalf, uint32_t lane)
    {
        const uint32_t m = m0 + q * 32 + lane;
        rowIn = m < M; kOff = khalf * 16; lda = ld;
        ptr = A + (size_t)(kBegin + kOff) * ld + m;
    }
    __device__ __forceinline__ void load(uint32_t (&r)[16], uint32_t k0, uint32_t kEnd)
    {
        const uint32_t kb = k0 + kOff;
        const uint32_t nv = (rowIn && kb < kEnd) ? min(16u, kEnd - kb) : 0u;
        if (nv == 16) {
#pragma unroll
            for (int e = 0; e < 16; e++) r[e] = ldg_nc_u32(ptr + e * lda);
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) r[e] = (uint32_t)e < nv ? ldg_nc_u32(ptr + e * lda) : 0u;
        }
        ptr += (size_t)BK * lda;
    }
    __device__ __forceinline__ void store(uint32_t ta, const uint32_t (&r)[16], bool lo) const
    {
        uint32_t t[16];
#pragma unroll
        for (int e = 0; e < 16; e++) t[e] = hi_of(r[e]);
        tmem_st16(ta, t);
        if (lo) {
#pragma unroll
            for (int e = 0; e < 16; e++) t[e] = lo_of(r[e]);
            tmem_st16(ta + BK, t);
        }
    }
};
struct PlanK {                                     // r[8 j + s]: row t / 4 + 8 * (2 j + ((s >> 1) & 1)), k = 16 khalf + 2 (t % 4) + 8 (s >> 2) + (s & 1)
    const float* ptr; size_t ld8; uint32_t kOff, rowMask; bool pair;
    __device__ __forceinline__ void init(const float* A, uint32_t ld, uint32_t m0, uint32_t kBegin, uint32_t M, uint32_t q, uint32_t khalf, uint32_t lane)
    {
        const uint32_t m = m0 + q * 32 + (lane >> 2);
        kOff = khalf * 16 + 2 * (lane & 3);
        ld8 = (size_t)8 * ld;
        rowMask = 0;
#pragma unroll
        for (uint32_t g = 0; g < 4; g++) if (m + 8 * g < M) rowMask |= 1u << g;
        ptr = A + (size_t)m * ld + kBegin + kOff;
        pair = ((ld & 1u) == 0) && ((((uintptr_t)A) & 7u) == 0) && ((kBegin & 1u) == 0);     // 64-bit loads stay aligned
    }
    __device__ __forceinline__ void load(uint32_t (&r)[16], uint32_t k0, uint32_t kEnd)
    {
        if (rowMask == 15u && k0 + BK <= kEnd && pair) {
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int s = 0; s < 8; s += 2) {
                    const uint2 t = ldg_nc_u32x2(ptr + (2 * j + ((s >> 1) & 1)) * ld8 + 8 * (s >> 2));
                    r[8 * j + s] = t.x; r[8 * j + s + 1] = t.y;
                }
        } else if (rowMask == 15u && k0 + BK <= kEnd) {                          // rows 4-byte aligned only (odd pitch: a column shard)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int s = 0; s < 8; s += 2) {
                    const float* p = ptr + (2 * j + ((s >> 1) & 1)) * ld8 + 8 * (s >> 2);
                    r[8 * j + s] = ldg_nc_u32(p); r[8 * j + s + 1] = ldg_nc_u32(p + 1);
                }
        } else {
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int s = 0; s < 8; s += 2) {
                    const uint32_t g = 2 * j + ((s >> 1) & 1), kk = k0 + kOff + 8 * (s >> 2);
                    const float* p = ptr + g * ld8 + 8 * (s >> 2);
                    const uint32_t n = (((rowMask >> g) & 1u) && kk < kEnd) ? min(2u, kEnd - kk) : 0u;
                    r[8 * j + s] = n > 0 ? ldg_nc_u32(p) : 0u;
                    r[8 * j + s + 1] = n > 1 ? ldg_nc_u32(p + 1) : 0u;
                }
        }
        ptr += BK;
    }
    __device__ __forceinline__ void store(uint32_t ta, const uint32_t (&r)[16], bool lo) const
    {
        uint32_t t[16];
#pragma unroll
        for (int e = 0; e < 16; e++) t[e] = hi_of(r[e]);
        tmem_st16x256_x2(ta, t); tmem_st16x256_x2(ta + (16u << 16), t + 8);
        if (lo) {
#pragma unroll
            for (int e = 0; e < 16; e++) t[e] = lo_of(r[e]);
            tmem_st16x256_x2(ta + BK, t); tmem_st16x256_x2(ta + BK + (16u << 16), t + 8);
        }
    }
};

template <int AMODE, int EPI>
__global__ void __launch_bounds__(S_THREADS, 1)
stream_kernel(const SArgs a, const __grid_constant__ CUtensorMap mapHi, const __grid_constant__ CUtensorMap mapLo)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t aFull[S_SLOTS], bFull[S_SLOTS], slotEmpty[S_SLOTS], accFull[2], accEmpty[2];
    __shared__ uint32_t tmemBase;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t numTiles = a.tilesM * a.tilesN * a.splits;
    const bool lo = a.passes == 3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S_SLOTS; s++) { mbar_init(&aFull[s], S_LOAD_WARPS); mbar_init(&bFull[s], 1); mbar_init(&slotEmpty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&accFull[s], 1); mbar_init(&accEmpty[s], S_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == S_TMA_WARP && lane == 0) { tma_prefetch_desc(&mapHi); tma_prefetch_desc(&mapLo); }
    if (warp == S_MMA_WARP) tmem_alloc(&tmemBase, TMEM_COLS);
    pdl_launch_dependents();
    pdl_wait();                                                                   // barriers, tensor memory and descriptors are set up under the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;
    const uint32_t ringAddr = smem_u32(smem);

    if (warp < S_LOAD_WARPS) {
        // ------------------------------------------------------------ A loaders: global -> registers -> tensor memory
        const uint32_t q = warp & 3, khalf = warp >> 2;
        typename std::conditional<AMODE == 0, PlanMN, PlanK>::type pa;
        uint32_t lt = blockIdx.x, lkt = 0, slot = 0, parity = 1;
        STile ltl = stile_of(a, min(lt, numTiles - 1));
        bool more = lt < numTiles;
        pa.init(a.A, a.lda, ltl.m0, ltl.kBegin, a.M, q, khalf, lane);
        auto fetch = [&](uint32_t (&r)[16]) {
            pa.load(r, ltl.kBegin + lkt * BK, (a.debug & 512) ? ltl.kBegin : ltl.kEnd);   // bring-up switch 512: no global loads (zero panels)
            if (++lkt == ltl.numK) {
                lkt = 0; lt += gridDim.x; more = lt < numTiles;
                if (more) { ltl = stile_of(a, lt); pa.init(a.A, a.lda, ltl.m0, ltl.kBegin, a.M, q, khalf, lane); }
            }
        };
        auto publish = [&](const uint32_t (&r)[16]) {
            mbar_wait(&slotEmpty[slot], parity);                                  // the MMAs that read this slot have retired
            tc_fence_after();
            pa.store(tmem + ((q * 32) << 16) + S_ACC_COLS + slot * S_A_COLS + khalf * 16, r, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&aFull[slot]);
            if (++slot == S_SLOTS) { slot = 0; parity ^= 1; }
        };
        // loads run S_DEPTH - 1 panels ahead of the tensor-memory stores
        uint32_t r[S_DEPTH][16];
        uint32_t pending = 0;
#pragma unroll
        for (int s = 0; s < S_DEPTH - 1; s++)
            if (more) { fetch(r[s]); pending++; }
        while (pending) {
#pragma unroll
            for (int s = 0; s < S_DEPTH; s++) {
                if (more) { fetch(r[(s + S_DEPTH - 1) % S_DEPTH]); pending++; }
                publish(r[s]);
                if (--pending == 0) break;
            }
        }
    } else if (warp == S_TMA_WARP) {
        // ------------------------------------------------------------ B producer: one thread drives the TMA engine.
        // (Measured and dropped, profiles/r2_gemm_stream.md: pulling the streamed A operand into L2 ahead of the loaders with
        // cp.async.bulk.prefetch.L2 doubled the kernel time -- 32 small bulk operations per k-iteration queue in front of the
        // B boxes in the TMA unit -- and per-line prefetch.global.L2 changed nothing: the loaders are not DRAM-latency bound.)
        if (lane == 0) {
            uint32_t slot = 0, parity = 1;
            for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x) {
                const STile tl = stile_of(a, t);
                for (uint32_t kt = 0; kt < tl.numK; kt++) {
                    mbar_wait(&slotEmpty[slot], parity);
                    const uint32_t st = ringAddr + slot * SLOT, k0 = tl.kBegin + kt * BK;
                    mbar_arrive_expect_tx(&bFull[slot], lo ? SLOT : PANEL);
                    tma_load_2d(st, &mapHi, k0, tl.n0, &bFull[slot]);
                    if (lo) tma_load_2d(st + PANEL, &mapLo, k0, tl.n0, &bFull[slot]);
                    if (++slot == S_SLOTS) { slot = 0; parity ^= 1; }
                }
            }
        }
    } else if (warp == S_MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t slot = 0, ph = 0, seq = 0;
            for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x) {
                const STile tl = stile_of(a, t);
                for (uint32_t kt = 0; kt < tl.numK; kt++) {
                    const uint32_t acc = seq & 1, d = tmem + acc * BN;
                    if (kt % S_CHUNK == 0) {
                        mbar_wait(&accEmpty[acc], ((seq >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
                        tc_fence_after();
                    }
                    mbar_wait(&aFull[slot], ph);
                    mbar_wait(&bFull[slot], ph);
                    tc_fence_after();
                    const uint32_t sb = ringAddr + slot * SLOT, ta = tmem + S_ACC_COLS + slot * S_A_COLS;
#pragma unroll
                    for (int j = 0; j < BK / 8; j++) {
                        const uint64_t bHi = kmajor_desc(sb, j);
                        const uint32_t aHi = ta + j * 8, first = (kt % S_CHUNK == 0 && j == 0) ? 0u : 1u;
                        if (a.debug & 1024) {                                     // bring-up switch 1024: no MMAs
                        } else if (lo) {
                            const uint64_t bLo = kmajor_desc(sb + PANEL, j);
                            mma_tf32_ts(d, aHi + BK, bHi, kIdesc, first);         // small terms first
                            mma_tf32_ts(d, aHi, bLo, kIdesc, 1u);
                            mma_tf32_ts(d, aHi, bHi, kIdesc, 1u);
                        } else {
                            mma_tf32_ts(d, aHi, bHi, kIdesc, first);
                        }
                    }
                    tc_commit(&slotEmpty[slot]);                                  // shared and tensor memory of the slot reusable once these retire
                    if (++slot == S_SLOTS) { slot = 0; ph ^= 1; }
                    if ((kt + 1) % S_CHUNK == 0 || kt + 1 == tl.numK) { tc_commit(&accFull[acc]); seq++; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: thread = accumulator row m
        const uint32_t q = warp & 3, row = q * 32 + lane;
        const uint32_t taddr = tmem + ((q * 32) << 16);
        const uint32_t sumAddr = ringAddr + S_SLOTS * SLOT + row * 16;            // + (column / 4) * (BM * 16)
        uint32_t seq = 0;
        for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x) {
            const STile tl = stile_of(a, t);
            const uint32_t chunks = (tl.numK + S_CHUNK - 1) / S_CHUNK;
            const uint32_t m = tl.m0 + row;
            for (uint32_t ch = 0; ch < chunks; ch++, seq++) {
                const uint32_t acc = seq & 1;
                const bool firstCh = ch == 0, lastCh = ch + 1 == chunks;
                __syncwarp();
                mbar_wait(&accFull[acc], (seq >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < BN / 16; c++) {
                    float v[16];
                    __syncwarp();                                                 // lanes diverge in the stores below (ragged edges)
                    tmem_ld16(taddr + acc * BN + c * 16, v);
                    if (c == BN / 16 - 1) {                                       // accumulator fully read: the MMA warp may reuse it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&accEmpty[acc]);
                    }
                    if (!firstCh) {                                               // running sum of the earlier chunks (only this thread touches its row)
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            float4 p;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w) : "r"(sumAddr + (c * 4 + g) * (BM * 16)));
                            v[4 * g] += p.x; v[4 * g + 1] += p.y; v[4 * g + 2] += p.z; v[4 * g + 3] += p.w;
                        }
                    }
                    if (!lastCh) {
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(sumAddr + (c * 4 + g) * (BM * 16)), "f"(v[4 * g]), "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3]) : "memory");
                        continue;
                    }
                    const uint32_t nb = tl.n0 + c * 16;
                    if (m >= a.M || nb >= a.N) continue;
                    const uint32_t ncol = min(16u, a.N - nb);
                    if (EPI == 0) {
                        float* o = a.C + (size_t)nb * a.ldc + m;
                        if (a.beta != 0.0f) {
                            // all 16 loads first: written as o[j] = f(o[j]) the compiler must assume that the rows alias and runs 16
                            // dependent round trips to L2 (measured: 121 us instead of 80 for the config-2 weight gradient with beta = 1)
                            float old[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) old[j] = (uint32_t)j < ncol ? __ldcg(o + (size_t)j * a.ldc) : 0.0f;
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                if ((uint32_t)j < ncol) o[(size_t)j * a.ldc] = fmaf(a.beta, old[j], a.alpha * v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                if ((uint32_t)j < ncol) o[(size_t)j * a.ldc] = a.alpha * v[j];
                        }
                    } else if (a.partial) {
                        float* o = a.partial + ((size_t)tl.split * a.M + m) * a.N + nb;
                        if (ncol == 16 && (a.N & 3u) == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j++) if ((uint32_t)j < ncol) o[j] = v[j];
                        }
                    } else {
                        float* o = a.C + (size_t)m * a.ldc + nb;
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            if ((uint32_t)j < ncol) o[j] = a.alpha * v[j] + (a.beta != 0.0f ? a.beta * o[j] : 0.0f);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == S_MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// C = alpha * sum_z partial[z] + beta * C, optionally times f'(unit) (kCalculateHadamardProduct, E/kDelta.cu:8979-9032: the delta of
// the layer below leaves the input-delta GEMM finished); fixed summation order -> deterministic
struct HadArgs { const float* unit; int act; float scale, slope, alpha, lambda; };
__device__ __forceinline__ float gs_hadamard(const HadArgs& h, float x, float d)
{
    switch (h.act) {
    case DSB200_ACT_SIGMOID: return x * (1.0f - x) * d;
    case DSB200_ACT_TANH:    { const float xs = x * (1.0f / h.scale); return h.scale * (1.0f - xs * xs) * d; }
    case DSB200_ACT_RELU:    return (x <= 0.0f) ? 0.0f : d;
    case DSB200_ACT_LRELU:   return (x <= 0.0f) ? d * h.slope : d;
    case DSB200_ACT_ELU:     return (x <= 0.0f) ? d * (x + h.alpha) : d;
    case DSB200_ACT_SELU:    return (x > 0.0f) ? d * h.lambda : d * (x + h.lambda * h.alpha);
    default:                 return d;
    }
}
__global__ void __launch_bounds__(256)
stream_reduce_kernel(const float* __restrict__ partial, uint32_t splits, uint32_t M, uint32_t N, uint32_t ldc, float alpha, float beta, float* __restrict__ C,
                     const HadArgs h)
{
    const size_t total = (size_t)M * N;
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t m = (uint32_t)(i / N), n = (uint32_t)(i % N);
        float s = 0.f;
        for (uint32_t z = 0; z < splits; z++) s += partial[(size_t)z * total + i];
        float* c = C + (size_t)m * ldc + n;
        float x = alpha * s;
        if (beta != 0.0f) x += beta * *c;
        if (h.unit) x = gs_hadamard(h, __ldg(h.unit + (size_t)m * ldc + n), x);
        *c = x;
    }
}

// ---------------------------------------------------------------- operand preparation (streaming, once per call)
// lo[r][c] = x - trunc_tf32(x) (and optionally hi[r][c] = x) into a pitch-padded copy; columns [cols, ldp) are zeroed
__global__ void __launch_bounds__(256)
split_pitch_kernel(const float* __restrict__ src, uint32_t ld, uint32_t rows, uint32_t cols, float* __restrict__ hi, float* __restrict__ lo, uint32_t ldp)
{
    const size_t total = (size_t)rows * ldp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / ldp), c = (uint32_t)(i % ldp);
        const float x = c < cols ? __ldg(src + (size_t)r * ld + c) : 0.0f;
        if (hi) hi[i] = x;
        lo[i] = __uint_as_float(lo_of(__float_as_uint(x)));
    }
}
// XT_hi[c][r] = X[r][c], XT_lo[c][r] = lo(X[r][c]); 32 x 32 tiles through shared memory, pitch ldt >= rows
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ X, uint32_t ldx, uint32_t rows, uint32_t cols, float* __restrict__ hi, float* __restrict__ lo, uint32_t ldt)
{
    __shared__ float tile[32][33];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t tilesC = (cols + 31) / 32, tilesR = (ldt + 31) / 32;
    for (uint32_t t = blockIdx.x; t < tilesC * tilesR; t += gridDim.x) {
        const uint32_t r0 = (t / tilesC) * 32, c0 = (t % tilesC) * 32;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t r = r0 + ty + 8 * i, c = c0 + tx;
            tile[ty + 8 * i][tx] = (r < rows && c < cols) ? __ldg(X + (size_t)r * ldx + c) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t c = c0 + ty + 8 * i, r = r0 + tx;
            if (c < cols && r < ldt) {
                const float x = tile[tx][ty + 8 * i];
                hi[(size_t)c * ldt + r] = x;
                lo[(size_t)c * ldt + r] = __uint_as_float(lo_of(__float_as_uint(x)));
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================================
// out_fwd_kernel: forward pass of a sigmoid output layer over Boolean sparse targets, loss and delta in the epilogue.
//   M dimension = output units (TMEM lane = thread = one output column n), N dimension = batch rows, K = hidden units <= 128.
//   A = W^T slice [128 outputs][K]: read once per work unit with coalesced 32-bit loads, split, RESIDENT in tensor memory
//       (columns 0..127 hi, 128..255 lo); accumulators in columns 256..511.
//   B = X batch tile [128 rows][K] (hi = X itself, lo = second array), TMA, ring of 32-k panels.
//   work unit = (output slice, group of batch tiles); units are dealt round robin to the persistent CTAs.
//   Epilogue (16 warps: quadrant = warp % 4, batch quarter of the tile = warp / 4; round-2 ncu with 8: 33 instructions per element,
//   issue slots 55 % busy -- latency bound with two warps per scheduler): z = acc + bias[n] -> a = sigmoid(z) -> loss and delta, the
//   target taken from the TRANSPOSED bitmap (one word = 32 batch rows of output n); delta[b][n] stored as 128-byte rows (32 lanes =
//   32 consecutive n).  Column sums of delta (the bias gradient, E/NNWeight.cpp:760-794) accumulate in the thread that owns the
//   column and leave as one partial per (group, quarter).
//   The W^T slice of the NEXT unit is requested right before the wait for the unit's last accumulator and stored right after it, so
//   its 32 registers never coexist with the epilogue's (the 96-register budget of 18 warps).
// Warps: 0-15 workers (A load + epilogue), 16 TMA producer, 17 MMA issuer.
// =====================================================================================================================
// bring-up cycle counters (gemm_debug & 65536): [CTA][16] clock64 sums, read back with dsb200_debug_counters
__device__ unsigned long long g_dbgCounters[256 * 16];
#define DSB_DBG_T0(on) const long long _t0 = (on) ? clock64() : 0
#define DSB_DBG_ADD(on, var) do { if (on) var += clock64() - _t0; } while (0)

constexpr int F_WORKERS = 16, F_TMA_WARP = 16, F_MMA_WARP = 17, F_THREADS = 18 * 32, F_SLOTS = 6;
constexpr int F_PARTS = F_WORKERS / 4;           // worker warps per TMEM lane quadrant = column-sum partials per (unit, column)
constexpr int F_SMEM_BYTES = F_SLOTS * SLOT + 1024;
constexpr uint32_t F_A_LO = 128, F_ACC0 = 256;

struct FArgs {
    const float* W; uint32_t ldw;
    const float* bias;
    uint32_t N, K, batch;
    uint32_t tilesM, tilesB, groups, tilesPerGroup;
    float* delta; float* unit; uint32_t ldd;
    const uint32_t* bitsT; uint32_t wordsB;
    const float* rowW;
    unsigned long long* acc;
    float* colPartials;                            // [groups * F_PARTS][N] or NULL
    int passes;
    int debug;                                     // option "gemm_debug" (bring-up, wrong results): 1024 = no MMAs, 2048 = no element math / stores, 8192 = no stores
    // element coefficients (see out_elem): gate thresholds on p, loss scales and signed delta scales per target value
    float thrZ, thrNz, lZ, lNz, dZ, dNz;
};

// The element, WITHOUT control flow and in the fewest instructions (round-2 probe, tools/fwd_probe.py: with the first branch-free form --
// 33 instructions per element -- the element math alone kept the kernel at 60 us against 40 us for the MMAs alone).
// With s = +1 at a zero target, -1 at a non-zero one, u = s * z, E = e^u, T = 1 + E, R = 1 / T, p = E * R = sigmoid(u):
//   x = sigmoid(z)            = p at a zero target, R at a non-zero one
//   the log the loss needs    = log(1 - x) resp. log(x) = -log(T) either way (the reference clamps its argument at 1e-12: T <= 1e12)
//   the delta's x resp. x - 1 = s * p
//   the gate of SMCE          : x > zeroTarget resp. x < oneTarget  <=>  p > thr,  thr = zeroTarget resp. 1 - oneTarget
// so, with the Raw and NonZero terms of the reference merged (the -wz*log(1-x) of the Raw kernel and the +wd*zeroScale*log(1-x) of the
// NonZero kernel cancel exactly at a non-zero target):
//   SMCE  loss += on * sc * wd * log T,  d = on * s * sc * wd * p        sc = zeroScale resp. oneScale
//   CE    loss += wd * log T,            d = s * boost * wd * p
//   L2    loss += wd / 2 * p^2,          d = s * boost * wd * p * (p * R)                    (x (1 - x) = p R)
// One ex2, one rcp, one lg2 per element whatever the target; the target only selects signs and coefficients (f.thr*, f.l*, f.d*).
// PLAIN = no gate and one loss scale (SMCE with zeroTarget <= 0, oneTarget >= 1, equal scales; CE): the loss is a plain sum of logs,
// scaled once per tile.  FAST: MUFU ex2 / lg2 / rcp in log2 units (the factor ln 2 is applied to the tile's sum).
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <bool ISL2, bool FAST, bool HASW, bool PLAIN, bool WANTX>
__device__ __forceinline__ void out_elem(const FArgs& f, float v, float cb, bool nz, float wd, float& loss, float& x, float& d)
{
    const float t0 = FAST ? fmaf(v, kLog2e, cb) : v + cb;                        // z (FAST: in log2 units, cb = bias * log2 e)
    const float t = fminf(nz ? -t0 : t0, FAST ? 39.8631371f : 27.6310211f);     // T <= 1e12
    const float E = FAST ? ex2_approx(t) : expf(t);
    const float T = 1.0f + E;
    const float R = FAST ? rcp_approx(T) : 1.0f / T;
    const float p = E * R;
    if (WANTX) x = nz ? R : p;
    if (ISL2) {
        const float dsc = nz ? f.dNz : f.dZ;
        loss = fmaf(HASW ? wd * p : p, p, loss);                                 // times 1/2 per tile
        d = (HASW ? dsc * wd : dsc) * p * (p * R);
    } else {
        const float L = FAST ? lg2_approx(T) : logf(T);
        if (PLAIN) {
            const float dsc = nz ? f.dNz : f.dZ;
            loss = HASW ? fmaf(wd, L, loss) : loss + L;                          // times lZ (and ln 2) per tile
            d = (HASW ? dsc * wd : dsc) * p;
        } else {
            const bool on = p > (nz ? f.thrNz : f.thrZ);
            const float lsc = nz ? f.lNz : f.lZ, dsc = nz ? f.dNz : f.dZ;
            loss = fmaf(on ? (HASW ? lsc * wd : lsc) : 0.0f, L, loss);
            d = on ? (HASW ? dsc * wd : dsc) * p : 0.0f;
        }
    }
}

template <bool ISL2, bool FAST, bool HASW, bool PLAIN, bool WANTX>
__global__ void __launch_bounds__(F_THREADS, 1)
out_fwd_kernel(const FArgs f, const __grid_constant__ CUtensorMap mapHi, const __grid_constant__ CUtensorMap mapLo)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bFull[F_SLOTS], bEmpty[F_SLOTS], accFull[2], accEmpty[2], aReady;
    __shared__ uint32_t tmemBase;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t numUnits = f.tilesM * f.groups;
    const uint32_t numK = (f.K + BK - 1) / BK;                                   // <= 4
    const bool lo = f.passes == 3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < F_SLOTS; s++) { mbar_init(&bFull[s], 1); mbar_init(&bEmpty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&accFull[s], 1); mbar_init(&accEmpty[s], F_WORKERS); }
        mbar_init(&aReady, F_WORKERS);
        mbar_fence_init();
    }
    if (warp == F_TMA_WARP && lane == 0) { tma_prefetch_desc(&mapHi); tma_prefetch_desc(&mapLo); }
    if (warp == F_MMA_WARP) tmem_alloc(&tmemBase, TMEM_COLS);
    pdl_launch_dependents();
    pdl_wait();                                                                   // barriers, tensor memory and descriptors are set up under the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;
    const uint32_t ringAddr = smem_u32(smem);

    if (warp < F_WORKERS) {
        // ------------------------------------------------------------ workers
        const uint32_t q = warp & 3, part = warp >> 2;
        const uint32_t laneBase = (q * 32) << 16;
        float loss = 0.0f;
        uint32_t seq = 0;
        // this thread's part of a unit's W^T slice: row n = m0 + 32 q + lane, k in [32 part, 32 part + 32)
        auto loadA = [&](uint32_t unit, uint32_t (&ra)[32]) {
            const uint32_t n = (unit / f.groups) * BM + q * 32 + lane;
            const float* p = f.W + (size_t)(32 * part) * f.ldw + n;
            const bool nIn = n < f.N;
#pragma unroll
            for (int e = 0; e < 32; e++) ra[e] = (nIn && 32 * part + e < f.K && !(f.debug & 32768)) ? ldg_nc_u32(p + (size_t)e * f.ldw) : 0u;
        };
        auto storeA = [&](const uint32_t (&ra)[32]) {
#pragma unroll
            for (int g = 0; g < 2; g++) {
                uint32_t t[16];
#pragma unroll
                for (int e = 0; e < 16; e++) t[e] = hi_of(ra[16 * g + e]);
                tmem_st16(tmem + laneBase + 32 * part + 16 * g, t);
                if (lo) {
#pragma unroll
                    for (int e = 0; e < 16; e++) t[e] = lo_of(ra[16 * g + e]);
                    tmem_st16(tmem + laneBase + F_A_LO + 32 * part + 16 * g, t);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&aReady);
        };
        uint32_t unit = blockIdx.x;
        const bool dbgOn = (f.debug & 65536) != 0;
        long long cTotal = 0, cWaitAcc = 0, cLd = 0, cMath = 0, cA = 0;
        const long long tStart = dbgOn ? clock64() : 0;
        if (unit < numUnits) { uint32_t ra[32]; loadA(unit, ra); storeA(ra); }
        if (dbgOn) cA += clock64() - tStart;
        for (; unit < numUnits; unit += gridDim.x) {
            const uint32_t next = unit + gridDim.x;
            const uint32_t mT = unit / f.groups, g = unit % f.groups;
            const uint32_t n = mT * BM + q * 32 + lane;
            const bool nIn = n < f.N;
            const float cb = ((nIn && f.bias) ? __ldg(f.bias + n) : 0.0f) * (FAST ? kLog2e : 1.0f);
            const uint32_t bt0 = g * f.tilesPerGroup, bt1 = min(f.tilesB, bt0 + f.tilesPerGroup);
            float colSum = 0.0f;
            for (uint32_t bt = bt0; bt < bt1; bt++, seq++) {
                const uint32_t acc = seq & 1;
                const uint32_t b0 = bt * BN + part * 32;
                // this thread's target word of the tile (32 batch rows), requested before the accumulator is waited for
                const uint32_t bits = (nIn && b0 < f.batch) ? __ldg(f.bitsT + (size_t)n * f.wordsB + (b0 >> 5)) : 0u;
                if (bt + 1 == bt1 && next < numUnits) {
                    // last tile of the unit: fetch the next unit's slice while the tensor core finishes, store it as soon as every MMA that
                    // reads the current slice has retired
                    uint32_t ra[32];
                    { DSB_DBG_T0(dbgOn); loadA(next, ra); DSB_DBG_ADD(dbgOn, cA); }
                    __syncwarp();
                    { DSB_DBG_T0(dbgOn); mbar_wait(&accFull[acc], (seq >> 1) & 1); DSB_DBG_ADD(dbgOn, cWaitAcc); }
                    tc_fence_after();
                    { DSB_DBG_T0(dbgOn); storeA(ra); DSB_DBG_ADD(dbgOn, cA); }
                } else {
                    __syncwarp();
                    { DSB_DBG_T0(dbgOn); mbar_wait(&accFull[acc], (seq >> 1) & 1); DSB_DBG_ADD(dbgOn, cWaitAcc); }
                    tc_fence_after();
                }
                float v[32];
                __syncwarp();
                { DSB_DBG_T0(dbgOn);
                tmem_ld32(tmem + laneBase + F_ACC0 + acc * BN + part * 32, v);
                tc_fence_before();                                            // this warp's quarter of the accumulator is in registers
                __syncwarp();
                if (lane == 0) mbar_arrive(&accEmpty[acc]);
                DSB_DBG_ADD(dbgOn, cLd); }
                if (!nIn || b0 >= f.batch || (f.debug & 2048)) continue;      // the host guarantees batch % 32 == 0: all 32 rows exist
                DSB_DBG_T0(dbgOn);
                float* o = f.delta + (size_t)b0 * f.ldd + n;
                float* u = WANTX ? f.unit + (size_t)b0 * f.ldd + n : nullptr;
                float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    float x[16], d[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float wd = HASW ? __ldg(f.rowW + b0 + 16 * hh + j) : 1.0f;
                        out_elem<ISL2, FAST, HASW, PLAIN, WANTX>(f, v[16 * hh + j], cb, (bits >> (16 * hh + j)) & 1u, wd, (j & 1) ? l1 : l0, x[j], d[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) { if (!(f.debug & 8192)) *o = d[j]; o += f.ldd; colSum += d[j]; }
                    if (WANTX) {
#pragma unroll
                        for (int j = 0; j < 16; j++) { *u = x[j]; u += f.ldd; }
                    }
                }
                loss += (l0 + l1) * ((ISL2 ? 0.5f : (PLAIN ? f.lZ : 1.0f)) * ((FAST && !ISL2) ? kLn2 : 1.0f));
                DSB_DBG_ADD(dbgOn, cMath);
            }
            if (f.colPartials && nIn) f.colPartials[(size_t)(g * F_PARTS + part) * f.N + n] = colSum;
        }
        if (f.acc) {
            const double e = warp_sum((double)loss);
            if (lane == 0 && e != 0.0) atomicAdd(f.acc, (unsigned long long)llrint(e * (double)kErrorScaleF));
        }
        if (dbgOn && (warp == 0 || warp == 15) && lane == 0 && blockIdx.x < 256) {
            unsigned long long* o = g_dbgCounters + blockIdx.x * 16 + (warp == 0 ? 0 : 5);
            cTotal = clock64() - tStart;
            o[0] = cTotal; o[1] = cWaitAcc; o[2] = cLd; o[3] = cMath; o[4] = cA;
        }
    } else if (warp == F_TMA_WARP) {
        // ------------------------------------------------------------ B producer
        if (lane == 0) {
            uint32_t slot = 0, parity = 1;
            for (uint32_t unit = blockIdx.x; unit < numUnits; unit += gridDim.x) {
                const uint32_t g = unit % f.groups;
                const uint32_t bt0 = g * f.tilesPerGroup, bt1 = min(f.tilesB, bt0 + f.tilesPerGroup);
                for (uint32_t bt = bt0; bt < bt1; bt++)
                    for (uint32_t kt = 0; kt < numK; kt++) {
                        mbar_wait(&bEmpty[slot], parity);
                        const uint32_t st = ringAddr + slot * SLOT;
                        if (f.debug & 16384) { mbar_arrive(&bFull[slot]); }       // bring-up switch 16384: no TMA loads
                        else {
                        mbar_arrive_expect_tx(&bFull[slot], lo ? SLOT : PANEL);
                        tma_load_2d(st, &mapHi, kt * BK, bt * BN, &bFull[slot]);
                        if (lo) tma_load_2d(st + PANEL, &mapLo, kt * BK, bt * BN, &bFull[slot]);
                        }
                        if (++slot == F_SLOTS) { slot = 0; parity ^= 1; }
                    }
            }
        }
    } else if (warp == F_MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t slot = 0, ph = 0, seq = 0, ui = 0;
            const bool dbgOn = (f.debug & 65536) != 0;
            long long cWaitA = 0, cWaitAcc = 0, cWaitB = 0, cIssue = 0;
            const long long tStart = dbgOn ? clock64() : 0;
            for (uint32_t unit = blockIdx.x; unit < numUnits; unit += gridDim.x, ui++) {
                const uint32_t g = unit % f.groups;
                const uint32_t bt0 = g * f.tilesPerGroup, bt1 = min(f.tilesB, bt0 + f.tilesPerGroup);
                { DSB_DBG_T0(dbgOn); mbar_wait(&aReady, ui & 1); DSB_DBG_ADD(dbgOn, cWaitA); }   // this unit's W^T slice is in tensor memory
                tc_fence_after();
                for (uint32_t bt = bt0; bt < bt1; bt++, seq++) {
                    const uint32_t acc = seq & 1, d = tmem + F_ACC0 + acc * BN;
                    { DSB_DBG_T0(dbgOn); mbar_wait(&accEmpty[acc], ((seq >> 1) & 1) ^ 1); DSB_DBG_ADD(dbgOn, cWaitAcc); }
                    tc_fence_after();
                    for (uint32_t kt = 0; kt < numK; kt++) {
                        { DSB_DBG_T0(dbgOn); mbar_wait(&bFull[slot], ph); DSB_DBG_ADD(dbgOn, cWaitB); }
                        tc_fence_after();
                        DSB_DBG_T0(dbgOn);
                        const uint32_t sb = ringAddr + slot * SLOT;
#pragma unroll
                        for (int j = 0; j < BK / 8; j++) {
                            const uint64_t bHi = kmajor_desc(sb, j);
                            const uint32_t aHi = tmem + kt * BK + j * 8, first = (kt == 0 && j == 0) ? 0u : 1u;
                            if (f.debug & 1024) {
                            } else if (lo) {
                                const uint64_t bLo = kmajor_desc(sb + PANEL, j);
                                mma_tf32_ts(d, aHi + F_A_LO, bHi, kIdesc, first); // small terms first
                                mma_tf32_ts(d, aHi, bLo, kIdesc, 1u);
                                mma_tf32_ts(d, aHi, bHi, kIdesc, 1u);
                            } else {
                                mma_tf32_ts(d, aHi, bHi, kIdesc, first);
                            }
                        }
                        tc_commit(&bEmpty[slot]);
                        DSB_DBG_ADD(dbgOn, cIssue);
                        if (++slot == F_SLOTS) { slot = 0; ph ^= 1; }
                    }
                    tc_commit(&accFull[acc]);
                }
            }
            if (dbgOn && blockIdx.x < 256) {
                unsigned long long* o = g_dbgCounters + blockIdx.x * 16 + 10;
                o[0] = clock64() - tStart; o[1] = cWaitA; o[2] = cWaitAcc; o[3] = cWaitB; o[4] = cIssue;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == F_MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// the three derived copies of the last hidden layer's units X[rows][cols] in one launch: xlo[r][c] = lo(X) (B operand of the forward
// pass), xtHi[c][r] = X, xtLo[c][r] = lo(X) (B operand of the weight gradient, pitch ldt)
__global__ void __launch_bounds__(256)
prep_x_kernel(const float* __restrict__ X, uint32_t rows, uint32_t cols, float* __restrict__ xlo, float* __restrict__ xtHi, float* __restrict__ xtLo, uint32_t ldt)
{
    __shared__ float tile[32][33];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t tilesC = (cols + 31) / 32, tilesR = (ldt + 31) / 32;
    pdl_launch_dependents();
    pdl_wait();
    for (uint32_t t = blockIdx.x; t < tilesC * tilesR; t += gridDim.x) {
        const uint32_t r0 = (t / tilesC) * 32, c0 = (t % tilesC) * 32;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t r = r0 + ty + 8 * i, c = c0 + tx;
            const float x = (r < rows && c < cols) ? __ldg(X + (size_t)r * cols + c) : 0.0f;
            tile[ty + 8 * i][tx] = x;
            if (r < rows && c < cols) xlo[(size_t)r * cols + c] = __uint_as_float(lo_of(__float_as_uint(x)));
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t c = c0 + ty + 8 * i, r = r0 + tx;
            if (c < cols && r < ldt) {
                const float x = tile[tx][ty + 8 * i];
                xtHi[(size_t)c * ldt + r] = x;
                xtLo[(size_t)c * ldt + r] = __uint_as_float(lo_of(__float_as_uint(x)));
            }
        }
        __syncthreads();
    }
}

// transposed target bitmap: bit (b % 32) of bitsT[c][b / 32] = output c is a non-zero target of batch row b (cleared by the caller)
__global__ void __launch_bounds__(256)
target_bitmap_t_kernel(const dsb200_params P, const dsb200_sparse S, uint32_t position, uint32_t batch, uint32_t width, uint32_t wordsB,
                       uint32_t* __restrict__ bitsT, float* __restrict__ rowW)
{
    for (uint32_t b = blockIdx.x; b < batch; b += gridDim.x) {
        const uint32_t ex = example_of(P, S.index, position, b);
        const uint64_t rs = __ldg(S.sparseStart + ex), re = __ldg(S.sparseEnd + ex);
        for (uint64_t j = rs + threadIdx.x; j < re; j += blockDim.x) {
            const uint32_t c = __ldg(S.sparseIndex + j);
            if (c < width) atomicOr(bitsT + (size_t)c * wordsB + (b >> 5), 1u << (b & 31));
        }
        if (threadIdx.x == 0 && rowW) rowW[b] = S.dataWeight ? __ldg(S.dataWeight + ex) : 1.0f;
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// K-major operand [rows][cols] (cols contiguous, pitch ld floats, ld % 4 == 0, base 16-byte aligned): boxes of 32 x 128
static bool make_map(CUtensorMap* m, const float* base, uint32_t rows, uint32_t cols, uint32_t ld)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {BK, BN};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline uint32_t up4(uint32_t x) { return (x + 3u) & ~3u; }
static inline bool tma_ok(const float* p, uint32_t ld) { return (((uintptr_t)p) & 15u) == 0 && (ld & 3u) == 0; }

}  // namespace gs

// grow-only scratch of the streamed GEMMs (operand copies, split-K partials, target bitmap), carved per call
static int gs_reserve(dsb200_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->gsWsBytes) return 0;
    if (ctx->dGsWs) { DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream)); DSB_CUDA_OK(cudaFree(ctx->dGsWs)); ctx->dGsWs = nullptr; ctx->gsWsBytes = 0; }
    bytes += bytes / 8;
    DSB_CUDA_OK(cudaMalloc(&ctx->dGsWs, bytes));
    ctx->gsWsBytes = bytes;
    return 0;
}
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static int prep_reserve(dsb200_ctx* ctx, dsb200_ctx::Prep& p, size_t bytes)
{
    p.valid = false;
    if (bytes <= p.bytes) return 0;
    if (p.buf) { DSB_CUDA_OK(cudaDeviceSynchronize()); DSB_CUDA_OK(cudaFree(p.buf)); p.buf = nullptr; p.bytes = 0; }
    bytes += bytes / 8;
    DSB_CUDA_OK(cudaMalloc(&p.buf, bytes));
    p.bytes = bytes;
    return 0;
}

bool gemm_stream_available() { return gs::encode_fn() != nullptr; }
int gemm_stream_debug_counters(unsigned long long* out, size_t count)
{
    return cudaMemcpyFromSymbol(out, gs::g_dbgCounters, std::min<size_t>(count, 256 * 16) * sizeof(unsigned long long)) == cudaSuccess ? 0 : DSB200_ESTATE;
}

// G[k][n] = beta * G + alpha * X[B][k]^T * D[B][n]      (swapped: M = n, N = k, K = B; A = D read with m contiguous)
int gemm_stream_dw(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, float alpha, const float* X, const float* D, uint32_t ldd, float beta,
                   float* G, uint32_t ldg)
{
    using namespace gs;
    static bool attrSet = false;
    if (!attrSet) {
        DSB_CUDA_OK(cudaFuncSetAttribute(stream_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES));
        attrSet = true;
    }
    const uint32_t ldt = up4(B);
    const size_t opBytes = al256((size_t)k * ldt * sizeof(float));
    float* xtHi; float* xtLo;
    if (ctx->prepX.valid && ctx->prepX.key == X && ctx->prepX.a == B && ctx->prepX.b == k) {
        // the forward pass of this step left X^T (hi | lo) behind (gemm_stream_out_fwd)
        ctx->prepX.valid = false;
        uint8_t* pb = reinterpret_cast<uint8_t*>(ctx->prepX.buf) + al256((size_t)B * k * sizeof(float));
        xtHi = reinterpret_cast<float*>(pb); xtLo = reinterpret_cast<float*>(pb + opBytes);
    } else {
        int rc = gs_reserve(ctx, 2 * opBytes);
        if (rc) return rc;
        xtHi = reinterpret_cast<float*>(ctx->dGsWs);
        xtLo = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->dGsWs) + opBytes);
        const uint32_t tiles = ((k + 31) / 32) * ((ldt + 31) / 32);
        transpose_split_kernel<<<std::min<uint32_t>(tiles, (uint32_t)ctx->numSMs * 8), 256, 0, ctx->stream>>>(X, k, B, k, xtHi, xtLo, ldt);
        count_launch();
    }
    CUtensorMap mh, ml;
    if (!make_map(&mh, xtHi, k, B, ldt) || !make_map(&ml, xtLo, k, B, ldt)) return fail(ctx, DSB200_ESTATE, "gemm_stream_dw: cuTensorMapEncodeTiled failed");
    SArgs a{};
    a.A = D; a.lda = ldd; a.M = n; a.N = k; a.K = B;
    a.tilesM = (n + BM - 1) / BM; a.tilesN = (k + BN - 1) / BN; a.splits = 1; a.kPerSplit = ((B + BK - 1) / BK) * BK;
    a.C = G; a.ldc = ldg; a.partial = nullptr; a.alpha = alpha; a.beta = beta;
    a.passes = (ctx->gemmMode == DSB200_GEMM_TF32) ? 1 : 3;
    a.debug = ctx->gemmDebug;
    const uint32_t grid = std::min<uint32_t>((uint32_t)ctx->numSMs, a.tilesM * a.tilesN);
    DSB_CUDA_OK(launch_pdl(stream_kernel<0, 0>, dim3(grid), dim3(S_THREADS), S_SMEM_BYTES, ctx->stream, a, mh, ml));
    DSB_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

// Dp[B][k] = beta * Dp + D[B][n] * W[k][n]^T            (M = B, N = k, K = n split; A = D read with k contiguous)
int gemm_stream_dx(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, uint32_t ldd, const float* W, uint32_t ldw, float beta,
                   float* Dp, uint32_t ldp, const float* hadUnit, int hadAct, float hadScale, float slope, float ealpha, float lambda)
{
    using namespace gs;
    static bool attrSet = false;
    if (!attrSet) {
        DSB_CUDA_OK(cudaFuncSetAttribute(stream_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES));
        attrSet = true;
    }
    SArgs a{};
    a.A = D; a.lda = ldd; a.M = B; a.N = k; a.K = n;
    a.tilesM = (B + BM - 1) / BM; a.tilesN = (k + BN - 1) / BN;
    const uint32_t tilesMN = a.tilesM * a.tilesN, kTiles = (n + BK - 1) / BK, sms = (uint32_t)ctx->numSMs;
    // split K so that the persistent grid runs whole rounds: minimise rounds * (k-iterations + fixed tile cost)
    uint32_t splits = 1;
    if (ctx->gemmSplits > 0) splits = std::min<uint32_t>((uint32_t)ctx->gemmSplits, kTiles);
    else {
        uint64_t best = ~0ull;
        for (uint32_t s = 1; s <= 64 && s * 8 <= std::max(kTiles, 8u); s++) {
            const uint32_t per = (kTiles + s - 1) / s, real = (kTiles + per - 1) / per;
            const uint64_t rounds = ((uint64_t)tilesMN * real + sms - 1) / sms;
            const uint64_t cost = rounds * (per + 12) + (real > 1 ? (uint64_t)real * tilesMN / sms + 4 : 0);
            if (cost < best) { best = cost; splits = real; }
        }
    }
    const uint32_t per = (kTiles + splits - 1) / splits;
    splits = (kTiles + per - 1) / per;
    a.splits = splits; a.kPerSplit = per * BK;
    // operand copies: lo always; hi only when W itself cannot be a tensor map (pitch not a multiple of 16 bytes).  Made here, or
    // ahead of time on another stream by dsb200_gemm_dx_prepare (W does not change between the optimizer step and this call)
    const bool direct = tma_ok(W, ldw);
    const uint32_t ldc = up4(n);
    const size_t opBytes = al256((size_t)k * ldc * sizeof(float));
    const size_t partBytes = splits > 1 ? al256((size_t)splits * B * k * sizeof(float)) : 0;
    const bool prepared = ctx->prepW.valid && ctx->prepW.key == W && ctx->prepW.a == k && ctx->prepW.b == n && ctx->prepW.c == ldw;
    int rc = gs_reserve(ctx, (prepared ? 0 : 2 * opBytes) + partBytes);
    if (rc) return rc;
    uint8_t* ws = reinterpret_cast<uint8_t*>(ctx->dGsWs);
    float* wLo; float* wHi;
    if (prepared) {
        ctx->prepW.valid = false;
        wLo = reinterpret_cast<float*>(ctx->prepW.buf);
        wHi = direct ? nullptr : reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->prepW.buf) + opBytes);
        a.partial = splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
    } else {
        wLo = reinterpret_cast<float*>(ws);
        wHi = direct ? nullptr : reinterpret_cast<float*>(ws + opBytes);
        a.partial = splits > 1 ? reinterpret_cast<float*>(ws + 2 * opBytes) : nullptr;
        const size_t total = (size_t)k * ldc;
        const uint32_t blocks = (uint32_t)std::min<size_t>((total + 255) / 256, (size_t)ctx->numSMs * 16);
        split_pitch_kernel<<<blocks, 256, 0, ctx->stream>>>(W, ldw, k, n, wHi, wLo, ldc);
        count_launch();
    }
    CUtensorMap mh, ml;
    const bool ok = (direct ? make_map(&mh, W, k, n, ldw) : make_map(&mh, wHi, k, n, ldc)) && make_map(&ml, wLo, k, n, ldc);
    if (!ok) return fail(ctx, DSB200_ESTATE, "gemm_stream_dx: cuTensorMapEncodeTiled failed");
    a.C = Dp; a.ldc = ldp; a.alpha = 1.0f; a.beta = beta;
    a.passes = (ctx->gemmMode == DSB200_GEMM_TF32) ? 1 : 3;
    a.debug = ctx->gemmDebug;
    const uint32_t grid = std::min<uint32_t>(sms, tilesMN * splits);
    DSB_CUDA_OK(launch_pdl(stream_kernel<1, 1>, dim3(grid), dim3(S_THREADS), S_SMEM_BYTES, ctx->stream, a, mh, ml));
    DSB_CUDA_OK(cudaGetLastError());
    count_launch();
    if (splits > 1) {
        const size_t total = (size_t)B * k;
        const uint32_t blocks = (uint32_t)std::min<size_t>((total + 255) / 256, (size_t)ctx->numSMs * 8);
        HadArgs h{hadUnit, hadAct, hadScale, slope, ealpha, lambda};
        DSB_CUDA_OK(launch_pdl(stream_reduce_kernel, dim3(blocks), dim3(256), 0, ctx->stream, (const float*)a.partial, splits, B, k, ldp, 1.0f, beta, Dp, h));
        DSB_CUDA_OK(cudaGetLastError());
        count_launch();
    } else if (hadUnit) {
        return dsb200_hadamard(ctx, hadAct, (uint64_t)B * k, hadScale, hadUnit, Dp, slope, ealpha, lambda);
    }
    return 0;
}

// transposed target bitmap (+ row weights) of the batch at `position` into the context's prepared-operand buffer
static int out_fwd_prepare_bits(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t n)
{
    using namespace gs;
    const uint32_t wordsB = (batch + 31) / 32;
    const size_t bitsBytes = al256((size_t)n * wordsB * sizeof(uint32_t)), rowBytes = al256((size_t)batch * sizeof(float));
    int rc = prep_reserve(ctx, ctx->prepBits, bitsBytes + rowBytes);
    if (rc) return rc;
    uint32_t* bitsT = reinterpret_cast<uint32_t*>(ctx->prepBits.buf);
    float* rowW = s->dataWeight ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->prepBits.buf) + bitsBytes) : nullptr;
    DSB_CUDA_OK(cudaMemsetAsync(bitsT, 0, (size_t)n * wordsB * sizeof(uint32_t), ctx->stream));
    target_bitmap_t_kernel<<<std::min<uint32_t>(batch, (uint32_t)ctx->numSMs * 8), 256, 0, ctx->stream>>>(ctx->params, *s, position, batch, n, wordsB, bitsT, rowW);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    ctx->prepBits.key = s->sparseIndex; ctx->prepBits.a = position; ctx->prepBits.b = batch; ctx->prepBits.c = n; ctx->prepBits.valid = true;
    return 0;
}
int gemm_stream_prepare_targets(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t n)
{
    return out_fwd_prepare_bits(ctx, s, position, batch, n);
}
// hi / lo pitch-padded copies of W[k][n] for the input-delta kernel, ahead of the call
int gemm_stream_prepare_dx(dsb200_ctx* ctx, uint32_t k, uint32_t n, const float* W, uint32_t ldw)
{
    using namespace gs;
    const bool direct = tma_ok(W, ldw);
    const uint32_t ldc = up4(n);
    const size_t opBytes = al256((size_t)k * ldc * sizeof(float));
    int rc = prep_reserve(ctx, ctx->prepW, 2 * opBytes);
    if (rc) return rc;
    float* wLo = reinterpret_cast<float*>(ctx->prepW.buf);
    float* wHi = direct ? nullptr : reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->prepW.buf) + opBytes);
    const size_t total = (size_t)k * ldc;
    const uint32_t blocks = (uint32_t)std::min<size_t>((total + 255) / 256, (size_t)ctx->numSMs * 16);
    split_pitch_kernel<<<blocks, 256, 0, ctx->stream>>>(W, ldw, k, n, wHi, wLo, ldc);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    ctx->prepW.key = W; ctx->prepW.a = k; ctx->prepW.b = n; ctx->prepW.c = ldw; ctx->prepW.valid = true;
    return 0;
}

// forward + loss + delta of a sigmoid output layer (see out_fwd_kernel).  pColPartials: optional [*pNumPartials][n] column sums of delta.
int gemm_stream_out_fwd(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, uint32_t position, uint32_t batch, uint32_t k, uint32_t n, const float* X,
                        const float* W, uint32_t ldw, const float* bias, float* unitOut, float* delta, uint32_t ldd, unsigned long long* acc,
                        float* pColPartials, uint32_t* pNumPartials)
{
    using namespace gs;
    if (k > 128 || !tma_ok(X, k) || (batch & 31u)) return DSB200_EUNSUPPORTED;
    FArgs f{};
    f.W = W; f.ldw = ldw; f.bias = bias; f.N = n; f.K = k; f.batch = batch;
    f.tilesM = (n + BM - 1) / BM; f.tilesB = (batch + BN - 1) / BN;
    // groups of batch tiles per output slice: minimise rounds * (tiles per unit + unit overhead)
    uint32_t bestG = 1; double bestCost = 1e30;
    for (uint32_t g = 1; g <= f.tilesB; g++) {
        const uint32_t tpg = (f.tilesB + g - 1) / g, real = (f.tilesB + tpg - 1) / tpg;
        const uint32_t rounds = (f.tilesM * real + (uint32_t)ctx->numSMs - 1) / (uint32_t)ctx->numSMs;
        const double cost = rounds * (tpg + 0.35);
        if (cost < bestCost - 1e-9) { bestCost = cost; bestG = real; }
    }
    f.tilesPerGroup = (f.tilesB + bestG - 1) / bestG;
    f.groups = (f.tilesB + f.tilesPerGroup - 1) / f.tilesPerGroup;
    f.delta = delta; f.unit = unitOut; f.ldd = ldd;
    f.wordsB = (batch + 31) / 32;
    const size_t bitsBytes = al256((size_t)n * f.wordsB * sizeof(uint32_t));
    const size_t rowBytes = al256((size_t)batch * sizeof(float));
    const bool bitsReady = ctx->prepBits.valid && ctx->prepBits.key == s->sparseIndex && ctx->prepBits.a == position && ctx->prepBits.b == batch && ctx->prepBits.c == n;
    if (!bitsReady) {
        const int rc = out_fwd_prepare_bits(ctx, s, position, batch, n);
        if (rc) return rc;
    }
    ctx->prepBits.valid = false;
    uint32_t* bitsT = reinterpret_cast<uint32_t*>(ctx->prepBits.buf);
    float* rowW = s->dataWeight ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->prepBits.buf) + bitsBytes) : nullptr;
    // derived copies of X: lo for this kernel's B operand, X^T (hi | lo) for the weight gradient that follows in the backward pass
    const uint32_t ldt = up4(batch);
    const size_t loBytes = al256((size_t)batch * k * sizeof(float)), xtBytes = al256((size_t)k * ldt * sizeof(float));
    f.passes = (ctx->gemmMode == DSB200_GEMM_TF32) ? 1 : 3;
    f.debug = ctx->gemmDebug;
    int rc = prep_reserve(ctx, ctx->prepX, loBytes + 2 * xtBytes);
    if (rc) return rc;
    float* xLo = reinterpret_cast<float*>(ctx->prepX.buf);
    {
        uint8_t* pb = reinterpret_cast<uint8_t*>(ctx->prepX.buf) + loBytes;
        const uint32_t tiles = ((k + 31) / 32) * ((ldt + 31) / 32);
        DSB_CUDA_OK(launch_pdl(prep_x_kernel, dim3(std::min<uint32_t>(tiles, (uint32_t)ctx->numSMs * 8)), dim3(256), 0, ctx->stream, X, batch, k, xLo,
                               reinterpret_cast<float*>(pb), reinterpret_cast<float*>(pb + xtBytes), ldt));
        count_launch();
        ctx->prepX.key = X; ctx->prepX.a = batch; ctx->prepX.b = k; ctx->prepX.valid = true;
    }
    CUtensorMap mh, ml;
    if (!make_map(&mh, X, batch, k, k) || !make_map(&ml, xLo, batch, k, k)) return fail(ctx, DSB200_ESTATE, "gemm_stream_out_fwd: cuTensorMapEncodeTiled failed");
    (void)rowBytes;
    f.bitsT = bitsT; f.rowW = rowW; f.acc = acc; f.colPartials = pColPartials;
    if (pNumPartials) *pNumPartials = f.groups * F_PARTS;
    // element coefficients (out_elem)
    const dsb200_params& P = ctx->params;
    bool plain = true;
    if (ef == DSB200_ERR_SMCE) {
        f.thrZ = P.SMCE_zeroTarget; f.thrNz = 1.0f - P.SMCE_oneTarget;
        f.lZ = P.SMCE_zeroScale; f.lNz = P.SMCE_oneScale; f.dZ = P.SMCE_zeroScale; f.dNz = -P.SMCE_oneScale;
        plain = P.SMCE_zeroTarget <= 0.0f && P.SMCE_oneTarget >= 1.0f && P.SMCE_zeroScale == P.SMCE_oneScale;   // p = 0 contributes nothing either way
    } else {
        f.thrZ = f.thrNz = -1.0f; f.lZ = f.lNz = 1.0f; f.dZ = P.deltaBoost_zero; f.dNz = -P.deltaBoost_one;
    }
    const uint32_t grid = std::min<uint32_t>((uint32_t)ctx->numSMs, f.tilesM * f.groups);
    typedef void (*Kern)(const FArgs, const CUtensorMap, const CUtensorMap);
    static const Kern table[24] = {
#define DSB_K4(L2, FAST, HASW) out_fwd_kernel<L2, FAST, HASW, false, false>, out_fwd_kernel<L2, FAST, HASW, false, true>, \
                               out_fwd_kernel<L2, FAST, HASW, true, false>, out_fwd_kernel<L2, FAST, HASW, true, true>
        DSB_K4(false, false, false), DSB_K4(false, false, true), DSB_K4(false, true, false), DSB_K4(false, true, true),
        // L2 has no gate: only its PLAIN instances exist
        out_fwd_kernel<true, false, false, true, false>, out_fwd_kernel<true, false, false, true, true>, out_fwd_kernel<true, false, true, true, false>,
        out_fwd_kernel<true, false, true, true, true>, out_fwd_kernel<true, true, false, true, false>, out_fwd_kernel<true, true, false, true, true>,
        out_fwd_kernel<true, true, true, true, false>, out_fwd_kernel<true, true, true, true, true>
#undef DSB_K4
    };
    static bool attrSet = false;
    if (!attrSet) {
        for (int i = 0; i < 24; i++) DSB_CUDA_OK(cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
        attrSet = true;
    }
    const int fast = ctx->fastMath ? 1 : 0, hasw = rowW ? 1 : 0, wantx = unitOut ? 1 : 0;
    const Kern kern = (ef == DSB200_ERR_L2) ? table[16 + fast * 4 + hasw * 2 + wantx] : table[fast * 8 + hasw * 4 + (plain ? 2 : 0) + wantx];
    DSB_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(F_THREADS), F_SMEM_BYTES, ctx->stream, f, mh, ml));
    DSB_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace dsb
