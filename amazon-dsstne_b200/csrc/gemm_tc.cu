// gemm_tc.cu -- hand-written tcgen05 / TMEM GEMM for the dense layers (hot-path row a11), sm_100a only.
//
// Replaces cublasSgemm at E/NNLayer.cpp:1073 (forward), :2223 (weight gradient), :2274 (input delta) when
// gemm_mode is DSB200_GEMM_TF32X3 (fp32-grade: every fp32 operand is split a = hi + lo, hi = the 19 bits the tf32
// tensor core reads, lo = the exact fp32 remainder, and lo*hi + hi*lo + hi*hi accumulate in fp32 in TMEM) or
// DSB200_GEMM_TF32 (hi*hi only).
//
// One 128x128 output tile per CTA, two CTAs per SM so that one CTA's epilogue overlaps the other's main loop.
//   warps 0-7  loaders, then epilogue.  Operand rows are not 16-byte aligned in general (N = 27,278 floats), so TMA
//              tensor maps are not usable; each thread copies its 16-byte chunks global -> shared with cp.async
//              (16 / 8 / 4-byte pieces by alignment, zero fill at the matrix edge) straight into the UMMA layouts,
//              ring depth - 1 panels ahead, then derives the "lo" panel of the 3xTF32 split from its own chunks.
//   warp 8     TMEM allocation; lane 0 issues every tcgen05.mma and signals stage reuse / accumulator completion
//              with tcgen05.commit on mbarriers.
// Shared-memory layouts of one 128 x 16 operand panel (16-byte chunks = 4 floats along the contiguous dimension):
//   K-major operand (k contiguous in memory): no swizzle, core matrix = 8 mn-rows x 16 bytes,
//       chunk(mn, kc) at (mn/8)*128 + (mn%8)*16 + kc*2048                                   LBO=2048 SBO=128
//       lanes of a load: 8 rows x 4 chunks -- a quarter-warp writes one whole core matrix (conflict free).
//   MN-major operand (mn contiguous in memory): the 32-bit "128B, 32B-base" swizzle, the only layout the tensor core
//       accepts for transposed tf32 operands (every other layout type returns zeros -- measured, tools/umma_probe.cu):
//       atom = 4 k-rows x 128 bytes (32 mn), 32-byte granules XOR-ed with k%4,
//       chunk(k, mc) at (mc/8)*2048 + (k/4)*512 + (k%4)*128 + ((((mc%8)/2) ^ (k%4))*32) + (mc%2)*16   LBO=2048 SBO=512
//       lanes of a load: 4 k-rows x 8 chunks (128 contiguous bytes per row) -- conflict free as well.
// Split-K (K = 27,278 for the input-delta GEMM of the output layer) writes raw partial tiles to a workspace that
// gemm_reduce_kernel sums in a fixed order -- deterministic, no float atomics.
#include "common.cuh"
#include "launch.h"

namespace dsb {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 16, MAX_STAGES = 6;
constexpr int PANEL = 128 * BK * 4;            // one operand panel (raw or lo): 8 KB
constexpr int STAGE_BYTES = 4 * PANEL;         // A raw | A lo | B raw | B lo
constexpr int LOADERS = 256, THREADS = 288, MMA_WARP = 8;   // warps 0-7 load + run the epilogue, warp 8 issues the MMAs
constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024; }

struct Args {
    const float* A; const float* B; float* C;
    uint32_t M, N, K;                          // C[M][N] (+)= A(M x K) * B(K x N)
    uint32_t lda, ldb, ldc;
    int aMN, bMN;                              // operand is MN-contiguous in memory (else K-contiguous)
    int vecA, vecB, vecC;                      // widest aligned access in floats (4, 2, 1)
    float alpha, beta;
    const float* bias; int act; float slope, ealpha, lambda;
    uint32_t kPerSplit;                        // multiple of BK; == K rounded up when not split
    float* partial;                            // split-K workspace [splits][M][N] or NULL
    int passes;                                // 3 = 3xTF32, 1 = TF32
    int debug;                                 // bring-up switches: 1 no global loads, 2 no MMA, 4 no epilogue stores, 8 no lo pass
    uint32_t depth;                            // panels of copies in flight ahead of the consumer
    uint32_t stages;                           // shared-memory ring depth: 3 (two CTAs per SM) or 6 (one CTA per SM)
};

__device__ __forceinline__ float act_apply(int act, float z, float slope, float alpha, float lambda)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    case DSB200_ACT_TANH:    return tanhf(z);
    case DSB200_ACT_RELU:    return fmaxf(0.0f, z);
    case DSB200_ACT_LRELU:   return fmaxf(z, z * slope);
    case DSB200_ACT_ELU:     return (z > 0.0f) ? z : alpha * (expf(z) - 1.0f);
    case DSB200_ACT_SELU:    return (z > 0.0f) ? lambda * z : lambda * alpha * (expf(z) - 1.0f);
    default:                 return z;
    }
}

// ---- tcgen05 wrappers ----
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" :: "r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor, version 1 (Blackwell); layout type 0 = no swizzle, 1 = 128B swizzle with 32B base
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes, uint32_t layoutType)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lboBytes >> 4) << 16) | ((uint64_t)(sboBytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layoutType << 61);
}
template <bool MN>
__device__ __forceinline__ uint64_t panel_desc(uint32_t panelAddr, int kStep)
{
    // one K = 8 step: two 16-byte chunks (K-major) or two groups of 4 k-rows (MN-major)
    return MN ? smem_desc(panelAddr + kStep * 1024, 2048, 512, 1) : smem_desc(panelAddr + kStep * 4096, 2048, 128, 0);
}
// 32 lanes x 32 consecutive fp32 columns: thread t of warp w gets row 32*(w%4)+t
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

// ---- operand staging: global -> shared with cp.async (LDGSTS), zero fill outside the matrix ----
// The raw fp32 panel doubles as the "hi" operand: the tf32 tensor core reads the upper 19 bits of each word and
// ignores the rest, so hi = trunc_tf32(a) needs no pass of its own; the "lo" panel (a - hi, exact in fp32) is
// produced by the thread that issued the copy, from its own chunks, once its copy group has landed.
__device__ __forceinline__ void cp_async_chunk(uint32_t dst, const float* __restrict__ base, uint32_t ld, uint32_t row, uint32_t col,
                                               uint32_t rowLimit, uint32_t colLimit, int vec)
{
    const bool in = row < rowLimit && col < colLimit;
    const uint32_t valid = in ? min(4u, colLimit - col) : 0u;                   // floats of this chunk inside the matrix
    const float* p = in ? base + (size_t)row * ld + col : base;
    if (vec == 4) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(p), "r"(valid * 4) : "memory");
    } else if (vec == 2) {
        const uint32_t b0 = min(valid, 2u) * 4, b1 = (valid > 2 ? valid - 2 : 0u) * 4;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(p), "r"(b0) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst + 8), "l"(b1 ? p + 2 : p), "r"(b1) : "memory");
    } else {
#pragma unroll
        for (uint32_t e = 0; e < 4; e++)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst + 4 * e), "l"(e < valid ? p + e : p), "r"(e < valid ? 4u : 0u) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait(int pending)
{
    switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    }
}

// Per-thread plan of one operand: CHUNKS 16-byte chunks of every 128 x BK panel, everything that does not depend on
// the k-iteration hoisted out of the main loop (the loaders are issue-bound: ~10 cycles per dependent instruction
// with one warp per scheduler, so instructions per byte is what matters).
constexpr int EPI_COLS = BN / 2;               // a warp's share of the tile in the epilogue: 32 rows x 64 columns
constexpr int EPI_LD = EPI_COLS + 4;           // padded row of the per-warp staging tile (floats)
constexpr int CHUNKS = 2;                      // 16 warp-level chunk groups per panel / 8 loader warps
template <bool MN>
struct OperandPlan {
    const float* ptr[CHUNKS];                  // global address of the chunk in the next k-iteration to issue
    uint32_t     soff[CHUNKS];                 // byte offset inside the panel
    uint32_t     valid[CHUNKS];                // floats inside the matrix along the contiguous dimension (full iterations)
    uint32_t     b0[CHUNKS], b1[CHUNKS];       // source bytes of the first / second copy of a full iteration
    uint32_t     kOff[CHUNKS];                 // first k of the chunk inside the panel
    size_t       step;                         // pointer advance per k-iteration (0 for chunks outside the matrix)
    const float* base;

    __device__ __forceinline__ void init(const float* b, uint32_t ld, uint32_t mn0, uint32_t kBegin, uint32_t mnLimit, int vec, uint32_t warp,
                                         uint32_t lane)
    {
        base = b;
        step = MN ? (size_t)BK * ld : (size_t)BK;
#pragma unroll
        for (int i = 0; i < CHUNKS; i++) {
            const uint32_t u = warp * CHUNKS + i;
            uint32_t mn, k;
            if (MN) {
                mn = mn0 + (u >> 2) * 32 + (lane >> 2) * 4; k = (u & 3) * 4 + (lane & 3);
                soff[i] = (u >> 2) * 2048 + (u & 3) * 512 + (lane & 3) * 128 + (((lane >> 3) ^ (lane & 3)) * 32) + ((lane >> 2) & 1) * 16;
                valid[i] = mn < mnLimit ? min(4u, mnLimit - mn) : 0u;
                ptr[i] = valid[i] ? b + (size_t)(kBegin + k) * ld + mn : b;
            } else {
                mn = mn0 + u * 8 + (lane & 7); k = (lane >> 3) * 4;
                soff[i] = u * 128 + (lane & 7) * 16 + (lane >> 3) * 2048;
                valid[i] = mn < mnLimit ? 4u : 0u;
                ptr[i] = valid[i] ? b + (size_t)mn * ld + kBegin + k : b;
            }
            kOff[i] = k;
            b0[i] = (vec == 4) ? valid[i] * 4 : min(valid[i], 2u) * 4;
            b1[i] = (valid[i] > 2 ? valid[i] - 2 : 0u) * 4;
        }
    }
    // a full panel (k0 + BK <= kEnd): nothing but the copies and the pointer advance
    __device__ __forceinline__ void issue_full(uint32_t panelAddr, int vec)
    {
#pragma unroll
        for (int i = 0; i < CHUNKS; i++) {
            const uint32_t dst = panelAddr + soff[i];
            if (vec == 4) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(ptr[i]), "r"(b0[i]) : "memory");
            } else if (vec == 2) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(ptr[i]), "r"(b0[i]) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst + 8), "l"(ptr[i] + (b1[i] ? 2 : 0)), "r"(b1[i]) : "memory");
            } else {
#pragma unroll
                for (uint32_t e = 0; e < 4; e++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst + 4 * e), "l"(e < valid[i] ? ptr[i] + e : base),
                                 "r"(e < valid[i] ? 4u : 0u) : "memory");
            }
            ptr[i] += valid[i] ? step : 0;
        }
    }
    // the last, partial panel of the K range: element-wise bounds
    __device__ __noinline__ void issue_tail(uint32_t panelAddr, uint32_t k0, uint32_t kEnd)
    {
#pragma unroll
        for (int i = 0; i < CHUNKS; i++) {
            uint32_t v = valid[i];
            const uint32_t k = k0 + kOff[i];
            if (MN) v = (k < kEnd) ? v : 0u;
            else    v = (v && k < kEnd) ? min(4u, kEnd - k) : 0u;
            const uint32_t dst = panelAddr + soff[i];
#pragma unroll
            for (uint32_t e = 0; e < 4; e++)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst + 4 * e), "l"(e < v ? ptr[i] + e : base), "r"(e < v ? 4u : 0u) : "memory");
        }
    }
    __device__ __forceinline__ void issue(uint32_t panelAddr, uint32_t k0, uint32_t kEnd, int vec)
    {
        if (k0 + BK <= kEnd) issue_full(panelAddr, vec);
        else issue_tail(panelAddr, k0, kEnd);
    }
    // lo = a - trunc_tf32(a) for this thread's own chunks
    __device__ __forceinline__ void make_lo(uint8_t* raw, uint8_t* lo) const
    {
        float4 r[CHUNKS];
#pragma unroll
        for (int i = 0; i < CHUNKS; i++) r[i] = *reinterpret_cast<const float4*>(raw + soff[i]);
#pragma unroll
        for (int i = 0; i < CHUNKS; i++) {
            float4 l;
            l.x = r[i].x - __uint_as_float(__float_as_uint(r[i].x) & 0xFFFFE000u);
            l.y = r[i].y - __uint_as_float(__float_as_uint(r[i].y) & 0xFFFFE000u);
            l.z = r[i].z - __uint_as_float(__float_as_uint(r[i].z) & 0xFFFFE000u);
            l.w = r[i].w - __uint_as_float(__float_as_uint(r[i].w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(lo + soff[i]) = l;
        }
    }
};

// row loop of the epilogue: staging tile (2 columns per lane) -> global, ACT < 0 = raw copy (split-K partials)
template <int ACT>
__device__ __forceinline__ void store_rows(const float* __restrict__ sp, float* __restrict__ o, uint32_t rows, uint32_t ldo, uint32_t ncol, bool vec2,
                                           float alpha, float beta, float bias0, float bias1, float slope, float ealpha, float lambda)
{
#pragma unroll 4
    for (uint32_t r = 0; r < rows; r++, o += ldo, sp += EPI_LD) {
        const float2 t = *reinterpret_cast<const float2*>(sp);
        float x0 = t.x, x1 = t.y;
        if (ACT >= 0) {
            x0 = alpha * x0 + bias0; x1 = alpha * x1 + bias1;
            if (beta != 0.0f) { x0 += beta * o[0]; if (ncol > 1) x1 += beta * o[1]; }
            x0 = act_apply(ACT, x0, slope, ealpha, lambda); x1 = act_apply(ACT, x1, slope, ealpha, lambda);
        }
        if (vec2) *reinterpret_cast<float2*>(o) = make_float2(x0, x1);
        else { o[0] = x0; if (ncol > 1) o[1] = x1; }
    }
}

template <bool AMN, bool BMN>
__global__ void __launch_bounds__(THREADS, 2)
gemm_tc_kernel(const Args a)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t fullBar[MAX_STAGES], emptyBar[MAX_STAGES], accumBar;
    __shared__ uint32_t tmemBase;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const uint32_t kBegin = blockIdx.z * a.kPerSplit;
    const uint32_t kEnd = min(a.K, kBegin + a.kPerSplit);
    const uint32_t numK = (kEnd > kBegin) ? (kEnd - kBegin + BK - 1) / BK : 0;
    // copies run `depth` panels ahead of the panel being handed to the MMA warp; depth <= stages - 2 leaves slack on both
    // handshakes (a refilled slot was released two iterations ago, a full slot is waiting before the MMA warp asks)
    const uint32_t stages = a.stages, depth = a.depth;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < stages; s++) { mbar_init(&fullBar[s], LOADERS / 32); mbar_init(&emptyBar[s], 1); }
        mbar_init(&accumBar, 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&tmemBase, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;

    if (warp < MMA_WARP) {
        // ---------------------------------------------------------------- loaders
        const uint32_t smemAddr = smem_u32(smem);
        OperandPlan<AMN> pa;
        OperandPlan<BMN> pb;
        pa.init(a.A, a.lda, m0, kBegin, a.M, a.vecA, warp, lane);
        pb.init(a.B, a.ldb, n0, kBegin, a.N, a.vecB, warp, lane);
        uint32_t issued = 0;                                                     // k-iterations whose copies have been issued
        for (; issued < depth; issued++) {
            if (issued < numK && !(a.debug & 1)) {
                const uint32_t st = smemAddr + issued * STAGE_BYTES, k0 = kBegin + issued * BK;
                pa.issue(st, k0, kEnd, a.vecA);
                pb.issue(st + 2 * PANEL, k0, kEnd, a.vecB);
            }
            cp_async_commit();
        }
        uint32_t s = 0, sNext = depth % stages, phNext = 1;                      // slot of kt; slot / wait parity of kt + depth
        for (uint32_t kt = 0; kt < numK; kt++) {
            cp_async_wait((int)depth - 1);                                       // this thread's copies of stage kt have landed
            uint8_t* st = smem + s * STAGE_BYTES;
            if (a.passes == 3 && !(a.debug & 8)) {
                pa.make_lo(st, st + PANEL);
                pb.make_lo(st + 2 * PANEL, st + 3 * PANEL);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullBar[s]);                             // one arrival per loader warp
            if (++s == stages) s = 0;
            // refill the slot the MMAs of iteration kt - 1 read (they were issued one iteration ago: normally retired)
            if (issued < numK) {
                mbar_wait(&emptyBar[sNext], phNext);
                const uint32_t sn = smemAddr + sNext * STAGE_BYTES, k0 = kBegin + issued * BK;
                if (!(a.debug & 1)) {
                    pa.issue(sn, k0, kEnd, a.vecA);
                    pb.issue(sn + 2 * PANEL, k0, kEnd, a.vecB);
                }
            }
            issued++;
            if (++sNext == stages) { sNext = 0; phNext ^= 1; }
            cp_async_commit();
        }
        // ---------------------------------------------------------------- epilogue
        // TMEM -> registers (thread = row) -> per-warp padded staging tile in the now idle pipeline memory ->
        // row-contiguous global stores (one warp instruction writes 256 contiguous bytes of an output row).
        // warp w: accumulator rows 32*(w%4) .. +31 (the TMEM lanes a warp may read), columns 64*(w/4) .. +63
        if (numK) { mbar_wait(&accumBar, 0); tc_fence_after(); }
        const uint32_t rowBase = (warp & 3) * 32, colBase = (warp >> 2) * EPI_COLS;
        float* stage = reinterpret_cast<float*>(smem) + warp * (32 * EPI_LD);
#pragma unroll 1
        for (int cb = 0; cb < EPI_COLS / 32; cb++) {
            float v[32];
            if (numK) tmem_ld32(tmem + (rowBase << 16) + colBase + cb * 32, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stage + lane * EPI_LD + cb * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncwarp();
        const uint32_t mBase = m0 + rowBase;
        const uint32_t rows = (mBase < a.M) ? min(32u, a.M - mBase) : 0u;
        const uint32_t c0 = colBase + lane * 2, nc = n0 + c0;                     // this lane's 2 columns
        const uint32_t ncol = (nc < a.N) ? min(2u, a.N - nc) : 0u;
        float bias2[2] = {0.f, 0.f};
        if (a.bias && !a.partial) {
            if (ncol > 0) bias2[0] = __ldg(a.bias + nc);
            if (ncol > 1) bias2[1] = __ldg(a.bias + nc + 1);
        }
        const uint32_t ldo = a.partial ? a.N : a.ldc;
        float* outBase = a.partial ? a.partial + (size_t)blockIdx.z * a.M * a.N : a.C;
        const bool vec2 = ncol == 2 && (a.partial ? ((a.N & 1) == 0) : (a.vecC >= 2));
        if (ncol && !(a.debug & 4)) {
            float* o = outBase + (size_t)mBase * ldo + nc;
            const float* sp = stage + lane * 2;
            if (a.partial)                           store_rows<-1>(sp, o, rows, ldo, ncol, vec2, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
            else if (a.act == DSB200_ACT_LINEAR)     store_rows<DSB200_ACT_LINEAR>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], 0.f, 0.f, 0.f);
            else if (a.act == DSB200_ACT_SIGMOID)    store_rows<DSB200_ACT_SIGMOID>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], 0.f, 0.f, 0.f);
            else if (a.act == DSB200_ACT_TANH)       store_rows<DSB200_ACT_TANH>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], 0.f, 0.f, 0.f);
            else if (a.act == DSB200_ACT_RELU)       store_rows<DSB200_ACT_RELU>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], 0.f, 0.f, 0.f);
            else if (a.act == DSB200_ACT_LRELU)      store_rows<DSB200_ACT_LRELU>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], a.slope, 0.f, 0.f);
            else if (a.act == DSB200_ACT_ELU)        store_rows<DSB200_ACT_ELU>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], 0.f, a.ealpha, 0.f);
            else                                     store_rows<DSB200_ACT_SELU>(sp, o, rows, ldo, ncol, vec2, a.alpha, a.beta, bias2[0], bias2[1], 0.f, a.ealpha, a.lambda);
        }
    } else if (lane == 0) {
        // ---------------------------------------------------------------- MMA issuer
        // instruction descriptor: D = F32, A = B = TF32, M = 128, N = 128, majors from the template
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (uint32_t kt = 0; kt < numK; kt++) {
            const uint32_t s = kt % stages, ph = (kt / stages) & 1;
            mbar_wait(&fullBar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < BK / 8; j++) {
                const uint64_t aHi = panel_desc<AMN>(sa, j), aLo = panel_desc<AMN>(sa + PANEL, j);
                const uint64_t bHi = panel_desc<BMN>(sa + 2 * PANEL, j), bLo = panel_desc<BMN>(sa + 3 * PANEL, j);
                const uint32_t first = (kt == 0 && j == 0) ? 0u : 1u;
                if (a.debug & 2) {
                } else if (a.passes == 3) {
                    tc_mma_tf32(tmem, aLo, bHi, idesc, first);      // small terms first
                    tc_mma_tf32(tmem, aHi, bLo, idesc, 1u);
                    tc_mma_tf32(tmem, aHi, bHi, idesc, 1u);
                } else {
                    tc_mma_tf32(tmem, aHi, bHi, idesc, first);
                }
            }
            tc_commit(&emptyBar[s]);                                 // stage reusable once these MMAs have read it
        }
        if (numK) tc_commit(&accumBar);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, BN);
}

// out = alpha * sum_z partial[z] (+ beta * out) (+ bias) -> activation; fixed summation order
__global__ void __launch_bounds__(256)
gemm_reduce_kernel(const float* __restrict__ partial, uint32_t splits, uint32_t M, uint32_t N, uint32_t ldc, float alpha, float beta,
                   const float* __restrict__ bias, int act, float slope, float ealpha, float lambda, float* __restrict__ C)
{
    const size_t total = (size_t)M * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t m = (uint32_t)(i / N), n = (uint32_t)(i % N);
        float s = 0.f;
        for (uint32_t z = 0; z < splits; z++) s += partial[(size_t)z * total + i];
        float x = alpha * s;
        if (bias) x += bias[n];
        float* c = C + (size_t)m * ldc + n;
        if (beta != 0.0f) x += beta * *c;
        *c = act_apply(act, x, slope, ealpha, lambda);
    }
}

static int vec_of(const void* p, uint32_t ld)
{
    const uintptr_t u = (uintptr_t)p;
    if ((u & 15) == 0 && (ld & 3) == 0) return 4;
    if ((u & 7) == 0 && (ld & 1) == 0) return 2;
    return 1;
}

}  // namespace tc

// C[M][N] = act(alpha * op(A) * op(B) + beta * C + bias); see Args for the operand conventions.
int gemm_tc_launch(dsb200_ctx* ctx, const float* A, int aMN, uint32_t lda, const float* B, int bMN, uint32_t ldb, float* C, uint32_t ldc,
                   uint32_t M, uint32_t N, uint32_t K, float alpha, float beta, const float* bias, int act, float slope, float ealpha,
                   float lambda)
{
    using namespace tc;
    static bool attrSet = false;
    if (!attrSet) {
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAX_STAGES)));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAX_STAGES)));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAX_STAGES)));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAX_STAGES)));
        attrSet = true;
    }
    Args a;
    a.A = A; a.B = B; a.C = C; a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.aMN = aMN; a.bMN = bMN;
    a.vecA = vec_of(A, lda); a.vecB = vec_of(B, ldb); a.vecC = vec_of(C, ldc);
    a.alpha = alpha; a.beta = beta; a.bias = bias; a.act = act; a.slope = slope; a.ealpha = ealpha; a.lambda = lambda;
    a.passes = (ctx->gemmMode == DSB200_GEMM_TF32) ? 1 : 3;
    a.debug = ctx->gemmDebug;
    const uint32_t tilesM = (M + BM - 1) / BM, tilesN = (N + BN - 1) / BN, kTiles = (K + BK - 1) / BK;
    // split K when the tile grid cannot fill the machine (two CTAs per SM)
    uint32_t splits = 1;
    const uint32_t target = (uint32_t)ctx->numSMs * 2;
    if (tilesM * tilesN < target / 2 && kTiles >= 16) {
        splits = min(min((target + tilesM * tilesN - 1) / (tilesM * tilesN), kTiles / 8), 64u);
        if (splits < 1) splits = 1;
    }
    uint32_t kTilesPerSplit = (kTiles + splits - 1) / splits;
    splits = (kTiles + kTilesPerSplit - 1) / kTilesPerSplit;
    a.kPerSplit = kTilesPerSplit * BK;
    a.partial = nullptr;
    if (splits > 1) {
        const size_t need = (size_t)splits * M * N;
        if (need > ctx->gemmWsCap) {
            if (ctx->dGemmWs) { DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream)); DSB_CUDA_OK(cudaFree(ctx->dGemmWs)); ctx->dGemmWs = nullptr; ctx->gemmWsCap = 0; }
            DSB_CUDA_OK(cudaMalloc(&ctx->dGemmWs, need * sizeof(float)));
            ctx->gemmWsCap = need;
        }
        a.partial = ctx->dGemmWs;
    }
    // short K (the forward GEMMs): shallow ring, two CTAs per SM so one CTA's epilogue overlaps the other's loads;
    // long K (gradient GEMMs): one CTA per SM with a deep ring, more bytes in flight
    a.stages = (ctx->gemmStages >= 3 && ctx->gemmStages <= MAX_STAGES) ? (uint32_t)ctx->gemmStages : ((kTilesPerSplit <= 16) ? 3u : 6u);
    a.depth = (ctx->gemmDepth >= 1 && ctx->gemmDepth < (int)a.stages) ? (uint32_t)ctx->gemmDepth : (a.stages > 3 ? a.stages - 2 : a.stages - 1);
    const int SMEM_BYTES = smem_bytes((int)a.stages);
    dim3 grid(tilesN, tilesM, splits);
    if (aMN) {
        if (bMN) gemm_tc_kernel<true, true><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(a);
        else     gemm_tc_kernel<true, false><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(a);
    } else {
        if (bMN) gemm_tc_kernel<false, true><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(a);
        else     gemm_tc_kernel<false, false><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(a);
    }
    DSB_CUDA_OK(cudaGetLastError());
    count_launch();
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        const uint32_t blocks = (uint32_t)min((total + 255) / 256, (size_t)ctx->numSMs * 8);
        gemm_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(a.partial, splits, M, N, ldc, alpha, beta, bias, act, slope, ealpha, lambda, C);
        DSB_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    return 0;
}

}  // namespace dsb
