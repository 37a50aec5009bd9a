// gemm_tc.cu -- hand-written tcgen05 / TMEM GEMM for the dense layers (hot-path row a11), sm_100a only.
//
// Replaces cublasSgemm at E/NNLayer.cpp:1073 (forward), :2223 (weight gradient), :2274 (input delta) when
// gemm_mode is DSB200_GEMM_TF32X3 (fp32-grade: every fp32 operand is split a = hi + lo, hi = the 19 bits the tf32
// tensor core reads, lo = the exact fp32 remainder, and lo*hi + hi*lo + hi*hi accumulate in fp32 in TMEM) or
// DSB200_GEMM_TF32 (hi*hi only).
//
// Persistent kernels, one CTA per SM, 128 x 128 output tiles, 128 x 32 operand panels per k-iteration, 21 warps with
// separate roles (loaders / one MMA-issuing warp / four epilogue warps), two 128-column accumulators in tensor memory so
// that the epilogue of a tile overlaps the main loop of the next.  Operand rows are not 16-byte aligned in general
// (N = 27,278 floats), so TMA tensor maps do not apply; three operand paths exist (option "gemm_loader"), in the order
// they were written -- each was measured against the one before (profiles/r1c_gemm_smem_pipe.md):
//   gemm_tc_kernel      (0) cp.async straight into the UMMA layouts (8 copy warps), 8 split warps derive the lo panels
//                           in shared memory
//   gemm_tc_reg_kernel  (1) 16 loader warps: ld.global.nc -> registers -> hi / lo -> st.shared.v4 (no cp.async, no
//                           second pass over shared memory)
//   gemm_tc_ts_kernel   (2, default) the A operand never enters shared memory: 8 A-loader warps write hi / lo with
//                           tcgen05.st into a ring in TENSOR MEMORY behind the accumulators and the MMAs take A from
//                           there; 8 B-loader warps fill shared memory as in (1)
// The default (-1) is (2), except for a K-major A whose rows are more than 256 KB apart (see gemm_tc_launch).
//
// Shared-memory layouts of one 128 x 32 operand panel (16-byte chunks = 4 floats along the contiguous dimension):
//   K-major operand (k contiguous in memory): the standard 128-byte swizzle, rows of 32 floats,
//       chunk(mn, kc) at (mn/8)*1024 + (mn%8)*128 + ((kc ^ (mn%8))*16)                      layout 2, SBO = 1024
//   MN-major operand (mn contiguous in memory): the "128B swizzle, 32-byte base" layout, the only one the tensor core
//       accepts for transposed tf32 operands (every other layout type returns zeros -- measured, tools/umma_probe.cu):
//       atom = 32 mn x 32 k = 4,096 bytes, 4 k-rows x 128 bytes per 512-byte group, 32-byte granules XOR-ed with k%4,
//       chunk(k, mc) at (mc/8)*4096 + (k/4)*512 + (k%4)*128 + ((((mc%8)/2) ^ (k%4))*32) + (mc%2)*16   layout 1, LBO = 4096, SBO = 512
// Tensor-memory layout of the A ring (tools/umma_ts_probe.cu): lane = row m, one tf32 word per column, consecutive
// columns = consecutive k; one K = 8 MMA step advances the A address by 8 columns.
// Split-K (K = 27,278 for the input-delta GEMM of the output layer) writes raw partial tiles to a workspace that
// gemm_reduce_kernel sums in a fixed order -- deterministic, no float atomics.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "launch.h"

namespace dsb {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int PANEL = 128 * BK * 4;            // one operand panel: 16 KB
constexpr int STAGE = 2 * PANEL;               // A panel | B panel (raw ring and lo ring alike)
constexpr int RAW_SLOTS = 4, LO_SLOTS = 2;
constexpr int COPY_WARPS = 8, SPLIT_GROUPS = 2, SPLIT_WARPS = 8, MMA_WARP = 16, EPI_WARP0 = 17, EPI_WARPS = 4;
constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int EPI_COLS = 32;                   // the epilogue drains the accumulator 32 columns at a time
constexpr int EPI_LD = EPI_COLS + 4;           // padded row of the per-warp staging tile (floats)
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4;
constexpr int SMEM_BYTES = (RAW_SLOTS + LO_SLOTS) * STAGE + EPI_BYTES + 1024;

struct Args {
    const float* A; const float* B; float* C;
    uint32_t M, N, K;                          // C[M][N] (+)= A(M x K) * B(K x N)
    uint32_t lda, ldb, ldc;
    int vecA, vecB, vecC;                      // widest aligned access in floats (4, 2, 1)
    float alpha, beta;
    const float* bias; int act; float slope, ealpha, lambda;
    uint32_t tilesM, tilesN, splits;
    uint32_t kPerSplit;                        // multiple of BK
    float* partial;                            // split-K workspace [splits][M][N] or NULL
    int passes;                                // 3 = 3xTF32, 1 = TF32
    int debug;                                 // bring-up: 1 no proxy fence, 2 no copies, 4 no MMA, 8 no lo pass, 16 no stores
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
    // one K = 8 step: 32 bytes along the 128-byte swizzled rows (K-major) or two groups of 4 k-rows (MN-major)
    return MN ? smem_desc(panelAddr + kStep * 1024, 4096, 512, 1) : smem_desc(panelAddr + kStep * 32, 16, 1024, 2);
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
// ignores the rest (measured: feeding raw words gives bit-identical results to feeding masked words).
__device__ __forceinline__ void cp_async_landed(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Per-thread plan of one operand of one tile.  A 128 x 32 panel is 1,024 16-byte chunks; one warp-level copy moves 32
// of them (16 in the 8-byte mode), the 8 copy warps take them round robin: piece i of warp w is group g = i * 8 + w.
//   K-major  (rows = mn, 128 bytes each): a group is 4 whole rows (2 in the 8-byte mode) -> the copy touches 4 (2)
//            cache lines, the fewest possible, and writes whole swizzled 128-byte rows (no bank conflicts)
//   MN-major (rows = k, 512 bytes each):  a group is 4 k-rows x 128 bytes of one 32-column atom
// Everything that does not depend on the k-iteration is hoisted: one base pointer and a constant stride per piece.
template <bool MN>
struct OperandPlan {
    const float* ptr;                          // global address of piece 0 in the next k-iteration to issue
    const float* base;
    size_t       pieceStride;                  // floats between consecutive pieces
    size_t       step;                         // pointer advance per k-iteration
    uint32_t     soff, spiece;                 // byte offset of piece 0 inside the panel / between pieces
    uint32_t     validPieces;                  // K-major: pieces whose row is inside the matrix
    uint32_t     valid0, valid1, valid2, valid3;  // MN-major: floats inside the matrix for the 4 atoms (pieces)
    uint32_t     kOff;                         // first k of this thread's chunk inside the panel
    int          vec;

    __device__ __forceinline__ void init(const float* b, uint32_t ld, uint32_t mn0, uint32_t kBegin, uint32_t mnLimit, int v, uint32_t warp, uint32_t lane)
    {
        base = b; vec = v;
        step = MN ? (size_t)BK * ld : (size_t)BK;
        if (MN) {
            // piece i: atom i (32 columns), k-group = warp (4 rows), row in group = lane & 3, 16-byte chunk = lane >> 2
            const uint32_t k = warp * 4 + (lane & 3), mn = mn0 + (lane >> 2) * 4;
            kOff = k;
            soff = warp * 512 + (lane & 3) * 128 + (((lane >> 3) ^ (lane & 3)) * 32) + ((lane >> 2) & 1) * 16;
            spiece = 4096;
            pieceStride = 32;
            valid0 = mn < mnLimit ? min(4u, mnLimit - mn) : 0u;
            valid1 = mn + 32 < mnLimit ? min(4u, mnLimit - mn - 32) : 0u;
            valid2 = mn + 64 < mnLimit ? min(4u, mnLimit - mn - 64) : 0u;
            valid3 = mn + 96 < mnLimit ? min(4u, mnLimit - mn - 96) : 0u;
            validPieces = 4;
            ptr = b + (size_t)(kBegin + k) * ld + mn;                            // may point past the matrix: only dereferenced when valid
        } else if (v == 4) {
            // piece i: rows (i * 8 + warp) * 4 + (lane >> 3), 16-byte chunk lane & 7
            const uint32_t r = warp * 4 + (lane >> 3), c = lane & 7;
            kOff = c * 4;
            soff = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) * 16);
            spiece = 4096;                                                       // 32 rows further
            pieceStride = (size_t)32 * ld;
            const uint32_t row = mn0 + r;
            validPieces = row < mnLimit ? min(4u, (mnLimit - row + 31) / 32) : 0u;
            valid0 = valid1 = valid2 = valid3 = 4;
            ptr = b + (size_t)row * ld + kBegin + c * 4;
        } else {
            // 8-byte (or 4-byte) copies: piece i: rows (i * 8 + warp) * 2 + (lane >> 4), 8-byte half-chunk lane & 15
            const uint32_t r = warp * 2 + (lane >> 4), h = lane & 15, c = h >> 1;
            kOff = h * 2;
            soff = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) * 16) + (h & 1) * 8;
            spiece = 2048;                                                       // 16 rows further
            pieceStride = (size_t)16 * ld;
            const uint32_t row = mn0 + r;
            validPieces = row < mnLimit ? min(8u, (mnLimit - row + 15) / 16) : 0u;
            valid0 = valid1 = valid2 = valid3 = 2;
            ptr = b + (size_t)row * ld + kBegin + h * 2;
        }
    }
    __device__ __forceinline__ uint32_t mn_valid(int i) const { return i == 0 ? valid0 : i == 1 ? valid1 : i == 2 ? valid2 : valid3; }

    // Issues this thread's copies of the panel that starts at k0.  `full`: the whole panel lies inside [kBegin, kEnd).
    __device__ __forceinline__ void issue(uint32_t panelAddr, uint32_t k0, uint32_t kEnd, bool full)
    {
        if (MN) {
            const bool rowIn = full || (k0 + kOff < kEnd);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t dst = panelAddr + soff + i * spiece, v = rowIn ? mn_valid(i) : 0u;
                const float* p = v ? ptr + i * pieceStride : base;
                if (vec == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(p), "r"(v * 4) : "memory");
                else if (vec == 2) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(p), "r"(min(v, 2u) * 4) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst + 8), "l"(v > 2 ? p + 2 : base), "r"((v > 2 ? v - 2 : 0u) * 4) : "memory");
                } else {
#pragma unroll
                    for (uint32_t e = 0; e < 4; e++)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst + 4 * e), "l"(e < v ? p + e : base), "r"(e < v ? 4u : 0u) : "memory");
                }
            }
        } else if (vec == 4) {
            const uint32_t kv = full ? 4u : (k0 + kOff < kEnd ? min(4u, kEnd - k0 - kOff) : 0u);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t v = (uint32_t)i < validPieces ? kv : 0u;
                const float* p = v ? ptr + i * pieceStride : base;
                if (v == 4 || v == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(panelAddr + soff + i * spiece), "l"(p), "r"(v * 4) : "memory");
                else {                                                           // ragged end of K: rows stay 16-byte aligned, sizes are not
#pragma unroll
                    for (uint32_t e = 0; e < 4; e++)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(panelAddr + soff + i * spiece + 4 * e), "l"(e < v ? p + e : base),
                                     "r"(e < v ? 4u : 0u) : "memory");
                }
            }
        } else {
            const uint32_t kv = full ? 2u : (k0 + kOff < kEnd ? min(2u, kEnd - k0 - kOff) : 0u);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t v = (uint32_t)i < validPieces ? kv : 0u;
                const float* p = v ? ptr + i * pieceStride : base;
                const uint32_t dst = panelAddr + soff + i * spiece;
                if (vec == 2 && v != 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(p), "r"(v * 4) : "memory");
                else {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(p), "r"(v ? 4u : 0u) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst + 4), "l"(v > 1 ? p + 1 : base), "r"(v > 1 ? 4u : 0u) : "memory");
                }
            }
        }
        ptr += step;
    }
};

// row loop of the epilogue: staging tile (2 columns per lane) -> global, ACT < 0 = raw copy (split-K partials)
// `sp` is a shared-space address: the staging tile is reached through pointer arithmetic the compiler cannot trace back to
// shared memory, and generic LD / ST on it were the slowest instructions of the epilogue (ncu source page, round 1c)
template <int ACT>
__device__ __forceinline__ void store_rows(uint32_t sp, float* __restrict__ o, uint32_t r0, uint32_t rows, uint32_t ldo, uint32_t ncol,
                                           bool vec2, float alpha, float beta, float bias0, float bias1, float slope, float ealpha, float lambda)
{
    // four row pairs per trip: the four staging reads are issued back to back, then the four results are finished and stored
    for (uint32_t r = r0; r < rows; r += 8, o += 8 * (size_t)ldo, sp += 8 * EPI_LD * 4) {
        float x[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x[u][0] = x[u][1] = 0.0f;
            if (r + 2 * u < 32) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[u][0]), "=f"(x[u][1]) : "r"(sp + u * 2 * EPI_LD * 4));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (r + 2 * u >= rows) break;
            float* ou = o + (size_t)(2 * u) * ldo;
            float x0 = x[u][0], x1 = x[u][1];
            if (ACT >= 0) {
                x0 = alpha * x0 + bias0; x1 = alpha * x1 + bias1;
                if (beta != 0.0f) { x0 += beta * ou[0]; if (ncol > 1) x1 += beta * ou[1]; }
                x0 = act_apply(ACT, x0, slope, ealpha, lambda); x1 = act_apply(ACT, x1, slope, ealpha, lambda);
            }
            if (vec2) *reinterpret_cast<float2*>(ou) = make_float2(x0, x1);
            else { ou[0] = x0; if (ncol > 1) ou[1] = x1; }
        }
    }
}

struct Tile { uint32_t m0, n0, split, kBegin, kEnd, numK; };
__device__ __forceinline__ Tile tile_of(const Args& a, uint32_t t);

__device__ __forceinline__ Tile tile_of(const Args& a, uint32_t t)
{
    // m fastest: CTAs running at the same time share the B (weight / delta column) tile through L2
    Tile x;
    const uint32_t mn = a.tilesM * a.tilesN;
    x.split = t / mn;
    const uint32_t r = t - x.split * mn;
    x.n0 = (r / a.tilesM) * BN;
    x.m0 = (r % a.tilesM) * BM;
    x.kBegin = x.split * a.kPerSplit;
    x.kEnd = min(a.K, x.kBegin + a.kPerSplit);
    x.numK = (x.kEnd - x.kBegin + BK - 1) / BK;               // >= 1: the host never creates an empty split
    return x;
}

// ---------------------------------------------------------------- epilogue warps (shared by both kernels)
__device__ __forceinline__ void epilogue_role(const Args& a, uint32_t tmem, float* epiStage, uint64_t* accFullBar, uint64_t* accEmptyBar,
                                              uint32_t warp, uint32_t lane, uint32_t numTiles)
{
    // ---------------------------------------------------------------- epilogue warps
    // warp w may read TMEM lanes 32*(w%4) .. +31 = accumulator rows; thread = row
    const uint32_t rowBase = (warp & 3) * 32;
    const uint32_t stageAddr = smem_u32(epiStage + (warp - EPI_WARP0) * (32 * EPI_LD));
    uint32_t seq = 0;
    for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x, seq++) {
        const Tile tl = tile_of(a, t);
        const uint32_t acc = seq & 1;
        mbar_wait(&accFullBar[acc], (seq >> 1) & 1);
        tc_fence_after();
        const uint32_t mBase = tl.m0 + rowBase;
        const uint32_t rows = (mBase < a.M) ? min(32u, a.M - mBase) : 0u;
        const uint32_t ldo = a.partial ? a.N : a.ldc;
        float* outBase = a.partial ? a.partial + (size_t)tl.split * a.M * a.N : a.C;
#pragma unroll 1
        for (int half = 0; half < BN / EPI_COLS; half++) {
            {
                float v[32];
                tmem_ld32(tmem + (rowBase << 16) + acc * BN + half * EPI_COLS, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(stageAddr + (lane * EPI_LD + j) * 4), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            }
            if (half == BN / EPI_COLS - 1) {                                  // accumulator fully read: the MMA warp may reuse it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&accEmptyBar[acc]);
            }
            __syncwarp();
            // 32 columns per pass: lanes 0-15 take the even rows, lanes 16-31 the odd rows, 2 columns each
            const uint32_t c0 = half * EPI_COLS + (lane & 15) * 2, nc = tl.n0 + c0;
            const uint32_t ncol = (nc < a.N) ? min(2u, a.N - nc) : 0u;
            const uint32_t rsel = lane >> 4;
            if (ncol && !(a.debug & 16)) {
                float bias0 = 0.f, bias1 = 0.f;
                if (a.bias && !a.partial) {
                    bias0 = __ldg(a.bias + nc);
                    if (ncol > 1) bias1 = __ldg(a.bias + nc + 1);
                }
                const bool vec2 = ncol == 2 && (a.partial ? ((a.N & 1) == 0) : (a.vecC >= 2));
                float* o = outBase + (size_t)(mBase + rsel) * ldo + nc;
                const uint32_t sp = stageAddr + (rsel * EPI_LD + (lane & 15) * 2) * 4;
                if (a.partial)                        store_rows<-1>(sp, o, rsel, rows, ldo, ncol, vec2, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
                else if (a.act == DSB200_ACT_LINEAR)  store_rows<DSB200_ACT_LINEAR>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, 0.f, 0.f, 0.f);
                else if (a.act == DSB200_ACT_SIGMOID) store_rows<DSB200_ACT_SIGMOID>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, 0.f, 0.f, 0.f);
                else if (a.act == DSB200_ACT_TANH)    store_rows<DSB200_ACT_TANH>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, 0.f, 0.f, 0.f);
                else if (a.act == DSB200_ACT_RELU)    store_rows<DSB200_ACT_RELU>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, 0.f, 0.f, 0.f);
                else if (a.act == DSB200_ACT_LRELU)   store_rows<DSB200_ACT_LRELU>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, a.slope, 0.f, 0.f);
                else if (a.act == DSB200_ACT_ELU)     store_rows<DSB200_ACT_ELU>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, 0.f, a.ealpha, 0.f);
                else                                  store_rows<DSB200_ACT_SELU>(sp, o, rsel, rows, ldo, ncol, vec2, a.alpha, a.beta, bias0, bias1, 0.f, a.ealpha, a.lambda);
            }
            __syncwarp();                                                     // staging tile free for the next half
        }
    }
}

template <bool AMN, bool BMN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const Args a)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    uint8_t* rawRing = smem;
    uint8_t* loRing = smem + RAW_SLOTS * STAGE;
    float* epiStage = reinterpret_cast<float*>(smem + (RAW_SLOTS + LO_SLOTS) * STAGE);
    __shared__ uint64_t landedBar[RAW_SLOTS], emptyRawBar[RAW_SLOTS], fullBar[LO_SLOTS], emptyLoBar[LO_SLOTS], accFullBar[2], accEmptyBar[2];
    __shared__ uint32_t tmemBase;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t numTiles = a.tilesM * a.tilesN * a.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RAW_SLOTS; s++) { mbar_init(&landedBar[s], COPY_WARPS * 32); mbar_init(&emptyRawBar[s], 1); }
        for (int s = 0; s < LO_SLOTS; s++) { mbar_init(&fullBar[s], SPLIT_WARPS / SPLIT_GROUPS); mbar_init(&emptyLoBar[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&accFullBar[s], 1); mbar_init(&accEmptyBar[s], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&tmemBase, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;

    if (warp < COPY_WARPS) {
        // ---------------------------------------------------------------- copy warps
        const uint32_t rawAddr = smem_u32(rawRing);
        uint32_t slot = 0, parity = 1;                                            // ring position / emptyRaw wait parity
        for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x) {
            const Tile tl = tile_of(a, t);
            OperandPlan<AMN> pa;
            OperandPlan<BMN> pb;
            pa.init(a.A, a.lda, tl.m0, tl.kBegin, a.M, a.vecA, warp, lane);
            pb.init(a.B, a.ldb, tl.n0, tl.kBegin, a.N, a.vecB, warp, lane);
            for (uint32_t kt = 0; kt < tl.numK; kt++) {
                mbar_wait(&emptyRawBar[slot], parity);                            // the MMAs that read this slot have retired
                const uint32_t st = rawAddr + slot * STAGE, k0 = tl.kBegin + kt * BK;
                if (!(a.debug & 2)) {
                    const bool full = k0 + BK <= tl.kEnd;
                    pa.issue(st, k0, tl.kEnd, full);
                    pb.issue(st + PANEL, k0, tl.kEnd, full);
                }
                cp_async_landed(&landedBar[slot]);                                // arrives when this thread's copies have landed
                if (++slot == RAW_SLOTS) { slot = 0; parity ^= 1; }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp < MMA_WARP) {
        // ---------------------------------------------------------------- split warps (two groups, alternate iterations)
        const uint32_t group = (warp - COPY_WARPS) >> 2, tid = threadIdx.x - (COPY_WARPS + group * 4) * 32;   // 0..127 inside the group
        uint32_t it = 0, rs = 0, rph = 0, ls = 0, lph = 1;
        for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x) {
            const Tile tl = tile_of(a, t);
            for (uint32_t kt = 0; kt < tl.numK; kt++, it++) {
                if ((it & 1) == group) {
                    mbar_wait(&landedBar[rs], rph);                               // every copy thread's pieces of the panel are in
                    mbar_wait(&emptyLoBar[ls], lph);
                    if (a.passes == 3 && !(a.debug & 8)) {
                        const uint8_t* src = rawRing + rs * STAGE + tid * 16;
                        uint8_t* dst = loRing + ls * STAGE + tid * 16;
#pragma unroll
                        for (int half = 0; half < 2; half++) {
                        float4 r[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) r[j] = *reinterpret_cast<const float4*>(src + (half * 8 + j) * 2048);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 l;
                            l.x = r[j].x - __uint_as_float(__float_as_uint(r[j].x) & 0xFFFFE000u);
                            l.y = r[j].y - __uint_as_float(__float_as_uint(r[j].y) & 0xFFFFE000u);
                            l.z = r[j].z - __uint_as_float(__float_as_uint(r[j].z) & 0xFFFFE000u);
                            l.w = r[j].w - __uint_as_float(__float_as_uint(r[j].w) & 0xFFFFE000u);
                            *reinterpret_cast<float4*>(dst + (half * 8 + j) * 2048) = l;
                        }
                        }
                    }
                    if (!(a.debug & 1)) fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&fullBar[ls]);
                }
                if (++rs == RAW_SLOTS) { rs = 0; rph ^= 1; }
                if (++ls == LO_SLOTS) { ls = 0; lph ^= 1; }
            }
        }
    } else if (warp == MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, M = 128, N = 128, majors from the template
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t rawAddr = smem_u32(rawRing), loAddr = smem_u32(loRing);
            uint32_t rs = 0, ls = 0, lph = 0, seq = 0;
            for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x, seq++) {
                const Tile tl = tile_of(a, t);
                const uint32_t acc = seq & 1, d = tmem + acc * BN;
                mbar_wait(&accEmptyBar[acc], ((seq >> 1) & 1) ^ 1);               // the epilogue has drained this accumulator
                tc_fence_after();
                for (uint32_t kt = 0; kt < tl.numK; kt++) {
                    mbar_wait(&fullBar[ls], lph);
                    tc_fence_after();
                    const uint32_t ra = rawAddr + rs * STAGE, la = loAddr + ls * STAGE;
#pragma unroll
                    for (int j = 0; j < BK / 8; j++) {
                        const uint64_t aHi = panel_desc<AMN>(ra, j), bHi = panel_desc<BMN>(ra + PANEL, j);
                        const uint32_t first = (kt == 0 && j == 0) ? 0u : 1u;
                        if (a.debug & 4) {
                        } else if (a.passes == 3) {
                            const uint64_t aLo = panel_desc<AMN>(la, j), bLo = panel_desc<BMN>(la + PANEL, j);
                            tc_mma_tf32(d, aLo, bHi, idesc, first);               // small terms first
                            tc_mma_tf32(d, aHi, bLo, idesc, 1u);
                            tc_mma_tf32(d, aHi, bHi, idesc, 1u);
                        } else {
                            tc_mma_tf32(d, aHi, bHi, idesc, first);
                        }
                    }
                    tc_commit(&emptyRawBar[rs]);                                  // slots reusable once these MMAs have read them
                    tc_commit(&emptyLoBar[ls]);
                    if (++rs == RAW_SLOTS) rs = 0;
                    if (++ls == LO_SLOTS) { ls = 0; lph ^= 1; }
                }
                tc_commit(&accFullBar[acc]);
            }
        }
    } else {
        epilogue_role(a, tmem, epiStage, accFullBar, accEmptyBar, warp, lane, numTiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, 2 * BN);
}

// =====================================================================================================================
// Register-staged loader (option "gemm_loader" = 1).
// ncu on the cp.async kernel above (profiles/r1c_gemm_smem_pipe.md): the shared-memory data pipe is the bound -- one
// 128 x 128 x 8 tf32 MMA reads its operands at the pipe's full 128 bytes / clock, so every other wavefront competes with
// the tensor core -- and LDGSTS is the worst customer: its shared-memory writes land sector by sector, 4.2 x the
// wavefronts of the same bytes written with 128-bit stores (7.3 M of the 15 M LSU wavefronts of the forward GEMM),
// before the split warps read every panel again.  Here 16 loader warps bring the operands through REGISTERS instead:
//   ld.global.nc (128 / 64 / 32-bit by alignment, one k-iteration ahead)  ->  hi = raw words, lo = a - trunc_tf32(a)
//   ->  two conflict-free st.shared.v4 per 16-byte chunk, straight into the UMMA layouts of a 3-deep ring
// i.e. 2 wavefront-bytes per operand byte instead of ~6, no cp.async, no separate split pass.  MMA and epilogue roles
// are unchanged; a ring slot is raw A | raw B | lo A | lo B (64 KB).
constexpr int RSLOTS = 3, RSLOT_BYTES = 4 * PANEL, LOAD_WARPS = 16;
constexpr int RSMEM_BYTES = RSLOTS * RSLOT_BYTES + EPI_BYTES + 1024;
static_assert(LOAD_WARPS == MMA_WARP, "the MMA warp follows the loader warps");

__device__ __forceinline__ float4 ldg_nc4(const float* p)
{
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_nc2(const float* p)
{
    float2 r;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_nc1(const float* p)
{
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void sts4(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// one 16-byte chunk of an operand: v = floats inside the matrix (0..4), vec = widest aligned access of the operand
__device__ __forceinline__ float4 load_chunk(const float* p, uint32_t v, int vec)
{
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v == 0) return r;
    if (vec == 4 && v == 4) return ldg_nc4(p);
    if (vec >= 2) {
        if (v >= 2) { const float2 t = ldg_nc2(p); r.x = t.x; r.y = t.y; } else r.x = ldg_nc1(p);
        if (v == 4) { const float2 t = ldg_nc2(p + 2); r.z = t.x; r.w = t.y; } else if (v == 3) r.z = ldg_nc1(p + 2);
        return r;
    }
    r.x = ldg_nc1(p);
    if (v > 1) r.y = ldg_nc1(p + 1);
    if (v > 2) r.z = ldg_nc1(p + 2);
    if (v > 3) r.w = ldg_nc1(p + 3);
    return r;
}
__device__ __forceinline__ float lo_of(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Per-thread plan of one operand of one tile: a 128 x 32 panel is 1,024 16-byte chunks, two per loader thread.
//   K-major  (rows = mn, 128 bytes each): piece i = row i * 64 + warp * 4 + lane / 8, chunk lane % 8 -- a warp reads 4
//            whole 128-byte rows and writes 4 whole swizzled rows (a quarter-warp = one row: conflict free)
//   MN-major (rows = k, 512 bytes each):  piece i = atom (warp / 8) * 2 + i (32 columns), k-row (warp % 8) * 4 + lane % 4,
//            chunk lane / 4 -- a warp reads 4 k-rows x 128 bytes; a quarter-warp writes 4 rows x 32 bytes on 32 distinct banks
template <bool MN, int NW>
struct RegPlan {
    static constexpr int PIECES = 32 / NW;                    // 16-byte chunks per thread and panel: 2 with 16 loader warps, 4 with 8
    static constexpr uint32_t SPIECE = MN ? 4096u : 512u * NW;   // shared-memory bytes between consecutive pieces
    const float* ptr;                          // global address of piece 0 in the next k-iteration to load
    size_t       pieceStride, step;            // floats between the pieces / pointer advance per k-iteration
    uint32_t     soff;                         // byte offset of piece 0 inside a panel (the same for every tile)
    uint32_t     v[PIECES];                    // floats of the chunk inside the matrix along mn (K-major: 4 or 0 by row)
    uint32_t     kOff;                         // first k of this thread's chunk inside the panel
    bool         inside;                       // every chunk of this thread lies inside the matrix along mn (interior tile)

    // `warp` = index inside the loader group, 0 .. NW-1
    __device__ __forceinline__ void init(const float* b, uint32_t ld, uint32_t mn0, uint32_t kBegin, uint32_t mnLimit, uint32_t warp, uint32_t lane)
    {
        if (MN) {
            const uint32_t kg = warp & 7, atom0 = (warp >> 3) * PIECES;
            const uint32_t k = kg * 4 + (lane & 3), mn = mn0 + atom0 * 32 + (lane >> 2) * 4;
            kOff = k;
            soff = atom0 * 4096 + kg * 512 + (lane & 3) * 128 + (((lane >> 3) ^ (lane & 3)) * 32) + ((lane >> 2) & 1) * 16;
            pieceStride = 32;
            step = (size_t)BK * ld;
#pragma unroll
            for (int i = 0; i < PIECES; i++) v[i] = mn + 32 * i < mnLimit ? min(4u, mnLimit - mn - 32 * i) : 0u;
            ptr = b + (size_t)(kBegin + k) * ld + mn;                            // may point past the matrix: only dereferenced when valid
        } else {
            const uint32_t r = warp * 4 + (lane >> 3), c = lane & 7, row = mn0 + r;
            kOff = c * 4;
            soff = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) * 16);
            pieceStride = (size_t)(4 * NW) * ld;
            step = BK;
#pragma unroll
            for (int i = 0; i < PIECES; i++) v[i] = row + 4 * NW * i < mnLimit ? 4u : 0u;
            ptr = b + (size_t)row * ld + kBegin + c * 4;
        }
        inside = true;
#pragma unroll
        for (int i = 0; i < PIECES; i++) inside = inside && v[i] == 4u;
    }
    // this thread's chunks of the panel that starts at k0; `full`: the whole panel lies inside [kBegin, kEnd)
    __device__ __forceinline__ void load(float4 (&r)[PIECES], uint32_t k0, uint32_t kEnd, bool full, int vec)
    {
        if (full && inside) {
            // interior panel (all but the edge tiles and the ragged last k-iteration): straight-line loads, one branch on the
            // alignment class -- the general path below costs ~10 instructions and 3 branches per chunk, and the loader
            // warps' serial instruction stream is what paces a k-iteration (ncu source page, round 1c)
            if (vec == 4) {
#pragma unroll
                for (int i = 0; i < PIECES; i++) r[i] = ldg_nc4(ptr + i * pieceStride);
            } else if (vec == 2) {
#pragma unroll
                for (int i = 0; i < PIECES; i++) {
                    const float2 t0 = ldg_nc2(ptr + i * pieceStride), t1 = ldg_nc2(ptr + i * pieceStride + 2);
                    r[i] = make_float4(t0.x, t0.y, t1.x, t1.y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < PIECES; i++) {
                    const float* p = ptr + i * pieceStride;
                    r[i] = make_float4(ldg_nc1(p), ldg_nc1(p + 1), ldg_nc1(p + 2), ldg_nc1(p + 3));
                }
            }
            ptr += step;
            return;
        }
        if (MN) {
            const bool rowIn = full || (k0 + kOff < kEnd);
#pragma unroll
            for (int i = 0; i < PIECES; i++) r[i] = load_chunk(ptr + i * pieceStride, rowIn ? v[i] : 0u, vec);
        } else {
            const uint32_t kv = full ? 4u : (k0 + kOff < kEnd ? min(4u, kEnd - k0 - kOff) : 0u);
#pragma unroll
            for (int i = 0; i < PIECES; i++) r[i] = load_chunk(ptr + i * pieceStride, v[i] ? kv : 0u, vec);
        }
        ptr += step;
    }
    __device__ __forceinline__ void store(uint32_t rawPanel, uint32_t loPanel, const float4 (&r)[PIECES], bool lo) const
    {
#pragma unroll
        for (int i = 0; i < PIECES; i++) {
            sts4(rawPanel + soff + i * SPIECE, r[i]);
            if (lo) sts4(loPanel + soff + i * SPIECE, make_float4(lo_of(r[i].x), lo_of(r[i].y), lo_of(r[i].z), lo_of(r[i].w)));
        }
    }
};

template <bool AMN, bool BMN, int DEPTH>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_reg_kernel(const Args a)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    float* epiStage = reinterpret_cast<float*>(smem + RSLOTS * RSLOT_BYTES);
    __shared__ uint64_t fullBar[RSLOTS], emptyBar[RSLOTS], accFullBar[2], accEmptyBar[2];
    __shared__ uint32_t tmemBase;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t numTiles = a.tilesM * a.tilesN * a.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RSLOTS; s++) { mbar_init(&fullBar[s], LOAD_WARPS); mbar_init(&emptyBar[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&accFullBar[s], 1); mbar_init(&accEmptyBar[s], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&tmemBase, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;
    const uint32_t ringAddr = smem_u32(smem);

    if (warp < LOAD_WARPS) {
        // ---------------------------------------------------------------- loader warps
        // The (tile, k-iteration) sequence of this CTA is walked by a load cursor that runs one iteration ahead of the
        // stores: the loads of iteration i + 1 are in flight while iteration i waits for its ring slot.
        RegPlan<AMN, LOAD_WARPS> pa;
        RegPlan<BMN, LOAD_WARPS> pb;
        uint32_t lt = blockIdx.x, lkt = 0, slot = 0, parity = 1;
        Tile ltl = tile_of(a, min(lt, numTiles - 1));
        bool more = lt < numTiles;
        pa.init(a.A, a.lda, ltl.m0, ltl.kBegin, a.M, warp, lane);
        pb.init(a.B, a.ldb, ltl.n0, ltl.kBegin, a.N, warp, lane);
        const bool lo = a.passes == 3 && !(a.debug & 8);
        auto fetch = [&](float4 (&ra)[2], float4 (&rb)[2]) {
            const uint32_t k0 = ltl.kBegin + lkt * BK;
            const bool full = k0 + BK <= ltl.kEnd;
            const uint32_t kLim = (a.debug & 32) ? k0 : ltl.kEnd;                  // bring-up switch 32: no global loads (zero panels)
            pa.load(ra, k0, kLim, full && !(a.debug & 32), a.vecA);
            pb.load(rb, k0, kLim, full && !(a.debug & 32), a.vecB);
            if (++lkt == ltl.numK) {
                lkt = 0; lt += gridDim.x; more = lt < numTiles;
                if (more) {
                    ltl = tile_of(a, lt);
                    pa.init(a.A, a.lda, ltl.m0, ltl.kBegin, a.M, warp, lane);
                    pb.init(a.B, a.ldb, ltl.n0, ltl.kBegin, a.N, warp, lane);
                }
            }
        };
        auto publish = [&](const float4 (&ra)[2], const float4 (&rb)[2]) {
            mbar_wait(&emptyBar[slot], parity);                                   // the MMAs that read this slot have retired
            const uint32_t st = ringAddr + slot * RSLOT_BYTES;
            pa.store(st, st + 2 * PANEL, ra, lo);
            pb.store(st + PANEL, st + 3 * PANEL, rb, lo);
            if (!(a.debug & 1)) fence_async_smem();                               // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullBar[slot]);
            if (++slot == RSLOTS) { slot = 0; parity ^= 1; }
        };
        float4 ra[DEPTH][2], rb[DEPTH][2];
        uint32_t pending = 0;
#pragma unroll
        for (int s = 0; s < DEPTH - 1; s++)
            if (more) { fetch(ra[s], rb[s]); pending++; }
        while (pending) {
#pragma unroll
            for (int s = 0; s < DEPTH; s++) {
                if (more) { fetch(ra[(s + DEPTH - 1) % DEPTH], rb[(s + DEPTH - 1) % DEPTH]); pending++; }
                publish(ra[s], rb[s]);
                if (--pending == 0) break;
            }
        }
    } else if (warp == MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t slot = 0, ph = 0, seq = 0;
            for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x, seq++) {
                const Tile tl = tile_of(a, t);
                const uint32_t acc = seq & 1, d = tmem + acc * BN;
                mbar_wait(&accEmptyBar[acc], ((seq >> 1) & 1) ^ 1);               // the epilogue has drained this accumulator
                tc_fence_after();
                for (uint32_t kt = 0; kt < tl.numK; kt++) {
                    mbar_wait(&fullBar[slot], ph);
                    tc_fence_after();
                    const uint32_t ra = ringAddr + slot * RSLOT_BYTES, la = ra + 2 * PANEL;
#pragma unroll
                    for (int j = 0; j < BK / 8; j++) {
                        const uint64_t aHi = panel_desc<AMN>(ra, j), bHi = panel_desc<BMN>(ra + PANEL, j);
                        const uint32_t first = (kt == 0 && j == 0) ? 0u : 1u;
                        if (a.debug & 4) {
                        } else if (a.passes == 3) {
                            const uint64_t aLo = panel_desc<AMN>(la, j), bLo = panel_desc<BMN>(la + PANEL, j);
                            tc_mma_tf32(d, aLo, bHi, idesc, first);               // small terms first
                            tc_mma_tf32(d, aHi, bLo, idesc, 1u);
                            tc_mma_tf32(d, aHi, bHi, idesc, 1u);
                        } else {
                            tc_mma_tf32(d, aHi, bHi, idesc, first);
                        }
                    }
                    tc_commit(&emptyBar[slot]);                                   // slot reusable once these MMAs have read it
                    if (++slot == RSLOTS) { slot = 0; ph ^= 1; }
                }
                tc_commit(&accFullBar[acc]);
            }
        }
    } else {
        epilogue_role(a, tmem, epiStage, accFullBar, accEmptyBar, warp, lane, numTiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, 2 * BN);
}

// =====================================================================================================================
// A operand through TENSOR MEMORY (option "gemm_loader" = 2, and the default -1).
// The register loader above still pushes 224 KB through the L1 / shared-memory data array per 128 x 128 x 32 k-iteration
// (fill, read to registers, raw + lo stores, 96 KB of MMA operand reads) and that array is the bound
// (profiles/r1c_gemm_smem_pipe.md).  tcgen05.mma accepts A from tensor memory (tools/umma_ts_probe.cu: lane = row m,
// one tf32 per column, consecutive columns = consecutive k), so here A never touches shared memory:
//   warps 0-7    A loaders: thread = (row m, 16 of the 32 k of the panel): ld.global.nc -> registers -> hi / lo split ->
//                tcgen05.st.32x32b.x16 into a 4-deep ring of (32 hi | 32 lo) columns behind the two accumulators
//   warps 8-15   B loaders: the register loader of gemm_tc_reg_kernel for the B panel only (raw | lo, 32 KB per slot)
//   warp 16      MMA: aLo*bHi + aHi*bLo + aHi*bHi with A addressed in TMEM, B by shared-memory descriptor
//   warps 17-20  epilogue (unchanged)
// Data-array bytes per k-iteration: A 16 (fill) + 16 (read); B 16 + 16 + 32 (stores); MMA 48 (B only) = 144 KB.
constexpr int TS_SLOTS = 4, TS_SLOT_BYTES = 2 * PANEL, TS_A_WARPS = 8, TS_B_WARPS = 8, TS_B_DEPTH = 3;
constexpr int TS_SMEM_BYTES = TS_SLOTS * TS_SLOT_BYTES + EPI_BYTES + 1024;
constexpr uint32_t TS_ACC_COLS = 2 * BN, TS_A_COLS = 2 * BK, TS_TMEM_COLS = 512;
static_assert(TS_ACC_COLS + TS_SLOTS * TS_A_COLS <= TS_TMEM_COLS, "tensor memory budget");
static_assert(TS_A_WARPS + TS_B_WARPS == MMA_WARP, "the MMA warp follows the loader warps");

__device__ __forceinline__ uint32_t ldg_nc1u(const float* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" :: "r"(tmemD), "r"(tmemA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}

// one thread's 16 consecutive k of one row of the A panel
template <bool MN>
struct TmemAPlan {
    const float* ptr;                          // (row, first k of this thread) in the next k-iteration to load
    size_t       step, ld;
    uint32_t     kOff;
    bool         rowIn;
    __device__ __forceinline__ void init(const float* A, uint32_t lda, uint32_t m0, uint32_t kBegin, uint32_t M, uint32_t quadrant, uint32_t khalf, uint32_t lane)
    {
        const uint32_t m = m0 + quadrant * 32 + lane;
        rowIn = m < M; kOff = khalf * 16; ld = lda;
        if (MN) { ptr = A + (size_t)(kBegin + kOff) * lda + m; step = (size_t)BK * lda; }    // lanes = consecutive m: coalesced
        else    { ptr = A + (size_t)m * lda + kBegin + kOff;   step = BK; }                   // a thread reads 64 contiguous bytes of its row
    }
    __device__ __forceinline__ void load(uint32_t (&r)[16], uint32_t k0, uint32_t kEnd, int vec)
    {
        const uint32_t kb = k0 + kOff;
        const uint32_t nv = (rowIn && kb < kEnd) ? min(16u, kEnd - kb) : 0u;
        if (!MN && nv == 16 && vec == 4) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 t = ldg_nc4(ptr + 4 * i);
                r[4 * i] = __float_as_uint(t.x); r[4 * i + 1] = __float_as_uint(t.y); r[4 * i + 2] = __float_as_uint(t.z); r[4 * i + 3] = __float_as_uint(t.w);
            }
        } else if (!MN && nv == 16 && vec == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 t = ldg_nc2(ptr + 2 * i);
                r[2 * i] = __float_as_uint(t.x); r[2 * i + 1] = __float_as_uint(t.y);
            }
        } else if (nv == 16) {
#pragma unroll
            for (int e = 0; e < 16; e++) r[e] = ldg_nc1u(MN ? ptr + e * ld : ptr + e);
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) r[e] = (uint32_t)e < nv ? ldg_nc1u(MN ? ptr + e * ld : ptr + e) : 0u;
        }
        ptr += step;
    }
};

// ---- option "gemm_loader" = 3 (EXPERIMENTAL, written at the end of round 1 without GPU time left: not yet run) -------------------
// K-major A read COALESCED for the tcgen05.st.16x256b shape, whose register mapping was probed on the B200
// (tools/umma_st16x256_probe.cu):  .x2 register s of thread t -> TMEM lane t / 4 + 8 * ((s >> 1) & 1), column 2 * (t % 4) + 8 * (s >> 2)
// + (s & 1); the instruction covers 16 lanes, lanes 16-31 of the quadrant take a second store at address + (16 << 16).
// Four consecutive threads therefore hold one 32-byte sector of a row and a load instruction touches 8 rows (8 pages when
// rows are far apart) instead of the 32 of the row-per-thread plan above -- the C4 input-delta case of gemm_tc_launch.
// r[8 j + s]: row t / 4 + 8 * (2 j + ((s >> 1) & 1)) of the quadrant, k = 16 khalf + 2 (t % 4) + 8 (s >> 2) + (s & 1).
__device__ __forceinline__ void tmem_st16x256_x2(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
struct TmemAPlanCoal {
    const float* ptr;                          // (row m0 + 32 q + t / 4, k = kBegin + 16 khalf + 2 (t % 4)) in the next k-iteration to load
    size_t       ld8;                          // floats between row groups (8 rows)
    uint32_t     kOff, rowMask;                // first k of the thread inside the panel / bit g: row t / 4 + 8 g is inside the matrix
    __device__ __forceinline__ void init(const float* A, uint32_t lda, uint32_t m0, uint32_t kBegin, uint32_t M, uint32_t quadrant, uint32_t khalf, uint32_t lane)
    {
        const uint32_t m = m0 + quadrant * 32 + (lane >> 2);
        kOff = khalf * 16 + 2 * (lane & 3);
        ld8 = (size_t)8 * lda;
        rowMask = 0;
#pragma unroll
        for (uint32_t g = 0; g < 4; g++) if (m + 8 * g < M) rowMask |= 1u << g;
        ptr = A + (size_t)m * lda + kBegin + kOff;                               // may point past the matrix: only dereferenced when valid
    }
    __device__ __forceinline__ void load(uint32_t (&r)[16], uint32_t k0, uint32_t kEnd, int vec)
    {
        if (rowMask == 15u && k0 + BK <= kEnd && vec >= 2) {                     // interior panel: eight straight 64-bit loads
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int s = 0; s < 8; s += 2) {
                    const float2 t = ldg_nc2(ptr + (2 * j + ((s >> 1) & 1)) * ld8 + 8 * (s >> 2));
                    r[8 * j + s] = __float_as_uint(t.x); r[8 * j + s + 1] = __float_as_uint(t.y);
                }
        } else {
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int s = 0; s < 8; s += 2) {
                    const uint32_t g = 2 * j + ((s >> 1) & 1), kk = k0 + kOff + 8 * (s >> 2);
                    const float* p = ptr + g * ld8 + 8 * (s >> 2);
                    const uint32_t n = (((rowMask >> g) & 1u) && kk < kEnd) ? min(2u, kEnd - kk) : 0u;
                    r[8 * j + s] = n > 0 ? ldg_nc1u(p) : 0u;
                    r[8 * j + s + 1] = n > 1 ? ldg_nc1u(p + 1) : 0u;
                }
        }
        ptr += BK;
    }
};

template <bool AMN, bool BMN, bool COAL = false>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_ts_kernel(const Args a)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    float* epiStage = reinterpret_cast<float*>(smem + TS_SLOTS * TS_SLOT_BYTES);
    __shared__ uint64_t fullBar[TS_SLOTS], emptyBar[TS_SLOTS], accFullBar[2], accEmptyBar[2];
    __shared__ uint32_t tmemBase;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t numTiles = a.tilesM * a.tilesN * a.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_SLOTS; s++) { mbar_init(&fullBar[s], TS_A_WARPS + TS_B_WARPS); mbar_init(&emptyBar[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&accFullBar[s], 1); mbar_init(&accEmptyBar[s], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&tmemBase, TS_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;
    const uint32_t ringAddr = smem_u32(smem);
    const bool lo = a.passes == 3 && !(a.debug & 8);

    if (warp < TS_A_WARPS) {
        // ---------------------------------------------------------------- A loaders: global -> registers -> tensor memory
        const uint32_t quadrant = warp & 3, khalf = warp >> 2;                    // a warp may only touch TMEM lanes 32 * (warp % 4) ..
        constexpr bool COALESCED = COAL && !AMN;                                  // MN-major A is coalesced in the row-per-thread plan already
        typename std::conditional<COALESCED, TmemAPlanCoal, TmemAPlan<AMN>>::type pa;
        uint32_t lt = blockIdx.x, lkt = 0, slot = 0, parity = 1;
        Tile ltl = tile_of(a, min(lt, numTiles - 1));
        bool more = lt < numTiles;
        pa.init(a.A, a.lda, ltl.m0, ltl.kBegin, a.M, quadrant, khalf, lane);
        auto fetch = [&](uint32_t (&r)[16]) {
            const uint32_t k0 = ltl.kBegin + lkt * BK;
            pa.load(r, k0, (a.debug & 32) ? k0 : ltl.kEnd, a.vecA);
            if (++lkt == ltl.numK) {
                lkt = 0; lt += gridDim.x; more = lt < numTiles;
                if (more) { ltl = tile_of(a, lt); pa.init(a.A, a.lda, ltl.m0, ltl.kBegin, a.M, quadrant, khalf, lane); }
            }
        };
        auto publish = [&](uint32_t (&r)[16]) {
            mbar_wait(&emptyBar[slot], parity);                                   // the MMAs that read this slot have retired
            tc_fence_after();
            const uint32_t ta = tmem + ((quadrant * 32) << 16) + TS_ACC_COLS + slot * TS_A_COLS + khalf * 16;
            uint32_t hi[16];
#pragma unroll
            for (int e = 0; e < 16; e++) hi[e] = r[e] & 0xFFFFE000u;
            if (COALESCED) { tmem_st16x256_x2(ta, hi); tmem_st16x256_x2(ta + (16u << 16), hi + 8); }
            else tmem_st16(ta, hi);
            if (lo) {
#pragma unroll
                for (int e = 0; e < 16; e++) hi[e] = __float_as_uint(__uint_as_float(r[e]) - __uint_as_float(hi[e]));
                if (COALESCED) { tmem_st16x256_x2(ta + BK, hi); tmem_st16x256_x2(ta + BK + (16u << 16), hi + 8); }
                else tmem_st16(ta + BK, hi);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullBar[slot]);
            if (++slot == TS_SLOTS) { slot = 0; parity ^= 1; }
        };
        uint32_t r0[16], r1[16];
        bool have0 = more;
        if (have0) fetch(r0);
        while (have0) {
            const bool have1 = more;
            if (have1) fetch(r1);
            publish(r0);
            if (!have1) break;
            have0 = more;
            if (have0) fetch(r0);
            publish(r1);
        }
    } else if (warp < MMA_WARP) {
        // ---------------------------------------------------------------- B loaders: global -> registers -> shared memory
        const uint32_t bw = warp - TS_A_WARPS;
        RegPlan<BMN, TS_B_WARPS> pb;
        constexpr int PB = RegPlan<BMN, TS_B_WARPS>::PIECES;
        uint32_t lt = blockIdx.x, lkt = 0, slot = 0, parity = 1;
        Tile ltl = tile_of(a, min(lt, numTiles - 1));
        bool more = lt < numTiles;
        pb.init(a.B, a.ldb, ltl.n0, ltl.kBegin, a.N, bw, lane);
        auto fetch = [&](float4 (&rb)[PB]) {
            const uint32_t k0 = ltl.kBegin + lkt * BK;
            const bool full = k0 + BK <= ltl.kEnd && !(a.debug & 32);
            pb.load(rb, k0, (a.debug & 32) ? k0 : ltl.kEnd, full, a.vecB);
            if (++lkt == ltl.numK) {
                lkt = 0; lt += gridDim.x; more = lt < numTiles;
                if (more) { ltl = tile_of(a, lt); pb.init(a.B, a.ldb, ltl.n0, ltl.kBegin, a.N, bw, lane); }
            }
        };
        auto publish = [&](const float4 (&rb)[PB]) {
            mbar_wait(&emptyBar[slot], parity);
            const uint32_t st = ringAddr + slot * TS_SLOT_BYTES;
            pb.store(st, st + PANEL, rb, lo);
            if (!(a.debug & 1)) fence_async_smem();                               // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullBar[slot]);
            if (++slot == TS_SLOTS) { slot = 0; parity ^= 1; }
        };
        // loads run TS_B_DEPTH - 1 k-iterations ahead of the stores (the first use of a loaded register was the B loaders'
        // largest single stall with one iteration of prefetch)
        float4 rb[TS_B_DEPTH][PB];
        uint32_t pending = 0;
#pragma unroll
        for (int s = 0; s < TS_B_DEPTH - 1; s++)
            if (more) { fetch(rb[s]); pending++; }
        while (pending) {
#pragma unroll
            for (int s = 0; s < TS_B_DEPTH; s++) {
                if (more) { fetch(rb[(s + TS_B_DEPTH - 1) % TS_B_DEPTH]); pending++; }
                publish(rb[s]);
                if (--pending == 0) break;
            }
        }
    } else if (warp == MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            // D = F32, A = B = TF32, A from tensor memory (K-major by construction), B major from the template
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((BMN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t slot = 0, ph = 0, seq = 0;
            for (uint32_t t = blockIdx.x; t < numTiles; t += gridDim.x, seq++) {
                const Tile tl = tile_of(a, t);
                const uint32_t acc = seq & 1, d = tmem + acc * BN;
                mbar_wait(&accEmptyBar[acc], ((seq >> 1) & 1) ^ 1);               // the epilogue has drained this accumulator
                tc_fence_after();
                for (uint32_t kt = 0; kt < tl.numK; kt++) {
                    mbar_wait(&fullBar[slot], ph);
                    tc_fence_after();
                    const uint32_t sb = ringAddr + slot * TS_SLOT_BYTES, ta = tmem + TS_ACC_COLS + slot * TS_A_COLS;
#pragma unroll
                    for (int j = 0; j < BK / 8; j++) {
                        const uint64_t bHi = panel_desc<BMN>(sb, j);
                        const uint32_t aHi = ta + j * 8, first = (kt == 0 && j == 0) ? 0u : 1u;
                        if (a.debug & 4) {
                        } else if (a.passes == 3) {
                            const uint64_t bLo = panel_desc<BMN>(sb + PANEL, j);
                            tc_mma_tf32_ts(d, aHi + BK, bHi, idesc, first);       // small terms first
                            tc_mma_tf32_ts(d, aHi, bLo, idesc, 1u);
                            tc_mma_tf32_ts(d, aHi, bHi, idesc, 1u);
                        } else {
                            tc_mma_tf32_ts(d, aHi, bHi, idesc, first);
                        }
                    }
                    tc_commit(&emptyBar[slot]);                                   // slot (shared and tensor memory) reusable once these MMAs retire
                    if (++slot == TS_SLOTS) { slot = 0; ph ^= 1; }
                }
                tc_commit(&accFullBar[acc]);
            }
        }
    } else {
        epilogue_role(a, tmem, epiStage, accFullBar, accEmptyBar, warp, lane, numTiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, TS_TMEM_COLS);
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

// C[M][N] = act(alpha * op(A) * op(B) + beta * C + bias); aMN / bMN: the operand is MN-contiguous in memory
int gemm_tc_launch(dsb200_ctx* ctx, const float* A, int aMN, uint32_t lda, const float* B, int bMN, uint32_t ldb, float* C, uint32_t ldc,
                   uint32_t M, uint32_t N, uint32_t K, float alpha, float beta, const float* bias, int act, float slope, float ealpha,
                   float lambda)
{
    using namespace tc;
    static bool attrSet = false;
    if (!attrSet) {
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_reg_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_reg_kernel<false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_reg_kernel<true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_reg_kernel<true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_ts_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_ts_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_ts_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_ts_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_ts_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        DSB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_ts_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        attrSet = true;
    }
    Args a;
    a.A = A; a.B = B; a.C = C; a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldb = ldb; a.ldc = ldc;
    a.vecA = vec_of(A, lda); a.vecB = vec_of(B, ldb); a.vecC = vec_of(C, ldc);
    a.alpha = alpha; a.beta = beta; a.bias = bias; a.act = act; a.slope = slope; a.ealpha = ealpha; a.lambda = lambda;
    a.passes = (ctx->gemmMode == DSB200_GEMM_TF32) ? 1 : 3;
    a.debug = ctx->gemmDebug;
    a.tilesM = (M + BM - 1) / BM; a.tilesN = (N + BN - 1) / BN;
    const uint32_t tilesMN = a.tilesM * a.tilesN, kTiles = (K + BK - 1) / BK, sms = (uint32_t)ctx->numSMs;
    // split K so that the persistent grid has whole rounds of work: minimise rounds * (k-iterations + fixed tile cost)
    uint32_t splits = 1;
    if (ctx->gemmSplits > 0) splits = min((uint32_t)ctx->gemmSplits, kTiles);
    else {
        uint64_t best = ~0ull;
        for (uint32_t s = 1; s <= 64 && s * 8 <= max(kTiles, 8u); s++) {
            const uint32_t per = (kTiles + s - 1) / s, real = (kTiles + per - 1) / per;
            const uint64_t rounds = ((uint64_t)tilesMN * real + sms - 1) / sms;
            const uint64_t cost = rounds * (per + 12) + (real > 1 ? (uint64_t)real * tilesMN / sms + 4 : 0);   // + partial write / reduce traffic
            if (cost < best) { best = cost; splits = real; }
        }
    }
    uint32_t kTilesPerSplit = (kTiles + splits - 1) / splits;
    splits = (kTiles + kTilesPerSplit - 1) / kTilesPerSplit;
    a.splits = splits;
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
    const uint32_t grid = min(sms, tilesMN * splits);
    // operand path (option "gemm_loader"): measured on the three output-layer GEMMs of BASELINE config 2
    // (profiles/r1c_logs/gemm_debug_matrix2.log, us per launch: forward / weight gradient / input delta)
    //   0 cp.async + split warps      92.5 / 96.6 / 85.1
    //   1 register loader             82.4 / 80.3 / 90.6
    //   2 A through tensor memory     81.0 / 67.7 / 65.5     <- default (-1)
    //   exception: a K-major A whose rows are far apart (the 1M-column delta of BASELINE config 4, 4 MB between rows): the
    //   tensor-memory loader reads one row per thread, i.e. 32 pages per load instruction, and measured 23.5 ms against
    //   15 ms for cp.async on the 1,024 x 1,024 x 1M input-delta GEMM (profiles/r1c_logs/bench_c4_1.json, two runs)
    const bool farRows = !aMN && lda > 65536u;
    const bool regLoader = ctx->gemmLoader == 1;
    if (ctx->gemmLoader == 3 && !aMN) {                                           // experimental coalesced tensor-memory A loader
        if (bMN) gemm_tc_ts_kernel<false, true, true><<<grid, THREADS, TS_SMEM_BYTES, ctx->stream>>>(a);
        else     gemm_tc_ts_kernel<false, false, true><<<grid, THREADS, TS_SMEM_BYTES, ctx->stream>>>(a);
    } else if (ctx->gemmLoader >= 2 || (ctx->gemmLoader < 0 && !farRows)) {
        if (aMN) {
            if (bMN) gemm_tc_ts_kernel<true, true><<<grid, THREADS, TS_SMEM_BYTES, ctx->stream>>>(a);
            else     gemm_tc_ts_kernel<true, false><<<grid, THREADS, TS_SMEM_BYTES, ctx->stream>>>(a);
        } else {
            if (bMN) gemm_tc_ts_kernel<false, true><<<grid, THREADS, TS_SMEM_BYTES, ctx->stream>>>(a);
            else     gemm_tc_ts_kernel<false, false><<<grid, THREADS, TS_SMEM_BYTES, ctx->stream>>>(a);
        }
    } else if (regLoader) {
        if (aMN) {
            if (bMN) gemm_tc_reg_kernel<true, true, 2><<<grid, THREADS, RSMEM_BYTES, ctx->stream>>>(a);
            else     gemm_tc_reg_kernel<true, false, 2><<<grid, THREADS, RSMEM_BYTES, ctx->stream>>>(a);
        } else {
            if (bMN) gemm_tc_reg_kernel<false, true, 2><<<grid, THREADS, RSMEM_BYTES, ctx->stream>>>(a);
            else     gemm_tc_reg_kernel<false, false, 2><<<grid, THREADS, RSMEM_BYTES, ctx->stream>>>(a);
        }
    } else if (aMN) {
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
