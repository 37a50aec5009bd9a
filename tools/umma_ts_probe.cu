// Bring-up probe: tcgen05.mma (kind::tf32, M = 128, N = 128, K = 8) with the A operand in TENSOR MEMORY.
// A(m, k) = 16 m + k + 1 is written with tcgen05.st.32x32b (thread = row / TMEM lane, consecutive registers =
// consecutive columns = consecutive k); B is a selector in shared memory (K-major, no swizzle, known-good layout) that
// routes k -> output column, so D[m][n] shows which TMEM word the tensor core used as A(m, k = n).  Two MMAs: columns
// 0-7 of the A block with selector k -> n = k, columns 8-15 with selector k -> n = 8 + k (accumulating).
// Expected: D[m][n] = 16 m + n + 1 for n < 16, 0 elsewhere.   Build: nvcc -arch=sm_100a -o umma_ts_probe umma_ts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0)
{
    return ((uint64_t)layout << 61) | (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void probe(float* out, int variant)
{
    extern __shared__ uint8_t raw[];
    float* sel0 = reinterpret_cast<float*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);   // K-major selector k -> n = k, 4 KB
    float* sel1 = sel0 + 1024;                                                            // K-major selector k -> n = 8 + k
    __shared__ uint64_t bar;
    __shared__ uint32_t tmemBase;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sel0[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 8) {
        int k = threadIdx.x;
        int mn = k;                                       // B element (n = mn, k) = 1
        sel0[((mn / 8) * 128 + (mn % 8) * 16 + (k / 4) * 2048 + (k % 4) * 4) / 4] = 1.0f;
        mn = 8 + k;
        sel1[((mn / 8) * 128 + (mn % 8) * 16 + (k / 4) * 2048 + (k % 4) * 4) / 4] = 1.0f;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x >= 128) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmemBase)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (threadIdx.x < 128) {
        // thread = row m = TMEM lane; 16 consecutive columns starting at column 128 of the allocation
        const uint32_t m = threadIdx.x, warp = threadIdx.x >> 5;
        uint32_t r[16];
        for (int k = 0; k < 16; k++) r[k] = __float_as_uint((float)(16 * m + k + 1));
        const uint32_t taddr = tmem + ((warp * 32) << 16) + 128;
        if (variant == 0) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        } else {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(taddr + 8), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 128) {
        // D = F32, A = B = TF32, both K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t d0 = smem_desc(smem_u32(sel0), 2048, 128), d1 = smem_desc(smem_u32(sel1), 2048, 128);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(tmem), "r"(tmem + 128), "l"(d0), "r"(idesc), "r"(0) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(tmem), "r"(tmem + 136), "l"(d1), "r"(idesc), "r"(1) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    if (threadIdx.x < 128) {
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t warp = threadIdx.x >> 5;
        for (int cb = 0; cb < 4; cb++) {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                           "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                           "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(tmem + ((warp * 32) << 16) + cb * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; j++) out[threadIdx.x * 128 + cb * 32 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x >= 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

int main()
{
    float* d; cudaMalloc(&d, 128 * 128 * 4);
    float* h = (float*)malloc(128 * 128 * 4);
    for (int variant = 0; variant < 2; variant++) {
        cudaMemset(d, 0xff, 128 * 128 * 4);
        probe<<<1, 160, 16384>>>(d, variant);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 128; n++) {
                const float want = n < 16 ? (float)(16 * m + n + 1) : 0.f;
                if (h[m * 128 + n] != want) bad++;
            }
        printf("variant %d (%s): %d mismatches of 16384\n", variant, variant ? "two x8 stores" : "one x16 store", bad);
        for (int m : {0, 1, 2, 31, 32, 33, 64, 127}) {
            printf("  m=%3d n=0..17:", m);
            for (int n = 0; n < 18; n++) printf(" %5.0f", h[m * 128 + n]);
            printf("\n");
        }
    }
    return 0;
}
