// Bring-up probe: which shared-memory word does tcgen05.mma (kind::tf32, no swizzle) read for operand element (k, mn)
// of an MN-major operand, as a function of LBO / SBO?  A is an identity selector (K-major, known-good layout), B holds
// its own word index, so D[m][n] = index of the word the hardware used as B(k = m, n).  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0)
{
    return ((uint64_t)layout << 61) | (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// probeA = 0: B is the MN-major pattern operand, A the selector; probeA = 1: roles exchanged
__global__ void probe(float* out, uint32_t lbo, uint32_t sbo, int probeA, int mnMajor, uint32_t layout)
{
    extern __shared__ uint8_t raw[];
    float* sel = reinterpret_cast<float*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);   // K-major selector, 4 KB
    float* pat = sel + 1024;                                                             // 16 KB pattern
    __shared__ uint64_t bar;
    __shared__ uint32_t tmemBase;
    // selector needs kc stride 2048 B -> 128 rows * 16 B = 2048: exactly 2 chunks * 2048 = 4096 B = 1024 floats
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sel[i] = 0.f;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) pat[i] = (float)(i & 2047);
    __syncthreads();
    if (threadIdx.x < 8) {
        int mn = threadIdx.x, k = threadIdx.x;            // element (mn, k) = 1
        sel[((mn / 8) * 128 + (mn % 8) * 16 + (k / 4) * 2048 + (k % 4) * 4) / 4] = 1.0f;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x >= 128) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmemBase)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = tmemBase;
    if (threadIdx.x == 128) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((probeA && mnMajor ? 1u : 0u) << 15) | ((!probeA && mnMajor ? 1u : 0u) << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t dSel = smem_desc(smem_u32(sel), 2048, 128);
        uint64_t dPat = smem_desc(smem_u32(pat), lbo, sbo, layout);
        uint64_t da = probeA ? dPat : dSel, db = probeA ? dSel : dPat;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
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
    if (threadIdx.x >= 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int main()
{
    float* d; cudaMalloc(&d, 128 * 128 * 4);
    float* h = (float*)malloc(128 * 128 * 4);
    const uint32_t combos[][4] = {{2048, 1024, 1, 2}, {1024, 2048, 1, 2}, {2048, 1024, 1, 4}, {2048, 1024, 1, 6}, {2048, 1024, 1, 1}, {16, 1024, 0, 2}};
    for (int probeA = 0; probeA < 1; probeA++)
        for (auto& c : combos) {
            cudaMemset(d, 0xff, 128 * 128 * 4);
            probe<<<1, 160, 22528>>>(d, c[0], c[1], probeA, (int)c[2], c[3]);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("probeA=%d lbo=%u sbo=%u: %s\n", probeA, c[0], c[1], cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            // probeA = 0: out[m][n] = word of B(k=m, n);  probeA = 1: out[m][n] = word of A(m, k=n) (n < 8)
            printf("probeA=%d lbo=%u sbo=%u mnMajor=%u layout=%u\n", probeA, c[0], c[1], c[2], c[3]);
            if (!probeA) {
                for (int k = 0; k < 8; k += 1) {
                    printf("  k=%d n=0..11:", k);
                    for (int n = 0; n < 12; n++) printf(" %4.0f", h[k * 128 + n]);
                    printf("  | n=32,33,64,124..127: %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f\n", h[k * 128 + 32], h[k * 128 + 33], h[k * 128 + 64], h[k * 128 + 124],
                           h[k * 128 + 125], h[k * 128 + 126], h[k * 128 + 127]);
                }
            } else {
                for (int k = 0; k < 8; k += 3) {
                    printf("  k=%d m=0..11:", k);
                    for (int m = 0; m < 12; m++) printf(" %4.0f", h[m * 128 + k]);
                    printf("  | m=32,33,64,124..127: %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f\n", h[32 * 128 + k], h[33 * 128 + k], h[64 * 128 + k], h[124 * 128 + k],
                           h[125 * 128 + k], h[126 * 128 + k], h[127 * 128 + k]);
                }
            }
        }
    return 0;
}
