// Bring-up probe for the next round's coalesced tensor-memory A loader: which (TMEM lane, column) does register i of
// thread t land in for tcgen05.st.16x256b.x1 / .x2?  Each thread stores the code 100 * t + i (+ 10000 for the second
// store at lane offset 16); the block is read back with the known 32x32b shape (thread = lane, register = column).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_st16x256_probe umma_st16x256_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(float* out, int variant)
{
    __shared__ uint32_t tmemBase;
    const uint32_t warp = threadIdx.x >> 5, t = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmemBase)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (warp == 0) {
        // clear lanes 0-31, columns 0-15
        uint32_t z = 0;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(tmem), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t r[8];
        for (int i = 0; i < 8; i++) r[i] = __float_as_uint((float)(100 * t + i));
        if (variant == 0) {
            asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" :: "r"(tmem), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
            for (int i = 0; i < 4; i++) r[i] = __float_as_uint((float)(10000 + 100 * t + i));
            asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" :: "r"(tmem + (16u << 16)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
        } else {
            asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(tmem), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(tmem) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) out[t * 16 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32));
}

int main()
{
    float* d; cudaMalloc(&d, 32 * 16 * 4);
    float h[32 * 16];
    for (int variant = 0; variant < 2; variant++) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 32>>>(d, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("variant %d (%s): value = 100 * thread + register (+10000: second store at lane offset 16)\n", variant, variant ? "16x256b.x2" : "two 16x256b.x1");
        for (int lane = 0; lane < 32; lane++) {
            printf("  lane %2d:", lane);
            for (int c = 0; c < 16; c++) printf(" %6.0f", h[lane * 16 + c]);
            printf("\n");
        }
    }
    return 0;
}
