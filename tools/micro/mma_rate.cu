// mma_rate.cu - measured issue-to-completion rate of tcgen05.mma by form (SS / TS), kind (f16 / tf32) and N on sm_100a:
// every CTA issues `iters` batches of 256 MMAs into TMEM and waits for each batch's commit; prints SM clocks per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rate tools/micro/mma_rate.cu      Run: build/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    const uint32_t a = smem_u32(b);
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 2000000000ll) __trap();
    } while (!done);
}
// mode: bit0 = TS form (A from TMEM), bit1 = tf32 kind, bit2 = alternate two B tiles, bit3 = M = 64
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int n, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 3 * 32768);
    uint32_t* slot = (uint32_t*)(bar + 1);
    for (int i = threadIdx.x; i < 3 * 32768 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        const bool ts = mode & 1, tf32 = mode & 2, alt = mode & 4;
        const int m = (mode & 8) ? 64 : 128;
        const uint32_t idesc = (1u << 4) | (tf32 ? ((2u << 7) | (2u << 10)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = sa + 32768, sb2 = sa + 65536;
        // operands of the 4 k-steps precomputed: the timed loop is nothing but tcgen05.mma instructions (an earlier version rebuilt the
        // descriptors per MMA and measured its own ~216 clk of address arithmetic)
        uint64_t da[4], db[4], db2[4];
        uint32_t at[4];
        for (int k = 0; k < 4; ++k) { da[k] = desc_kmajor(sa + k * 32); db[k] = desc_kmajor(sb + k * 32); db2[k] = desc_kmajor(sb2 + k * 32); at[k] = tm + k * 8; }
        const uint32_t d0 = tm + 256, d1 = (n <= 128) ? tm + 384 : tm + 256;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int i = 0; i < 256; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = j & 3;
                    const uint64_t b = (alt && (j & 1)) ? db2[k] : db[k];
                    const uint32_t d = (j & 1) ? d1 : d0;
                    if (ts) {
                        if (tf32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at[k]), "l"(b), "r"(idesc), "r"(1u) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at[k]), "l"(b), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        if (tf32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da[k]), "l"(b), "r"(idesc), "r"(1u) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da[k]), "l"(b), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
            mbar_wait(bar, (uint32_t)(it & 1));   // (one batch in flight: the drain bubble is < 2 % of a 256-MMA batch)
        }
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d;
    cudaMalloc(&d, sizeof(long long) * 256);
    const int SMEM = 3 * 32768 + 64 + 1024;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const int iters = 40;
    const char* names[] = {"SS f16", "TS f16", "SS tf32", "TS tf32", "SS f16 alt-B", "TS f16 alt-B", "SS tf32 alt-B", "TS tf32 alt-B"};
    for (int grid : {1, sms}) {
        for (int mode = 0; mode < 8; ++mode) {
            for (int n : {256, 128, 64}) {
                k_rate<<<grid, 128, SMEM>>>(mode, n, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
                long long h[256];
                cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
                double avg = 0, mx = 0;
                for (int i = 0; i < grid; ++i) { avg += (double)h[i]; mx = h[i] > mx ? (double)h[i] : mx; }
                avg /= grid;
                printf("grid %3d  %-14s M=128 N=%3d : %7.1f clk / MMA (max CTA %7.1f)\n", grid, names[mode], n, avg / (iters * 256.0), mx / (iters * 256.0));
            }
        }
    }
    return 0;
}
