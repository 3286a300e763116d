// mma_peak.cu — measured tcgen05.mma issue-rate ceilings on this box, the denominators the conv kernel's roofline is read
// against (VERDICT r1 "make the roofline reproducible": the only driver-measured tensor peak is cuBLAS bf16 at a
// power-capped 1.25 GHz, while the conv kernel runs its 3-MMA split at ~1.7 GHz).
//
// One CTA per SM, operands resident in shared memory (A 128 x 128 B, B N x 128 B, K-major SWIZZLE_128B, random finite
// data so that the power draw is realistic), one thread issues `iters` x 4 back-to-back MMAs of 128 x N x {8 tf32 | 16 f16}
// into one TMEM accumulator, commits, waits.  Prints one JSON line per (kind, N): TFLOP/s over all SMs, cycles per MMA
// (clock64) and the SM clock derived from clock64 / globaltimer.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/bin/mma_peak profiles/mma_peak.cu && profiles/bin/mma_peak
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

template <int KIND /*0 tf32, 1 f16*/, int TS /*A operand from tensor memory*/>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int N, unsigned long long* out /*[grid][3]: cycles, ns, mmas*/) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_tile = smem;                 // 128 rows x 128 B
    uint8_t* b_tile = smem + 16384;         // N rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    // finite pseudo-random operands
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        uint32_t v;
        if (KIND == 0) v = 0x3F000000u | (h & 0x007FE000u) | ((h & 1u) << 31);                       // +-[0.5, 1) with a tf32 mantissa
        else v = (0x3800u | (h & 0x03FFu) | ((h & 0x8000u))) | ((0x3800u | ((h >> 16) & 0x03FFu) | ((h >> 1) & 0x8000u)) << 16);
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (TS) {   // A operand: 32 columns of tensor memory at column 256 (whatever bits are there are finite after this store)
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) r[j] = KIND == 0 ? (0x3F000000u | ((threadIdx.x * 37 + j * 101) << 13 & 0x007FE000u)) : 0x38003C00u;
        const uint32_t dst = tmem + 256 + ((uint32_t)(threadIdx.x & ~31) << 16);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(dst), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
              "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
              "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
              "r"(r[30]), "r"(r[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (KIND ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = umma_desc(smem_u32(a_tile)), bd = umma_desc(smem_u32(b_tile));
        unsigned long long t0n, t1n;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0n));
        const long long c0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (it | k) != 0;
                if (TS) {
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                                     ::"r"(tmem), "r"(tmem + 256 + 8 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                                     ::"r"(tmem), "r"(tmem + 256 + 8 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
                } else {
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                                     ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                     ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long c1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1n));
        out[blockIdx.x * 3 + 0] = (unsigned long long)(c1 - c0);
        out[blockIdx.x * 3 + 1] = t1n - t0n;
        out[blockIdx.x * 3 + 2] = (unsigned long long)iters * 4;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND, int TS>
void run(int N, int iters, int sms) {
    unsigned long long* out;
    cudaMalloc(&out, sizeof(unsigned long long) * 3 * sms);
    const int smem = 1024 + 16384 + N * 128;
    cudaFuncSetAttribute(peak_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<KIND, TS><<<sms, 128, smem>>>(iters / 8, N, out);      // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    peak_kernel<KIND, TS><<<sms, 128, smem>>>(iters, N, out);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[3 * 256];
    cudaMemcpy(h, out, sizeof(unsigned long long) * 3 * sms, cudaMemcpyDeviceToHost);
    double cyc = 0, ns = 0;
    for (int i = 0; i < sms; ++i) { cyc += (double)h[i * 3]; ns += (double)h[i * 3 + 1]; }
    cyc /= sms; ns /= sms;
    const double kstep = KIND ? 16 : 8, mmas = (double)iters * 4;
    const double flops = 2.0 * 128 * N * kstep * mmas * sms;
    printf("{\"kind\": \"%s\", \"a_operand\": \"%s\", \"M\": 128, \"N\": %d, \"K\": %d, \"sms\": %d, \"mmas_per_sm\": %.0f, \"ms\": %.3f, "
           "\"tflops\": %.1f, \"cycles_per_mma\": %.2f, \"sm_mhz\": %.0f}\n",
           KIND ? "f16" : "tf32", TS ? "tmem" : "smem", N, (int)kstep, sms, mmas, ms, flops / (ms * 1e-3) / 1e12, cyc / mmas, cyc / ns * 1e3);
    cudaFree(out);
}

int main(int argc, char** argv) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = argc > 1 ? atoi(argv[1]) : 200000;     // x4 MMAs: ~50-100 ms per launch, long enough for the power cap to settle
    for (int N : {128, 256, 64, 32}) {
        run<0, 0>(N, iters, sms);
        run<0, 1>(N, iters, sms);
        run<1, 0>(N, iters, sms);
        run<1, 1>(N, iters, sms);
    }
    return 0;
}
