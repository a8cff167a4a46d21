// umma_bench.cu — microbenchmark: issue/execution rate of back-to-back tcgen05.mma (kind::f16, M=128,
// cta_group::1) on one SM, for several N, operand sources and accumulator patterns.  Development aid
// for k_mlp_tc.cu (measures what bounds its layer-1 / layer-2 MMA streams).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode 5/6/7: a layer-1-like stream (SS, N=128, D at col 0) and a layer-2-like stream (TS, N=144, D at col 256)
//             alternating every 1 / 4 / 11+8 instructions;  mode 8 = mode 7 while warps 1..3 hammer tcgen05.ld on cols 128..255
// mode: 0 = SS same accumulator, 1 = SS alternating 2 accumulators, 2 = TS (A in TMEM) same accumulator,
//       3 = SS, commit after every 4 MMAs, 4 = SS same accumulator but k-steps walk 4 slices of 4 smem blocks
__global__ void __launch_bounds__(128, 1) k_bench(int n, int mode_, int reps, long long *out_)
{
    const int mode = mode_ == 10 ? 7 : mode_;          // mode 10 = mode 7 on every SM of the chip
    if (mode == 11) {                                   // mode 11: the two streams of mode 7 issued by TWO warps (warp 0: SS, warp 1: TS)
        extern __shared__ uint8_t raw2[];
        uint8_t *smem2 = raw2 + ((1024u - (smem_u32(raw2) & 1023u)) & 1023u);
        __shared__ uint64_t bar2[2];
        __shared__ uint32_t slot2;
        for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem2)[i] = 0x3c003c00u;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot2)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tm = slot2;
        const int w = threadIdx.x >> 5;
        if (w < 2) {
            const uint64_t dB2 = make_sw128_desc(smem_u32(smem2 + 4 * 16384)), dA2 = make_sw128_desc(smem_u32(smem2));
            const uint32_t i1 = make_idesc(128), i2 = make_idesc(144);
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
                const int mine = w == 0 ? reps * 11 / 19 : reps * 8 / 19;
                const long long t0 = clock64();
                for (int r = 0; r < mine; ++r) {
                    if (w == 0) umma_ss(tm, dA2 + 2 * (r & 3), dB2 + 2 * (r & 3), i1, r > 0);
                    else umma_ts(tm + 256u, tm + 448u + 8u * (r & 3), dB2 + 2 * (r & 3), i2, r > 0);
                }
                const long long t1 = clock64();
                commit(&bar2[w]); mbar_wait(&bar2[w], 0);
                const long long t2 = clock64();
                out_[4 * blockIdx.x + 2 * w] = (t1 - t0) * 1000 / mine;
                out_[4 * blockIdx.x + 2 * w + 1] = (t2 - t0) * 1000 / mine;
            }
            __syncwarp();
        }
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
        return;
    }
    long long *out = out_ + 4 * blockIdx.x;
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int lane_done_s;
    volatile int *lane_done = &lane_done_s;
    if (threadIdx.x == 0) lane_done_s = 0;
    uint8_t *sA = smem, *sB = smem + 4 * 16384;   // 4 blocks of A (128 x 64 fp16), up to 4 x 32 KB of B (256 x 64)
    for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {   // converged warp, one elected lane issues (the pattern k_mlp_tc.cu uses)
        const uint32_t idesc = make_idesc(n);
        const uint64_t dA = make_sw128_desc(smem_u32(sA)), dB = make_sw128_desc(smem_u32(sB));
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        long long t0 = 0, t1 = 0, t2 = 0;
        if (pred) {
            t0 = clock64();
            if (mode >= 5) {
                const uint32_t i1 = make_idesc(128), i2 = make_idesc(144);
                const int g1 = mode == 5 ? 1 : (mode == 6 ? 4 : 11), g2 = mode == 5 ? 1 : (mode == 6 ? 4 : 8);
                for (int r = 0; r < reps;) {
                    for (int i = 0; i < g1; ++i, ++r) umma_ss(tmem, dA + 2 * (i & 3), dB + 2 * (i & 3), i1, r > 1);
                    for (int i = 0; i < g2; ++i, ++r) umma_ts(tmem + 256u, tmem + 448u + 8u * (i & 3), dB + 2 * (i & 3), i2, r > 1);
                }
            } else
            for (int r = 0; r < reps; r += 4) {
                const int blk = (r >> 2) & 3;
                const uint64_t a = dA + (mode == 4 ? (uint64_t)(blk * 16384 >> 4) : 0);
                const uint64_t b = dB + (mode == 4 ? (uint64_t)(blk * 32768 >> 4) : 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t d = tmem + ((mode == 1 && (ks & 1)) ? 256u : 0u);
                    if (mode == 2) umma_ts(d, tmem + 448u + 8u * ks, b + 2 * ks, idesc, (r | ks) > 1);
                    else umma_ss(d, a + 2 * ks, b + 2 * ks, idesc, (r | ks) > 1);
                }
                if (mode == 3) commit(&bar);   // extra arrivals only advance phases
            }
            t1 = clock64();
            if (mode != 3) { commit(&bar); mbar_wait(&bar, 0); }
            t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
        __syncwarp();
        if (lane_done) *lane_done = 1;
    } else if (mode == 9) {   // epilogue-like ALU/MUFU load from the other three warps
        float x = (float)threadIdx.x, y = 1.0f;
        while (*lane_done == 0) {
#pragma unroll
            for (int i = 0; i < 64; ++i) { x = fmaf(x, 1.0001f, 0.5f); y = __frcp_rn(y + x); }
        }
        if (x + y == 0.123f) out[2] = 1;
    } else if (mode == 8) {   // epilogue-like TMEM reads from the other three warps (their own lane quarters)
        uint32_t v[32], acc = 0;
        const uint32_t la = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
        while (*lane_done == 0) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tmem + la + 128u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
        if (acc == 0x12345678u) out[2] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main()
{
    long long *d_out, h[2];
    cudaMalloc(&d_out, 32 * 148);
    const int smem = 4 * 16384 + 4 * 32768 + 1024;
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 19 * 16;
    const char *names[] = {"SS same acc", "SS 2 accs", "TS same acc", "SS commit/4", "SS walk blocks", "L1/L2 alt 1", "L1/L2 alt 4",
                           "L1/L2 alt 11/8", "alt 11/8 + LDTM", "alt 11/8 + ALU", "alt 11/8 x148 SMs"};
    {   // two issuing warps
        for (int it = 0; it < 2; ++it) { k_bench<<<1, 128, smem>>>(128, 11, reps, d_out); cudaDeviceSynchronize(); }
        long long h4[4];
        cudaMemcpy(h4, d_out, 32, cudaMemcpyDeviceToHost);
        printf("two issuers      SS warp: issue %6.1f complete %6.1f clk/MMA | TS warp: issue %6.1f complete %6.1f clk/MMA\n",
               h4[0] / 1000.0, h4[1] / 1000.0, h4[2] / 1000.0, h4[3] / 1000.0);
    }
    for (int mode = 0; mode < 11; ++mode)
        for (int n : {64, 128, 144, 256}) {
            if (mode >= 5 && n != 128) continue;
            if (mode == 1 && n > 256) continue;
            for (int it = 0; it < 2; ++it) {
                k_bench<<<mode == 10 ? 148 : 1, 128, smem>>>(n, mode, mode == 10 ? reps * 64 : reps, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
            const int rr = mode == 10 ? reps * 64 : reps;
            printf("%-16s N=%3d : issue %6.1f clk/MMA, complete %6.1f clk/MMA (nominal %d)\n", names[mode], n, (double)h[0] / rr,
                   (double)h[1] / rr, n / 2);
        }
    return 0;
}
