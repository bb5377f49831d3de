// Microbenchmark: what one tcgen05.mma (kind::f16, bf16 operands, K = 16) costs on sm_100a as a function of
//   * where A comes from (shared memory descriptor vs tensor memory),
//   * N (64 / 128 / 256),
//   * cta_group (1 CTA, M = 128; CTA pair, M = 256, each CTA holding half of B),
//   * and how many CTAs run at once (1 vs one per SM: the part is power capped).
// No loads, no epilogue: one thread issues `iters` x 4 MMAs on fixed operands, commits, waits; cycles = clock64 around it.
// Also times tcgen05.cp (shared -> tensor memory, 128x256b) of one K = 64 A tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I handmvnet_b200/csrc tools/mma_issue_bench.cu -o tools/_mma_issue_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
constexpr int kTcBlockM = 128;
#include "tc_ptx.cuh"

using namespace hmv;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

// mode 0: A from shared memory; 1: A from tensor memory; 2: tcgen05.cp of the A tile then A from tensor memory (per 4 MMAs);
// 3: only the tcgen05.cp's
template <int PAIR>
__global__ void __launch_bounds__(256) mma_bench(int mode, int n, int iters, int bg, long long* out, const uint8_t* gsrc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                 // 128 rows x 64 bf16, 128B swizzle atoms (16 KB)
    uint8_t* b_tile = smem + 16384;         // up to 256 rows x 64 bf16 (32 KB)
    __shared__ uint64_t bar;
    __shared__ uint64_t cbar[4];
    __shared__ volatile int done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    if (tid == 0) done = 0;
    for (int i = tid; i < (16384 + 32768 + 65536) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // bf16 pairs, finite
    uint32_t rank = 0;
    if (PAIR) rank = cluster_ctarank();
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&cbar[i]), 1);
        fence_barrier_init();
    }
    if (tid < 32) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long cycles = 0;
    if (mode == 4 && tid < 32 && rank == 0) {
        // lean issue path: the whole warp walks the loop (uniform control flow), one elected lane issues; descriptors are
        // built once and advanced with a 64-bit add
        const uint32_t idesc = PAIR ? make_idesc_mn(256, n) : make_idesc_mn(128, n);
        const uint64_t ad = make_sw128_desc(smem_u32(a_tile));
        const uint64_t bd = make_sw128_desc(smem_u32(b_tile));
        const uint32_t d_tmem = tmem;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (PAIR) umma_f16_2sm(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
                    else umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) {
            if (PAIR) umma_commit_2sm_mc(smem_u32(&bar), 3);
            else umma_commit(smem_u32(&bar));
        }
        while (!mbar_try_wait(smem_u32(&bar), 0) && clock64() - t0 < 3000000000LL) {}
        cycles = clock64() - t0;
        if (tid == 0) { out[blockIdx.x / (PAIR ? 2 : 1)] = cycles; done = 1; }
    } else if (mode != 4 && tid == 0 && rank == 0) {
        const uint32_t idesc = PAIR ? make_idesc_mn(256, n) : make_idesc_mn(128, n);
        const uint64_t ad = make_sw128_desc(smem_u32(a_tile));
        const uint64_t bd = make_sw128_desc(smem_u32(b_tile));
        const uint32_t d_tmem = tmem;            // columns 0..n-1
        const uint32_t a_tmem = tmem + 256;      // A tile in tensor memory: 128 lanes x 32 columns per K = 64
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (mode == 2 || mode == 3) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tmem_cp_128x256b(a_tmem + (it & 1) * 32 + k * 8, ad + (k * 32 >> 4));
            }
            if (mode != 3) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (PAIR) umma_f16_2sm(d_tmem, ad + (k * 32 >> 4), bd + (k * 32 >> 4), idesc, 1u);
                    else if (mode == 0) umma_f16(d_tmem, ad + (k * 32 >> 4), bd + (k * 32 >> 4), idesc, 1u);
                    else umma_f16_ts(d_tmem, a_tmem + (it & 1) * 32 + k * 8, bd + (k * 32 >> 4), idesc, 1u);
                }
            }
        }
        if (PAIR) umma_commit_2sm_mc(smem_u32(&bar), 3);
        else umma_commit(smem_u32(&bar));
        while (!mbar_try_wait(smem_u32(&bar), 0) && clock64() - t0 < 3000000000LL) {}
        cycles = clock64() - t0;
        out[blockIdx.x / (PAIR ? 2 : 1)] = cycles;
        done = 1;
    } else if (PAIR && tid == 0) {
        const long long t0 = clock64();
        while (!mbar_try_wait(smem_u32(&bar), 0) && clock64() - t0 < 3000000000LL) {}
        done = 1;
    } else if ((bg & 64) && tid >= 128) {
        // background ALU work that never stalls: 8 independent FMA chains per thread (warp 4 shares the issuer's scheduler)
        float a[8];
        for (int j = 0; j < 8; ++j) a[j] = static_cast<float>(tid + j);
        const float x = 1.0000001f, y = 1e-9f;
        long long n_ops = 0;
        const long long t0 = clock64();
        while (!done && clock64() - t0 < 3000000000LL) {
#pragma unroll
            for (int r = 0; r < 32; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], x, y);
            n_ops += 256;
        }
        float sum = 0.f;
        for (int j = 0; j < 8; ++j) sum += a[j];
        if ((tid & 31) == 0) out[128 + blockIdx.x % 64] = n_ops + (sum == 0.12345f ? 1 : 0);
    } else if ((bg & 16) && tid >= 128) {
        // background tensor-memory reads: 4 warps, each tcgen05.ld.32x32b.x32 of its own 32 lanes (128 B per lane per op)
        const uint32_t taddr = tmem + (static_cast<uint32_t>((tid >> 5) & 3) << 21) + 320;
        uint32_t r[32];
        uint32_t acc = 0;
        long long n_ops = 0;
        const long long t0 = clock64();
        while (!done && clock64() - t0 < 3000000000LL) {
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                tmem_ld32(taddr + (j & 3) * 32, r);
                tmem_ld_wait();
                acc ^= r[0] ^ r[31];
            }
            n_ops += 8;
        }
        if ((tid & 31) == 0) out[128 + blockIdx.x % 64] = n_ops * 4096 + (acc == 0x12345 ? 1 : 0);
    } else if ((bg & 32) && tid == 128) {
        // background bulk copies global (L2 resident) -> shared memory, 16 KB each, 4 in flight
        const uint32_t dst0 = smem_u32(smem + 16384 + 32768);
        long long n_ops = 0;
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int i = 0; i < 4; ++i) {
            mbar_arrive_expect_tx(smem_u32(&cbar[i]), 16384);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst0 + i * 16384), "l"(gsrc + ((blockIdx.x * 4 + i) & 63) * 16384), "r"(16384), "r"(smem_u32(&cbar[i])) : "memory");
        }
        while (!done && clock64() - t0 < 3000000000LL) {
            for (int i = 0; i < 4; ++i) {
                while (!mbar_try_wait(smem_u32(&cbar[i]), ph) && clock64() - t0 < 3000000000LL) {}
                mbar_arrive_expect_tx(smem_u32(&cbar[i]), 16384);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst0 + i * 16384), "l"(gsrc + ((blockIdx.x * 4 + i + n_ops) & 63) * 16384), "r"(16384), "r"(smem_u32(&cbar[i])) : "memory");
            }
            ph ^= 1;
            n_ops += 4;
        }
        for (int i = 0; i < 4; ++i)
            while (!mbar_try_wait(smem_u32(&cbar[i]), ph) && clock64() - t0 < 3000000000LL) {}
        out[128 + blockIdx.x % 64] = n_ops * 16384;
    } else if (!(bg & 112) && tid >= 128 && tid < 128 + 32 * (bg & 7)) {
        // background shared-memory traffic: every warp streams 512 B per instruction through a separate 64 KB region
        const uint32_t base = smem_u32(smem + 16384 + 32768) + (tid - 128) * 16;
        uint32_t acc = 0;
        const long long t0 = clock64();
        long long n_ops = 0;
        while (!done && clock64() - t0 < 3000000000LL) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t a, b, c, d;
                const uint32_t addr = base + ((j * 4096) & 65535);
                if (bg & 8) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(acc) : "memory");
                else {
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
                    acc += a ^ b ^ c ^ d;
                }
            }
            n_ops += 16;
        }
        if ((tid & 31) == 0) out[128 + blockIdx.x % 64] = n_ops * 512 + (acc == 0x12345 ? 1 : 0);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (tid < 32) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int PAIR>
static void run(const char* name, int mode, int n, int ctas, int iters, int bg = 0) {
    long long* out;
    CK(cudaMallocManaged(&out, sizeof(long long) * 256));
    static uint8_t* gsrc = nullptr;
    if (!gsrc) { CK(cudaMalloc(&gsrc, 64 * 16384)); CK(cudaMemset(gsrc, 0, 64 * 16384)); }
    const size_t smem = 16384 + 32768 + 65536 + 1024;
    CK(cudaFuncSetAttribute(mma_bench<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelEx(&cfg, mma_bench<PAIR>, mode, n, iters, bg, out, static_cast<const uint8_t*>(gsrc)));
        CK(cudaDeviceSynchronize());
    }
    const int groups = ctas / (PAIR ? 2 : 1);
    std::vector<long long> v(out, out + groups);
    std::sort(v.begin(), v.end());
    const double per = static_cast<double>(v[groups / 2]) / (iters * 4.0);
    printf("%-44s N=%3d ctas=%3d : %7.1f cycles per K=16 step (median CTA; min %.1f max %.1f)", name, n, ctas, per,
           v[0] / (iters * 4.0), v[groups - 1] / (iters * 4.0));
    if (bg & 64) printf("  background FMA warps x4: %.2f FMA instr/clk per warp", static_cast<double>(out[128]) / v[groups / 2]);
    else if (bg & 16) printf("  background tcgen05.ld x4 warps: %.1f B/clk per warp", static_cast<double>(out[128]) / v[groups / 2]);
    else if (bg & 32) printf("  background bulk copies into shared memory: %.1f B/clk", static_cast<double>(out[128]) / v[groups / 2]);
    else if (bg) printf("  background %s x%d warps: %.1f B/clk per warp", (bg & 8) ? "st.shared" : "ld.shared", bg & 7, static_cast<double>(out[128]) / v[groups / 2]);
    printf("\n");
    CK(cudaFree(out));
}

int main() {
    const int iters = 4000;
    for (int n : {64, 128, 192, 256}) run<1>("CTA pair M=256 A, B from shared memory", 0, n, 148, iters);
    for (int bg : {0, 64}) {
        for (int n : {64, 128, 256}) run<0>("1 CTA  M=128 A from shared memory", 0, n, 148, iters, bg);
        for (int n : {64, 128, 256}) run<0>("1 CTA  M=128 lean issue (warp + elect)", 4, n, 148, iters, bg);
        for (int n : {64, 128, 256}) run<1>("CTA pair M=256 A, B from shared memory", 0, n, 148, iters, bg);
        for (int n : {64, 128, 256}) run<1>("CTA pair M=256 lean issue (warp + elect)", 4, n, 148, iters, bg);
    }
    for (int ctas : {1}) {
        for (int n : {64, 128, 256}) run<0>("1 CTA  M=128 A from shared memory", 0, n, ctas, iters);
        for (int n : {64, 128, 256}) run<0>("1 CTA  M=128 A from tensor memory", 1, n, ctas, iters);
        for (int n : {64, 128, 256}) run<0>("1 CTA  M=128 tcgen05.cp A + A from tmem", 2, n, ctas, iters);
        run<0>("1 CTA  tcgen05.cp 128x256b x4 only", 3, 64, ctas, iters);
        const int pc = ctas == 1 ? 2 : 148;
        for (int n : {128, 256}) run<1>("CTA pair M=256 A, B from shared memory", 0, n, pc, iters);
    }
    return 0;
}
