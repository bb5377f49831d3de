// One fusion-transformer layer after its QKV projection as ONE kernel (bf16 tensor-core path):
//   softmax(Q K^T * scale) V  ->  to_out (+bias) + residual  ->  LayerNorm(norm1)  ->  LayerNorm(ff.net.0)
//   ->  Linear(d,128) + GELU  ->  Linear(128,d) (+bias) + residual  ->  LayerNorm(norm2)
// (reference MultiHeadAttention.forward, src/models/layers.py:217-236, FeedForward :165-170).
//
// A CTA owns 32 query rows of one sample (everything after QKV is row-local except attention, which is per sample):
//   phase 1  attention, mma.sync m16n8k16 bf16 / fp32 accumulate, flash-style online softmax over 64-key blocks;
//            8 warps = 4 heads in flight x 2 query m-tiles, two rounds for the 8 heads; the normalised head outputs
//            overwrite the Q tile in shared memory, which becomes the A operand [32 x 1024] of the out-projection
//   phase 2  out-projection: warp w owns output columns [72w, 72w+72) (d = 524 padded to 576) for all 32 rows and
//            streams ITS rows of W_out from L2 through a private 3-stage cp.async ring (no block-wide barriers)
//   phase 3  + bias + residual (fp32 master copy of the tokens), LayerNorm x2 with block-wide row statistics
//   phase 4/5 the feed-forward GEMMs with the same streaming helper, hidden activations in shared memory
//   phase 6  + residual, LayerNorm, fp32 master + bf16 copy of the layer output (pad columns written as zero)
// so the token stream makes one HBM round trip per layer and the layer is 2 launches (QKV GEMM + this) instead of 7.
#include "kernels.cuh"
#include "conv_gemm_tc.cuh"
#include "tc_ptx.cuh"      // cluster rank / barrier / mapa wrappers

namespace hmv {

namespace {

constexpr int kFbRows = 32;                      // query rows per CTA
constexpr int kFbHeads = 8, kFbD = 128, kFbInner = kFbHeads * kFbD;
constexpr int kFbKeys = 64;                      // key block
constexpr int kFbDp = 576;                       // d_model padded: 8 warps x 72 columns
constexpr int kFbNt = 9;                         // n8 tiles per warp in the d_model-wide GEMMs
constexpr int kFbHid = 128;                      // feed-forward hidden width
constexpr int kFbQPitch = kFbInner + 8;          // bf16 elements; (pitch * 2) % 128 == 16 -> conflict-free ldmatrix
constexpr int kFbKvPitch = kFbD + 8;
constexpr int kFbHPitch = kFbDp + 8;
constexpr int kFbFPitch = kFbHid + 8;
constexpr int kFbStageBytes = 72 * (32 + 8) * 2; // largest per-warp weight stage: 72 rows x 32 k
constexpr int kFbQoBytes = kFbRows * kFbQPitch * 2;                       // 66048
constexpr int kFbKvBytes = 4 * 2 * kFbKeys * kFbKvPitch * 2;              // 139264 (aliased by the weight rings)
constexpr int kFbRedBytes = 6 * 8 * kFbRows * 4;                          // 6 reductions x [8 warps][32 rows]
constexpr int kFbSmem = kFbQoBytes + kFbKvBytes + kFbRedBytes;
static_assert(8 * 3 * kFbStageBytes <= kFbKvBytes, "weight rings alias the K/V blocks");
static_assert(kFbRows * kFbHPitch * 2 + kFbRows * kFbFPitch * 2 <= kFbQoBytes, "H and F alias the Q/O tile");

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bar_pair(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// acc[2][NT][4] += A[32 x K] * W[n0 .. n0 + 8*NT, 0 .. K)^T.   A: shared memory, row-major bf16 (byte pitch a_pitch);
// W: global, row-major bf16 with leading dimension ldw (already offset to row n0).  The warp streams its W rows in
// K chunks of KCH through its private 3-stage ring at `ring` (shared address); no block-level synchronisation.
template <int NT, int KCH>
__device__ __forceinline__ void warp_stream_gemm(float (&acc)[2][NT][4], uint32_t a_addr, int a_pitch, const bf16* __restrict__ w,
                                                 int ldw, int K, uint32_t ring, int lane) {
    constexpr int kPitch = (KCH + 8) * 2;                 // bytes per staged W row
    constexpr int kStage = NT * 8 * kPitch;
    constexpr int kPieces = NT * 8 * (KCH / 8);           // 16-byte pieces per chunk
    static_assert(kStage <= kFbStageBytes && kPieces % 32 == 0, "stage geometry");
    const int nchunks = K / KCH;
    auto issue = [&](int c) {
        if (c < nchunks) {
            const uint32_t dst = ring + (c % 3) * kStage;
#pragma unroll
            for (int i = 0; i < kPieces / 32; ++i) {
                const int idx = lane + 32 * i;
                const int row = idx / (KCH / 8), pc = idx % (KCH / 8);
                cp16(dst + row * kPitch + pc * 16, w + static_cast<size_t>(row) * ldw + c * KCH + pc * 8);
            }
        }
        cp_commit();
    };
    issue(0);
    issue(1);
    const uint32_t a_lane = a_addr + (lane & 15) * a_pitch + (lane >> 4) * 16;
    const uint32_t b_lane = ((lane & 7) + (lane >> 4) * 8) * kPitch + ((lane >> 3) & 1) * 16;
    const uint32_t b_lane2 = (lane & 7) * kPitch + ((lane >> 3) & 1) * 16;           // x2 form: lanes 0-15 address
    for (int c = 0; c < nchunks; ++c) {
        issue(c + 2);
        cp_wait<2>();
        __syncwarp();
        const uint32_t wst = ring + (c % 3) * kStage;
#pragma unroll
        for (int kk = 0; kk < KCH / 16; ++kk) {
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                ldsm4(a_lane + mt * 16 * a_pitch + (c * KCH + kk * 16) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t b0, b1, b2, b3;
                ldsm4(wst + b_lane + np * 16 * kPitch + kk * 32, b0, b1, b2, b3);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma16816(acc[mt][2 * np], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
                    mma16816(acc[mt][2 * np + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
                }
            }
            if (NT & 1) {
                uint32_t b0, b1;
                ldsm2(wst + b_lane2 + (NT - 1) * 8 * kPitch + kk * 32, b0, b1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma16816(acc[mt][NT - 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
            }
        }
        __syncwarp();                                     // the stage is rewritten two iterations from now
    }
    cp_wait<0>();
}

// Block-wide sum over the columns of each of the 32 rows.  part[mt*2 + hi] is this thread's partial for row
// mt*16 + g + 8*hi; `red` is a [8 warps][32 rows] scratch used by exactly one call.
__device__ __forceinline__ void row_sums(float (&part)[4], float* red, int warp, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1);
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2);
    }
    const int g = lane >> 2;
    if ((lane & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) red[warp * kFbRows + (i >> 1) * 16 + g + (i & 1) * 8] = part[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * kFbRows + (i >> 1) * 16 + g + (i & 1) * 8];
        part[i] = s;
    }
}

// Row statistics (mean, 1/std) of the 32 x d tile held in the accumulator layout; entry mt*2 + hi is row mt*16 + g + 8*hi.
__device__ __forceinline__ void layer_norm_stats(const float (&v)[2][kFbNt][4], int col0, int d, float* red, int warp, int lane,
                                                 float (&mean)[4], float (&rstd)[4]) {
    const int t = lane & 3;
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < kFbNt; ++nt) {
            part[mt * 2] += v[mt][nt][0] + v[mt][nt][1];        // columns >= d hold exact zeros
            part[mt * 2 + 1] += v[mt][nt][2] + v[mt][nt][3];
        }
    row_sums(part, red, warp, lane);
    const float inv_d = 1.f / static_cast<float>(d);
#pragma unroll
    for (int i = 0; i < 4; ++i) { mean[i] = part[i] * inv_d; part[i] = 0.f; }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < kFbNt; ++nt) {
            const int c = col0 + nt * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float dv = (c + (e & 1)) < d ? v[mt][nt][e] - mean[mt * 2 + (e >> 1)] : 0.f;
                part[mt * 2 + (e >> 1)] += dv * dv;
            }
        }
    row_sums(part, red + 8 * kFbRows, warp, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) rstd[i] = rsqrtf(part[i] * inv_d + 1e-5f);
}

// LayerNorm of one accumulator element (row statistics from layer_norm_stats); pad columns stay zero.
__device__ __forceinline__ float ln_apply(float x, float mean, float rstd, float gamma, float beta, bool valid) {
    return valid ? (x - mean) * rstd * gamma + beta : 0.f;
}

// In-place LayerNorm of the tile.
__device__ __forceinline__ void layer_norm_tile(float (&v)[2][kFbNt][4], const float* __restrict__ gamma, const float* __restrict__ beta,
                                                int col0, int d, float* red, int warp, int lane) {
    float mean[4], rstd[4];
    layer_norm_stats(v, col0, d, red, warp, lane, mean, rstd);
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < kFbNt; ++nt) {
        const int c = col0 + nt * 8 + 2 * t;
        const float2 gq = __ldg(reinterpret_cast<const float2*>(gamma + c)), bq = __ldg(reinterpret_cast<const float2*>(beta + c));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int e = 0; e < 4; ++e)
                v[mt][nt][e] = ln_apply(v[mt][nt][e], mean[mt * 2 + (e >> 1)], rstd[mt * 2 + (e >> 1)], (e & 1) ? gq.y : gq.x,
                                        (e & 1) ? bq.y : bq.x, (c + (e & 1)) < d);
    }
}

__global__ void __launch_bounds__(256, 1) fusion_block_kernel(const FusionBlockParams p) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(128) uint8_t fb_smem[];
    bf16* qo = reinterpret_cast<bf16*>(fb_smem);
    uint8_t* region = fb_smem + kFbQoBytes;
    float* red = reinterpret_cast<float*>(fb_smem + kFbQoBytes + kFbKvBytes);
    const uint32_t qo_addr = static_cast<uint32_t>(__cvta_generic_to_shared(qo));
    const uint32_t region_addr = static_cast<uint32_t>(__cvta_generic_to_shared(region));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y, q0 = blockIdx.x * kFbRows;
    const int rows_valid = p.nq - q0 < kFbRows ? p.nq - q0 : kFbRows;
    const size_t in_row0 = static_cast<size_t>(b) * p.s_in + q0;         // first query row in the qkv / residual streams
    const size_t out_row0 = static_cast<size_t>(b) * p.nq + q0;

    // ---------------- phase 1: attention ----------------
    {
        const bf16* qsrc = p.qkv + in_row0 * p.ld_qkv;
        for (int i = threadIdx.x; i < kFbRows * (kFbInner / 8); i += 256) {
            const int r = i / (kFbInner / 8), c = (i % (kFbInner / 8)) * 8;
            const uint32_t dst = qo_addr + (r * kFbQPitch + c) * 2;
            if (r < rows_valid) cp16(dst, qsrc + static_cast<size_t>(r) * p.ld_qkv + c);
            else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
        cp_commit();
        cp_wait<0>();
        __syncthreads();

        const int hs = warp >> 1, mt = warp & 1, pl = threadIdx.x & 63;    // head slot, query m-tile, lane within the pair
        const uint32_t ks_addr = region_addr + hs * (2 * kFbKeys * kFbKvPitch * 2);
        const uint32_t vs_addr = ks_addr + kFbKeys * kFbKvPitch * 2;
        const uint32_t bk_off = (((lane & 7) + (lane >> 4) * 8) * kFbKvPitch + ((lane >> 3) & 1) * 8) * 2;
        const uint32_t bv_off = (((lane & 7) + ((lane >> 3) & 1) * 8) * kFbKvPitch + (lane >> 4) * 8) * 2;
        const bf16* kv0 = p.qkv + (static_cast<size_t>(b) * p.s_in + p.kv_row0) * p.ld_qkv + kFbInner;
        for (int rd = 0; rd < 2; ++rd) {
            const int head = hs + 4 * rd;
            const uint32_t a_off = qo_addr + ((mt * 16 + (lane & 15)) * kFbQPitch + head * kFbD + (lane >> 4) * 8) * 2;
            float o[16][4];
#pragma unroll
            for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
            float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
            for (int j0 = 0; j0 < p.nk; j0 += kFbKeys) {
                bar_pair(1 + hs);                                          // both warps are done with the previous block
                const int kvalid = p.nk - j0 < kFbKeys ? p.nk - j0 : kFbKeys;
                const bf16* ksrc = kv0 + static_cast<size_t>(j0) * p.ld_qkv + head * kFbD;
                for (int i = pl; i < kFbKeys * 16; i += 64) {
                    const int r = i >> 4, c = (i & 15) * 8;
                    const uint32_t off = (r * kFbKvPitch + c) * 2;
                    if (r < kvalid) {
                        cp16(ks_addr + off, ksrc + static_cast<size_t>(r) * p.ld_qkv + c);
                        cp16(vs_addr + off, ksrc + static_cast<size_t>(r) * p.ld_qkv + kFbInner + c);
                    } else {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ks_addr + off), "r"(0u) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(vs_addr + off), "r"(0u) : "memory");
                    }
                }
                cp_commit();
                cp_wait<0>();
                bar_pair(1 + hs);
                float sc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
                for (int kk = 0; kk < kFbD / 16; ++kk) {
                    uint32_t a0, a1, a2, a3;
                    ldsm4(a_off + kk * 32, a0, a1, a2, a3);
#pragma unroll
                    for (int np = 0; np < 4; ++np) {
                        uint32_t b0, b1, b2, b3;
                        ldsm4(ks_addr + bk_off + np * 16 * kFbKvPitch * 2 + kk * 32, b0, b1, b2, b3);
                        mma16816(sc[2 * np], a0, a1, a2, a3, b0, b1);
                        mma16816(sc[2 * np + 1], a0, a1, a2, a3, b2, b3);
                    }
                }
                float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int c = nt * 8 + 2 * t;
                    if (c >= kvalid) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
                    if (c + 1 >= kvalid) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
                    bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
                    bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
                }
                bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
                bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
                const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
                const float al0 = exp2f((m0 - mn0) * p.scale_log2e), al1 = exp2f((m1 - mn1) * p.scale_log2e);
                m0 = mn0; m1 = mn1;
                float rs0 = 0.f, rs1 = 0.f;
                uint32_t pa[8][2];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float p0 = exp2f((sc[nt][0] - mn0) * p.scale_log2e), p1 = exp2f((sc[nt][1] - mn0) * p.scale_log2e);
                    const float p2 = exp2f((sc[nt][2] - mn1) * p.scale_log2e), p3 = exp2f((sc[nt][3] - mn1) * p.scale_log2e);
                    rs0 += p0 + p1; rs1 += p2 + p3;
                    pa[nt][0] = pack_bf16x2(p0, p1);
                    pa[nt][1] = pack_bf16x2(p2, p3);
                }
                rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
                rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
                l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
#pragma unroll
                for (int i = 0; i < 16; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
#pragma unroll
                for (int kk = 0; kk < kFbKeys / 16; ++kk) {
                    const uint32_t a0 = pa[2 * kk][0], a1 = pa[2 * kk][1], a2 = pa[2 * kk + 1][0], a3 = pa[2 * kk + 1][1];
#pragma unroll
                    for (int dp = 0; dp < 8; ++dp) {
                        uint32_t b0, b1, b2, b3;
                        ldsm4_trans(vs_addr + bv_off + kk * 16 * kFbKvPitch * 2 + dp * 32, b0, b1, b2, b3);
                        mma16816(o[2 * dp], a0, a1, a2, a3, b0, b1);
                        mma16816(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
                    }
                }
            }
            // normalised head output overwrites this warp's own Q rows / head columns (nobody else reads them)
            const float inv0 = 1.f / l0, inv1 = 1.f / l1;
            __syncwarp();
            bf16* orow = qo + (mt * 16 + g) * kFbQPitch + head * kFbD + 2 * t;
#pragma unroll
            for (int dt = 0; dt < 16; ++dt) {
                *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
                *reinterpret_cast<uint32_t*>(orow + 8 * kFbQPitch + dt * 8) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
            }
        }
    }
    __syncthreads();                                          // O complete; K/V blocks are dead -> weight rings

    // ---------------- phase 2: out-projection ----------------
    const int col0 = warp * 72;
    const uint32_t ring = region_addr + warp * 3 * kFbStageBytes;
    float acc[2][kFbNt][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < kFbNt; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
    warp_stream_gemm<kFbNt, 32>(acc, qo_addr, kFbQPitch * 2, p.wo + static_cast<size_t>(col0) * kFbInner, kFbInner, kFbInner, ring, lane);

    // ---------------- phase 3: + bias + residual, norm1, ff.net.0 ----------------
#pragma unroll
    for (int nt = 0; nt < kFbNt; ++nt) {
        const int c = col0 + nt * 8 + 2 * t;
        const float2 bq = __ldg(reinterpret_cast<const float2*>(p.bo + c));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
                const int r = mt * 16 + g + hi * 8;
                float2 rv = make_float2(0.f, 0.f);
                if (r < rows_valid) rv = *reinterpret_cast<const float2*>(p.res_in + (in_row0 + r) * p.pitch + c);
                acc[mt][nt][hi * 2] += bq.x + rv.x;
                acc[mt][nt][hi * 2 + 1] += bq.y + rv.y;
            }
    }
    layer_norm_tile(acc, p.g1, p.b1, col0, p.d, red, warp, lane);                      // h = norm1(out + q)
    __syncthreads();                                          // every warp has finished reading O: H may overwrite it
    {
        float mean[4], rstd[4];
        layer_norm_stats(acc, col0, p.d, red + 2 * 8 * kFbRows, warp, lane, mean, rstd);    // ff.net.0 (h itself stays in acc)
        bf16* hs = qo;                                        // H [32 x 576] bf16, pitch kFbHPitch
#pragma unroll
        for (int nt = 0; nt < kFbNt; ++nt) {
            const int c = col0 + nt * 8 + 2 * t;
            const float2 gq = __ldg(reinterpret_cast<const float2*>(p.gff + c)), bq = __ldg(reinterpret_cast<const float2*>(p.bff + c));
            const bool v0 = c < p.d, v1 = c + 1 < p.d;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                *reinterpret_cast<uint32_t*>(hs + (mt * 16 + g) * kFbHPitch + c) =
                    pack_bf16x2(ln_apply(acc[mt][nt][0], mean[mt * 2], rstd[mt * 2], gq.x, bq.x, v0),
                                ln_apply(acc[mt][nt][1], mean[mt * 2], rstd[mt * 2], gq.y, bq.y, v1));
                *reinterpret_cast<uint32_t*>(hs + (mt * 16 + g + 8) * kFbHPitch + c) =
                    pack_bf16x2(ln_apply(acc[mt][nt][2], mean[mt * 2 + 1], rstd[mt * 2 + 1], gq.x, bq.x, v0),
                                ln_apply(acc[mt][nt][3], mean[mt * 2 + 1], rstd[mt * 2 + 1], gq.y, bq.y, v1));
            }
        }
    }
    __syncthreads();

    // ---------------- phase 4: ff.net.1 + GELU (warp w owns hidden columns [16w, 16w+16)) ----------------
    bf16* fs = qo + kFbRows * kFbHPitch;                      // F [32 x 128] bf16, pitch kFbFPitch
    {
        float f1[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { f1[mt][nt][0] = f1[mt][nt][1] = f1[mt][nt][2] = f1[mt][nt][3] = 0.f; }
        warp_stream_gemm<2, 96>(f1, qo_addr, kFbHPitch * 2, p.w1 + static_cast<size_t>(warp) * 16 * kFbDp, kFbDp, kFbDp, ring, lane);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int c = warp * 16 + nt * 8 + 2 * t;
            const float2 bq = __ldg(reinterpret_cast<const float2*>(p.bf1 + c));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                *reinterpret_cast<uint32_t*>(fs + (mt * 16 + g) * kFbFPitch + c) =
                    pack_bf16x2(gelu_erf(f1[mt][nt][0] + bq.x), gelu_erf(f1[mt][nt][1] + bq.y));
                *reinterpret_cast<uint32_t*>(fs + (mt * 16 + g + 8) * kFbFPitch + c) =
                    pack_bf16x2(gelu_erf(f1[mt][nt][2] + bq.x), gelu_erf(f1[mt][nt][3] + bq.y));
            }
        }
    }
    __syncthreads();

    // ---------------- phase 5: ff.net.4 accumulated onto h + bias ----------------
#pragma unroll
    for (int nt = 0; nt < kFbNt; ++nt) {
        const int c = col0 + nt * 8 + 2 * t;
        const float2 bq = __ldg(reinterpret_cast<const float2*>(p.bf2 + c));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            acc[mt][nt][0] += bq.x; acc[mt][nt][1] += bq.y; acc[mt][nt][2] += bq.x; acc[mt][nt][3] += bq.y;
        }
    }
    warp_stream_gemm<kFbNt, 32>(acc, static_cast<uint32_t>(__cvta_generic_to_shared(fs)), kFbFPitch * 2,
                                p.w2 + static_cast<size_t>(col0) * kFbHid, kFbHid, kFbHid, ring, lane);

    // ---------------- phase 6: norm2, store fp32 master + bf16 copy ----------------
    layer_norm_tile(acc, p.g2, p.b2, col0, p.d, red + 4 * 8 * kFbRows, warp, lane);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
            const int r = mt * 16 + g + hi * 8;
            if (r >= rows_valid) continue;
            float* of = p.out_f32 + (out_row0 + r) * p.pitch;
            bf16* ol = p.out_lp + (out_row0 + r) * p.pitch;
#pragma unroll
            for (int nt = 0; nt < kFbNt; ++nt) {
                const int c = col0 + nt * 8 + 2 * t;
                *reinterpret_cast<float2*>(of + c) = make_float2(acc[mt][nt][hi * 2], acc[mt][nt][hi * 2 + 1]);
                *reinterpret_cast<uint32_t*>(ol + c) = pack_bf16x2(acc[mt][nt][hi * 2], acc[mt][nt][hi * 2 + 1]);
            }
        }
}

// ------------------------------------------------------------------------------------------------
// Small passes (B = 1: 4 CTAs for layers 0-1, ONE for layers 2-4): the kernel above is a latency chain -- every warp pulls
// its 120 KB share of the three weight matrices from L2 two 4.6 KB chunks at a time, 62 us per layer whatever the row count,
// a third of the B = 1 forward.  The same layer as a CLUSTER OF 8 CTAs per 32 query rows shortens the chain 8 x:
//   CTA c = head c of the attention = output columns [72c, 72c + 72) of to_out / ff.net.4 = hidden columns [16c, 16c + 16)
//   of ff.net.1.  Inside a CTA the 8 warps split K: warp w multiplies head w's output tile -- copied from CTA w's shared
//   memory (ld.shared::cluster) -- with W_out[72c.., 128w .. 128w + 128) (4 ring chunks instead of 32); the partial tiles
//   are summed through shared memory and re-distributed, thread = (row, 9 columns).  LayerNorm statistics are combined
//   across the cluster from (sum, M2) pairs with Chan's formula (exactly the two-pass variance), one exchange per
//   LayerNorm; the LayerNorm'd H slices and the GELU'd F slices are all-gathered through distributed shared memory.
// 7 cluster barriers per layer.  Used when the pass has <= kFbMaxClusters row blocks (HMV_FUSION_CLUSTER=0 disables).
// ------------------------------------------------------------------------------------------------
constexpr int kFcTilePitch = kFbD + 8;                                   // 136: head tiles, Q / K / V blocks, F
constexpr int kFcOtBytes = kFbHeads * kFbRows * kFcTilePitch * 2;        // 69632: the 8 gathered head-output tiles
constexpr int kFcRingBytes = 8 * 3 * kFbStageBytes;                      // 138240: per-warp weight rings
constexpr int kFcSlice = kFbDp / 8;                                      // 72 output columns per CTA
constexpr int kFcHidSlice = kFbHid / 8;                                  // 16 hidden columns per CTA
constexpr int kFcOff_ring = kFcOtBytes;
constexpr int kFcOff_opub = kFcOff_ring + kFcRingBytes;                  // this CTA's head output [32][136] bf16
constexpr int kFcOff_hpub = kFcOff_opub + kFbRows * kFcTilePitch * 2;    // this CTA's H slice [32][72] bf16
constexpr int kFcOff_fpub = kFcOff_hpub + kFbRows * kFcSlice * 2;        // this CTA's F slice [32][16] bf16
constexpr int kFcOff_stat = kFcOff_fpub + kFbRows * kFcHidSlice * 2;     // [3 LayerNorms][32 rows] (mean, M2) of the slice
constexpr int kFcOff_cvec = kFcOff_stat + 3 * kFbRows * 8;               // [8 vectors][72] fp32: bo g1 b1 gff bff bf2 g2 b2
constexpr int kFcSmem = kFcOff_cvec + 8 * kFcSlice * 4;
constexpr int kFcHPitch = kFbDp + 8;                                     // 584: gathered H [32][576] (aliases the head tiles)
constexpr int kFcOff_hfull = 0;
constexpr int kFcOff_ffull = kFbRows * kFcHPitch * 2;                    // gathered F [32][128], pitch 136
static_assert(kFcOff_ffull + kFbRows * kFcTilePitch * 2 <= kFcOtBytes, "H and F alias the head tiles");
static_assert(kFbRows * kFcTilePitch * 2 + 2 * kFbKeys * kFcTilePitch * 2 <= kFcRingBytes, "Q / K / V alias the rings");
static_assert(8 * kFbRows * kFcSlice * 4 <= kFcRingBytes, "partial tiles alias the rings");
static_assert(kFcSmem <= 227 * 1024, "shared memory");
constexpr int kFbMaxClusters = 16;

__device__ __forceinline__ uint4 ld_cluster_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_cluster_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// LayerNorm statistics of a row whose 576 columns live in 8 CTAs x 8 threads x 9 columns: every thread gets (mean, 1/std).
// Slice-local (mean, M2) -> own shared memory -> cluster barrier -> Chan's combination of the 8 slices.
__device__ __forceinline__ void cluster_row_stats(const float (&v)[9], int gc0, int d, int slice_valid, uint32_t stat_addr, float* stat,
                                                  int row, int l7, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) s += v[j];                               // pad columns hold exact zeros
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float lm = slice_valid > 0 ? s / static_cast<float>(slice_valid) : 0.f;
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const float dv = (gc0 + j) < d ? v[j] - lm : 0.f;
        m2 = fmaf(dv, dv, m2);
    }
    m2 += __shfl_xor_sync(0xffffffffu, m2, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
    if (l7 == 0) { stat[2 * row] = lm; stat[2 * row + 1] = m2; }
    cluster_sync_all();
    float lms[8], m2s[8];
    float tot = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float2 q = ld_cluster_f2(mapa_u32(stat_addr + row * 8, c));
        lms[c] = q.x; m2s[c] = q.y;
        int n = d - c * kFcSlice; n = n < 0 ? 0 : (n > kFcSlice ? kFcSlice : n);
        tot = fmaf(static_cast<float>(n), q.x, tot);
    }
    mean = tot / static_cast<float>(d);
    float m2t = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int n = d - c * kFcSlice; n = n < 0 ? 0 : (n > kFcSlice ? kFcSlice : n);
        const float dm = lms[c] - mean;
        m2t += m2s[c] + static_cast<float>(n) * dm * dm;
    }
    rstd = rsqrtf(m2t / static_cast<float>(d) + 1e-5f);
}

__global__ void __launch_bounds__(256, 1) fusion_block_cluster_kernel(const FusionBlockParams p) {
    extern __shared__ __align__(128) uint8_t fb_smem[];
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(fb_smem));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int c = static_cast<int>(cluster_ctarank());                   // head / column slice of this CTA
    const int b = blockIdx.y, q0 = static_cast<int>(blockIdx.x >> 3) * kFbRows;
    const int rows_valid = p.nq - q0 < kFbRows ? p.nq - q0 : kFbRows;
    const size_t in_row0 = static_cast<size_t>(b) * p.s_in + q0;
    const size_t out_row0 = static_cast<size_t>(b) * p.nq + q0;
    const int row = tid >> 3, l7 = tid & 7;                              // epilogue ownership: (row, 9 columns)
    const int lc0 = l7 * 9, gc0 = c * kFcSlice + lc0;                    // first owned column: in the slice / global
    int slice_valid = p.d - c * kFcSlice; slice_valid = slice_valid < 0 ? 0 : (slice_valid > kFcSlice ? kFcSlice : slice_valid);

    float* cvec = reinterpret_cast<float*>(fb_smem + kFcOff_cvec);
    {   // the layer's own vectors (weights: independent of the previous kernel) -- before the dependency wait
        const float* src[8] = {p.bo, p.g1, p.b1, p.gff, p.bff, p.bf2, p.g2, p.b2};
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (tid < kFcSlice) cvec[i * kFcSlice + tid] = __ldg(src[i] + c * kFcSlice + tid);
    }
    pdl_wait();
    pdl_launch_dependents();

    float res[9];                                                        // residual: fp32 master copy of the layer input
#pragma unroll
    for (int j = 0; j < 9; ++j) res[j] = row < rows_valid ? p.res_in[(in_row0 + row) * p.pitch + gc0 + j] : 0.f;

    // ---------------- phase 1: attention of head c (warps 0 / 1 = query m-tiles; all threads load) ----------------
    const uint32_t region = base + kFcOff_ring;
    const uint32_t qs_addr = region, ks_addr = region + kFbRows * kFcTilePitch * 2, vs_addr = ks_addr + kFbKeys * kFcTilePitch * 2;
    {
        const bf16* qsrc = p.qkv + in_row0 * p.ld_qkv + c * kFbD;
        for (int i = tid; i < kFbRows * 16; i += 256) {
            const int r = i >> 4, cc = (i & 15) * 8;
            const uint32_t dst = qs_addr + (r * kFcTilePitch + cc) * 2;
            if (r < rows_valid) cp16(dst, qsrc + static_cast<size_t>(r) * p.ld_qkv + cc);
            else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
        const int mt = warp & 1;
        const uint32_t a_off = qs_addr + ((mt * 16 + (lane & 15)) * kFcTilePitch + (lane >> 4) * 8) * 2;
        const uint32_t bk_off = (((lane & 7) + (lane >> 4) * 8) * kFcTilePitch + ((lane >> 3) & 1) * 8) * 2;
        const uint32_t bv_off = (((lane & 7) + ((lane >> 3) & 1) * 8) * kFcTilePitch + (lane >> 4) * 8) * 2;
        const bf16* kv0 = p.qkv + (static_cast<size_t>(b) * p.s_in + p.kv_row0) * p.ld_qkv + kFbInner + c * kFbD;
        float o[16][4];
#pragma unroll
        for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int j0 = 0; j0 < p.nk; j0 += kFbKeys) {
            __syncthreads();                                             // the previous block has been consumed
            const int kvalid = p.nk - j0 < kFbKeys ? p.nk - j0 : kFbKeys;
            const bf16* ksrc = kv0 + static_cast<size_t>(j0) * p.ld_qkv;
            for (int i = tid; i < kFbKeys * 16; i += 256) {
                const int r = i >> 4, cc = (i & 15) * 8;
                const uint32_t off = (r * kFcTilePitch + cc) * 2;
                if (r < kvalid) {
                    cp16(ks_addr + off, ksrc + static_cast<size_t>(r) * p.ld_qkv + cc);
                    cp16(vs_addr + off, ksrc + static_cast<size_t>(r) * p.ld_qkv + kFbInner + cc);
                } else {
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ks_addr + off), "r"(0u) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(vs_addr + off), "r"(0u) : "memory");
                }
            }
            cp_commit();
            cp_wait<0>();
            __syncthreads();
            if (warp < 2) {
                float sc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
                for (int kk = 0; kk < kFbD / 16; ++kk) {
                    uint32_t a0, a1, a2, a3;
                    ldsm4(a_off + kk * 32, a0, a1, a2, a3);
#pragma unroll
                    for (int np = 0; np < 4; ++np) {
                        uint32_t b0, b1, b2, b3;
                        ldsm4(ks_addr + bk_off + np * 16 * kFcTilePitch * 2 + kk * 32, b0, b1, b2, b3);
                        mma16816(sc[2 * np], a0, a1, a2, a3, b0, b1);
                        mma16816(sc[2 * np + 1], a0, a1, a2, a3, b2, b3);
                    }
                }
                float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int cc = nt * 8 + 2 * t;
                    if (cc >= kvalid) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
                    if (cc + 1 >= kvalid) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
                    bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
                    bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
                }
                bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
                bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
                const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
                const float al0 = exp2f((m0 - mn0) * p.scale_log2e), al1 = exp2f((m1 - mn1) * p.scale_log2e);
                m0 = mn0; m1 = mn1;
                float rs0 = 0.f, rs1 = 0.f;
                uint32_t pa[8][2];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float p0 = exp2f((sc[nt][0] - mn0) * p.scale_log2e), p1 = exp2f((sc[nt][1] - mn0) * p.scale_log2e);
                    const float p2 = exp2f((sc[nt][2] - mn1) * p.scale_log2e), p3 = exp2f((sc[nt][3] - mn1) * p.scale_log2e);
                    rs0 += p0 + p1; rs1 += p2 + p3;
                    pa[nt][0] = pack_bf16x2(p0, p1);
                    pa[nt][1] = pack_bf16x2(p2, p3);
                }
                rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
                rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
                l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
#pragma unroll
                for (int i = 0; i < 16; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
#pragma unroll
                for (int kk = 0; kk < kFbKeys / 16; ++kk) {
                    const uint32_t a0 = pa[2 * kk][0], a1 = pa[2 * kk][1], a2 = pa[2 * kk + 1][0], a3 = pa[2 * kk + 1][1];
#pragma unroll
                    for (int dp = 0; dp < 8; ++dp) {
                        uint32_t b0, b1, b2, b3;
                        ldsm4_trans(vs_addr + bv_off + kk * 16 * kFcTilePitch * 2 + dp * 32, b0, b1, b2, b3);
                        mma16816(o[2 * dp], a0, a1, a2, a3, b0, b1);
                        mma16816(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
                    }
                }
            }
        }
        if (warp < 2) {                                                  // normalised head output -> the tile the cluster reads
            const float inv0 = 1.f / l0, inv1 = 1.f / l1;
            bf16* orow = reinterpret_cast<bf16*>(fb_smem + kFcOff_opub) + (mt * 16 + g) * kFcTilePitch + 2 * t;
#pragma unroll
            for (int dt = 0; dt < 16; ++dt) {
                *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
                *reinterpret_cast<uint32_t*>(orow + 8 * kFcTilePitch + dt * 8) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
            }
        }
    }
    cluster_sync_all();                                       // B1: every head output is published; Q / K / V are dead -> rings

    // ---------------- phase 2: out-projection, K split over the warps (warp w <- head w's tile from CTA w) ----------------
    const uint32_t ring = region + warp * 3 * kFbStageBytes;
    float* part = reinterpret_cast<float*>(fb_smem + kFcOff_ring);       // [8 warps][32][72] fp32 partial tiles (alias the rings)
    {
        const uint32_t src = mapa_u32(base + kFcOff_opub, static_cast<uint32_t>(warp));
        const uint32_t dst = base + warp * (kFbRows * kFcTilePitch * 2);
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int idx = lane + 32 * i, r = idx >> 4, pc = idx & 15;
            const uint32_t off = (r * kFcTilePitch + pc * 8) * 2;
            st_shared_v4(dst + off, ld_cluster_v4(src + off));
        }
        __syncwarp();
        float acc[2][kFbNt][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < kFbNt; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
        warp_stream_gemm<kFbNt, 32>(acc, dst, kFcTilePitch * 2, p.wo + static_cast<size_t>(c * kFcSlice) * kFbInner + warp * kFbD, kFbInner, kFbD, ring, lane);
        __syncthreads();                                      // every ring is idle: the partial tiles may overwrite them
        float* pw = part + warp * (kFbRows * kFcSlice);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < kFbNt; ++nt) {
                *reinterpret_cast<float2*>(pw + (mt * 16 + g) * kFcSlice + nt * 8 + 2 * t) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<float2*>(pw + (mt * 16 + g + 8) * kFcSlice + nt * 8 + 2 * t) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            }
        __syncthreads();
    }

    // ---------------- phase 3: + bias + residual, norm1, ff.net.0 ----------------
    float* stat = reinterpret_cast<float*>(fb_smem + kFcOff_stat);
    const uint32_t stat_addr = base + kFcOff_stat;
    float h[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += part[w * (kFbRows * kFcSlice) + row * kFcSlice + lc0 + j];
        h[j] = a + cvec[0 * kFcSlice + lc0 + j] + res[j];
    }
    float mean, rstd;
    cluster_row_stats(h, gc0, p.d, slice_valid, stat_addr, stat, row, l7, mean, rstd);                       // B2
#pragma unroll
    for (int j = 0; j < 9; ++j)
        h[j] = ln_apply(h[j], mean, rstd, cvec[1 * kFcSlice + lc0 + j], cvec[2 * kFcSlice + lc0 + j], gc0 + j < p.d);    // h = norm1(out + q)
    cluster_row_stats(h, gc0, p.d, slice_valid, stat_addr + kFbRows * 8, stat + 2 * kFbRows, row, l7, mean, rstd);   // B3
    {
        bf16* hp = reinterpret_cast<bf16*>(fb_smem + kFcOff_hpub) + row * kFcSlice + lc0;
#pragma unroll
        for (int j = 0; j < 9; ++j)
            hp[j] = __float2bfloat16(ln_apply(h[j], mean, rstd, cvec[3 * kFcSlice + lc0 + j], cvec[4 * kFcSlice + lc0 + j], gc0 + j < p.d));
    }
    cluster_sync_all();                                       // B4: every H slice is published

    // ---------------- phase 4: gather H, ff.net.1 + GELU (hidden columns [16c, 16c + 16); warps 0-5 split K = 6 x 96) ----------------
    for (int idx = tid; idx < 8 * kFbRows * 9; idx += 256) {
        const int w = idx / (kFbRows * 9), rem = idx - w * (kFbRows * 9), r = rem / 9, pc = rem - r * 9;
        const uint4 v = ld_cluster_v4(mapa_u32(base + kFcOff_hpub + (r * kFcSlice + pc * 8) * 2, static_cast<uint32_t>(w)));
        st_shared_v4(base + kFcOff_hfull + (r * kFcHPitch + w * kFcSlice + pc * 8) * 2, v);
    }
    __syncthreads();
    {
        float f1[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { f1[mt][nt][0] = f1[mt][nt][1] = f1[mt][nt][2] = f1[mt][nt][3] = 0.f; }
        if (warp < 6)
            warp_stream_gemm<2, 96>(f1, base + kFcOff_hfull + warp * 96 * 2, kFcHPitch * 2,
                                    p.w1 + static_cast<size_t>(c * kFcHidSlice) * kFbDp + warp * 96, kFbDp, 96, ring, lane);
        __syncthreads();
        if (warp < 6) {
            float* pw = part + warp * (kFbRows * kFcHidSlice);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    *reinterpret_cast<float2*>(pw + (mt * 16 + g) * kFcHidSlice + nt * 8 + 2 * t) = make_float2(f1[mt][nt][0], f1[mt][nt][1]);
                    *reinterpret_cast<float2*>(pw + (mt * 16 + g + 8) * kFcHidSlice + nt * 8 + 2 * t) = make_float2(f1[mt][nt][2], f1[mt][nt][3]);
                }
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < 6; ++w) {
            const float2 q = *reinterpret_cast<const float2*>(part + w * (kFbRows * kFcHidSlice) + row * kFcHidSlice + l7 * 2);
            a0 += q.x; a1 += q.y;
        }
        const float2 bq = __ldg(reinterpret_cast<const float2*>(p.bf1 + c * kFcHidSlice + l7 * 2));
        *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(fb_smem + kFcOff_fpub) + row * kFcHidSlice + l7 * 2) =
            pack_bf16x2(gelu_erf(a0 + bq.x), gelu_erf(a1 + bq.y));
    }
    cluster_sync_all();                                       // B5: every F slice is published

    // ---------------- phase 5: gather F, ff.net.4 (warps 0-3 split K = 4 x 32) accumulated onto h + bias ----------------
    for (int idx = tid; idx < 8 * kFbRows * 2; idx += 256) {
        const int w = idx >> 6, rem = idx & 63, r = rem >> 1, pc = rem & 1;
        const uint4 v = ld_cluster_v4(mapa_u32(base + kFcOff_fpub + (r * kFcHidSlice + pc * 8) * 2, static_cast<uint32_t>(w)));
        st_shared_v4(base + kFcOff_ffull + (r * kFcTilePitch + w * kFcHidSlice + pc * 8) * 2, v);
    }
    __syncthreads();
    {
        float acc[2][kFbNt][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < kFbNt; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
        if (warp < 4)
            warp_stream_gemm<kFbNt, 32>(acc, base + kFcOff_ffull + warp * 32 * 2, kFcTilePitch * 2,
                                        p.w2 + static_cast<size_t>(c * kFcSlice) * kFbHid + warp * 32, kFbHid, 32, ring, lane);
        __syncthreads();
        if (warp < 4) {
            float* pw = part + warp * (kFbRows * kFcSlice);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < kFbNt; ++nt) {
                    *reinterpret_cast<float2*>(pw + (mt * 16 + g) * kFcSlice + nt * 8 + 2 * t) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                    *reinterpret_cast<float2*>(pw + (mt * 16 + g + 8) * kFcSlice + nt * 8 + 2 * t) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) a += part[w * (kFbRows * kFcSlice) + row * kFcSlice + lc0 + j];
        h[j] += cvec[5 * kFcSlice + lc0 + j] + a;
    }

    // ---------------- phase 6: norm2, store fp32 master + bf16 copy (pad columns written as zero) ----------------
    cluster_row_stats(h, gc0, p.d, slice_valid, stat_addr + 2 * kFbRows * 8, stat + 4 * kFbRows, row, l7, mean, rstd);   // B6
    if (row < rows_valid) {
        float* of = p.out_f32 + (out_row0 + row) * p.pitch + gc0;
        bf16* ol = p.out_lp + (out_row0 + row) * p.pitch + gc0;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float y = ln_apply(h[j], mean, rstd, cvec[6 * kFcSlice + lc0 + j], cvec[7 * kFcSlice + lc0 + j], gc0 + j < p.d);
            of[j] = y;
            ol[j] = __float2bfloat16(y);
        }
    }
    cluster_sync_all();                                       // B7: nobody reads this CTA's shared memory any more
}

}  // namespace

int fusion_block_launch(const FusionBlockParams& p, int batch, cudaStream_t s) {
    if (batch == 0) return 0;
    HMV_CHECK(p.pitch == kFbDp && p.d <= kFbDp && p.d % 2 == 0 && p.ld_qkv == 3 * kFbInner, "fusion block: d_model must be <= 576 (even) with 8 heads of 128");
    HMV_CHECK(p.nq > 0 && p.nk > 0, "fusion block: empty attention");
    static unsigned long long configured = 0;
    if (first_use_on_this_device(configured)) {
        HMV_CUDA(cudaFuncSetAttribute(fusion_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFbSmem));
        HMV_CUDA(cudaFuncSetAttribute(fusion_block_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFcSmem));
    }
    const int row_blocks = (p.nq + kFbRows - 1) / kFbRows;
    static const bool cluster_env = [] { const char* e = getenv("HMV_FUSION_CLUSTER"); return !(e && e[0] == '0'); }();
    if (cluster_env && clusters_enabled() && row_blocks * batch <= kFbMaxClusters) {
        HMV_CUDA(launch_kernel_cluster(fusion_block_cluster_kernel, dim3(8 * row_blocks, batch), dim3(256), 8, kFcSmem, s, p));
        HMV_CUDA(cudaGetLastError());
        return 0;
    }
    HMV_CUDA(launch_kernel(fusion_block_kernel, dim3(row_blocks, batch), dim3(256), kFbSmem, s, p));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hmv
