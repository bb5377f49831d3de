// Fused stem of the ResNet-50 "paper" backbone for the bf16 path (reference backbones/resnet.py:218-221):
//   x fp32 NCHW [n,3,256,256] -> conv 7x7 / stride 2 / pad 3 (3 -> 64, BatchNorm folded) -> ReLU
//     -> max-pool 3x3 / stride 2 / pad 1 -> bf16 NHWC [n,64,64,64]
// in ONE kernel: the 128x128x64 conv map (2 MB / image in bf16) never touches HBM and the input is read as fp32.
//
// A work item is (image, strip of `strip` pooled rows = 2 * strip + 1 conv rows; 16 for large passes, shorter strips when a
// small pass would otherwise leave most SMs idle).  Per conv row (128 pixels x 64 channels):
//   * converter warps keep a rolling ring of zero-padded input rows in shared memory as bf16 NHWC4
//     (8 bytes / pixel, row pitch 2176 B) - two new rows per conv row;
//   * the MMA warp issues 7 taps x 2 tcgen05.mma (M=128, N=64, K=16).  The A operand of tap r is the RAW padded
//     input row 2*oh + r: with the no-swizzle K-major descriptor (core matrix = 8 rows x 16 B) a row pitch of
//     16 bytes (= 2 pixels = the conv stride) and a K-chunk distance of 16 bytes make row m of the operand the
//     8-pixel window starting at pixel 2*m - no im2col copy exists anywhere.  Weights [64][7][8px][4ch] stay
//     resident in shared memory;
//   * 8 epilogue warps read the accumulator from TMEM, add the folded-BN bias, apply ReLU and park the row in a
//     4-row shared-memory ring; every second row they emit one pooled row (post-ReLU values are >= 0, so the
//     pooling padding can be 0).
#include "conv_gemm_tc.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace hmv {

namespace {

constexpr int kStemThreads = 416;          // warps 0-3 converters, 4-11 epilogue, 12 MMA issuer + TMEM owner
constexpr int kImg = 256, kConv = 128, kPool = 64, kC = 64;
constexpr int kRowPitch = 2176;            // 272 padded pixels x 8 B
constexpr int kPairBytes = 2 * kRowPitch;  // two input rows per ring slot
constexpr int kRingSlots = 8;
constexpr int kWBytes = 7 * 4 * kC * 16;   // 7 taps x 4 K-chunks x 64 couts x 16 B = 28 KiB
constexpr int kRowBufBytes = kConv * kC * 2;   // one ReLU'd conv row, bf16 [128 px][64 ch] = 16 KiB
constexpr int kRowBufs = 4;
constexpr int kStripMax = 16;              // pooled rows per work item (large passes); small passes use shorter strips
constexpr int kSmemBytes = kWBytes + kRingSlots * kPairBytes + kRowBufs * kRowBufBytes + 256 + 1024;

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint4 max4(const uint4& a, const uint4& b) {
    return make_uint4(bf16x2_max(a.x, b.x), bf16x2_max(a.y, b.y), bf16x2_max(a.z, b.z), bf16x2_max(a.w, b.w));
}

// U8 = true: x is uint8 NCHW and is normalised on the fly as ((v / 255) - mean[c]) / std[c] (torchvision ToTensor +
// Normalize, reference datasets/ho3d.py:35-40), with the same IEEE operations in the same order as the host transform.
template <bool U8>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_pool_kernel(const void* __restrict__ x_raw, const bf16* __restrict__ wpack, const float* __restrict__ bias,
                 bf16* __restrict__ out, int n_img, int* err_flag, StemNorm norm, int strip /* pooled rows per work item, divides 64 */) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* wsm = smem;                                   // resident weights
    uint8_t* ring = wsm + kWBytes;                         // input row pairs
    uint8_t* rowbuf = ring + kRingSlots * kPairBytes;      // ReLU'd conv rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(rowbuf + kRowBufs * kRowBufBytes);
    const uint32_t pfull0 = smem_u32(bars);                // [8] pair written by the converters
    const uint32_t pempty0 = pfull0 + 8 * kRingSlots;      // [8] pair no longer needed by the tensor core
    const uint32_t afull0 = pempty0 + 8 * kRingSlots;      // [2] accumulator ready
    const uint32_t aempty0 = afull0 + 16;                  // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRingSlots + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

    // ---- one-time setup (overlaps the previous kernel under PDL) ----
    for (int i = tid; i < kRingSlots * kPairBytes / 16; i += kStemThreads)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);          // horizontal padding pixels stay zero
    if (tid == 0) {
        for (int i = 0; i < kRingSlots; ++i) {
            mbar_init(pfull0 + 8 * i, 1);                  // the converter warp that owns the pair
            mbar_init(pempty0 + 8 * i, 1);                 // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(afull0 + 8 * i, 1);
            mbar_init(aempty0 + 8 * i, 8);                 // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    for (int i = tid; i < kWBytes / 16; i += kStemThreads)                      // weights: global -> smem, once
        reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const int num_items = n_img * (kPool / strip);

    if (warp < 4) {
        // ===================== converters: fp32 NCHW rows -> bf16 NHWC4 padded rows =====================
        // Warp w owns every 4th row pair (whole rows: lane -> 8 pixels), so four pairs' global loads are in flight.
        uint32_t gq = 0;                                   // running pair counter (ring position / phase)
        bool alive = true;
        for (int item = blockIdx.x; item < num_items && alive; item += gridDim.x) {
            const int n = item / (kPool / strip), p0 = (item % (kPool / strip)) * strip;
            const int c_lo = p0 > 0 ? 2 * p0 - 1 : 0, c_hi = 2 * p0 + 2 * strip - 1;
            const size_t img_off = static_cast<size_t>(n) * 3 * kImg * kImg;
            for (int j = c_lo; j <= c_hi + 3 && alive; ++j, ++gq) {              // pair j = padded rows 2j, 2j+1
                if ((gq & 3) != static_cast<uint32_t>(warp)) continue;
                const uint32_t slot = gq % kRingSlots;
                float4 v[2][3][2];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int row = 2 * j + rr - 3;                              // original input row
                    const bool ok = row >= 0 && row < kImg;
                    const size_t off = img_off + static_cast<size_t>(ok ? row : 0) * kImg + 8 * lane;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        if constexpr (U8) {
                            uint2 q = make_uint2(0, 0);
                            if (ok) q = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(x_raw) + off + ch * kImg * kImg));
                            const float m = norm.mean[ch], sd = norm.std[ch];
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float b = static_cast<float>(((e < 4 ? q.x : q.y) >> (8 * (e & 3))) & 0xffu);
                                f[e] = ok ? __fdiv_rn(__fsub_rn(__fdiv_rn(b, 255.f), m), sd) : 0.f;     // padding rows are zeros AFTER normalisation
                            }
                            v[rr][ch][0] = make_float4(f[0], f[1], f[2], f[3]);
                            v[rr][ch][1] = make_float4(f[4], f[5], f[6], f[7]);
                        } else {
                            const float* pr = static_cast<const float*>(x_raw) + off;
#pragma unroll
                            for (int hq = 0; hq < 2; ++hq)
                                v[rr][ch][hq] = ok ? __ldg(reinterpret_cast<const float4*>(pr + ch * kImg * kImg) + hq)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                if (!mbar_wait(pempty0 + 8 * slot, ((gq / kRingSlots) & 1) ^ 1, err_flag, 11)) { alive = false; break; }
                uint8_t* dst = ring + slot * kPairBytes;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    uint2* d = reinterpret_cast<uint2*>(dst + rr * kRowPitch + (8 * lane + 3) * 8);   // +3: left padding
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const float c0[4] = {v[rr][0][hq].x, v[rr][0][hq].y, v[rr][0][hq].z, v[rr][0][hq].w};
                        const float c1[4] = {v[rr][1][hq].x, v[rr][1][hq].y, v[rr][1][hq].z, v[rr][1][hq].w};
                        const float c2[4] = {v[rr][2][hq].x, v[rr][2][hq].y, v[rr][2][hq].z, v[rr][2][hq].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            uint2 o;
                            o.x = pack_bf16x2(c0[e], c1[e]);
                            o.y = pack_bf16x2(c2[e], 0.f);
                            d[hq * 4 + e] = o;
                        }
                    }
                }
                fence_async_smem();                        // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull0 + 8 * slot);
            }
        }
    } else if (warp == 12) {
        // ===================== MMA issuer =====================
        {   // the whole warp walks the schedule and waits; one elected lane issues (see elect_one() in tc_ptx.cuh for why)
            constexpr uint32_t idesc = make_idesc_mn(128, kC);
            const uint32_t ring_addr = smem_u32(ring), w_addr = smem_u32(wsm);
            uint32_t gq = 0, gt = 0;                       // pair counter at item start, running conv-row counter
            bool alive = true;
            for (int item = blockIdx.x; item < num_items && alive; item += gridDim.x) {
                const int p0 = (item % (kPool / strip)) * strip;
                const int c_lo = p0 > 0 ? 2 * p0 - 1 : 0, c_hi = 2 * p0 + 2 * strip - 1;
                const int nrows = c_hi - c_lo + 1;
                for (int t = 0; t < nrows && alive; ++t, ++gt) {
                    const uint32_t acc = gt & 1;
                    if (!mbar_wait(aempty0 + 8 * acc, ((gt >> 1) & 1) ^ 1, err_flag, 12)) { alive = false; break; }
                    for (int q = (t == 0 ? 0 : 3); q < 4; ++q) {                 // only pair t+3 is new after the first row
                        const uint32_t pq = gq + t + q;
                        if (!mbar_wait(pfull0 + 8 * (pq % kRingSlots), (pq / kRingSlots) & 1, err_flag, 13)) { alive = false; break; }
                    }
                    if (!alive) break;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kC;
                    if (elect_one()) {
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            const uint32_t pq = gq + t + (r >> 1);
                            const uint32_t row_addr = ring_addr + (pq % kRingSlots) * kPairBytes + (r & 1) * kRowPitch;
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const uint64_t adesc = make_nosw_desc(row_addr + kk * 32, 16, 128);
                                const uint64_t bdesc = make_nosw_desc(w_addr + (r * 4 + kk * 2) * (kC * 16), kC * 16, 128);
                                umma_f16(d_tmem, adesc, bdesc, idesc, (r | kk) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(pempty0 + 8 * ((gq + t) % kRingSlots));          // pair t is not needed by later rows
                        if (t == nrows - 1)
                            for (int q = 1; q < 4; ++q) umma_commit(pempty0 + 8 * ((gq + t + q) % kRingSlots));
                        umma_commit(afull0 + 8 * acc);
                    }
                }
                gq += nrows + 3;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue + pooling: 8 warps = 4 TMEM lane quarters x 2 channel halves =====================
        const int ew = warp - 4, quarter = warp & 3, half = ew >> 2;
        const int et = ew * 32 + lane;                     // 0..255 among the epilogue threads
        const int px = quarter * 32 + lane;                // conv pixel owned by this thread
        float4 bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(bias + half * 32) + j);
        uint32_t gt = 0;
        bool alive = true;
        for (int item = blockIdx.x; item < num_items && alive; item += gridDim.x) {
            const int n = item / (kPool / strip), p0 = (item % (kPool / strip)) * strip;
            const int c_lo = p0 > 0 ? 2 * p0 - 1 : 0, c_hi = 2 * p0 + 2 * strip - 1;
            for (int c = c_lo; c <= c_hi; ++c, ++gt) {
                const uint32_t acc = gt & 1;
                if (alive && !mbar_wait(afull0 + 8 * acc, (gt >> 1) & 1, err_flag, 14)) alive = false;
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kC + half * 32, r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(aempty0 + 8 * acc);
                uint8_t* dst = rowbuf + (c & 3) * kRowBufBytes + px * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 v0 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])), make_float2(bq[2 * j].x, bq[2 * j].y));
                    const float2 v1 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), make_float2(bq[2 * j].z, bq[2 * j].w));
                    const float2 v2 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), make_float2(bq[2 * j + 1].x, bq[2 * j + 1].y));
                    const float2 v3 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])), make_float2(bq[2 * j + 1].z, bq[2 * j + 1].w));
                    uint4 o;
                    o.x = cvt_bf16x2(v0.x, v0.y, true); o.y = cvt_bf16x2(v1.x, v1.y, true);
                    o.z = cvt_bf16x2(v2.x, v2.y, true); o.w = cvt_bf16x2(v3.x, v3.y, true);
                    *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(half * 4 + j) ^ (px & 7)) << 4)) = o;   // chunk swizzle
                }
                named_bar_sync(1, 256);                    // conv row c is complete in shared memory
                const int p = (c - 1) >> 1;
                if ((c & 1) && p >= p0) {                  // rows c-2, c-1, c -> pooled row p (row 2*p0-1 only feeds p0)
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int idx = et + it * 256;     // 64 pooled pixels x 8 channel groups
                        const int q = idx >> 3, gch = idx & 7;
                        uint4 m = make_uint4(0, 0, 0, 0);  // post-ReLU values are >= 0: zero is the pooling identity
#pragma unroll
                        for (int dr = 0; dr < 3; ++dr) {
                            const int cr = c - 2 + dr;
                            if (cr < c_lo) continue;       // conv row -1 (image top); strips below start at 2*p0-1
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const int cp = 2 * q - 1 + dx;
                                if (cp < 0) continue;
                                const uint8_t* src = rowbuf + (cr & 3) * kRowBufBytes + cp * 128 + ((static_cast<uint32_t>(gch) ^ (cp & 7)) << 4);
                                m = max4(m, *reinterpret_cast<const uint4*>(src));
                            }
                        }
                        *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * kPool + p) * kPool + q) * kC + gch * 8) = m;
                    }
                }
            }
            named_bar_sync(1, 256);                        // pooled reads done before the next item reuses the row ring
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

}  // namespace

// wpack: [7 taps][4 K-chunks][64 couts][8 elements] bf16 with element e of chunk kc = (pixel kc*2 + e/4, channel e%4)
int stem_pool_launch(const void* x, bool x_is_u8, const StemNorm& norm, const bf16* wpack, const float* bias, bf16* out, int n_img,
                     int num_sms, int* err_flag, cudaStream_t s) {
    if (n_img == 0) return 0;
    static unsigned long long configured = 0;
    if (first_use_on_this_device(configured)) {
        HMV_CUDA(cudaFuncSetAttribute(stem_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        HMV_CUDA(cudaFuncSetAttribute(stem_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    // Work items are (image, strip of pooled rows).  16-row strips (33 conv rows for 32 needed) for large passes; a small pass
    // takes the longest strip that still gives at least half a wave of CTAs (B = 1: 5 images x 16 strips of 4 rows instead of
    // 20 CTAs walking 33 conv rows each, 39 us of the forward) at the price of one extra conv row per strip.
    int strip = kStripMax;
    while (strip > 2 && n_img * (kPool / strip) * 2 < num_sms) strip >>= 1;
    const int items = n_img * (kPool / strip);
    const dim3 grid(items < num_sms ? items : num_sms);
    if (x_is_u8)
        HMV_CUDA(launch_kernel(stem_pool_kernel<true>, grid, dim3(kStemThreads), kSmemBytes, s, x, wpack, bias, out, n_img, err_flag, norm, strip));
    else
        HMV_CUDA(launch_kernel(stem_pool_kernel<false>, grid, dim3(kStemThreads), kSmemBytes, s, x, wpack, bias, out, n_img, err_flag, norm, strip));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

// uint8 NCHW -> normalised fp32 NCHW (fp32 check mode of the uint8 entry points)
__global__ void u8_to_f32_norm_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t total, int hw, StemNorm norm) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = static_cast<int>((i / hw) % 3);
    out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(in[i]), 255.f), norm.mean[ch]), norm.std[ch]);
}
int u8_to_f32_norm_launch(const uint8_t* in, float* out, size_t total, int hw, const StemNorm& norm, cudaStream_t s) {
    if (total == 0) return 0;
    u8_to_f32_norm_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, total, hw, norm);
    HMV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hmv
