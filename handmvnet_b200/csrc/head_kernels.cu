// Memory-/latency-bound kernels around the GEMMs: input packing, max-pool, soft-argmax,
// bilinear neighbour gather, token assembly (+ positional encodings), attention core,
// LayerNorm and the Chebyshev graph-conv head.  Templated on the activation type
// (bf16 on the tensor-core path, float in fp32 check mode).
#include "kernels.cuh"

namespace hmv {

// ------------------------------------------------------------------------------------------------
// input packing: fp32 NCHW -> padded NHWC4
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ x, T* __restrict__ out, int H, int W, int Hp, int Wp,
                                  int pad, size_t total) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int wp = static_cast<int>(idx % Wp);
    const int hp = static_cast<int>((idx / Wp) % Hp);
    const size_t n = idx / (static_cast<size_t>(Wp) * Hp);
    const int h = hp - pad, w = wp - pad;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) {
        const float* px = x + (n * 3 * H + h) * W + w;
        v0 = __ldg(px);
        v1 = __ldg(px + static_cast<size_t>(H) * W);
        v2 = __ldg(px + 2 * static_cast<size_t>(H) * W);
    }
    if constexpr (sizeof(T) == 2) {
        uint2 q;
        q.x = pack_bf16x2(v0, v1);
        q.y = pack_bf16x2(v2, 0.f);
        reinterpret_cast<uint2*>(out)[idx] = q;
    } else {
        reinterpret_cast<float4*>(out)[idx] = make_float4(v0, v1, v2, 0.f);
    }
}

template <typename T>
int pack_input_launch(const float* x, T* out, int n_img, int H, int W, int Hp, int Wp, int pad, cudaStream_t s) {
    const size_t total = static_cast<size_t>(n_img) * Hp * Wp;
    if (total == 0) return 0;
    HMV_CUDA(launch_kernel(pack_input_kernel<T>, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, x, out, H, W, Hp, Wp, pad, total));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int pack_input_launch<bf16>(const float*, bf16*, int, int, int, int, int, int, cudaStream_t);
template int pack_input_launch<float>(const float*, float*, int, int, int, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// max-pool 3x3 / 2 / pad 1, NHWC, 16-byte vectors along C
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int Hin, int Win, int Hout, int Wout,
                               int C, size_t total) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int VEC = 16 / sizeof(T);
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = C / VEC;
    const int c = static_cast<int>(idx % cv) * VEC;
    const int ow = static_cast<int>((idx / cv) % Wout);
    const int oh = static_cast<int>((idx / (static_cast<size_t>(cv) * Wout)) % Hout);
    const size_t n = idx / (static_cast<size_t>(cv) * Wout * Hout);
    float m[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int ih = oh * 2 - 1 + r;
        if (ih < 0 || ih >= Hin) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int iw = ow * 2 - 1 + s;
            if (iw < 0 || iw >= Win) continue;
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((n * Hin + ih) * Win + iw) * C + c));
            if constexpr (sizeof(T) == 2) {
                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack_bf16x2(w4[i]);
                    m[2 * i] = fmaxf(m[2 * i], f.x);
                    m[2 * i + 1] = fmaxf(m[2 * i + 1], f.y);
                }
            } else {
                m[0] = fmaxf(m[0], __uint_as_float(q.x)); m[1] = fmaxf(m[1], __uint_as_float(q.y));
                m[2] = fmaxf(m[2], __uint_as_float(q.z)); m[3] = fmaxf(m[3], __uint_as_float(q.w));
            }
        }
    }
    uint4 o;
    if constexpr (sizeof(T) == 2) {
        o.x = pack_bf16x2(m[0], m[1]); o.y = pack_bf16x2(m[2], m[3]);
        o.z = pack_bf16x2(m[4], m[5]); o.w = pack_bf16x2(m[6], m[7]);
    } else {
        o.x = __float_as_uint(m[0]); o.y = __float_as_uint(m[1]); o.z = __float_as_uint(m[2]); o.w = __float_as_uint(m[3]);
    }
    *reinterpret_cast<uint4*>(out + ((n * Hout + oh) * Wout + ow) * C + c) = o;
}

template <typename T>
int maxpool_launch(const T* in, T* out, int n_img, int Hin, int Win, int C, cudaStream_t s) {
    constexpr int VEC = 16 / sizeof(T);
    HMV_CHECK(C % VEC == 0, "maxpool: C must be a multiple of the 16-byte vector");
    const int Hout = (Hin + 2 - 3) / 2 + 1, Wout = (Win + 2 - 3) / 2 + 1;
    const size_t total = static_cast<size_t>(n_img) * Hout * Wout * (C / VEC);
    if (total == 0) return 0;
    HMV_CUDA(launch_kernel(maxpool_kernel<T>, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, in, out, Hin, Win, Hout, Wout, C, total));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int maxpool_launch<bf16>(const bf16*, bf16*, int, int, int, int, cudaStream_t);
template int maxpool_launch<float>(const float*, float*, int, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// soft-argmax: one warp per heatmap, fp32 throughout
// ------------------------------------------------------------------------------------------------
__global__ void softargmax_kernel(const float* __restrict__ hm, float* __restrict__ xy, float* __restrict__ xy_scaled,
                                  int n_maps, int H, int W, float temperature, float scale) {
    pdl_wait();
    pdl_launch_dependents();
    const int map = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (map >= n_maps) return;
    const int hw = H * W;
    const float4* p = reinterpret_cast<const float4*>(hm + static_cast<size_t>(map) * hw);
    const int nvec = hw >> 2;
    float mx = -INFINITY;
    for (int i = lane; i < nvec; i += 32) {
        const float4 q = __ldg(p + i);
        // __fmul_rn: the product must be ROUNDED exactly as in the second pass (an FMA-contracted `v * T - mx` keeps the
        // exact product and leaves a residual of up to half an ulp of mx - thousands when the heat-map is ~1e8 as with
        // random-init HRNet weights - in the exponent of the maximum element: exp -> inf -> NaN)
        mx = fmaxf(mx, fmaxf(fmaxf(__fmul_rn(q.x, temperature), __fmul_rn(q.y, temperature)),
                             fmaxf(__fmul_rn(q.z, temperature), __fmul_rn(q.w, temperature))));
    }
    mx = warp_max(mx);
    float se = 0.f, sx = 0.f, sy = 0.f;
    for (int i = lane; i < nvec; i += 32) {
        const float4 q = __ldg(p + i);
        const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = i * 4 + k;
            const float e = expf(__fsub_rn(__fmul_rn(v[k], temperature), mx));
            se += e;
            sx += e * static_cast<float>(idx % W);
            sy += e * static_cast<float>(idx / W);
        }
    }
    se = warp_sum(se); sx = warp_sum(sx); sy = warp_sum(sy);
    if (lane == 0) {
        const float x = sx / se, y = sy / se;
        xy[2 * map] = x; xy[2 * map + 1] = y;
        if (xy_scaled) { xy_scaled[2 * map] = x * scale; xy_scaled[2 * map + 1] = y * scale; }
    }
}

int softargmax_launch(const float* hm, float* xy, float* xy_scaled, int n_maps, int H, int W, float temperature,
                      float scale, cudaStream_t s) {
    if (n_maps == 0) return 0;
    HMV_CHECK((H * W) % 4 == 0, "softargmax: H*W must be a multiple of 4");
    HMV_CUDA(launch_kernel(softargmax_kernel, dim3((n_maps + 7) / 8), dim3(256), 0, s, hm, xy, xy_scaled, n_maps, H, W, temperature, scale));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// bilinear neighbour gather (grid_sample align_corners=True, zeros padding)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sample_gather_kernel(const T* __restrict__ feat, const float* __restrict__ xy, T* __restrict__ rows,
                                     float* __restrict__ wts, int H, int W, int C) {
    pdl_wait();
    pdl_launch_dependents();
    const int nj = blockIdx.x;                 // n * 21 + j
    const int n = nj / kJoints;
    const float x = xy[2 * nj], y = xy[2 * nj + 1];
    // same op order as the reference: nets.py:48-49 then grid_sampler unnormalize (align_corners)
    const float gx = x / static_cast<float>(W - 1) * 2.f - 1.f;
    const float gy = y / static_cast<float>(H - 1) * 2.f - 1.f;
    const float ix = ((gx + 1.f) / 2.f) * static_cast<float>(W - 1);
    const float iy = ((gy + 1.f) / 2.f) * static_cast<float>(H - 1);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float wq[4] = {(x0 + 1.f - ix) * (y0 + 1.f - iy), (ix - x0) * (y0 + 1.f - iy),
                         (x0 + 1.f - ix) * (iy - y0), (ix - x0) * (iy - y0)};
    const int vec_per_row = C * static_cast<int>(sizeof(T)) / 16;
    for (int q = 0; q < 4; ++q) {
        const int xi = static_cast<int>(x0) + (q & 1), yi = static_cast<int>(y0) + (q >> 1);
        const bool ok = xi >= 0 && xi < W && yi >= 0 && yi < H;     // also false for NaN coords
        uint4* dst = reinterpret_cast<uint4*>(rows + (static_cast<size_t>(nj) * 4 + q) * C);
        if (ok) {
            const uint4* src = reinterpret_cast<const uint4*>(feat + ((static_cast<size_t>(n) * H + yi) * W + xi) * C);
            for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) dst[i] = __ldg(src + i);
        } else {
            for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) dst[i] = make_uint4(0, 0, 0, 0);
        }
        if (threadIdx.x == 0) wts[nj * 4 + q] = ok ? wq[q] : 0.f;
    }
}

template <typename T>
int sample_gather_launch(const T* feat, const float* xy, T* rows, float* wts, int n_img, int H, int W, int C,
                         cudaStream_t s) {
    if (n_img == 0) return 0;
    HMV_CHECK((C * sizeof(T)) % 16 == 0, "sample_gather: row must be a multiple of 16 bytes");
    HMV_CUDA(launch_kernel(sample_gather_kernel<T>, dim3(n_img * kJoints), dim3(128), 0, s, feat, xy, rows, wts, H, W, C));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int sample_gather_launch<bf16>(const bf16*, const float*, bf16*, float*, int, int, int, int, cudaStream_t);
template int sample_gather_launch<float>(const float*, const float*, float*, float*, int, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// token assembly: bilinear blend | xy | crop fov | + sinusoidal PE  (reference handmvnet.py:185-225,
// models/utils.py:134-171, layers.py:152-158)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void tokens_kernel(const TokenParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const int row = blockIdx.x;                // n * 21 + j
    const int n = row / kJoints;
    const int pos = row % p.tokens_per_sample; // token index inside the sample (view-major)
    float* of = p.tok_f32 + static_cast<size_t>(row) * p.pitch;
    T* ol = p.tok_lp ? static_cast<T*>(p.tok_lp) + static_cast<size_t>(row) * p.pitch : nullptr;
    const float* pe = p.pe ? p.pe + static_cast<size_t>(pos) * p.d : nullptr;
    for (int c = threadIdx.x; c < p.d; c += blockDim.x) {
        float v;
        if (c < p.feat) {
            int l = 0, cl = c;                   // feature level and column inside it (levels are concatenated: handmvnet.py:187)
            while (l + 1 < p.n_src && cl >= p.src[l].width) { cl -= p.src[l].width; ++l; }
            const TokenSource& sl = p.src[l];
            const float* g = sl.g + static_cast<size_t>(row) * 4 * sl.ldg;
            const float* w = sl.wts + row * 4;
            v = 0.f;
            v += g[cl] * w[0];
            v += g[sl.ldg + cl] * w[1];
            v += g[2 * sl.ldg + cl] * w[2];
            v += g[3 * sl.ldg + cl] * w[3];
        } else {
            int e = c - p.feat;
            if (p.use_pos2d && e < 2) {
                v = p.xy[row * 2 + e];
            } else {
                if (p.use_pos2d) e -= 2;
                // e in [0,10): point e/2 of (x1,y1),(x1,y2),(x2,y1),(x2,y2),centre ; axis e%2
                const float* bb = p.bbox + n * 4;
                const float* k = p.intr + n * 4;
                const int pt = e >> 1, ax = e & 1;
                float coord;
                if (pt == 4) coord = (bb[ax] + bb[2 + ax]) / 2.f;
                else coord = ax == 0 ? bb[(pt >> 1) * 2] : bb[1 + (pt & 1) * 2];
                v = atanf((coord - k[2 + ax]) / k[ax]);
            }
        }
        if (pe) v += pe[c];
        of[c] = v;
        if (ol) ol[c] = from_f<T>(v);
    }
}

template <typename T>
int tokens_launch(const TokenParams& p, cudaStream_t s) {
    if (p.n_img == 0) return 0;
    HMV_CHECK(!p.use_crop || (p.bbox && p.intr), "tokens: 'crop' positional encoding needs bbox and intrinsics");
    HMV_CUDA(launch_kernel(tokens_kernel<T>, dim3(p.n_img * kJoints), dim3(128), 0, s, p));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int tokens_launch<bf16>(const TokenParams&, cudaStream_t);
template int tokens_launch<float>(const TokenParams&, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// attention core: one CTA per (sample, head); K/V staged in shared memory, one warp per query row
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
attention_kernel(const T* __restrict__ qkv, int ld, T* __restrict__ out, int ld_out, int tokens_per_sample,
                 int q_row0, int nq, int kv_row0, int nk, int heads, float scale) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int DH = 128;
    constexpr int PK = DH + (sizeof(T) == 2 ? 2 : 1);      // padded K pitch: conflict-free lane-per-key reads
    constexpr int MAXJ = 11;                               // keys per lane: nk <= 352
    extern __shared__ __align__(16) uint8_t att_smem[];
    const int b = blockIdx.x, h = blockIdx.y;
    const int inner = heads * DH;
    T* ks = reinterpret_cast<T*>(att_smem);
    T* vs = ks + static_cast<size_t>(nk) * PK + (sizeof(T) == 2 ? ((nk * PK) & 1 ? 1 : 0) : 0);
    // align V to 16 B
    vs = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(vs) + 15) & ~static_cast<uintptr_t>(15));
    float* qs = reinterpret_cast<float*>(vs + static_cast<size_t>(nk) * DH);
    float* ps = qs + 8 * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const T* kbase = qkv + (static_cast<size_t>(b) * tokens_per_sample + kv_row0) * ld + inner + h * DH;
    const T* vbase = kbase + inner;
    for (int i = threadIdx.x; i < nk * DH; i += blockDim.x) {
        const int j = i / DH, d = i % DH;
        ks[j * PK + d] = kbase[static_cast<size_t>(j) * ld + d];
        vs[j * DH + d] = vbase[static_cast<size_t>(j) * ld + d];
    }
    __syncthreads();

    float* q = qs + warp * DH;
    float* pw = ps + warp * (MAXJ * 32);
    for (int i = warp; i < nq; i += 8) {
        const T* qrow = qkv + (static_cast<size_t>(b) * tokens_per_sample + q_row0 + i) * ld + h * DH;
        for (int d = lane; d < DH; d += 32) q[d] = to_f(qrow[d]) * scale;
        __syncwarp();
        float sc[MAXJ];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float s = -INFINITY;
            if (j < nk) {
                s = 0.f;
                if constexpr (sizeof(T) == 2) {
                    const uint32_t* kr = reinterpret_cast<const uint32_t*>(ks + j * PK);
#pragma unroll 8
                    for (int d2 = 0; d2 < DH / 2; ++d2) {
                        const float2 kf = unpack_bf16x2(kr[d2]);
                        s = fmaf(q[2 * d2], kf.x, s);
                        s = fmaf(q[2 * d2 + 1], kf.y, s);
                    }
                } else {
                    const float* kr = reinterpret_cast<const float*>(ks) + j * PK;
#pragma unroll 8
                    for (int d = 0; d < DH; ++d) s = fmaf(q[d], kr[d], s);
                }
            }
            sc[jj] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < nk) {
                const float e = expf(sc[jj] - mx);
                pw[j] = e;
                sum += e;
            }
        }
        sum = warp_sum(sum);
        __syncwarp();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int j = 0; j < nk; ++j) {
            const float pj = pw[j];
            if constexpr (sizeof(T) == 2) {
                const uint2 vv = *reinterpret_cast<const uint2*>(vs + j * DH + lane * 4);
                const float2 f0 = unpack_bf16x2(vv.x), f1 = unpack_bf16x2(vv.y);
                a0 = fmaf(pj, f0.x, a0); a1 = fmaf(pj, f0.y, a1); a2 = fmaf(pj, f1.x, a2); a3 = fmaf(pj, f1.y, a3);
            } else {
                const float4 vv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(vs) + j * DH + lane * 4);
                a0 = fmaf(pj, vv.x, a0); a1 = fmaf(pj, vv.y, a1); a2 = fmaf(pj, vv.z, a2); a3 = fmaf(pj, vv.w, a3);
            }
        }
        const float inv = 1.f / sum;
        T* orow = out + (static_cast<size_t>(b) * nq + i) * ld_out + h * DH + lane * 4;
        orow[0] = from_f<T>(a0 * inv); orow[1] = from_f<T>(a1 * inv);
        orow[2] = from_f<T>(a2 * inv); orow[3] = from_f<T>(a3 * inv);
        __syncwarp();
    }
}

template <typename T>
int attention_launch(const T* qkv, int ld, T* out, int ld_out, int batch, int tokens_per_sample, int q_row0, int nq,
                     int kv_row0, int nk, int heads, int dim_head, float scale, cudaStream_t s) {
    if (batch == 0) return 0;
    HMV_CHECK(dim_head == 128, "attention: dim_head must be 128");
    HMV_CHECK(nk <= 352 && nk > 0 && nq > 0, "attention: key count out of range (<= 352)");
    constexpr int PK = 128 + (sizeof(T) == 2 ? 2 : 1);
    const size_t smem = static_cast<size_t>(nk) * PK * sizeof(T) + 32 + static_cast<size_t>(nk) * 128 * sizeof(T) +
                        8 * 128 * sizeof(float) + 8 * 11 * 32 * sizeof(float);
    HMV_CHECK(smem <= 227 * 1024, "attention: K/V do not fit in shared memory at this view count and precision");
    static size_t configured[2][64] = {};                  // per element type and device
    int dev = 0;
    HMV_CUDA(cudaGetDevice(&dev));
    size_t& cfg = configured[sizeof(T) == 2 ? 0 : 1][dev & 63];
    if (smem > cfg) {
        HMV_CUDA(cudaFuncSetAttribute(attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        cfg = smem;
    }
    HMV_CUDA(launch_kernel(attention_kernel<T>, dim3(batch, heads), dim3(256), smem, s, qkv, ld, out, ld_out, tokens_per_sample, q_row0, nq,
                                                             kv_row0, nk, heads, scale));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int attention_launch<bf16>(const bf16*, int, bf16*, int, int, int, int, int, int, int, int, int, float, cudaStream_t);
template int attention_launch<float>(const float*, int, float*, int, int, int, int, int, int, int, int, int, float, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// bf16 attention core on the tensor cores (mma.sync m16n8k16, fp32 accumulate), flash-style:
// one CTA = 64 queries of one (sample, head); K/V stream through shared memory in blocks of 64 keys with an
// online softmax, so any key count (21 .. 336 for 1 .. 16 views) runs with bounded registers.
// (reference layers.py:217-223; the fp32 check mode keeps the scalar kernel above.)
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kAttQ = 64, kAttK = 64, kAttD = 128, kAttPitch = kAttD + 8;     // +8 bf16: conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 64 rows x 128 bf16 from global (row pitch ld) into padded smem; rows >= valid are zero filled
__device__ __forceinline__ void att_load_tile(bf16* dst, const bf16* src, int ld, int valid) {
    for (int i = threadIdx.x; i < 64 * 16; i += 128) {
        const int r = i >> 4, c = (i & 15) * 8;
        const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst + r * kAttPitch + c));
        if (r < valid) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + static_cast<size_t>(r) * ld + c) : "memory");
        } else {
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0u) : "memory");
        }
    }
}
}  // namespace

__global__ void __launch_bounds__(128)
attention_mma_kernel(const bf16* __restrict__ qkv, int ld, bf16* __restrict__ out, int ld_out, int tokens_per_sample,
                     int q_row0, int nq, int kv_row0, int nk, int heads, float scale_log2e) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) uint8_t att_smem[];
    bf16* qs = reinterpret_cast<bf16*>(att_smem);
    bf16* ks = qs + kAttQ * kAttPitch;
    bf16* vs = ks + kAttK * kAttPitch;
    const int b = blockIdx.x, h = blockIdx.y, q0 = blockIdx.z * kAttQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int inner = heads * kAttD;
    const bf16* qbase = qkv + (static_cast<size_t>(b) * tokens_per_sample + q_row0 + q0) * ld + h * kAttD;
    const bf16* kbase = qkv + (static_cast<size_t>(b) * tokens_per_sample + kv_row0) * ld + inner + h * kAttD;
    const bf16* vbase = kbase + inner;

    att_load_tile(qs, qbase, ld, nq - q0 < kAttQ ? nq - q0 : kAttQ);
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // rows g and g+8 of this warp's 16-row slab

    const uint32_t qs_addr = static_cast<uint32_t>(__cvta_generic_to_shared(qs));
    const uint32_t ks_addr = static_cast<uint32_t>(__cvta_generic_to_shared(ks));
    const uint32_t vs_addr = static_cast<uint32_t>(__cvta_generic_to_shared(vs));
    // ldmatrix lane addressing: A (Q): lanes 0-15 rows 0-15 @k, lanes 16-31 rows 0-15 @k+8
    const uint32_t a_off = ((warp * 16 + (lane & 15)) * kAttPitch + (lane >> 4) * 8) * 2;
    // B (K): matrices (keys j..j+7 @k), (j..j+7 @k+8), (j+8..15 @k), (j+8..15 @k+8)
    const uint32_t bk_off = (((lane & 7) + (lane >> 4) * 8) * kAttPitch + ((lane >> 3) & 1) * 8) * 2;
    // B (V, transposed): matrices (keys j..j+7 @d), (j+8..15 @d), (j..j+7 @d+8), (j+8..15 @d+8)
    const uint32_t bv_off = (((lane & 7) + ((lane >> 3) & 1) * 8) * kAttPitch + (lane >> 4) * 8) * 2;

    for (int j0 = 0; j0 < nk; j0 += kAttK) {
        __syncthreads();                                           // previous block fully consumed
        const int kvalid = nk - j0 < kAttK ? nk - j0 : kAttK;
        att_load_tile(ks, kbase + static_cast<size_t>(j0) * ld, ld, kvalid);
        att_load_tile(vs, vbase + static_cast<size_t>(j0) * ld, ld, kvalid);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        // ---- S = Q K^T for this warp's 16 queries x 64 keys ----
        float sc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < kAttD / 16; ++kk) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(qs_addr + a_off + kk * 32, a0, a1, a2, a3);
#pragma unroll
            for (int np = 0; np < 4; ++np) {                       // pairs of 8-key tiles
                uint32_t b0, b1, b2, b3;
                ldsm_x4(ks_addr + bk_off + np * 16 * kAttPitch * 2 + kk * 32, b0, b1, b2, b3);
                mma_bf16(sc[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16(sc[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        // ---- online softmax (rows g, g+8; key columns nt*8 + 2t, +1) ----
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
        const float al0 = exp2f((m0 - mn0) * scale_log2e), al1 = exp2f((m1 - mn1) * scale_log2e);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pa[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f((sc[nt][0] - mn0) * scale_log2e), p1 = exp2f((sc[nt][1] - mn0) * scale_log2e);
            const float p2 = exp2f((sc[nt][2] - mn1) * scale_log2e), p3 = exp2f((sc[nt][3] - mn1) * scale_log2e);
            rs0 += p0 + p1; rs1 += p2 + p3;
            pa[nt][0] = pack_bf16x2(p0, p1);
            pa[nt][1] = pack_bf16x2(p2, p3);
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
#pragma unroll
        for (int i = 0; i < 16; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < kAttK / 16; ++kk) {
            const uint32_t a0 = pa[2 * kk][0], a1 = pa[2 * kk][1], a2 = pa[2 * kk + 1][0], a3 = pa[2 * kk + 1][1];
#pragma unroll
            for (int dp = 0; dp < 8; ++dp) {                       // pairs of 8-wide d tiles
                uint32_t b0, b1, b2, b3;
                ldsm_x4_trans(vs_addr + bv_off + kk * 16 * kAttPitch * 2 + dp * 32, b0, b1, b2, b3);
                mma_bf16(o[2 * dp], a0, a1, a2, a3, b0, b1);
                mma_bf16(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    }
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int dt = 0; dt < 16; ++dt) {
        const int d = dt * 8 + 2 * t;
        if (r0 < nq)
            *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * nq + r0) * ld_out + h * kAttD + d) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
        if (r1 < nq)
            *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * nq + r1) * ld_out + h * kAttD + d) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

int attention_mma_launch(const bf16* qkv, int ld, bf16* out, int ld_out, int batch, int tokens_per_sample, int q_row0,
                         int nq, int kv_row0, int nk, int heads, int dim_head, float scale, cudaStream_t s) {
    if (batch == 0) return 0;
    HMV_CHECK(dim_head == kAttD, "attention: dim_head must be 128");
    HMV_CHECK(nk > 0 && nq > 0 && ld % 8 == 0 && ld_out % 2 == 0, "attention: bad shape");
    const int smem = (kAttQ + 2 * kAttK) * kAttPitch * 2;
    static unsigned long long configured = 0;
    if (first_use_on_this_device(configured)) {
        HMV_CUDA(cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    HMV_CUDA(launch_kernel(attention_mma_kernel, dim3(batch, heads, (nq + kAttQ - 1) / kAttQ), dim3(128), smem, s, qkv, ld, out, ld_out, tokens_per_sample, q_row0, nq, kv_row0, nk, heads, scale * 1.4426950408889634f));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (optionally two chained LayerNorms): one warp per row, values kept in registers
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void layernorm_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ g1,
                                 const float* __restrict__ b1, float* __restrict__ out_f32, int ld_out,
                                 const float* __restrict__ g2, const float* __restrict__ b2, T* __restrict__ out_lp,
                                 int ld_lp, int rows, int d, float eps) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int MAXPER = 20;                 // d <= 640
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = in + static_cast<size_t>(row) * ld_in;
    float v[MAXPER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPER; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < d ? x[c] : 0.f;
        s += v[i];
    }
    const float inv_d = 1.f / static_cast<float>(d);
    float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPER; ++i) {
        const int c = lane + 32 * i;
        const float t = c < d ? v[i] - mean : 0.f;
        ss += t * t;
    }
    float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
    s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPER; ++i) {
        const int c = lane + 32 * i;
        if (c < d) {
            v[i] = (v[i] - mean) * rstd * g1[c] + b1[c];
            if (out_f32) out_f32[static_cast<size_t>(row) * ld_out + c] = v[i];
            s += v[i];
        }
    }
    if (!out_lp) return;
    if (g2) {
        mean = warp_sum(s) * inv_d;
        ss = 0.f;
#pragma unroll
        for (int i = 0; i < MAXPER; ++i) {
            const int c = lane + 32 * i;
            const float t = c < d ? v[i] - mean : 0.f;
            ss += t * t;
        }
        rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
#pragma unroll
        for (int i = 0; i < MAXPER; ++i) {
            const int c = lane + 32 * i;
            if (c < d) out_lp[static_cast<size_t>(row) * ld_lp + c] = from_f<T>((v[i] - mean) * rstd * g2[c] + b2[c]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < MAXPER; ++i) {
            const int c = lane + 32 * i;
            if (c < d) out_lp[static_cast<size_t>(row) * ld_lp + c] = from_f<T>(v[i]);
        }
    }
}

template <typename T>
int layernorm_launch(const float* in, int ld_in, const float* g1, const float* b1, float* out_f32, int ld_out,
                     const float* g2, const float* b2, T* out_lp, int ld_lp, int rows, int d, float eps,
                     cudaStream_t s) {
    if (rows == 0) return 0;
    HMV_CHECK(d <= 640, "layernorm: d_model > 640 not supported");
    HMV_CUDA(launch_kernel(layernorm_kernel<T>, dim3((rows + 7) / 8), dim3(256), 0, s, in, ld_in, g1, b1, out_f32, ld_out, g2, b2, out_lp, ld_lp, rows, d, eps));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int layernorm_launch<bf16>(const float*, int, const float*, const float*, float*, int, const float*,
                                    const float*, bf16*, int, int, int, float, cudaStream_t);
template int layernorm_launch<float>(const float*, int, const float*, const float*, float*, int, const float*,
                                     const float*, float*, int, int, int, float, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// Chebyshev graph-conv head (reference nets.py:133-139, layers.py:387-403), fp32.
// Layer: out = sum_k T_k (X W_k) + b.  A CTA owns 64 output columns; its 256 threads are 64 columns x 4 K-splits,
// weights stream from L2 in double-buffered batches of 8 (the loop is L2-latency bound otherwise), partial
// products are reduced through shared memory and T_k is applied per column in registers.
//   gcn_l1_kernel : grid (batch, 256/64)   X[21, d_in] -> H1[21, 256]          (95 % of the head's FLOPs)
//   gcn_l23_kernel: grid (batch)           H1 -> H2[21, 64] -> joints[21, 3]
// ------------------------------------------------------------------------------------------------
constexpr int kGcnPad = 24;                    // 21 joints padded to 6 float4
// COLS output columns x KS K-splits = the CTA's threads: 64 x 4 (256 threads), grid (batch, 4) for layer 1, (batch) for
// layers 2+3.  At B = 1 that is 4 + 1 CTAs, each thread walking 131 (64) sequential L2-latency-bound weight loads per T_k
// (62 + 50 us of a 0.93 ms forward): passes of <= kGcnSmallBatch samples use the *_small kernels further down.
constexpr int kGcnSmallBatch = 8;

// X[21, cin] (row pitch ld, global) -> xt[cin][24] (shared, joint-minor, pad joints zero); coalesced along the channels.
__device__ __forceinline__ void gcn_load_xt(float* __restrict__ xt, const float* __restrict__ x, int cin, int ld, int tid, int nthreads) {
    for (int i = tid; i < cin * kGcnPad; i += nthreads) {
        const int r = i / cin, c = i - r * cin;
        xt[c * kGcnPad + r] = r < kJoints ? x[static_cast<size_t>(r) * ld + c] : 0.f;
    }
}

// Computes, for column `col` (global) and K-split `ks`, the partial z_k[s] = sum_{i in split} X[s, i] W_k[i, col],
// reduces over the splits through `red` and returns (in threads with ks == 0) o[r] = bias + sum_k (T_k z_k)[r].
template <int COLS, int KS>
__device__ __forceinline__ void gcn_layer_split(const float* __restrict__ xt /*[cin][24] smem*/, int cin,
                                                const float* __restrict__ w /*[3][cin][cout]*/,
                                                const float* __restrict__ bias, int cout, int col, int lcol, int ks,
                                                const float* __restrict__ basis /*[3][21][21] smem*/,
                                                float* __restrict__ red /*[KS][24][COLS] smem*/, bool leaky,
                                                float (&o)[kJoints], bool active) {
    const int per_max = (cin + KS - 1) / KS, i_lo = ks * per_max;
    const int per = i_lo >= cin ? 0 : (cin - i_lo < per_max ? cin - i_lo : per_max);
    if (ks == 0 && active) {
#pragma unroll
        for (int r = 0; r < kJoints; ++r) o[r] = bias[col];
    }
    for (int k = 0; k < 3; ++k) {
        float z[kGcnPad];
#pragma unroll
        for (int r = 0; r < kGcnPad; ++r) z[r] = 0.f;
        if (active) {
            const float* wk = w + (static_cast<size_t>(k) * cin + i_lo) * cout + col;
            float wc[8], wn[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) wc[u] = u < per ? __ldg(wk + static_cast<size_t>(u) * cout) : 0.f;
            for (int i0 = 0; i0 < per; i0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) wn[u] = (i0 + 8 + u) < per ? __ldg(wk + static_cast<size_t>(i0 + 8 + u) * cout) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (i0 + u < per) {
                        const float4* xr = reinterpret_cast<const float4*>(xt + (i_lo + i0 + u) * kGcnPad);
                        const float wv = wc[u];
#pragma unroll
                        for (int q = 0; q < kGcnPad / 4; ++q) {
                            const float4 xv = xr[q];
                            z[4 * q] = fmaf(xv.x, wv, z[4 * q]);         z[4 * q + 1] = fmaf(xv.y, wv, z[4 * q + 1]);
                            z[4 * q + 2] = fmaf(xv.z, wv, z[4 * q + 2]); z[4 * q + 3] = fmaf(xv.w, wv, z[4 * q + 3]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) wc[u] = wn[u];
            }
        }
        __syncthreads();                       // previous use of `red` is over
        if (active) {
#pragma unroll
            for (int r = 0; r < kJoints; ++r) red[(ks * kGcnPad + r) * COLS + lcol] = z[r];
        }
        __syncthreads();
        if (ks == 0 && active) {
#pragma unroll
            for (int r = 0; r < kJoints; ++r) {
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < KS; ++q) a += red[(q * kGcnPad + r) * COLS + lcol];
                z[r] = a;
            }
            const float* tk = basis + k * kJoints * kJoints;
#pragma unroll
            for (int r = 0; r < kJoints; ++r) {
                float a = 0.f;
#pragma unroll
                for (int s2 = 0; s2 < kJoints; ++s2) a = fmaf(tk[r * kJoints + s2], z[s2], a);
                o[r] += a;
            }
        }
    }
    if (leaky && ks == 0 && active) {
#pragma unroll
        for (int r = 0; r < kJoints; ++r) o[r] = o[r] > 0.f ? o[r] : 0.01f * o[r];
    }
}

template <int COLS, int KS>
__global__ void __launch_bounds__(COLS * KS)
gcn_l1_kernel(const GcnParams p, float* __restrict__ h1 /*[batch][21][256]*/) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float gsm[];
    float* xt = gsm;                                       // [d_in][24]
    float* basis = xt + static_cast<size_t>(p.d_in) * kGcnPad;
    float* red = basis + 3 * kJoints * kJoints + 1;        // scalar access only
    const int b = blockIdx.x, tid = threadIdx.x;
    gcn_load_xt(xt, p.x + static_cast<size_t>(b) * kJoints * p.ld, p.d_in, p.ld, tid, COLS * KS);
    for (int i = tid; i < 3 * kJoints * kJoints; i += COLS * KS) basis[i] = p.basis[i];
    __syncthreads();
    const int lcol = tid % COLS, ks = tid / COLS;
    const int col = blockIdx.y * COLS + lcol;
    float o[kJoints];
    gcn_layer_split<COLS, KS>(xt, p.d_in, p.w[0], p.b[0], 256, col, lcol, ks, basis, red, true, o, true);
    if (ks == 0) {
#pragma unroll
        for (int r = 0; r < kJoints; ++r) h1[(static_cast<size_t>(b) * kJoints + r) * 256 + col] = o[r];
    }
}

template <int KS>
__global__ void __launch_bounds__(64 * KS)
gcn_l23_kernel(const GcnParams p, const float* __restrict__ h1) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float gsm[];
    float* xt = gsm;                                       // [256][24]
    float* basis = xt + 256 * kGcnPad;
    float* red = basis + 3 * kJoints * kJoints + 1;
    const int b = blockIdx.x, tid = threadIdx.x;
    gcn_load_xt(xt, h1 + static_cast<size_t>(b) * kJoints * 256, 256, 256, tid, 64 * KS);
    for (int i = tid; i < 3 * kJoints * kJoints; i += 64 * KS) basis[i] = p.basis[i];
    __syncthreads();
    const int lcol = tid % 64, ks = tid / 64;
    float o[kJoints];
    gcn_layer_split<64, KS>(xt, 256, p.w[1], p.b[1], 64, lcol, lcol, ks, basis, red, true, o, true);
    __syncthreads();
    if (ks == 0) {
#pragma unroll
        for (int r = 0; r < kJoints; ++r) xt[lcol * kGcnPad + r] = o[r];
        xt[lcol * kGcnPad + 21] = 0.f; xt[lcol * kGcnPad + 22] = 0.f; xt[lcol * kGcnPad + 23] = 0.f;
    }
    __syncthreads();
    const bool active = lcol < 3;
    gcn_layer_split<64, KS>(xt, 64, p.w[2], p.b[2], 3, lcol, lcol, ks, basis, red, false, o, active);
    if (ks == 0 && active) {
#pragma unroll
        for (int r = 0; r < kJoints; ++r) p.out[(static_cast<size_t>(b) * kJoints + r) * 3 + lcol] = o[r];
    }
}

// ---- small passes (batch <= kGcnSmallBatch): the loop above is a chain of L2 latencies (a thread waits for 8 weights, uses
// them, waits for the next 8: 15 round trips in layer 1).  Here a thread issues ALL of its weight loads at once -- before the
// dependency wait of the launch, they are parameters -- and the split / Chebyshev reductions are spread over the whole CTA.
//   gcn_l1_small_kernel : grid (batch, 16), 256 threads = 16 columns x 16 K-splits, <= 33 input rows per split (d_in <= 528)
//   gcn_l23_small_kernel: grid (batch), 512 threads = 64 columns x 8 K-splits of 32 rows (layer 2), then layer 3 (64 -> 3) ----
constexpr int kGsCols = 16, kGsKs = 16, kGsPer = 33;
constexpr int kGs2Ks = 8, kGs2Per = 32;

__global__ void __launch_bounds__(kGsCols * kGsKs)
gcn_l1_small_kernel(const GcnParams p, float* __restrict__ h1 /*[batch][21][256]*/) {
    constexpr int NT = kGsCols * kGsKs;
    extern __shared__ __align__(16) float gsm[];
    float* xt = gsm;                                           // [kGsKs * kGsPer][24], rows >= d_in are zero
    float* basis = xt + kGsKs * kGsPer * kGcnPad;              // [3][21][21]
    float* red = basis + 3 * kJoints * kJoints + 1;            // [3][kGsKs][21][kGsCols]
    float* zs = red + 3 * kGsKs * kJoints * kGsCols;           // [3][21][kGsCols]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int lcol = tid % kGsCols, ks = tid / kGsCols;
    const int col = blockIdx.y * kGsCols + lcol;
    const int cin = p.d_in;
    const int per_max = (cin + kGsKs - 1) / kGsKs, i_lo = ks * per_max;
    const int per = i_lo >= cin ? 0 : (cin - i_lo < per_max ? cin - i_lo : per_max);
    float wv[3][kGsPer];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int u = 0; u < kGsPer; ++u)
            wv[k][u] = u < per ? __ldg(p.w[0] + (static_cast<size_t>(k) * cin + i_lo + u) * 256 + col) : 0.f;
    for (int i = tid; i < 3 * kJoints * kJoints; i += NT) basis[i] = p.basis[i];
    pdl_wait();
    pdl_launch_dependents();
    {   // X[21, d_in] -> xt[c][r]; 8-byte loads along the channels (d_in is even), everything else zero
        const float* x = p.x + static_cast<size_t>(b) * kJoints * p.ld;
        for (int i = tid; i < kGsKs * kGsPer * kGcnPad; i += NT) xt[i] = 0.f;
        __syncthreads();
        const int half = cin >> 1;
        for (int i = tid; i < kJoints * half; i += NT) {
            const int r = i / half, c2 = i - r * half;
            const float2 v = *reinterpret_cast<const float2*>(x + static_cast<size_t>(r) * p.ld + 2 * c2);
            xt[(2 * c2) * kGcnPad + r] = v.x;
            xt[(2 * c2 + 1) * kGcnPad + r] = v.y;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float z[kGcnPad];
#pragma unroll
        for (int r = 0; r < kGcnPad; ++r) z[r] = 0.f;
#pragma unroll
        for (int u = 0; u < kGsPer; ++u) {
            const int i = i_lo + u < kGsKs * kGsPer ? i_lo + u : kGsKs * kGsPer - 1;       // (weights beyond `per` are zero)
            const float4* xr = reinterpret_cast<const float4*>(xt + i * kGcnPad);
            const float w = wv[k][u];
#pragma unroll
            for (int q = 0; q < kGcnPad / 4; ++q) {
                const float4 xv = xr[q];
                z[4 * q] = fmaf(xv.x, w, z[4 * q]);         z[4 * q + 1] = fmaf(xv.y, w, z[4 * q + 1]);
                z[4 * q + 2] = fmaf(xv.z, w, z[4 * q + 2]); z[4 * q + 3] = fmaf(xv.w, w, z[4 * q + 3]);
            }
        }
#pragma unroll
        for (int r = 0; r < kJoints; ++r) red[((k * kGsKs + ks) * kJoints + r) * kGsCols + lcol] = z[r];
    }
    __syncthreads();
    for (int idx = tid; idx < 3 * kJoints * kGsCols; idx += NT) {       // sum over the K-splits
        const int k = idx / (kJoints * kGsCols), rem = idx - k * (kJoints * kGsCols);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kGsKs; ++q) a += red[(k * kGsKs + q) * kJoints * kGsCols + rem];
        zs[idx] = a;
    }
    __syncthreads();
    for (int idx = tid; idx < kJoints * kGsCols; idx += NT) {           // o = b + sum_k T_k z_k, LeakyReLU
        const int r = idx / kGsCols, c = idx - r * kGsCols;
        float o = p.b[0][blockIdx.y * kGsCols + c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < kJoints; ++s2) a = fmaf(basis[(k * kJoints + r) * kJoints + s2], zs[(k * kJoints + s2) * kGsCols + c], a);
            o += a;
        }
        o = o > 0.f ? o : 0.01f * o;
        h1[(static_cast<size_t>(b) * kJoints + r) * 256 + blockIdx.y * kGsCols + c] = o;
    }
}

__device__ __forceinline__ void gcn2_load_w(float (&w)[kGs2Per], const float* __restrict__ w2, int k, int i_lo, int lcol) {
#pragma unroll
    for (int u = 0; u < kGs2Per; ++u) w[u] = __ldg(w2 + (static_cast<size_t>(k) * 256 + i_lo + u) * 64 + lcol);
}
__device__ __forceinline__ void gcn2_partial(const float (&w)[kGs2Per], const float* __restrict__ xt, int i_lo, float* __restrict__ red_k /*[21][64] of this split*/, int lcol) {
    float z[kGcnPad];
#pragma unroll
    for (int r = 0; r < kGcnPad; ++r) z[r] = 0.f;
#pragma unroll
    for (int u = 0; u < kGs2Per; ++u) {
        const float4* xr = reinterpret_cast<const float4*>(xt + (i_lo + u) * kGcnPad);
        const float wv = w[u];
#pragma unroll
        for (int q = 0; q < kGcnPad / 4; ++q) {
            const float4 xv = xr[q];
            z[4 * q] = fmaf(xv.x, wv, z[4 * q]);         z[4 * q + 1] = fmaf(xv.y, wv, z[4 * q + 1]);
            z[4 * q + 2] = fmaf(xv.z, wv, z[4 * q + 2]); z[4 * q + 3] = fmaf(xv.w, wv, z[4 * q + 3]);
        }
    }
#pragma unroll
    for (int r = 0; r < kJoints; ++r) red_k[r * 64 + lcol] = z[r];
}

__global__ void __launch_bounds__(64 * kGs2Ks, 1)
gcn_l23_small_kernel(const GcnParams p, const float* __restrict__ h1) {
    constexpr int NT = 64 * kGs2Ks;
    extern __shared__ __align__(16) float gsm[];
    float* xt = gsm;                                           // [256][24]
    float* basis = xt + 256 * kGcnPad;                         // [3][21][21]
    float* w3 = basis + 3 * kJoints * kJoints + 1;             // [3][64][3]
    float* x3 = w3 + 3 * 64 * 3 + 3;                           // [64][24]  (16-byte aligned: 1324 + 579 = 1903 floats after xt... scalar access only)
    float* z3 = x3 + 64 * kGcnPad;                             // [3][21][3]
    float* zs = z3 + 3 * kJoints * 3 + 3;                      // [3][21][64]
    float* red = zs + 3 * kJoints * 64;                        // [3][kGs2Ks][21][64]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int lcol = tid % 64, ks = tid / 64;
    const int i_lo = ks * kGs2Per;
    float wa[kGs2Per], wb[kGs2Per];                            // weights of T_k, double-buffered: k = 0 / 2 in wa, k = 1 in wb
    gcn2_load_w(wa, p.w[1], 0, i_lo, lcol);
    for (int i = tid; i < 3 * kJoints * kJoints; i += NT) basis[i] = p.basis[i];
    for (int i = tid; i < 3 * 64 * 3; i += NT) w3[i] = p.w[2][i];
    pdl_wait();
    pdl_launch_dependents();
    {   // H1[21, 256] -> xt[c][r]
        const float* x = h1 + static_cast<size_t>(b) * kJoints * 256;
        for (int i = tid; i < kJoints * 64; i += NT) {
            const int r = i >> 6, c4 = i & 63;
            const float4 v = *reinterpret_cast<const float4*>(x + r * 256 + 4 * c4);
            xt[(4 * c4) * kGcnPad + r] = v.x;     xt[(4 * c4 + 1) * kGcnPad + r] = v.y;
            xt[(4 * c4 + 2) * kGcnPad + r] = v.z; xt[(4 * c4 + 3) * kGcnPad + r] = v.w;
        }
        for (int i = tid; i < 256 * 3; i += NT) xt[(i / 3) * kGcnPad + kJoints + i % 3] = 0.f;
    }
    __syncthreads();
    // (the barriers keep the compiler from hoisting all 96 weight loads to the top: 64 weights + 24 sums live at most)
    gcn2_load_w(wb, p.w[1], 1, i_lo, lcol);
    gcn2_partial(wa, xt, i_lo, red + (0 * kGs2Ks + ks) * kJoints * 64, lcol);
    __syncthreads();
    gcn2_load_w(wa, p.w[1], 2, i_lo, lcol);
    gcn2_partial(wb, xt, i_lo, red + (1 * kGs2Ks + ks) * kJoints * 64, lcol);
    __syncthreads();
    gcn2_partial(wa, xt, i_lo, red + (2 * kGs2Ks + ks) * kJoints * 64, lcol);
    __syncthreads();
    for (int idx = tid; idx < 3 * kJoints * 64; idx += NT) {
        const int k = idx / (kJoints * 64), rem = idx - k * (kJoints * 64);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kGs2Ks; ++q) a += red[(k * kGs2Ks + q) * kJoints * 64 + rem];
        zs[idx] = a;
    }
    __syncthreads();
    for (int idx = tid; idx < kGcnPad * 64; idx += NT) {                // layer-2 output -> x3[c][r] (pad joints zero)
        const int c = idx / kGcnPad, r = idx - c * kGcnPad;
        float o = 0.f;
        if (r < kJoints) {
            o = p.b[1][c];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float a = 0.f;
#pragma unroll
                for (int s2 = 0; s2 < kJoints; ++s2) a = fmaf(basis[(k * kJoints + r) * kJoints + s2], zs[(k * kJoints + s2) * 64 + c], a);
                o += a;
            }
            o = o > 0.f ? o : 0.01f * o;
        }
        x3[idx] = o;
    }
    __syncthreads();
    if (tid < 3 * kJoints * 3) {                                        // layer 3: z_k[s][c] = sum_i X3[s, i] W_k[i, c]
        const int k = tid / (kJoints * 3), rem = tid - k * (kJoints * 3), s = rem / 3, c = rem - s * 3;
        float a = 0.f;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) a = fmaf(x3[i * kGcnPad + s], w3[(k * 64 + i) * 3 + c], a);
        z3[tid] = a;
    }
    __syncthreads();
    if (tid < kJoints * 3) {
        const int r = tid / 3, c = tid - r * 3;
        float o = p.b[2][c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < kJoints; ++s2) a = fmaf(basis[(k * kJoints + r) * kJoints + s2], z3[(k * kJoints + s2) * 3 + c], a);
            o += a;
        }
        p.out[(static_cast<size_t>(b) * kJoints + r) * 3 + c] = o;
    }
}

static int gcn_launch_small(const GcnParams& p, float* h1_scratch, cudaStream_t s) {
    constexpr size_t fixed = 3 * kJoints * kJoints + 1;
    constexpr size_t smem = (static_cast<size_t>(kGsKs) * kGsPer * kGcnPad + fixed + 3 * kGsKs * kJoints * kGsCols + 3 * kJoints * kGsCols) * sizeof(float);
    constexpr size_t smem2 = (256 * kGcnPad + fixed + (3 * 64 * 3 + 3) + 64 * kGcnPad + (3 * kJoints * 3 + 3) + 3 * kJoints * 64 +
                              3 * kGs2Ks * kJoints * 64) * sizeof(float);
    static unsigned long long configured = 0;
    if (first_use_on_this_device(configured)) {
        HMV_CUDA(cudaFuncSetAttribute(gcn_l1_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        HMV_CUDA(cudaFuncSetAttribute(gcn_l23_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
    }
    HMV_CUDA(launch_kernel(gcn_l1_small_kernel, dim3(p.batch, 256 / kGsCols), dim3(kGsCols * kGsKs), smem, s, p, h1_scratch));
    HMV_CUDA(cudaGetLastError());
    HMV_CUDA(launch_kernel(gcn_l23_small_kernel, dim3(p.batch), dim3(64 * kGs2Ks), smem2, s, p, h1_scratch));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

static int gcn_launch_large(const GcnParams& p, float* h1_scratch, cudaStream_t s) {
    constexpr int COLS1 = 64, KS1 = 4, KS2 = 4;
    const size_t fixed = 3 * kJoints * kJoints + 1;
    const size_t smem = (static_cast<size_t>(p.d_in) * kGcnPad + fixed + static_cast<size_t>(KS1) * kGcnPad * COLS1) * sizeof(float);
    const size_t smem2 = (static_cast<size_t>(256) * kGcnPad + fixed + static_cast<size_t>(KS2) * kGcnPad * 64) * sizeof(float);
    static size_t configured_dev[64] = {};                 // per device
    int dev = 0;
    HMV_CUDA(cudaGetDevice(&dev));
    size_t& configured = configured_dev[dev & 63];
    if (smem > configured) {
        HMV_CUDA(cudaFuncSetAttribute(gcn_l1_kernel<COLS1, KS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        HMV_CUDA(cudaFuncSetAttribute(gcn_l23_kernel<KS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
        configured = smem;
    }
    HMV_CUDA(launch_kernel(gcn_l1_kernel<COLS1, KS1>, dim3(p.batch, 256 / COLS1), dim3(COLS1 * KS1), smem, s, p, h1_scratch));
    HMV_CUDA(cudaGetLastError());
    HMV_CUDA(launch_kernel(gcn_l23_kernel<KS2>, dim3(p.batch), dim3(64 * KS2), smem2, s, p, h1_scratch));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

int gcn_launch(const GcnParams& p, float* h1_scratch, cudaStream_t s) {
    if (p.batch == 0) return 0;
    HMV_CHECK(p.d_in <= 1024, "gcn: d_in must be <= 1024");
    static const bool small_env = [] { const char* e = getenv("HMV_GCN_SMALL"); return !(e && e[0] == '0'); }();
    if (small_env && p.batch <= kGcnSmallBatch && p.d_in <= kGsKs * kGsPer && p.d_in % 2 == 0 && p.ld % 2 == 0)
        return gcn_launch_small(p, h1_scratch, s);
    return gcn_launch_large(p, h1_scratch, s);
}

// ------------------------------------------------------------------------------------------------
// Image transform in front of the model (reference datasets/utils.py:40-77 crop_and_pad_image, datasets/ho3d.py:35-40
// ToTensor -> Resize((S, S), antialias=True) -> Normalize): full camera frames (uint8 HWC) + integer boxes ->
// normalised fp32 NCHW crops.  The resize is torch's separable anti-aliased bilinear filter
// (aten UpSampleKernel.cpp, _compute_indices_weights_aa): scale = in / out, support = max(scale, 1),
// taps [int(c - support + .5), int(c + support + .5)) around c = scale * (i + .5), triangle weights normalised to 1;
// horizontal pass first, then vertical, as the CPU kernel does.  One thread per output pixel.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kPreMaxTaps = 20;                 // crop side <= ~9 x the output side

struct AaTaps { int lo, n; float w[kPreMaxTaps]; };

__device__ __forceinline__ void aa_taps(int in_size, int out_size, int i, AaTaps& t, bool& overflow) {
    const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
    const float support = scale >= 1.f ? scale : 1.f;
    const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
    const float center = scale * (static_cast<float>(i) + 0.5f);
    int lo = static_cast<int>(center - support + 0.5f);
    lo = lo < 0 ? 0 : lo;
    int hi = static_cast<int>(center + support + 0.5f);
    hi = hi > in_size ? in_size : hi;
    int n = hi - lo;
    if (n > kPreMaxTaps) { n = kPreMaxTaps; overflow = true; }
    float total = 0.f;
#pragma unroll
    for (int j = 0; j < kPreMaxTaps; ++j) {
        float w = 0.f;
        if (j < n) {
            const float x = fabsf((static_cast<float>(j + lo) - center + 0.5f) * invscale);
            w = x < 1.f ? 1.f - x : 0.f;
        }
        t.w[j] = w;
        total += w;
    }
    if (total != 0.f) {
#pragma unroll
        for (int j = 0; j < kPreMaxTaps; ++j) t.w[j] = t.w[j] / total;
    }
    t.lo = lo; t.n = n;
}
}  // namespace

__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ bbox, float* __restrict__ out, int frame_h,
                  int frame_w, int size, StemNorm norm, int* err_flag) {
    pdl_wait();
    pdl_launch_dependents();
    const int n = blockIdx.y, oy = blockIdx.x;
    const int x1 = bbox[4 * n], y1 = bbox[4 * n + 1], x2 = bbox[4 * n + 2], y2 = bbox[4 * n + 3];
    const int cw = x2 - x1, ch = y2 - y1;
    const uint8_t* img = frames + static_cast<size_t>(n) * frame_h * frame_w * 3;
    bool overflow = cw <= 0 || ch <= 0;
    AaTaps ty;
    aa_taps(ch > 0 ? ch : 1, size, oy, ty, overflow);
    for (int ox = threadIdx.x; ox < size; ox += blockDim.x) {
        AaTaps tx;
        aa_taps(cw > 0 ? cw : 1, size, ox, tx, overflow);
        float acc[3] = {0.f, 0.f, 0.f};
        for (int jy = 0; jy < ty.n; ++jy) {
            const int fy = y1 + ty.lo + jy;                   // frame row of this crop row (outside the frame: zero padding)
            float hacc[3] = {0.f, 0.f, 0.f};
            if (fy >= 0 && fy < frame_h && !overflow) {
                const uint8_t* row = img + static_cast<size_t>(fy) * frame_w * 3;
#pragma unroll
                for (int jx = 0; jx < kPreMaxTaps; ++jx) {
                    if (jx < tx.n) {
                        const int fx = x1 + tx.lo + jx;
                        if (fx >= 0 && fx < frame_w) {
                            const float w = tx.w[jx];
#pragma unroll
                            for (int c = 0; c < 3; ++c) hacc[c] += w * __fdiv_rn(static_cast<float>(row[fx * 3 + c]), 255.f);   // ToTensor
                        }
                    }
                }
            }
            float wy = 0.f;
#pragma unroll
            for (int j = 0; j < kPreMaxTaps; ++j) wy = j == jy ? ty.w[j] : wy;      // register select (no local-memory indexing)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] += wy * hacc[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[((static_cast<size_t>(n) * 3 + c) * size + oy) * size + ox] = __fdiv_rn(__fsub_rn(acc[c], norm.mean[c]), norm.std[c]);
    }
    if (overflow && threadIdx.x == 0) {
        *reinterpret_cast<volatile int*>(err_flag) = 41;       // box empty or more than ~9x larger than the output
        __threadfence_system();
    }
}

// ------------------------------------------------------------------------------------------------
// HRNet pieces (reference backbones/hrnet.py)
// ------------------------------------------------------------------------------------------------
// One CTA = one output row of one image; one thread = one output pixel, all 64 output channels in registers.  The three
// input rows (zero padded) and the 64 x 27 weights live in shared memory; weights are read as warp-wide broadcasts.
template <typename T>
__global__ void __launch_bounds__(128) hr_stem_kernel(const void* __restrict__ xin, int x_is_u8, StemNorm norm, const float* __restrict__ w,
                                                      const float* __restrict__ bias, T* __restrict__ out, int size) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float hs_smem[];
    const int wout = size / 2;
    float* sw = hs_smem;                         // [27][64]  (tap-major so that a thread's 64 outputs read consecutive floats)
    float* sb = sw + 27 * 64;                    // [64]
    float* sx = sb + 64;                         // [3 channels][3 rows][size + 2]
    const int oy = blockIdx.x, n = blockIdx.y;
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) { const int co = i / 27, t = i % 27; sw[t * 64 + co] = w[i]; }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = bias[i];
    const int pitch = size + 2;
    for (int i = threadIdx.x; i < 9 * pitch; i += blockDim.x) {
        const int c = i / (3 * pitch), r = (i / pitch) % 3, col = i % pitch;
        const int iy = 2 * oy - 1 + r, ix = col - 1;
        float v = 0.f;
        if (iy >= 0 && iy < size && ix >= 0 && ix < size) {
            const size_t idx = ((static_cast<size_t>(n) * 3 + c) * size + iy) * size + ix;
            if (x_is_u8) v = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(static_cast<const uint8_t*>(xin)[idx]), 255.f), norm.mean[c]), norm.std[c]);
            else v = static_cast<const float*>(xin)[idx];
        }
        sx[i] = v;
    }
    __syncthreads();
    for (int ox = threadIdx.x; ox < wout; ox += blockDim.x) {
        float acc[64];
#pragma unroll
        for (int co = 0; co < 64; ++co) acc[co] = 0.f;
        // (a fully unrolled tap loop makes ptxas hoist all 432 weight loads: 6 KB of spills per thread; one tap at a time)
#pragma unroll 1
        for (int t = 0; t < 27; ++t) {
            const int r = t / 9, s2 = (t / 3) % 3, c = t % 3;
            const float a = sx[(c * 3 + r) * pitch + 2 * ox + s2];
            const float4* wt = reinterpret_cast<const float4*>(sw + t * 64);      // weight layout [cout][r][s][cin] -> tap index (r*3+s)*3+c = t
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 wv = wt[q];
                acc[4 * q] = fmaf(a, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(a, wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(a, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(a, wv.w, acc[4 * q + 3]);
            }
        }
        // 16-byte stores (64 scalar 2-byte stores per thread made this kernel 30 % of the HRNet step)
        constexpr int VEC = 16 / sizeof(T);
        uint4* o = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * wout + oy) * wout + ox) * 64);
#pragma unroll
        for (int v = 0; v < 64 / VEC; ++v) {
            uint4 q;
            T* e = reinterpret_cast<T*>(&q);
#pragma unroll
            for (int i = 0; i < VEC; ++i) e[i] = from_f<T>(fmaxf(acc[v * VEC + i] + sb[v * VEC + i], 0.f));
            o[v] = q;
        }
    }
}

template <typename T>
int hr_stem_launch(const void* x, bool x_is_u8, const StemNorm& norm, const float* w, const float* bias, T* out, int n_img, int size,
                   cudaStream_t s) {
    if (n_img == 0) return 0;
    const size_t smem = (27 * 64 + 64 + 9 * (size + 2)) * sizeof(float);
    HMV_CUDA(launch_kernel(hr_stem_kernel<T>, dim3(size / 2, n_img), dim3(128), smem, s, x, x_is_u8 ? 1 : 0, norm, w, bias, out, size));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int hr_stem_launch<bf16>(const void*, bool, const StemNorm&, const float*, const float*, bf16*, int, int, cudaStream_t);
template int hr_stem_launch<float>(const void*, bool, const StemNorm&, const float*, const float*, float*, int, int, cudaStream_t);

template <typename T>
__global__ void fuse_sum_kernel(const FuseSumParams p, size_t total) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int VEC = 16 / sizeof(T);
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = p.C / VEC;
    const int c = static_cast<int>(idx % cv) * VEC;
    const int x = static_cast<int>((idx / cv) % p.W);
    const int y = static_cast<int>((idx / (static_cast<size_t>(cv) * p.W)) % p.H);
    const size_t n = idx / (static_cast<size_t>(cv) * p.W * p.H);
    float v[VEC];
    auto add = [&](const T* src, bool first) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
        const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = first ? to_f(e[i]) : v[i] + to_f(e[i]);
    };
    add(static_cast<const T*>(p.base) + ((n * p.H + y) * p.W + x) * p.C + c, true);
    for (int k = 0; k < p.n_up; ++k) {               // hrnet.py:226-230 adds the branches in ascending order
        const int sh = p.shift[k], hs = p.H >> sh, ws = p.W >> sh;
        add(static_cast<const T*>(p.up[k]) + ((n * hs + (y >> sh)) * ws + (x >> sh)) * p.C + c, false);
    }
    uint4 o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int i = 0; i < VEC; ++i) oe[i] = from_f<T>(p.relu ? fmaxf(v[i], 0.f) : v[i]);
    *reinterpret_cast<uint4*>(static_cast<T*>(p.out) + ((n * p.H + y) * p.W + x) * p.C + c) = o;
}

template <typename T>
int fuse_sum_launch(const FuseSumParams& p, cudaStream_t s) {
    if (p.n_img == 0) return 0;
    constexpr int VEC = 16 / sizeof(T);
    HMV_CHECK(p.C % VEC == 0 && p.n_up >= 0 && p.n_up <= 3, "fuse_sum: bad geometry");
    const size_t total = static_cast<size_t>(p.n_img) * p.H * p.W * (p.C / VEC);
    HMV_CUDA(launch_kernel(fuse_sum_kernel<T>, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, p, total));
    HMV_CUDA(cudaGetLastError());
    return 0;
}
template int fuse_sum_launch<bf16>(const FuseSumParams&, cudaStream_t);
template int fuse_sum_launch<float>(const FuseSumParams&, cudaStream_t);

int preprocess_launch(const uint8_t* frames, const int* bbox, float* out, int n_img, int frame_h, int frame_w, int size,
                      const StemNorm& norm, int* err_flag, cudaStream_t s) {
    if (n_img == 0) return 0;
    HMV_CHECK(frame_h > 0 && frame_w > 0 && size > 0, "preprocess: bad geometry");
    HMV_CUDA(launch_kernel(preprocess_kernel, dim3(size, n_img), dim3(256), 0, s, frames, bbox, out, frame_h, frame_w, size, norm, err_flag));
    HMV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hmv
