// fp32 "check mode" implicit-GEMM convolution (CUDA cores, FFMA).  Same math and the same
// epilogue as the tensor-core kernel, used when the handle is created with precision = fp32:
// it is the <=1e-4 validation path named by the north star, not the performance path.
// Covers every conv / linear of the forward (reference backbones/resnet.py:124-144,218;
// layers.py:318-334; layers.py:213-215,224,165-170).
#include "kernels.cuh"

namespace hmv {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
conv_f32_kernel(const ConvF32Params p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = tid % 16, ty = tid / 16;          // 16 x 16 threads, 4 x 4 outputs each
    const int lrow = tid / 4;                        // 0..63 : tile row (A) / tile col (B) this thread loads
    const int lk = (tid % 4) * 4;                    // 4 consecutive k

    // decode the output pixel this thread loads for A
    const int m = m0 + lrow;
    const bool m_ok = m < p.M;
    int img = 0, oh = 0, ow = 0;
    if (m_ok) {
        img = m / (p.Hout * p.Wout);
        const int r = m % (p.Hout * p.Wout);
        oh = r / p.Wout;
        ow = r % p.Wout;
    }
    const int n_ld = n0 + lrow;
    const bool n_ok = n_ld < p.Nalloc;

    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        const int k = k0 + lk;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m_ok && k < p.K) {
            const int tap = k / p.Cin, c = k % p.Cin;        // Cin % 4 == 0 -> 4 k's share a tap
            const int r = tap / p.kw, s = tap % p.kw;
            const int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
            if (ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win)
                a = __ldg(reinterpret_cast<const float4*>(
                    p.in + ((static_cast<size_t>(img) * p.Hin + ih) * p.Win + iw) * p.Cin + c));
        }
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n_ok && k < p.K) b = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(n_ld) * p.K + k));
        As[lk][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
        Bs[lk][lrow] = b.x; Bs[lk + 1][lrow] = b.y; Bs[lk + 2][lrow] = b.z; Bs[lk + 3][lrow] = b.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }

    const Epilogue& ep = p.ep;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        if (row >= ep.M) continue;
        const size_t rrow = ep.res_mode == RES_NONE
                                ? 0
                                : static_cast<size_t>(row / ep.res_group) * ep.res_stride + row % ep.res_group;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx * 4 + j;
            if (col >= ep.N) continue;
            float v = acc[i][j] + ep.bias[col];
            if (ep.res_mode == RES_F32) v += static_cast<const float*>(ep.residual)[rrow * ep.res_ld + col];
            if (ep.act == ACT_RELU) v = fmaxf(v, 0.f);
            else if (ep.act == ACT_GELU) v = gelu_erf(v);
            if (ep.out_mode == OUT_F32_ROWMAJOR) {
                static_cast<float*>(ep.out)[static_cast<size_t>(row) * ep.ldc + col] = v;
            } else if (ep.out_mode == OUT_F32_NCHW) {
                {
                    const int im = row / ep.hw, pix = row % ep.hw;
                    static_cast<float*>(ep.out)[(static_cast<size_t>(im) * ep.N + col) * ep.hw + pix] = v;
                }
            }
        }
    }
}

}  // namespace

int conv_f32_launch(const ConvF32Params& p, cudaStream_t stream) {
    if (p.M <= 0) return 0;
    HMV_CHECK(p.Cin % 4 == 0 && p.K % 4 == 0, "conv_f32: Cin and K must be multiples of 4");
    HMV_CHECK(p.ep.out_mode != OUT_BF16_ROWMAJOR && p.ep.res_mode != RES_BF16, "conv_f32: fp32 tensors only");
    dim3 grid((p.M + BM - 1) / BM, (p.Nalloc + BN - 1) / BN);
    conv_f32_kernel<<<grid, 256, 0, stream>>>(p);
    HMV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hmv
