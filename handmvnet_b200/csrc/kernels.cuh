// Launchers for the non-GEMM kernels of the path and the fp32 check-mode conv.
#pragma once
#include "common.cuh"

namespace hmv {

constexpr int kJoints = 21;

struct ConvF32Params {
    const float* in;     // NHWC [n_img, Hin, Win, Cin]
    const float* w;      // [Nalloc, K], K = kh*kw*Cin ordered (r, s, c)
    int Hin, Win, Cin, Hout, Wout, kh, kw, stride, pad;
    int M, K, Nalloc;
    Epilogue ep;
};
int conv_f32_launch(const ConvF32Params& p, cudaStream_t stream);

// x fp32 NCHW [n,3,H,W] -> zero-padded NHWC4 [n, Hp, Wp, 4]; pixel (h,w) lands at (h+pad, w+pad).
template <typename T>
int pack_input_launch(const float* x, T* out, int n_img, int H, int W, int Hp, int Wp, int pad, cudaStream_t s);

// Fused bf16 stem: conv7x7/2 + BN + ReLU + maxpool3x3/2 straight from fp32 NCHW input (stem_pool.cu).
// wpack: [7 taps][4 K-chunks][64 couts][8] bf16, element e of chunk kc = (pixel 2*kc + e/4, channel e%4).
// x: fp32 NCHW, or uint8 NCHW normalised on the fly with `norm` (ToTensor + Normalize of the reference data pipeline).
struct StemNorm { float mean[3]; float std[3]; };
int stem_pool_launch(const void* x, bool x_is_u8, const StemNorm& norm, const bf16* wpack, const float* bias, bf16* out, int n_img,
                     int num_sms, int* err_flag, cudaStream_t s);
// crop_and_pad_image + ToTensor + Resize(antialias) + Normalize (reference datasets/utils.py:40-77, ho3d.py:35-40):
// frames [n, frame_h, frame_w, 3] uint8, bbox [n, 4] int32 xyxy -> out [n, 3, size, size] fp32
int preprocess_launch(const uint8_t* frames, const int* bbox, float* out, int n_img, int frame_h, int frame_w, int size,
                      const StemNorm& norm, int* err_flag, cudaStream_t s);
int u8_to_f32_norm_launch(const uint8_t* in, float* out, size_t total, int hw, const StemNorm& norm, cudaStream_t s);

// 3x3 / stride 2 / pad 1 max pooling, NHWC (reference resnet.py:165,221).
template <typename T>
int maxpool_launch(const T* in, T* out, int n_img, int Hin, int Win, int C, cudaStream_t s);

// soft-argmax over H*W with temperature (reference models/utils.py:35-62); xy in heatmap pixels,
// xy_scaled = xy * scale (reference handmvnet.py:252).
int softargmax_launch(const float* hm, float* xy, float* xy_scaled, int n_maps, int H, int W, float temperature,
                      float scale, cudaStream_t s);

// Bilinear neighbours of every joint (reference nets.py:46-53 == grid_sample align_corners=True,
// zero padding): copies the 4 neighbour feature rows to rows[(n*21+j)*4+q][C] and writes the
// bilinear weights (0 for out-of-range neighbours) to wts.
template <typename T>
int sample_gather_launch(const T* feat, const float* xy, T* rows, float* wts, int n_img, int H, int W, int C,
                         cudaStream_t s);

struct TokenSource {       // one feature level's sampled rows (the ResNet configs have one level, HRNet four: handmvnet.py:185-187)
    const float* g;        // [n_img*21*4, ldg] conv+BN+ReLU of the 4 bilinear neighbours of every joint (fp32)
    const float* wts;      // [n_img*21*4] bilinear weights on THIS level (0 for neighbours outside the map)
    int ldg;
    int width;             // feature columns this level contributes (c_level / 2)
};
struct TokenParams {
    TokenSource src[4];
    int n_src;
    const float* xy;       // [n_img*21*2]
    const float* bbox;     // [n_img*4] xyxy or null
    const float* intr;     // [n_img*4] fx fy cx cy or null
    const float* pe;       // [tokens_per_sample, d] or null
    float* tok_f32;        // [n_img*21, pitch]
    void* tok_lp;          // [n_img*21, pitch] low-precision copy (bf16) or null in fp32 mode
    int n_img, feat, d, pitch, tokens_per_sample, use_pos2d, use_crop;
};
template <typename T>
int tokens_launch(const TokenParams& p, cudaStream_t s);

// per (sample, head) softmax(Q K^T * scale) V  (reference layers.py:217-223)
template <typename T>
int attention_launch(const T* qkv, int ld, T* out, int ld_out, int batch, int tokens_per_sample, int q_row0, int nq,
                     int kv_row0, int nk, int heads, int dim_head, float scale, cudaStream_t s);

// bf16 tensor-core (mma.sync) flash-style variant used on the bf16 path
int attention_mma_launch(const bf16* qkv, int ld, bf16* out, int ld_out, int batch, int tokens_per_sample, int q_row0,
                         int nq, int kv_row0, int nk, int heads, int dim_head, float scale, cudaStream_t s);

// h = LN(in; g1,b1) -> out_f32 ; out_lp = g2 ? LN(h; g2,b2) : h   (reference layers.py:228,165,233)
template <typename T>
int layernorm_launch(const float* in, int ld_in, const float* g1, const float* b1, float* out_f32, int ld_out,
                     const float* g2, const float* b2, T* out_lp, int ld_lp, int rows, int d, float eps,
                     cudaStream_t s);

// One fusion-transformer layer after its QKV projection, fused (fusion_block.cu; reference layers.py:217-236):
// attention -> to_out + residual -> norm1 -> ff (LayerNorm, Linear, GELU, Linear) + residual -> norm2.
// All d_model-wide vectors / weight rows are zero padded to 576; bf16 tensor-core path only.
struct FusionBlockParams {
    const bf16* qkv;       // [batch * s_in, 3072]  (q | k | v, 8 heads x 128 each)
    int ld_qkv;
    const float* res_in;   // [batch * s_in, pitch] fp32 master copy of the layer input (residual)
    int s_in, nq, nk, kv_row0;     // tokens per sample in; queries = tokens [0, nq); keys = tokens [kv_row0, kv_row0 + nk)
    int pitch, d;
    float scale_log2e;     // dim_head^-0.5 * log2(e)
    const bf16* wo;        // [576, 1024]
    const float* bo;       // [576]
    const float *g1, *b1, *gff, *bff, *g2, *b2;   // LayerNorm affine, [576] each
    const bf16* w1;        // [128, 576]
    const float* bf1;      // [128]
    const bf16* w2;        // [576, 128]
    const float* bf2;      // [576]
    float* out_f32;        // [batch * nq, pitch]
    bf16* out_lp;          // [batch * nq, pitch]
};
int fusion_block_launch(const FusionBlockParams& p, int batch, cudaStream_t s);

struct GcnParams {
    const float* x;        // [batch, 21, ld]
    int ld, d_in;
    const float* w[3];     // [3, cin, cout] each (reference layers.py:372 weight [K+1,1,cin,cout])
    const float* b[3];
    const float* basis;    // [3, 21, 21] Chebyshev T_k
    float* out;            // [batch, 21, 3]
    int batch;
};
// h1_scratch: [batch, 21, 256] fp32 workspace for the first layer's output
int gcn_launch(const GcnParams& p, float* h1_scratch, cudaStream_t s);

// ---- HRNet backbone pieces (reference backbones/hrnet.py) ---------------------------------------------------------
// First stem conv: 3x3 / stride 2 / pad 1, 3 -> 64 channels, + folded BN + ReLU (hrnet.py:244-245,380-382), straight from
// the fp32 (or uint8, normalised on the fly) NCHW input to an NHWC activation.  w: [64][27] (tap-major (r, s), channel
// minor) fp32, bias [64].  CUDA cores: K = 27 is no tensor-core shape.
template <typename T>
int hr_stem_launch(const void* x, bool x_is_u8, const StemNorm& norm, const float* w, const float* bias, T* out, int n_img, int size,
                   cudaStream_t s);
// Branch fusion (hrnet.py:222-231): out = [relu](base + sum_k nearest_upsample(up[k], 2^shift[k])), NHWC with `C` channels
// (a multiple of 8); up[k] is [n, H >> shift, W >> shift, C].
struct FuseSumParams {
    const void* base;
    const void* up[3];
    int shift[3];
    int n_up;
    void* out;
    int n_img, H, W, C, relu;
};
template <typename T>
int fuse_sum_launch(const FuseSumParams& p, cudaStream_t s);

}  // namespace hmv
