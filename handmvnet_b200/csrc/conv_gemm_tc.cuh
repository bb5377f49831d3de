// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// One persistent, warp-specialised kernel covers every GEMM-shaped op of the path
// (reference call sites: backbones/resnet.py:124-144 Bottleneck convs, :218 stem,
// layers.py:318-334 pose_net / SampleNet convs, layers.py:213-215,224,165-170 the fusion
// linears):   D[m, n] = act( sum_taps sum_c A_tap[m, c] * W[n, tap, c] + bias[n] (+ res[m, n]) )
//
//   * A (activations, NHWC bf16) is described by ONE 5-D TMA tensor map
//     (C, W', A, H', N); a "tap" is a coordinate offset (c_off, dw, a, dh) into that map, so
//     3x3/pad-1 convs are 9 shifted box loads with hardware zero fill, stride-2 convs address
//     the input through its (row-parity, col-parity) view, and 1x1 convs / linears are the
//     single-tap "flat" case (the 7x7/2 stem has its own fused kernel, stem_pool.cu).
//   * W is [N_alloc, K] K-major bf16 (BN scale folded in), one 2-D tensor map.
//   * Both land in 128B-swizzled shared memory (4..8 stage mbarrier ring) and feed
//     tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) issued by one thread; the fp32
//     accumulator lives in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of
//     tile i+1.  Epilogue warps read TMEM with tcgen05.ld (32 lanes x 32b), add the folded-BN
//     bias, the optional residual, apply ReLU/GELU and write bf16 NHWC through a swizzled
//     shared-memory staging ring + TMA stores (or fp32 / fp32-NCHW straight from registers).
#pragma once
#include "common.cuh"

namespace hmv {

constexpr int kTcBlockM = 128;
constexpr int kTcBlockK = 64;          // bf16 elements -> 128 B = one swizzle-128B row
constexpr int kTcUmmaK = 16;
constexpr int kTcThreads = 384;        // warp0 TMA, warp1 MMA + TMEM alloc, warp2 residual TMA, warps 4-11 epilogue
constexpr int kTcMaxTaps = 9;

struct TcTap { int c_off, dw, a, dh; };

struct TcParams {
    int num_m_tiles, num_n_tiles;
    int num_taps, cblks;               // K blocks per tile = num_taps * cblks
    int flat;                          // 1: A is [M, K] (coord1 = m_tile*128); 0: spatial tile
    int tpi, hbox;                     // spatial: tiles per image, rows of H' per tile
    int a_bytes;                       // bytes one A box delivers (16 KiB; less when an image has fewer than 128 pixels: 8 x 8 maps)
    int tile_rows;                     // output rows a tile owns (128; 64 for 8 x 8 maps: the upper half of the MMA tile is unused)
    TcTap taps[kTcMaxTaps];
    Epilogue ep;
    int* err_flag;                     // set to non-zero on an mbarrier timeout
    int bias_in_params;                // 1: biases come from the BiasBank kernel parameter (N_alloc <= kBiasBankFloats), 0: from ep.bias
};

// Epilogue variants.  DIRECT: registers -> global (fp32 / NCHW / remapped-residual outputs).
// STORE*: bf16 NHWC outputs staged through 128B-swizzled shared memory and written with TMA stores
// (64-column x 32-row slabs per epilogue warp).  STORE = 2 staging slots (mainloop-bound layers, keeps 4
// operand stages at BN=256); STORE_DEEP = 4 slots (store-bound small-K layers); STORE_RES = 2 store slots plus
// a separate 3-slot ring into which a dedicated producer warp prefetches the bf16 residual tile with TMA.
enum TcMode { TC_DIRECT = 0, TC_STORE = 1, TC_STORE_DEEP = 2, TC_STORE_RES = 3 };

struct TcLaunch {                      // everything needed to enqueue one layer
    BiasBank bank;                     // [N_alloc] biases when p.bias_in_params
    CUtensorMap tmA, tmB;
    CUtensorMap tmC, tmR;              // output / residual maps (TC_STORE* only)
    TcParams p;
    int bn;
    int mode;
    int cluster;                       // 1, or 2: CTA pairs with multicast weight tiles (BN = 256 only; tmB box is BN/2 rows)
};

// Host API -------------------------------------------------------------------------------------
int tc_init();                                               // resolves cuTensorMapEncodeTiled, sets smem attrs
int tc_make_tmap_act(CUtensorMap* out, const void* base, const uint64_t dims[5],
                     const uint64_t strides_bytes[4], const uint32_t box[5]);
int tc_make_tmap_wgt(CUtensorMap* out, const void* base, uint64_t k_total, uint64_t n_alloc, int bn);
// [rows, cols] bf16 row-major tensor, box = 64 columns x box_rows rows (output slabs: 32, residual tiles: 128)
int tc_make_tmap_out(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, int box_rows);
int tc_pick_bn(int n);                                       // tile width for a given output width
int tc_launch(const TcLaunch& l, int num_sms, cudaStream_t stream);

// ---- fused bottleneck tail (bottleneck_tc.cu): conv3x3+BN+ReLU -> conv1x1+BN+residual+ReLU in one kernel ----------
struct BtParams {
    int num_m_tiles;
    int tpi, hbox, cblks;              // conv2 tile geometry (as TcParams) ; cblks = P / 64
    TcTap taps[kTcMaxTaps];
    int* err_flag;
    long long* prof;                   // optional [grid][16] stall counters (HMV_BT_PROF=1), else null
    int prefetch;                      // 1: L2-prefetch the next tile's residual (HMV_BN_PREFETCH=0 disables)
    int ds_kb;                         // > 0: the block's 1x1 stride-1 downsample is folded into conv3 as ds_kb extra K blocks per chunk
                                       //      (A = the block input through `tmRes`, B = `tmWd`); no residual is read then
};
constexpr int kBtBias3Off = 256;       // BiasBank layout of the fused tail: conv2 biases at [0, P), conv3 biases at [256, 256 + 4P)
struct BtLaunch {
    BiasBank bank;
    CUtensorMap tmA, tmW2, tmY2s, tmY2l, tmW3, tmOut, tmRes, tmWd;
    BtParams p;
    int planes;                        // P in {64, 128, 256}
    int pair;                          // 1: CTA pairs issuing cta_group::2 MMAs (tmW2 / tmW3 / tmWd carry half-height boxes)
};
int bt_init();                                               // shared-memory attributes of the three instantiations
int bt_launch(const BtLaunch& l, int num_sms, cudaStream_t stream);

// ---- fused bottleneck seam (bottleneck_next_tc.cu): conv3 + residual + ReLU of block b -> conv1 + ReLU of block b+1 (P = 256) ----
struct BnParams {
    int num_m_tiles;
    int* err_flag;
    long long* prof;                   // optional [grid][24] stall counters (HMV_BN_PROF=1), else null
};
constexpr int kBnBias1Off = 1024;      // BiasBank layout of the fused seam: conv3 biases at [0, 1024), next conv1 biases at [1024, 1280)
struct BnLaunch {
    BiasBank bank;
    CUtensorMap tmY2, tmW3, tmRes, tmOut, tmW1, tmY1;
    BnParams p;
    int pair;                          // 1: CTA pairs issuing cta_group::2 MMAs (tmW3 / tmW1 are half-height boxes)
};
int bn_init();
int bn_launch(const BnLaunch& l, int num_sms, cudaStream_t stream);

}  // namespace hmv
