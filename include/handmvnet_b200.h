/*
 * handmvnet_b200 - C ABI of the B200-native HandMvNet inference forward path.
 *
 * The reference (pyxploiter/HandMvNet) is pure Python/PyTorch and has no FFI of its own; this
 * header is the boundary a maintainer binds instead of calling the torch modules.  Each entry
 * point names the reference interface it replaces (paths relative to the reference root).
 * Conventions: plain pointers and sizes only (no torch / C++ types), every function returns 0 on
 * success and non-zero on failure with the message available from hmv_last_error(); no exception
 * crosses this boundary; all device pointers are fp32 and caller-owned; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  One handle per device, re-entrant per
 * handle, not thread-safe on one handle.
 */
#ifndef HANDMVNET_B200_H
#define HANDMVNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hmv_handle hmv_handle;

enum { HMV_PRECISION_BF16 = 0, HMV_PRECISION_FP32 = 1 };

/* Mirrors the keys HandMvNet.__init__ reads from cfg["model"], cfg["data"], cfg["train"]
 * (src/models/handmvnet.py:28-125, configs/release/*.yaml). */
typedef struct hmv_config {
    int32_t num_views;        /* len(model.selected_views)                          */
    int32_t image_size;       /* data.image_size   (256)                            */
    int32_t heatmap_size;     /* data.heatmap_size (32)                             */
    int32_t use_pos2d;        /* "pos2d" in model.pos_enc                           */
    int32_t use_crop;         /* "crop"  in model.pos_enc  (needs bbox + intrinsic)  */
    int32_t use_sin;          /* "sin"   in model.pos_enc                           */
    int32_t fusion_layers;    /* model.fusion_layers (odd)                          */
    int32_t precision;        /* HMV_PRECISION_BF16 (tcgen05 path) | HMV_PRECISION_FP32 (check mode) */
    int32_t micro_batch;      /* samples processed per internal pass (workspace is sized for it) */
    int32_t device;           /* CUDA device ordinal                                */
    int32_t backbone;         /* model.backbone: HMV_BACKBONE_RESNET50_PAPER (0, default) | HMV_BACKBONE_HRNET */
    int32_t hr_channels[4];   /* model.backbone_channels of the HRNet configs (w40: 40, 80, 160, 320)          */
} hmv_config;
enum { HMV_BACKBONE_RESNET50_PAPER = 0, HMV_BACKBONE_HRNET = 1 };

/* HandMvNet(train_params, model_params, data_params)  - src/models/handmvnet.py:28 */
int hmv_create(const hmv_config* cfg, hmv_handle** out);
int hmv_destroy(hmv_handle* h);

/* model.load_state_dict(strict=True) - src/eval.py:46-50.  `name` is the reference state_dict key
 * (355 keys, e.g. "backbone.layer3.0.conv2.weight"); data is host fp32, dims as in the reference.
 * Two extra optional constants the reference builds in its constructors may be supplied:
 * "pe" [21*V, feat_dim] (layers.py:136-150) and "cheb_basis" [3,21,21] (layers.py:405-445). */
int hmv_set_weight(hmv_handle* h, const char* name, const float* data, const int64_t* dims, int32_t ndim);
/* Folds BatchNorm, repacks to the kernel layouts, uploads, builds TMA descriptors.  Fails if a
 * key is missing (strict). */
int hmv_prepare(hmv_handle* h);

/* model(x, bbox, cam_params) - src/models/handmvnet.py:158-266 (callers: src/eval_fps.py:83,91;
 * validation_step/test_step :468-516).  Device pointers:
 *   x        [B, V, 3, S, S] fp32 NCHW          bbox [B, V, 4] xyxy (may be NULL without "crop")
 *   intr     [B, V, 4] (fx, fy, cx, cy)  (cam_params["intrinsic"]; may be NULL without "crop")
 *   heatmap  [B, V, 21, 32, 32]   joints_crop_img [B, V, 21, 2]   joints_cam [B, 21, 3]
 * Asynchronous on `stream`. */
int hmv_forward(hmv_handle* h, const float* x, const float* bbox, const float* intr, int32_t batch,
                float* heatmap, float* joints_crop_img, float* joints_cam, void* stream);

/* Same call with HOST buffers (pinned recommended): chunks of micro_batch samples are copied
 * host->device on a copy stream overlapped with compute, results are copied back; returns after
 * everything has completed.  Any output pointer may be NULL to skip that copy. */
int hmv_forward_host(hmv_handle* h, const float* x, const float* bbox, const float* intr, int32_t batch,
                     float* heatmap, float* joints_crop_img, float* joints_cam);

/* Streaming form of hmv_forward_host for a caller that feeds batch after batch (the loop of
 * src/eval_fps.py:79-92 / the Lightning test loop): enqueues the copies and kernels of one batch and returns
 * a ticket without waiting, so the host->device copies of call k+1 overlap the compute of call k.  At most 4
 * calls may be in flight; host buffers (inputs AND outputs) must stay valid until hmv_host_wait(ticket) has
 * returned.  Tickets complete in issue order. */
int hmv_forward_host_async(hmv_handle* h, const float* x, const float* bbox, const float* intr, int32_t batch,
                           float* heatmap, float* joints_crop_img, float* joints_cam, int64_t* ticket);
int hmv_host_wait(hmv_handle* h, int64_t ticket);

/* uint8 inputs: x is [B, V, 3, S, S] uint8 NCHW (what torchvision's ToTensor sees before its /255) and the stem kernel
 * applies ((v / 255) - mean[c]) / std[c] on the fly - the ToTensor + Normalize step of the reference data pipeline
 * (src/datasets/ho3d.py:35-40; defaults are its ImageNet constants) - so the host->device traffic of a batch drops 4x.
 * Same IEEE operations in the same order as the host transform: results are bit-identical to hmv_forward on the
 * host-normalised fp32 tensor.  hmv_forward_u8: device pointers, asynchronous on `stream`;
 * hmv_forward_host_u8_async: host pointers, ticket semantics of hmv_forward_host_async. */
int hmv_set_input_norm(hmv_handle* h, const float* mean3, const float* std3);

/* The image transform the reference's dataset applies in front of the model, on the device: crop_and_pad_image
 * (src/datasets/utils.py:40-77; box parts outside the frame are zero) + ToTensor + Resize((S, S), antialias=True) +
 * Normalize (src/datasets/ho3d.py:35-40, 139-147).  frames [n_img, frame_h, frame_w, 3] uint8 HWC, bbox [n_img, 4]
 * int32 xyxy (x2 > x1, y2 > y1, side <= 9 * S), x_out [n_img, 3, S, S] fp32 - ready for hmv_forward.  Device
 * pointers, asynchronous on `stream`; an invalid box is reported by the next call / hmv_synchronize. */
int hmv_preprocess(hmv_handle* h, const uint8_t* frames, const int32_t* bbox, int32_t n_img, int32_t frame_h,
                   int32_t frame_w, float* x_out, void* stream);
int hmv_forward_u8(hmv_handle* h, const uint8_t* x, const float* bbox, const float* intr, int32_t batch,
                   float* heatmap, float* joints_crop_img, float* joints_cam, void* stream);
int hmv_forward_host_u8_async(hmv_handle* h, const uint8_t* x, const float* bbox, const float* intr, int32_t batch,
                              float* heatmap, float* joints_crop_img, float* joints_cam, int64_t* ticket);

/* Blocks until the handle's work is done and reports device-side pipeline errors. */
int hmv_synchronize(hmv_handle* h);

/* ---- per-stage entry points (teacher-forced parity tests, micro-benchmarks) ------------------ */
enum {
    HMV_STAGE_BACKBONE = 0,   /* resnet.py:216-239       x -> FEAT                          */
    HMV_STAGE_POSE = 1,       /* handmvnet.py:180-182    FEAT -> HEATMAP, XY                */
    HMV_STAGE_SAMPLE = 2,     /* handmvnet.py:185-225 + layers.py:157  FEAT, XY -> TOKENS   */
    HMV_STAGE_FUSION = 3,     /* fusion.py:26-30         TOKENS -> FUSED                    */
    HMV_STAGE_GCN = 4,        /* nets.py:133-139         FUSED -> JOINTS                    */
    HMV_STAGE_SOFTARGMAX = 5  /* models/utils.py:35-62   HEATMAP -> XY (soft_argmax_2d alone) */
};
enum {
    HMV_T_FEAT = 0,           /* [n*V, 1024, 32, 32] NCHW (HRNet: level 0, [n*V, C_0, 64, 64])  */
    HMV_T_HEATMAP = 1,        /* [n*V, 21, 32, 32]                                          */
    HMV_T_XY = 2,             /* [n*V, 21, 2] heatmap pixels                                 */
    HMV_T_TOKENS = 3,         /* [n, 21*V, feat_dim] (positional encoding already added)     */
    HMV_T_FUSED = 4,          /* [n, 21, feat_dim]                                           */
    HMV_T_JOINTS = 5,         /* [n, 21, 3]                                                  */
    HMV_T_FEAT1 = 6,          /* HRNet only: feature levels 1..3, [n*V, C_l, 64 >> l, 64 >> l] (HMV_T_FEAT is level 0 there) */
    HMV_T_FEAT2 = 7,
    HMV_T_FEAT3 = 8
};
/* Runs one stage on the handle's internal tensors for `batch` <= micro_batch samples.
 * x / bbox / intr are only read by the stages that need them (others may pass NULL). */
int hmv_stage_run(hmv_handle* h, int32_t stage, const float* x, const float* bbox, const float* intr,
                  int32_t batch, void* stream);
/* Copy an internal tensor to / from a dense fp32 device buffer in the reference's layout. */
int hmv_tensor_get(hmv_handle* h, int32_t tensor, float* dst, int32_t batch, void* stream);
int hmv_tensor_set(hmv_handle* h, int32_t tensor, const float* src, int32_t batch, void* stream);

/* Backbone bisect helper: run the first `num_steps` entries of the backbone plan on n_img images
 * and export the output of the last one as NCHW fp32; chw receives (C, H, W). */
int hmv_debug_backbone(hmv_handle* h, const float* x, int32_t n_img, int32_t num_steps, float* out,
                       int32_t* chw, void* stream);
int hmv_debug_num_steps(hmv_handle* h);
const char* hmv_debug_step_name(hmv_handle* h, int32_t step);
/* One backbone plan step alone, TEACHER-FORCED (per-kernel parity of the fused bottleneck kernels against
 * backbones/resnet.py:124-144): hmv_debug_step_io describes the activations the step reads and writes as
 * "in <oracle tap> C H W;...;out <oracle tap> C H W;" (empty "in" list: the step reads the network input and
 * cannot be forced); hmv_debug_step_run imports `inputs[i]` (NCHW fp32 device buffers, in the order listed),
 * runs the step's kernel and exports every output to `outputs[i]` (NCHW fp32). */
int hmv_debug_step_io(hmv_handle* h, int32_t step, char* buf, int32_t buflen);
int hmv_debug_step_run(hmv_handle* h, int32_t step, int32_t n_img, const float* const* inputs,
                       float* const* outputs, void* stream);

/* One implicit-GEMM convolution through the same kernels the model uses (micro-benchmark /
 * kernel unit test):  out = act(conv(in, w) * scale + shift (+ residual)),  NCHW fp32 in/out,
 * w [cout, cin, k, k], scale/shift per output channel (the folded BatchNorm), residual optional.
 * Replaces nn.Conv2d + nn.BatchNorm2d + ReLU call sites (backbones/resnet.py:127-141). */
int hmv_conv_bn_act(int32_t precision, const float* in, const float* w, const float* scale, const float* shift,
                    const float* residual, float* out, int32_t n_img, int32_t cin, int32_t hin, int32_t win,
                    int32_t cout, int32_t ksize, int32_t stride, int32_t relu, float* elapsed_ms, int32_t iters,
                    void* stream);

/* Measurement support (bench.py roofline leg): when enabled every launch of the tensor-core implicit-GEMM
 * kernel is bracketed by CUDA events on its stream.  hmv_profile_read() synchronises, returns the summed device
 * time, the algorithmic FLOPs (2*M*N*K over the real extents) and the launch count since the last read, and
 * writes one CSV line per launch when csv_path is non-NULL. */
int hmv_profile_enable(hmv_handle* h, int32_t enable);
int hmv_profile_read(hmv_handle* h, double* tc_ms, double* tc_flops, int64_t* tc_launches, const char* csv_path);
/* Stream time (ms, launch gaps included) of the profiled passes per phase: backbone, heads, fusion, graph head. */
int hmv_profile_phases(hmv_handle* h, double* out4);

/* Number of kernels enqueued by the handle so far (bench.py's gpu_launches claim). */
int64_t hmv_launch_count(hmv_handle* h);
int hmv_num_sms(hmv_handle* h);
const char* hmv_last_error(void);
const char* hmv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HANDMVNET_B200_H */
