// Handle, weight repacking (BN fold), execution plan and the C ABI of handmvnet_b200.
// Reference behaviour being replaced: src/models/handmvnet.py:158-266 (forward) and the module
// constructors it relies on (see include/handmvnet_b200.h for the per-entry-point citations).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/handmvnet_b200.h"
#include "conv_gemm_tc.cuh"
#include "kernels.cuh"

namespace hmv {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }
bool clusters_enabled() {
    static const bool on = [] { const char* e = getenv("HMV_CLUSTER"); return !(e && e[0] == '0'); }();
    return on;
}
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("HMV_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

// ------------------------------------------------------------------------------------------------
// layout conversion utilities (stage import/export, unit-test entry point)
// ------------------------------------------------------------------------------------------------
// `ld` = channel pitch of the NHWC buffer (>= C; HRNet-w40 buffers are wider than the layer, extra channels are zero)
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int C, int HW, size_t total, int ld = 0) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over NCHW output
    if (idx >= total) return;
    if (ld == 0) ld = C;
    const int p = static_cast<int>(idx % HW);
    const int c = static_cast<int>((idx / HW) % C);
    const size_t n = idx / (static_cast<size_t>(HW) * C);
    out[idx] = to_f(in[(n * HW + p) * ld + c]);
}
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int C, int HW, size_t total, int ld = 0) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over the (pitched) NHWC output; total = n * HW * ld
    if (idx >= total) return;
    if (ld == 0) ld = C;
    const int c = static_cast<int>(idx % ld);
    const int p = static_cast<int>((idx / ld) % HW);
    const size_t n = idx / (static_cast<size_t>(HW) * ld);
    out[idx] = c < C ? from_f<T>(in[(n * C + c) * HW + p]) : from_f<T>(0.f);
}
// dense fp32 [rows, d] <-> pitched fp32 (+ optional low-precision copy)
template <typename T>
__global__ void rows_import_kernel(const float* __restrict__ src, float* __restrict__ dst_f32, T* __restrict__ dst_lp,
                                   int d, int pitch, size_t total) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % d);
    const size_t r = idx / d;
    const float v = src[idx];
    dst_f32[r * pitch + c] = v;
    if (dst_lp) dst_lp[r * pitch + c] = from_f<T>(v);
}
__global__ void rows_export_kernel(const float* __restrict__ src, float* __restrict__ dst, int d, int pitch, size_t total) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    dst[idx] = src[(idx / d) * pitch + idx % d];
}
static inline unsigned nblk(size_t total) { return static_cast<unsigned>((total + 255) / 256); }

// ------------------------------------------------------------------------------------------------
// one GEMM-shaped layer
// ------------------------------------------------------------------------------------------------
enum LayerKind { LK_FLAT = 0, LK_CONV3_S1 = 1, LK_CONV_S2 = 2, LK_STEM = 3 };

struct Layer {
    std::string name;
    int kind = LK_FLAT;
    int cin = 0, cout = 0, ksize = 1, stride = 1, pad = 0;
    int cin_real = 0;                             // input channels of the layer itself (cin is the buffer's, possibly zero-padded, width)
    int hin = 1, win = 1, hout = 1, wout = 1;     // per image (1x1 for linears)
    int K = 0;                                    // GEMM K in the active precision
    int n_alloc = 0, bn = 0;
    int max_units = 0;                            // images (convs) or rows (linears) the maps cover
    void* w = nullptr;                            // device [n_alloc, K] (bf16 | fp32)
    float* bias = nullptr;                        // device [n_alloc]
    std::vector<float> bias_host;                 // the same values on the host (BiasBank kernel parameters)
    const void* in = nullptr;
    Epilogue ep{};
    TcLaunch tc{};
    int rows_per_unit() const { return hout * wout; }
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> dims;
};

enum StepKind { SK_PACK = 0, SK_GEMM = 1, SK_MAXPOOL = 2, SK_STEM_POOL = 3, SK_TAIL = 4, SK_SEAM = 5, SK_HR_STEM = 6, SK_FUSE_SUM = 7 };
struct StepIO {                                    // one activation tensor a backbone step reads / writes (NHWC in the workspace)
    void* ptr;
    std::string tap;                              // name of the oracle tap holding the same tensor (oracle.backbone(per_layer=True))
    int C, H, W;
    int ld = 0;                                   // channel pitch of the buffer (0 = C)
};
struct Step {
    int kind;
    std::string name;
    int layer = -1;                               // index into layers (SK_GEMM) or tails (SK_TAIL)
    const void* in = nullptr;
    void* out = nullptr;
    int C = 0, H = 0, W = 0;                      // output geometry (NHWC)
    std::vector<StepIO> ins, outs;                // teacher-forced single-step runs (hmv_debug_step_run)
    FuseSumParams fuse{};                         // SK_FUSE_SUM
};

// conv2 + conv3 (+residual) of one bottleneck as a single launch (bottleneck_tc.cu)
struct FusedTail {
    std::string name;
    int l2 = -1, l3 = -1;                         // the two layers it covers (weights, biases, maps live there)
    int l_ds = -1;                                // the block's downsample when it is folded into conv3 (layer1.0)
    BtLaunch bt{};
};

// conv3 (+residual) of block b + conv1 of block b+1 as a single launch (bottleneck_next_tc.cu, layer3 only)
struct FusedSeam {
    std::string name;
    int l3 = -1, l1 = -1;
    BnLaunch bn{};
};

struct FusionLayerPlan {
    int qkv, outp, ff1, ff2;                      // layer indices
    int s_in, nq, nk, kv_row0;
    const float *g1, *b1, *gff, *bff, *g2, *b2;
    float *res_in, *out_f32;                      // token stream in / out (fp32 master)
    void *in_lp, *out_lp;
    // bf16 path, fused layer kernel (fusion_block.cu): d_model-wide vectors / weight rows zero padded to the pitch
    void *wo_p = nullptr, *w2_p = nullptr;
    float *bo_p = nullptr, *b2_p = nullptr, *ln_p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace hmv

using namespace hmv;

struct hmv_handle {
    hmv_config cfg{};
    bool bf16 = true;
    int V = 0, S = 0, d = 0, pitch = 576, feat = 512, hm = 32, img = 256;
    int mb = 0, mb_img = 0, num_sms = 148, esz = 2;
    int fcap = 0;                                 // samples the fusion + graph-head stage handles per pass (>= mb)
    bool prepared = false;
    bool use_clusters = hmv::clusters_enabled();  // HMV_CLUSTER=0: no CTA pairs / multicast weight tiles
    bool fuse_block = true;                       // HMV_FUSION_UNFUSED=1: attention / projections / LayerNorms as separate kernels
    int64_t launches = 0;
    std::map<std::string, HostTensor> weights;
    std::vector<void*> allocs;
    std::vector<Layer> layers;
    std::vector<Step> backbone;
    std::vector<FusedTail> tails;
    std::vector<FusedSeam> seams;
    bool fuse_next = true;                        // HMV_FUSE_NEXT=0: layer3 conv3(b) and conv1(b+1) as separate kernels
    int fuse_mask = 3;                            // bottleneck widths whose conv2+conv3 run fused: bit0 P=64, bit1 P=128, bit2 P=256 (HMV_FUSE_TAIL=<mask>)
    std::vector<FusionLayerPlan> fusion;
    int pose0 = -1, pose3 = -1, samp = -1;
    // buffers
    void *xpad = nullptr, *bufX = nullptr, *bufY = nullptr, *bufT1 = nullptr, *bufT2 = nullptr, *bufDS = nullptr;
    void* featbuf = nullptr;                      // where the backbone output lives (bufX or bufY)
    float *hm_int = nullptr, *xy = nullptr, *xy_scaled = nullptr, *wts = nullptr, *pe = nullptr, *basis = nullptr;
    float *tok0_f32 = nullptr, *tokA_f32 = nullptr, *tokB_f32 = nullptr, *ybuf = nullptr, *hbuf = nullptr, *y2buf = nullptr;
    void *tok0_lp = nullptr, *tokA_lp = nullptr, *tokB_lp = nullptr, *qkvbuf = nullptr, *attbuf = nullptr, *hnbuf = nullptr, *f1buf = nullptr;
    float* fused_f32 = nullptr;                   // final fusion output (one of tokA/tokB)
    float* joints_int = nullptr;
    float* gcn_h1 = nullptr;
    void* stem_w = nullptr;                       // fused stem (bf16 path): packed weights + folded-BN bias
    float* stem_b = nullptr;
    // HRNet backbone (cfg.backbone == HMV_BACKBONE_HRNET): four feature levels, level l is [n, 64 >> l, 64 >> l, hr_cp[l]]
    bool hr = false;
    int hr_c[4] = {0, 0, 0, 0}, hr_cp[4] = {0, 0, 0, 0};   // real / buffer channel counts (tensor-core path pads 40/80/160 to 64/128/192)
    void* hr_lvl[4] = {nullptr, nullptr, nullptr, nullptr};
    float* hr_stem_w = nullptr;                   // first stem conv [64][27] fp32 (CUDA-core kernel)
    int hr_samp[4] = {-1, -1, -1, -1};            // per-level SampleNet conv layers
    void* hr_rows[4] = {nullptr, nullptr, nullptr, nullptr};    // gathered neighbour rows per level
    float* hr_g[4] = {nullptr, nullptr, nullptr, nullptr};      // sampled conv outputs per level (fp32)
    float* hr_wts[4] = {nullptr, nullptr, nullptr, nullptr};    // bilinear weights per level
    // uint8 inputs (hmv_forward_u8 / hmv_forward_host_u8_async): normalised inside the stem kernel
    bool x_u8 = false;                            // element type of the x pointer of the call being enqueued
    StemNorm norm{{0.485f, 0.456f, 0.406f}, {0.229f, 0.224f, 0.225f}};   // datasets/ho3d.py:35-40
    float* u8_f32 = nullptr;                      // fp32 check mode: normalised copy of a uint8 micro-batch
    float* gcn_w[3] = {nullptr, nullptr, nullptr};
    float* gcn_b[3] = {nullptr, nullptr, nullptr};
    float *bbox_int = nullptr, *intr_int = nullptr;
    int* err_flag_host = nullptr;                 // pinned + mapped
    int* err_flag_dev = nullptr;
    // host-buffer pipeline
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    float* xstage[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    float *d_bbox = nullptr, *d_intr = nullptr, *d_hm = nullptr, *d_xy = nullptr, *d_j = nullptr;
    int host_cap = 0;
    // asynchronous host calls (hmv_forward_host_async / hmv_host_wait): the staging ring runs across calls, so the
    // copies of call k+1 overlap the compute of call k; one completion event per in-flight call
    int64_t chunk_seq = 0;                        // staging chunks issued so far (buffer = chunk_seq & 1)
    static constexpr int kMaxInflight = 4;
    cudaEvent_t ev_done[kMaxInflight] = {nullptr, nullptr, nullptr, nullptr};
    int64_t tickets_issued = 0, tickets_waited = 0;
    // CUDA graphs for small batches (launch-bound regime, B=1 latency): one instantiated graph per batch size over
    // internal I/O buffers; calls 1 and 2 with a batch size run eagerly / capture, later calls replay.
    struct GraphSlot { int calls = 0; int kernels = 0; cudaGraphExec_t exec = nullptr; };
    cudaStream_t graph_stream = nullptr;          // capture / replay stream (the caller's may be the legacy default stream)
    cudaEvent_t graph_in = nullptr, graph_out = nullptr;
    std::map<int, GraphSlot> graphs;
    int graph_max_batch = 0;                      // 0 = disabled
    // Larger batches: a caller that runs the same buffers step after step (a benchmark / serving loop) gets a graph captured
    // over ITS pointers (no staging copies): first sight of a pointer set runs eagerly, the second captures, later ones replay.
    struct PtrGraph {
        const void *x, *bbox, *intr; void *hm, *j2d, *j3d; int batch;
        int seen; int kernels; cudaGraphExec_t exec; int64_t last_use;
    };
    std::vector<PtrGraph> ptr_graphs;
    bool ptr_graphs_ok = true;
    int64_t ptr_graph_clock = 0;
    float *g_x = nullptr, *g_bbox = nullptr, *g_intr = nullptr, *g_hm = nullptr, *g_xy = nullptr, *g_j = nullptr;
    // optional per-launch profiling of the tensor-core GEMM kernel (bench.py roofline leg)
    bool profiling = false;
    struct ProfRec { int layer; int units; cudaEvent_t e0, e1; int tail; int seam = -1; };   // tail / seam >= 0: fused launches
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, cudaEvent_t>> phase_marks;   // (phase id, event) recorded while profiling
    // The entry points share ONE workspace but may be called on different streams (the caller's, graph_stream,
    // compute_stream): every entry point first makes its stream wait for the last work that touched the workspace.
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_valid = false;
};

namespace hmv {

static int dev_alloc(hmv_handle* h, void** p, size_t bytes, bool zero = true) {
    HMV_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    h->allocs.push_back(*p);
    if (zero) HMV_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16));
    return 0;
}
template <typename P>
static int dev_alloc_t(hmv_handle* h, P** p, size_t bytes, bool zero = true) {
    return dev_alloc(h, reinterpret_cast<void**>(p), bytes, zero);
}

// Switches to the handle's device for the duration of a C entry point and restores the caller's device afterwards.
struct DeviceGuard {
    int prev = -1, want = -1;
    explicit DeviceGuard(int dev) : want(dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// workspace ordering across streams (see hmv_handle::ws_event)
static int ws_acquire(hmv_handle* h, cudaStream_t s) {
    if (h->ws_valid && h->ws_stream != s) HMV_CUDA(cudaStreamWaitEvent(s, h->ws_event, 0));
    return 0;
}
static int ws_release(hmv_handle* h, cudaStream_t s) {
    if (!h->ws_event) HMV_CUDA(cudaEventCreateWithFlags(&h->ws_event, cudaEventDisableTiming));
    HMV_CUDA(cudaEventRecord(h->ws_event, s));
    h->ws_stream = s;
    h->ws_valid = true;
    return 0;
}

static int upload_f32(hmv_handle* h, float** dst, const std::vector<float>& v) {
    if (dev_alloc_t(h, dst, v.size() * sizeof(float), false)) return 1;
    HMV_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
static uint16_t f2bf(float f) {              // round-to-nearest-even, same as __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
static int upload_weights(hmv_handle* h, void** dst, const std::vector<float>& v) {
    if (h->bf16) {
        std::vector<uint16_t> b(v.size());
        for (size_t i = 0; i < v.size(); ++i) b[i] = f2bf(v[i]);
        if (dev_alloc(h, dst, b.size() * 2, false)) return 1;
        HMV_CUDA(cudaMemcpy(*dst, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    } else {
        if (dev_alloc(h, dst, v.size() * 4, false)) return 1;
        HMV_CUDA(cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

static const HostTensor* find_w(hmv_handle* h, const std::string& name) {
    auto it = h->weights.find(name);
    return it == h->weights.end() ? nullptr : &it->second;
}
#define NEED(var, name)                                                         \
    const HostTensor* var = find_w(h, name);                                    \
    HMV_CHECK(var != nullptr, std::string("missing state_dict key: ") + (name))

// conv weight [cout,cin,k,k] (+bias) with an optional BatchNorm folded in (eval mode, eps 1e-5):
//   w' = w * g / sqrt(var + eps) ; b' = beta + (b - mean) * g / sqrt(var + eps)
// Output: wf[cout][k*k][cin] (tap-major, channel-minor) and bf[cout].
static int fold_conv(hmv_handle* h, const std::string& conv, const std::string& bn, bool has_bias, int cout, int cin,
                     int k, std::vector<float>& wf, std::vector<float>& bf) {
    NEED(w, conv + ".weight");
    HMV_CHECK(static_cast<int64_t>(w->data.size()) == static_cast<int64_t>(cout) * cin * k * k,
              "unexpected weight shape for " + conv);
    std::vector<double> scale(cout, 1.0), shift(cout, 0.0);
    if (has_bias) {
        NEED(b, conv + ".bias");
        HMV_CHECK(static_cast<int>(b->data.size()) == cout, "unexpected bias shape for " + conv);
        for (int c = 0; c < cout; ++c) shift[c] = b->data[c];
    }
    if (!bn.empty()) {
        NEED(g, bn + ".weight");
        NEED(be, bn + ".bias");
        NEED(mu, bn + ".running_mean");
        NEED(var, bn + ".running_var");
        HMV_CHECK(static_cast<int>(g->data.size()) == cout && static_cast<int>(var->data.size()) == cout,
                  "unexpected BatchNorm shape for " + bn);
        for (int c = 0; c < cout; ++c) {
            const double s = static_cast<double>(g->data[c]) / sqrt(static_cast<double>(var->data[c]) + 1e-5);
            scale[c] = s;
            shift[c] = static_cast<double>(be->data[c]) + (shift[c] - static_cast<double>(mu->data[c])) * s;
        }
    }
    wf.assign(static_cast<size_t>(cout) * k * k * cin, 0.f);
    bf.assign(cout, 0.f);
    for (int co = 0; co < cout; ++co) {
        bf[co] = static_cast<float>(shift[co]);
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < k * k; ++t)
                wf[(static_cast<size_t>(co) * k * k + t) * cin + ci] =
                    static_cast<float>(static_cast<double>(w->data[(static_cast<size_t>(co) * cin + ci) * k * k + t]) * scale[co]);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layer construction
// ------------------------------------------------------------------------------------------------
static int finish_layer(hmv_handle* h, Layer& L, const std::vector<float>& wmat, const std::vector<float>& bias) {
    // wmat is [cout][K]; pad rows to n_alloc
    std::vector<float> wp(static_cast<size_t>(L.n_alloc) * L.K, 0.f), bp(L.n_alloc, 0.f);
    for (int n = 0; n < L.cout; ++n) {
        memcpy(&wp[static_cast<size_t>(n) * L.K], &wmat[static_cast<size_t>(n) * L.K], sizeof(float) * L.K);
        bp[n] = bias[n];
    }
    if (upload_weights(h, &L.w, wp)) return 1;
    if (upload_f32(h, &L.bias, bp)) return 1;
    L.bias_host = bp;
    L.ep.bias = L.bias;
    L.ep.N = L.cout;
    // (a row-major pitch below the padded tile width is only legal on the TMA-store path, which clips at the tensor's
    //  columns: checked in build_tc once the epilogue variant is known)
    return 0;
}

static int build_tc(hmv_handle* h, Layer& L) {
    TcLaunch& t = L.tc;
    memset(&t.p, 0, sizeof(t.p));
    memset(&t.bank, 0, sizeof(t.bank));
    t.p.bias_in_params = L.n_alloc <= kBiasBankFloats ? 1 : 0;      // (the 3072-wide QKV projection keeps the pointer path)
    if (t.p.bias_in_params) memcpy(t.bank.v, L.bias_host.data(), sizeof(float) * L.n_alloc);
    t.bn = L.bn;
    t.p.num_n_tiles = L.n_alloc / L.bn;
    t.p.err_flag = h->err_flag_dev;
    t.p.a_bytes = kTcBlockM * kTcBlockK * 2;
    t.p.tile_rows = kTcBlockM;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    const uint64_t N = static_cast<uint64_t>(L.max_units);
    if (L.kind == LK_FLAT) {
        const uint64_t M = N * L.rows_per_unit();
        dims[0] = L.K; dims[1] = M; dims[2] = 1; dims[3] = 1; dims[4] = 1;
        strides[0] = static_cast<uint64_t>(L.K) * 2; strides[1] = strides[2] = strides[3] = M * L.K * 2;
        box[0] = 64; box[1] = 128; box[2] = box[3] = box[4] = 1;
        t.p.flat = 1; t.p.tpi = 1; t.p.hbox = 1;
        t.p.num_taps = 1; t.p.cblks = L.K / 64;
        t.p.taps[0] = TcTap{0, 0, 0, 0};
    } else if (L.kind == LK_CONV3_S1) {
        const uint64_t C = L.cin, W = L.win, H = L.hin;
        const bool small = L.hin * L.win < 128;            // an image is smaller than a tile (8 x 8 maps): one image per tile, 64 rows used
        HMV_CHECK(128 % L.win == 0 && (small || L.hin % (128 / L.win) == 0) && L.cin % 64 == 0 && (!small || L.hin * L.win == 64),
                  "conv3x3: unsupported geometry");
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
        strides[0] = C * 2; strides[1] = W * C * 2; strides[2] = W * C * 2; strides[3] = H * W * C * 2;
        box[0] = 64; box[1] = L.win; box[2] = 1; box[3] = small ? L.hin : 128 / L.win; box[4] = 1;
        t.p.flat = 0; t.p.hbox = static_cast<int>(box[3]); t.p.tpi = L.hin / t.p.hbox;
        if (small) { t.p.tile_rows = L.hin * L.win; t.p.a_bytes = L.hin * L.win * kTcBlockK * 2; }
        t.p.num_taps = 9; t.p.cblks = L.cin / 64;
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) t.p.taps[r * 3 + s] = TcTap{0, s - 1, 0, r - 1};
    } else if (L.kind == LK_CONV_S2) {
        // input [N, H, W, C] addressed as (2C | W/2 | row parity | H/2 | N)
        const uint64_t C = L.cin, W = L.win, H = L.hin;
        const bool small = L.hout * L.wout < 128;
        HMV_CHECK(L.win % 2 == 0 && L.hin % 2 == 0 && 128 % L.wout == 0 && (small || L.hout % (128 / L.wout) == 0) && L.cin % 64 == 0 &&
                      (!small || L.hout * L.wout == 64),
                  "stride-2 conv: unsupported geometry");
        dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
        strides[0] = 2 * C * 2; strides[1] = W * C * 2; strides[2] = 2 * W * C * 2; strides[3] = H * W * C * 2;
        box[0] = 64; box[1] = L.wout; box[2] = 1; box[3] = small ? L.hout : 128 / L.wout; box[4] = 1;
        t.p.flat = 0; t.p.hbox = static_cast<int>(box[3]); t.p.tpi = L.hout / t.p.hbox;
        if (small) { t.p.tile_rows = L.hout * L.wout; t.p.a_bytes = L.hout * L.wout * kTcBlockK * 2; }
        t.p.cblks = L.cin / 64;
        if (L.ksize == 1) {
            t.p.num_taps = 1;
            t.p.taps[0] = TcTap{0, 0, 0, 0};
        } else {
            t.p.num_taps = 9;      // input row 2*oh + r - 1 : r=0 -> (odd row, oh-1), r=1 -> (even, oh), r=2 -> (odd, oh)
            for (int r = 0; r < 3; ++r)
                for (int s = 0; s < 3; ++s)
                    t.p.taps[r * 3 + s] = TcTap{(s == 1 ? 0 : 1) * L.cin, s == 0 ? -1 : 0, r == 1 ? 0 : 1, r == 0 ? -1 : 0};
        }
    } else {
        HMV_CHECK(false, "the 7x7 stem has its own fused kernel on the tensor-core path (stem_pool.cu)");
    }
    if (tc_make_tmap_act(&t.tmA, L.in, dims, strides, box)) {
        set_error(std::string(get_error()) + " [A map of " + L.name + "]");
        return 1;
    }
    // wide, K-deep layers run as 2-CTA clusters that share multicast weight tiles (conv_gemm_tc.cu, CL = 2)
    static const int cluster_min_k = [] { const char* e = getenv("HMV_CLUSTER_MINK"); return e ? atoi(e) : 512; }();
    t.cluster = (L.bn == 256 && L.K >= cluster_min_k && h->use_clusters && t.p.tile_rows == kTcBlockM) ? 2 : 1;
    // cta_group::2 CTA pairs (one M = 256 MMA per K step for two tiles, half of B per CTA, 6 operand stages) instead of the
    // multicast clusters, for the K-deep layers of large passes: measured at B = 64 l3.conv2 0.279 -> 0.267 ms, pose_net.0
    // 0.284 -> 0.250; short-K layers (K = 512 / 576: downsample, l3.0.conv1, QKV) get 5-10 % slower because the pair couples
    // two epilogues to one MMA issuer.  HMV_PAIR=0 disables, HMV_PAIR_MINK=<K> moves the threshold.
    static const bool pair_mma = [] { const char* e = getenv("HMV_PAIR"); return !(e && e[0] == '0'); }();
    static const int pair_min_k = [] { const char* e = getenv("HMV_PAIR_MINK"); return e ? atoi(e) : 512; }();
    if (t.cluster == 2 && pair_mma && L.K >= pair_min_k && static_cast<int64_t>(L.max_units) * L.rows_per_unit() >= 65536) t.cluster = 4;
    if (tc_make_tmap_wgt(&t.tmB, L.w, L.K, L.n_alloc, t.cluster >= 2 ? L.bn / 2 : L.bn)) {
        set_error(std::string(get_error()) + " [B map of " + L.name + "]");
        return 1;
    }
    // epilogue variant: bf16 NHWC outputs go through shared memory + TMA stores (see conv_gemm_tc.cuh)
    t.mode = TC_DIRECT;
    t.tmC = t.tmA; t.tmR = t.tmA;      // valid placeholders (never dereferenced in TC_DIRECT)
    const bool identity_res = L.ep.res_mode == RES_NONE || (L.ep.res_mode == RES_BF16 && L.ep.res_group >= (1 << 30));
    if (L.ep.out_mode == OUT_BF16_ROWMAJOR && L.bn % 64 == 0 && identity_res && L.ep.ldc % 8 == 0) {
        const uint64_t rows = N * L.rows_per_unit();
        t.mode = L.ep.res_mode == RES_BF16 ? TC_STORE_RES : (L.K >= 512 ? TC_STORE : TC_STORE_DEEP);
        if (tc_make_tmap_out(&t.tmC, L.ep.out, L.ep.ldc, rows, 32)) {
            set_error(std::string(get_error()) + " [C map of " + L.name + "]");
            return 1;
        }
        if (L.ep.res_mode == RES_BF16) {
            HMV_CHECK(L.ep.res_ld % 8 == 0, "residual pitch must be a multiple of 8 in " + L.name);
            if (tc_make_tmap_out(&t.tmR, L.ep.residual, L.ep.res_ld, rows, 128)) {
                set_error(std::string(get_error()) + " [R map of " + L.name + "]");
                return 1;
            }
        }
    }
    HMV_CHECK(t.mode != TC_DIRECT || L.ep.out_mode == OUT_F32_NCHW || L.ep.ldc >= L.n_alloc,
              "row-major output pitch must cover the padded tile width in " + L.name);
    return 0;
}

// Enqueue a layer for `units` images / rows.  out_override (optional) redirects the output.
static int run_layer(hmv_handle* h, Layer& L, int units, cudaStream_t s, void* out_override = nullptr) {
    const int M = units * L.rows_per_unit();
    if (M == 0) return 0;
    HMV_CHECK(units <= L.max_units, "run_layer: batch exceeds the workspace of " + L.name);
    ++h->launches;
    if (h->bf16) {
        TcLaunch t = L.tc;
        t.p.ep = L.ep;
        t.p.ep.M = M;
        if (out_override) t.p.ep.out = out_override;
        t.p.num_m_tiles = L.kind == LK_FLAT ? (M + 127) / 128 : units * t.p.tpi;
        if (!h->profiling) return tc_launch(t, h->num_sms, s);
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!h->ev_pool.empty()) { ev[i] = h->ev_pool.back(); h->ev_pool.pop_back(); }
            else HMV_CUDA(cudaEventCreate(&ev[i]));
        }
        HMV_CUDA(cudaEventRecord(ev[0], s));
        const int rc = tc_launch(t, h->num_sms, s);
        HMV_CUDA(cudaEventRecord(ev[1], s));
        h->prof.push_back({static_cast<int>(&L - h->layers.data()), units, ev[0], ev[1], -1});
        return rc;
    }
    ConvF32Params p{};
    p.in = static_cast<const float*>(L.in);
    p.w = static_cast<const float*>(L.w);
    p.M = M; p.K = L.K; p.Nalloc = L.n_alloc;
    if (L.kind == LK_FLAT) {
        p.Hin = p.Win = p.Hout = p.Wout = 1; p.Cin = L.K; p.kh = p.kw = 1; p.stride = 1; p.pad = 0;
    } else {
        p.Hin = L.hin; p.Win = L.win; p.Hout = L.hout; p.Wout = L.wout; p.Cin = L.kind == LK_STEM ? 4 : L.cin;
        p.kh = p.kw = L.ksize; p.stride = L.stride; p.pad = L.pad;
    }
    p.ep = L.ep;
    p.ep.M = M;
    if (out_override) p.ep.out = out_override;
    return conv_f32_launch(p, s);
}

// Enqueue a fused conv2+conv3 bottleneck tail for `units` images.
static int run_tail(hmv_handle* h, int tail, int units, cudaStream_t s) {
    FusedTail& T = h->tails[tail];
    if (units == 0) return 0;
    HMV_CHECK(units <= h->layers[T.l2].max_units, "run_tail: batch exceeds the workspace of " + T.name);
    ++h->launches;
    BtLaunch b = T.bt;
    b.p.num_m_tiles = units * b.p.tpi;
    static const bool bt_prof = [] { const char* e = getenv("HMV_BT_PROF"); return e && e[0] == '1'; }();
    if (bt_prof) {                                    // bring-up aid: per-role stall cycles of one launch, printed to stderr
        static long long* dbuf = nullptr;
        static int printed = 0;
        if (!dbuf) HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbuf), 148 * 24 * sizeof(long long)));
        HMV_CUDA(cudaMemsetAsync(dbuf, 0, 148 * 24 * sizeof(long long), s));
        b.p.prof = dbuf;
        const int rc = bt_launch(b, h->num_sms, s);
        if (rc == 0 && units >= 64 && printed < 40) {
            std::vector<long long> host(148 * 24);
            HMV_CUDA(cudaStreamSynchronize(s));
            HMV_CUDA(cudaMemcpy(host.data(), dbuf, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double a[24] = {0};
            const int grid = b.p.num_m_tiles < h->num_sms ? b.p.num_m_tiles : h->num_sms;
            for (int c = 0; c < grid; ++c) for (int k = 0; k < 24; ++k) a[k] += static_cast<double>(host[c * 24 + k]) / grid;
            fprintf(stderr, "[bt_prof] %s tiles/cta %.1f total %.0f | mma: t1empty %.0f full2 %.0f t2empty %.0f full3 %.0f | prod: empty %.0f y2ready %.0f | "
                    "epi: t2full %.0f cfull %.0f t1full %.0f bulk %.0f namedbar %.0f | res: cempty %.0f | mma warp: issue blocks %.0f, of which conv2: elect+descriptors %.0f, 4 x tcgen05.mma %.0f, commits %.0f (cycles, mean over CTAs)\n",
                    T.name.c_str(), a[13], a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[14], a[16], a[17], a[18]);
            ++printed;
        }
        return rc;
    }
    if (!h->profiling) return bt_launch(b, h->num_sms, s);
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
        if (!h->ev_pool.empty()) { ev[i] = h->ev_pool.back(); h->ev_pool.pop_back(); }
        else HMV_CUDA(cudaEventCreate(&ev[i]));
    }
    HMV_CUDA(cudaEventRecord(ev[0], s));
    const int rc = bt_launch(b, h->num_sms, s);
    HMV_CUDA(cudaEventRecord(ev[1], s));
    h->prof.push_back({T.l2, units, ev[0], ev[1], tail});
    return rc;
}

// Enqueue a fused conv3(b) + conv1(b+1) seam for `units` images.
static int run_seam(hmv_handle* h, int seam, int units, cudaStream_t s) {
    FusedSeam& S = h->seams[seam];
    if (units == 0) return 0;
    const Layer& L3 = h->layers[S.l3];
    HMV_CHECK(units <= L3.max_units, "run_seam: batch exceeds the workspace of " + S.name);
    ++h->launches;
    BnLaunch b = S.bn;
    b.p.num_m_tiles = units * L3.rows_per_unit() / kTcBlockM;
    static const bool bn_prof = [] { const char* e = getenv("HMV_BN_PROF"); return e && e[0] == '1'; }();
    if (bn_prof) {                                    // bring-up aid: per-role stall cycles of one launch, printed to stderr
        static long long* dbuf = nullptr;
        static int printed = 0;
        if (!dbuf) HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbuf), 148 * 24 * sizeof(long long)));
        HMV_CUDA(cudaMemsetAsync(dbuf, 0, 148 * 24 * sizeof(long long), s));
        b.p.prof = dbuf;
        const int rc = bn_launch(b, h->num_sms, s);
        if (rc == 0 && units >= 64 && printed < 10) {
            std::vector<long long> host(148 * 24);
            HMV_CUDA(cudaStreamSynchronize(s));
            HMV_CUDA(cudaMemcpy(host.data(), dbuf, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double a[24] = {0};
            const int grid = b.p.num_m_tiles < h->num_sms ? b.p.num_m_tiles : h->num_sms;
            for (int c = 0; c < grid; ++c) for (int k = 0; k < 24; ++k) a[k] += static_cast<double>(host[c * 24 + k]) / grid;
            fprintf(stderr, "[bn_prof] %s tiles/cta %.1f | mma total %.0f: t3empty %.0f full3 %.0f t1empty %.0f aready %.0f full1 %.0f y2full %.0f | prod: empty3 %.0f empty1 %.0f y2empty %.0f | "
                    "slots: sfree %.0f | epi total %.0f: t3full %.0f sres %.0f t1full %.0f | store warp: waits for slabs %.0f, for store reads %.0f | tmem_ld %.0f | per-slot phases (incl. the waits above): head %.0f ldwait %.0f math+sts %.0f fence %.0f issue %.0f (cycles, mean over CTAs)\n",
                    S.name.c_str(), a[15], a[0], a[1], a[2], a[3], a[4], a[5], a[21], a[6], a[7], a[22], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[16],
                    a[17], a[16], a[18], a[19], a[20]);
            ++printed;
        }
        return rc;
    }
    if (!h->profiling) return bn_launch(b, h->num_sms, s);
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
        if (!h->ev_pool.empty()) { ev[i] = h->ev_pool.back(); h->ev_pool.pop_back(); }
        else HMV_CUDA(cudaEventCreate(&ev[i]));
    }
    HMV_CUDA(cudaEventRecord(ev[0], s));
    const int rc = bn_launch(b, h->num_sms, s);
    HMV_CUDA(cudaEventRecord(ev[1], s));
    hmv_handle::ProfRec r{S.l3, units, ev[0], ev[1], -1};
    r.seam = seam;
    h->prof.push_back(r);
    return rc;
}

static Epilogue make_ep(void* out, int ldc, int out_mode, int act) {
    Epilogue e{};
    e.out = out; e.ldc = ldc; e.out_mode = out_mode; e.act = act;
    e.res_mode = RES_NONE; e.res_group = 1 << 30; e.res_stride = 0; e.res_ld = 0;
    e.hw = 1; e.N = 0; e.M = 0;
    return e;
}

// A conv layer from already-folded weights wf[cout][k*k][cin], bf[cout].  in: NHWC activations.
// cin_pad / cout_pad (0 = none): channel counts of the NHWC buffers when they are wider than the layer (HRNet-w40 widths
// 40 / 80 / 160 live in 64 / 128 / 192-channel buffers on the tensor-core path; the extra weights and biases are zero, so
// the extra output channels stay exactly zero).
static int add_conv_raw(hmv_handle* h, const std::string& name, const std::vector<float>& wf_in, const std::vector<float>& bf,
                        int cin, int cout, int k, int stride, int hin, int win, const void* in, void* out, int act,
                        const void* residual, int* index, int nchw_hw = 0, int cin_pad = 0, int cout_pad = 0, int bn_max = 0) {
    if (cin_pad < cin) cin_pad = cin;
    if (cout_pad < cout) cout_pad = cout;
    std::vector<float> wpad;
    if (cin_pad != cin) {                             // [cout][k*k][cin] -> [cout][k*k][cin_pad]
        wpad.assign(static_cast<size_t>(cout) * k * k * cin_pad, 0.f);
        for (int co = 0; co < cout; ++co)
            for (int t = 0; t < k * k; ++t)
                memcpy(&wpad[(static_cast<size_t>(co) * k * k + t) * cin_pad], &wf_in[(static_cast<size_t>(co) * k * k + t) * cin], sizeof(float) * cin);
    }
    const std::vector<float>& wf = cin_pad != cin ? wpad : wf_in;
    Layer L;
    L.name = name;
    L.cin = cin_pad; L.cin_real = cin; L.cout = cout; L.ksize = k; L.stride = stride; L.pad = k / 2;
    L.hin = hin; L.win = win; L.hout = hin / stride; L.wout = win / stride;
    L.max_units = h->mb_img;
    L.in = in;
    if (k == 1 && stride == 1) { L.kind = LK_FLAT; }
    else if (k == 3 && stride == 1) { L.kind = LK_CONV3_S1; }
    else if (stride == 2 && (k == 1 || k == 3)) { L.kind = LK_CONV_S2; }
    else { HMV_CHECK(false, "unsupported conv geometry for " + name); }
    L.K = k * k * cin_pad;
    HMV_CHECK(!h->bf16 || L.K % 64 == 0, "tensor-core path needs Cin to be a multiple of 64 in " + name);
    HMV_CHECK(cin_pad % 4 == 0, "Cin must be a multiple of 4 in " + name);
    L.bn = tc_pick_bn(nchw_hw > 0 ? cout : cout_pad);
    if (L.bn == 0) L.bn = tc_pick_bn((cout_pad + 191) / 192 * 192);
    HMV_CHECK(L.bn > 0, "no tile width for " + name);
    if (bn_max > 0 && L.bn > bn_max && L.bn % bn_max == 0) L.bn = bn_max;      // narrower N tiles: more CTAs for a small pass
    L.n_alloc = ((nchw_hw > 0 ? cout : cout_pad) + L.bn - 1) / L.bn * L.bn;
    if (!h->bf16) L.n_alloc = (cout_pad + 3) / 4 * 4;
    L.ep = make_ep(out, cout_pad, h->bf16 ? OUT_BF16_ROWMAJOR : OUT_F32_ROWMAJOR, act);
    if (nchw_hw > 0) { L.ep.out_mode = OUT_F32_NCHW; L.ep.hw = nchw_hw; }     // fp32 [n_img, cout, hw] output
    if (residual) {
        L.ep.residual = residual; L.ep.res_mode = h->bf16 ? RES_BF16 : RES_F32; L.ep.res_ld = cout_pad;
    }
    if (finish_layer(h, L, wf, bf)) return 1;
    if (h->bf16 && build_tc(h, L)) return 1;
    *index = static_cast<int>(h->layers.size());
    h->layers.push_back(L);
    return 0;
}

// A backbone / head conv with its BatchNorm folded from the state_dict.
static int add_conv(hmv_handle* h, const std::string& name, const std::string& conv_key, const std::string& bn_key,
                    bool has_bias, int cin, int cout, int k, int stride, int hin, int win, const void* in, void* out,
                    int act, const void* residual, int* index, int cin_pad = 0, int cout_pad = 0, int bn_max = 0) {
    std::vector<float> wf, bf;
    if (fold_conv(h, conv_key, bn_key, has_bias, cout, cin, k, wf, bf)) return 1;
    return add_conv_raw(h, name, wf, bf, cin, cout, k, stride, hin, win, in, out, act, residual, index, 0, cin_pad, cout_pad, bn_max);
}

// A linear layer y = x W^T + b on [rows, K] with K zero-padded to k_pad.
static int add_linear(hmv_handle* h, const std::string& name, const std::vector<const HostTensor*>& ws,
                      const HostTensor* bias, int in_features, int k_pad, const void* in, int max_rows, Epilogue ep,
                      int* index) {
    Layer L;
    L.name = name;
    L.kind = LK_FLAT;
    L.cin = in_features; L.K = k_pad;
    L.max_units = max_rows;
    L.in = in;
    int cout = 0;
    for (auto* w : ws) {
        HMV_CHECK(w->dims.size() == 2 && w->dims[1] == in_features, "unexpected linear weight shape in " + name);
        cout += static_cast<int>(w->dims[0]);
    }
    L.cout = cout;
    L.bn = tc_pick_bn(cout > 256 && cout % 128 != 0 ? (cout + 175) / 176 * 176 : cout);
    HMV_CHECK(L.bn > 0, "no tile width for " + name);
    {   // small passes (the QKV projection of a B = 1 pass is ONE M tile x 12 N tiles): N tiles of 128 while twice the tiles fit in a wave
        static const bool narrow_env = [] { const char* e = getenv("HMV_NARROW_SMALL"); return !(e && e[0] == '0'); }();
        const int m_tiles = (max_rows + kTcBlockM - 1) / kTcBlockM;
        if (narrow_env && h->bf16 && L.bn == 256 && m_tiles * ((cout + 255) / 256) * 2 <= h->num_sms) L.bn = 128;
    }
    L.n_alloc = (cout + L.bn - 1) / L.bn * L.bn;
    if (!h->bf16) L.n_alloc = (cout + 3) / 4 * 4;
    std::vector<float> wm(static_cast<size_t>(cout) * k_pad, 0.f), bv(cout, 0.f);
    int row = 0;
    for (auto* w : ws)
        for (int r = 0; r < w->dims[0]; ++r, ++row)
            memcpy(&wm[static_cast<size_t>(row) * k_pad], &w->data[static_cast<size_t>(r) * in_features], sizeof(float) * in_features);
    if (bias) {
        HMV_CHECK(static_cast<int>(bias->data.size()) == cout, "unexpected bias shape in " + name);
        for (int i = 0; i < cout; ++i) bv[i] = bias->data[i];
    }
    L.ep = ep;
    if (finish_layer(h, L, wm, bv)) return 1;
    if (h->bf16 && build_tc(h, L)) return 1;
    *index = static_cast<int>(h->layers.size());
    h->layers.push_back(L);
    return 0;
}

static int add_gemm_step(hmv_handle* h, const std::string& name, int layer, void* out, int C, int H, int W,
                         std::vector<StepIO> ins = {}) {
    Step st;
    st.kind = SK_GEMM; st.name = name; st.layer = layer; st.out = out; st.C = C; st.H = H; st.W = W;
    st.ins = std::move(ins);
    st.outs.push_back({out, name, C, H, W});
    h->backbone.push_back(st);
    return 0;
}

// Fused launch plan for the conv2 (3x3) / conv3 (1x1 + residual) pair of one bottleneck; reuses the tensor maps the
// two layers already own and adds the reload map of the conv2 output and the chunked conv3 weight map.
static int add_tail(hmv_handle* h, const std::string& name, int l2, int l3, int l_ds = -1) {
    const Layer& A = h->layers[l2];
    const Layer& B = h->layers[l3];
    const int P = A.cout;
    HMV_CHECK((P == 64 || P == 128 || P == 256) && A.bn == P && A.cin == P && B.cin == P && B.cout == 4 * P && B.n_alloc == 4 * P,
              "fused bottleneck tail: unexpected widths in " + name);
    HMV_CHECK(A.tc.mode != TC_DIRECT && B.tc.mode == TC_STORE_RES && A.kind != LK_FLAT, "fused bottleneck tail: unexpected layer plan in " + name);
    FusedTail T;
    T.name = name; T.l2 = l2; T.l3 = l3;
    BtLaunch& b = T.bt;
    memset(&b.p, 0, sizeof(b.p));
    b.planes = P;
    const uint64_t rows = static_cast<uint64_t>(A.max_units) * A.rows_per_unit();
    // CTA-pair variant (cta_group::2, each CTA holds half of every weight tile): large passes of the P = 64 / 128 tails
    static const bool tail_pair_env = [] { const char* e = getenv("HMV_TAIL_PAIR"); return !(e && e[0] == '0'); }();
    const bool pair = tail_pair_env && P <= 128 && rows >= 65536 && (rows / 128) % 2 == 0;
    b.pair = pair ? 1 : 0;
    b.tmA = A.tc.tmA; b.tmY2s = A.tc.tmC;
    if (tc_make_tmap_wgt(&b.tmW2, A.w, A.K, A.n_alloc, pair ? P / 2 : P)) return 1;        // (the layer's own map may be a half box)
    b.tmOut = B.tc.tmC; b.tmRes = B.tc.tmR;
    if (tc_make_tmap_out(&b.tmY2l, A.ep.out, P, rows, 128) || tc_make_tmap_out(&b.tmW3, B.w, P, 4 * P, pair ? 64 : 128)) {
        set_error(std::string(get_error()) + " [fused-tail maps of " + name + "]");
        return 1;
    }
    b.p.tpi = A.tc.p.tpi; b.p.hbox = A.tc.p.hbox; b.p.cblks = A.tc.p.cblks;
    HMV_CHECK(A.tc.p.num_taps == 9 && b.p.cblks * 64 == P, "fused bottleneck tail: conv2 must be a 3x3 with P input channels");
    for (int t = 0; t < 9; ++t) b.p.taps[t] = A.tc.p.taps[t];
    memset(&b.bank, 0, sizeof(b.bank));
    memcpy(b.bank.v, A.bias_host.data(), sizeof(float) * P);
    memcpy(b.bank.v + kBtBias3Off, B.bias_host.data(), sizeof(float) * 4 * P);
    b.tmWd = b.tmW3;                                   // valid placeholder
    T.l_ds = l_ds;
    if (l_ds >= 0) {
        // The block's 1x1 stride-1 downsample (resnet.py:145-146, `identity = self.downsample(x)`) folded into conv3: its K
        // blocks (block input x folded weights) accumulate into the same TMEM accumulator, the biases add up, and neither the
        // downsample launch nor the write + re-read of its 4P-channel output exist any more.
        const Layer& D = h->layers[l_ds];
        HMV_CHECK(D.kind == LK_FLAT && D.cout == 4 * P && D.K % 64 == 0 && D.K == D.cin && D.max_units * D.rows_per_unit() == static_cast<int64_t>(rows),
                  "fused bottleneck tail: the folded downsample must be a 1x1 stride-1 conv over the tile's rows in " + name);
        b.p.ds_kb = D.K / 64;
        if (tc_make_tmap_out(&b.tmRes, D.in, D.K, rows, 128) || tc_make_tmap_out(&b.tmWd, D.w, D.K, 4 * P, pair ? 64 : 128)) {
            set_error(std::string(get_error()) + " [fused-tail downsample maps of " + name + "]");
            return 1;
        }
        for (int j = 0; j < 4 * P; ++j) b.bank.v[kBtBias3Off + j] += D.bias_host[j];
    }
    b.p.err_flag = h->err_flag_dev;
    {
        const char* e = getenv("HMV_BN_PREFETCH");         // measured neutral (P = 64) to harmful (P = 128): off
        b.p.prefetch = e && e[0] == '1';
    }
    h->tails.push_back(T);
    return 0;
}

// Fused launch plan for conv3 of one layer3 block + conv1 of the next one (both 1x1, P = 256).
static int add_seam(hmv_handle* h, const std::string& name, int l3, int l1) {
    const Layer& A = h->layers[l3];
    const Layer& B = h->layers[l1];
    HMV_CHECK(A.kind == LK_FLAT && B.kind == LK_FLAT && A.cin == 256 && A.cout == 1024 && B.cin == 1024 && B.cout == 256 &&
                  A.rows_per_unit() % kTcBlockM == 0 && A.tc.mode == TC_STORE_RES && B.tc.mode != TC_DIRECT && B.in == A.ep.out,
              "fused bottleneck seam: unexpected layer plan in " + name);
    FusedSeam S;
    S.name = name; S.l3 = l3; S.l1 = l1;
    BnLaunch& b = S.bn;
    memset(&b.p, 0, sizeof(b.p));
    const uint64_t rows = static_cast<uint64_t>(A.max_units) * A.rows_per_unit();
    // CTA-pair variant of the seam kernel (cta_group::2): 0.408 -> 0.396 ms per B = 64 launch; small passes keep one CTA per tile
    static const bool seam_pair_env = [] { const char* e = getenv("HMV_SEAM_PAIR"); return !(e && e[0] == '0'); }();
    const bool seam_pair = seam_pair_env && rows >= 65536;
    b.pair = seam_pair ? 1 : 0;
    if (tc_make_tmap_out(&b.tmY2, A.in, 256, rows, 128) || tc_make_tmap_out(&b.tmW3, A.w, 256, 1024, seam_pair ? 64 : 128) ||
        tc_make_tmap_wgt(&b.tmW1, B.w, 1024, 256, seam_pair ? 128 : 256)) {
        set_error(std::string(get_error()) + " [fused-seam maps of " + name + "]");
        return 1;
    }
    b.tmRes = A.tc.tmR;
    // whole-slot stores (128 rows x 64 columns) issued by the kernel's store warp
    if (tc_make_tmap_out(&b.tmOut, A.ep.out, A.ep.ldc, rows, 128) || tc_make_tmap_out(&b.tmY1, B.ep.out, B.ep.ldc, rows, 128)) {
        set_error(std::string(get_error()) + " [fused-seam store maps of " + name + "]");
        return 1;
    }
    memset(&b.bank, 0, sizeof(b.bank));
    memcpy(b.bank.v, A.bias_host.data(), sizeof(float) * 1024);
    memcpy(b.bank.v + kBnBias1Off, B.bias_host.data(), sizeof(float) * 256);
    b.p.err_flag = h->err_flag_dev;
    h->seams.push_back(S);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// plan: buffers, layers, descriptors (hmv_prepare)
// ------------------------------------------------------------------------------------------------
static int build_backbone(hmv_handle* h) {
    const size_t e = h->esz;
    const size_t act_bytes = static_cast<size_t>(h->mb_img) * 64 * 64 * 256 * e;      // largest activation
    if (dev_alloc(h, &h->bufX, act_bytes) || dev_alloc(h, &h->bufY, act_bytes) || dev_alloc(h, &h->bufT1, act_bytes) ||
        dev_alloc(h, &h->bufT2, act_bytes) || dev_alloc(h, &h->bufDS, act_bytes))
        return 1;
    if (h->bf16) {
        // ---- fused stem: conv 7x7/2 + BN + ReLU + maxpool 3x3/2 in one kernel (stem_pool.cu) ----
        std::vector<float> wf, bf;
        if (fold_conv(h, "backbone.conv1", "backbone.bn1", false, 64, 3, 7, wf, bf)) return 1;
        std::vector<float> wp(static_cast<size_t>(7) * 4 * 64 * 8, 0.f);
        for (int r = 0; r < 7; ++r)
            for (int kc = 0; kc < 4; ++kc)
                for (int co = 0; co < 64; ++co)
                    for (int e2 = 0; e2 < 8; ++e2) {
                        const int sx = 2 * kc + (e2 >> 2), c = e2 & 3;
                        if (sx < 7 && c < 3)
                            wp[((static_cast<size_t>(r) * 4 + kc) * 64 + co) * 8 + e2] = wf[(static_cast<size_t>(co) * 49 + r * 7 + sx) * 3 + c];
                    }
        if (upload_weights(h, &h->stem_w, wp) || upload_f32(h, &h->stem_b, bf)) return 1;
        Step st; st.kind = SK_STEM_POOL; st.name = "maxpool"; st.out = h->bufX; st.C = 64; st.H = h->img / 4; st.W = h->img / 4;
        h->backbone.push_back(st);
    } else {
    const int Hp = h->img + 6, Wp = h->img + 16;                                       // 262 x 272
    if (dev_alloc(h, &h->xpad, static_cast<size_t>(h->mb_img) * Hp * Wp * 4 * e)) return 1;

    Step pk; pk.kind = SK_PACK; pk.name = "pack_input"; pk.out = h->xpad; pk.C = 4; pk.H = Hp; pk.W = Wp;
    h->backbone.push_back(pk);

    // ---- stem: conv 7x7/2 + BN + ReLU (resnet.py:218-220), fp32 check mode ----
    {
        Layer L;
        L.name = "conv1"; L.kind = LK_STEM;
        L.cin = 3; L.cout = 64; L.ksize = 7; L.stride = 2; L.pad = 0;    // padding is materialised by pack_input
        L.hin = Hp; L.win = Wp; L.hout = h->img / 2; L.wout = h->img / 2;
        L.max_units = h->mb_img; L.in = h->xpad;
        std::vector<float> wf, bf;
        if (fold_conv(h, "backbone.conv1", "backbone.bn1", false, 64, 3, 7, wf, bf)) return 1;
        std::vector<float> wm;
        L.K = 49 * 4;            // K = 7 x 7 x 4 (channel 3 is zero)
        wm.assign(static_cast<size_t>(64) * L.K, 0.f);
        for (int co = 0; co < 64; ++co)
            for (int t = 0; t < 49; ++t)
                for (int c = 0; c < 3; ++c) wm[static_cast<size_t>(co) * L.K + t * 4 + c] = wf[(static_cast<size_t>(co) * 49 + t) * 3 + c];
        L.bn = 64; L.n_alloc = 64;
        L.ep = make_ep(h->bufT1, 64, OUT_F32_ROWMAJOR, ACT_RELU);
        if (finish_layer(h, L, wm, bf)) return 1;
        h->layers.push_back(L);
        add_gemm_step(h, "conv1", static_cast<int>(h->layers.size()) - 1, h->bufT1, 64, L.hout, L.wout);
    }
    Step mp; mp.kind = SK_MAXPOOL; mp.name = "maxpool"; mp.in = h->bufT1; mp.out = h->bufX; mp.C = 64;
    mp.H = h->img / 4; mp.W = h->img / 4;
    mp.ins.push_back({h->bufT1, "conv1", 64, h->img / 2, h->img / 2});
    mp.outs.push_back({h->bufX, "maxpool", 64, mp.H, mp.W});
    h->backbone.push_back(mp);
    }

    // ---- layer1..3 (resnet.py:124-144, 189-203; paper variant: layer3 stride 1) ----
    void* cur = h->bufX;
    void* nxt = h->bufY;
    int C = 64, H = h->img / 4, W = h->img / 4;
    std::string cur_name = "maxpool";                 // oracle tap of the tensor in `cur`
    const int blocks[3] = {3, 4, 6}, planes[3] = {64, 128, 256}, strides[3] = {1, 2, 1};
    for (int li = 0; li < 3; ++li) {
        for (int b = 0; b < blocks[li]; ++b) {
            const std::string p = "backbone.layer" + std::to_string(li + 1) + "." + std::to_string(b);
            const std::string sp = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
            const int pl = planes[li], st = b == 0 ? strides[li] : 1;
            const bool ds = b == 0 && (st != 1 || C != pl * 4);
            int idx;
            const bool fuse = h->bf16 && ((h->fuse_mask >> li) & 1);
            // Small passes (B = 1: 40 M tiles of layer3 on 148 SMs): the stand-alone N = 256 layers run with N tiles of 128 when
            // twice the tiles still fit in one wave -- half the MMA time per CTA (N / 2 cycles per MMA) on twice the SMs.
            // (Not the conv1 of blocks >= 1 -- it lives in the seam kernel -- nor a conv2 that a fused tail consumes.)
            static const bool narrow_env = [] { const char* e = getenv("HMV_NARROW_SMALL"); return !(e && e[0] == '0'); }();
            const bool narrow = narrow_env && h->bf16 && pl == 256 &&
                                static_cast<int64_t>(h->mb_img) * (H / st) * (W / st) / kTcBlockM * 2 <= h->num_sms;
            if (add_conv(h, sp + ".conv1", p + ".conv1", p + ".bn1", false, C, pl, 1, 1, H, W, cur, h->bufT1, ACT_RELU, nullptr, &idx, 0, 0,
                         narrow && b == 0 ? 128 : 0)) return 1;
            if (h->bf16 && h->fuse_next && li == 2 && b >= 1 && !h->backbone.empty() && h->backbone.back().kind == SK_GEMM) {
                // the previous block's conv3 step also produces this conv1 (bottleneck_next_tc.cu): no step of its own
                Step& prev = h->backbone.back();
                if (add_seam(h, prev.name, prev.layer, idx)) return 1;
                prev.kind = SK_SEAM;
                prev.layer = static_cast<int>(h->seams.size()) - 1;
                prev.outs.push_back({h->bufT1, sp + ".conv1", pl, H, W});
            } else {
                add_gemm_step(h, sp + ".conv1", idx, h->bufT1, pl, H, W, {{cur, cur_name, C, H, W}});
            }
            int idx2;
            if (add_conv(h, sp + ".conv2", p + ".conv2", p + ".bn2", false, pl, pl, 3, st, H, W, h->bufT1, h->bufT2, ACT_RELU, nullptr, &idx2, 0, 0,
                         narrow && !fuse ? 128 : 0)) return 1;
            if (!fuse) add_gemm_step(h, sp + ".conv2", idx2, h->bufT2, pl, H / st, W / st, {{h->bufT1, sp + ".conv1", pl, H, W}});
            const void* res = cur;
            StepIO res_io{cur, cur_name, C, H, W};
            int idx_ds = -1;
            if (ds) {
                if (add_conv(h, sp + ".downsample", p + ".downsample.0", p + ".downsample.1", false, C, pl * 4, 1, st, H, W, cur, h->bufDS, ACT_NONE, nullptr, &idx)) return 1;
                static const bool fold_ds = [] { const char* e = getenv("HMV_FOLD_DS"); return !(e && e[0] == '0'); }();
                if (fuse && fold_ds && st == 1 && C % 64 == 0) {
                    idx_ds = idx;                            // folded into the fused tail below: no launch, no step of its own
                } else {
                    add_gemm_step(h, sp + ".downsample", idx, h->bufDS, pl * 4, H / st, W / st, {{cur, cur_name, C, H, W}});
                    res_io = StepIO{h->bufDS, sp + ".downsample", pl * 4, H / st, W / st};
                }
                res = h->bufDS;
            }
            if (add_conv(h, sp + ".conv3", p + ".conv3", p + ".bn3", false, pl, pl * 4, 1, 1, H / st, W / st, h->bufT2, nxt, ACT_RELU, res, &idx)) return 1;
            if (fuse) {
                if (add_tail(h, sp + ".conv3", idx2, idx, idx_ds)) return 1;
                Step ts; ts.kind = SK_TAIL; ts.name = sp + ".conv3"; ts.layer = static_cast<int>(h->tails.size()) - 1;
                ts.out = nxt; ts.C = pl * 4; ts.H = H / st; ts.W = W / st;
                ts.ins.push_back({h->bufT1, sp + ".conv1", pl, H, W});
                ts.ins.push_back(res_io);
                ts.outs.push_back({nxt, sp + ".conv3", pl * 4, H / st, W / st});
                h->backbone.push_back(ts);
            } else {
                add_gemm_step(h, sp + ".conv3", idx, nxt, pl * 4, H / st, W / st, {{h->bufT2, sp + ".conv2", pl, H / st, W / st}, res_io});
            }
            void* t = cur; cur = nxt; nxt = t;
            C = pl * 4; H /= st; W /= st;
            cur_name = sp + ".conv3";
        }
    }
    h->featbuf = cur;
    HMV_CHECK(C == 1024 && H == h->hm && W == h->hm, "backbone output geometry mismatch");
    return 0;
}


// ------------------------------------------------------------------------------------------------
// HRNet-w40 / w64 backbone plan (reference backbones/hrnet.py:241-409; the `*_HR*` release configs)
// ------------------------------------------------------------------------------------------------
// Every convolution runs through the same implicit-GEMM kernels as the ResNet path (3x3 stride 1 / 2 and 1x1 classes,
// BN folded, ReLU / residual in the epilogue); the branch-fusion sums (hrnet.py:222-231) are built as
//   out_i = relu( x_i + sum_{j<i} chain_ij(x_j)  [each chain's last conv adds the running sum as its residual]
//                     + sum_{j>i} nearest_up(conv1x1_ij(x_j)) [one elementwise kernel] )
// - the reference's summation order up to fp32 re-association.  On the tensor-core path the w40 widths 40 / 80 / 160
// live in 64 / 128 / 192-channel NHWC buffers whose extra channels are exactly zero (zero weights and biases).
static int hr_alloc(hmv_handle* h, void** p, int level, int cp) {
    const size_t res = static_cast<size_t>(64 >> level);
    return dev_alloc(h, p, static_cast<size_t>(h->mb_img) * res * res * cp * h->esz);
}

static int add_fuse_step(hmv_handle* h, const std::string& name, const void* base, std::vector<std::pair<const void*, int>> ups, void* out,
                         int level, int cp, int c_real, bool relu, std::vector<StepIO> ins) {
    Step st;
    st.kind = SK_FUSE_SUM; st.name = name; st.out = out; st.C = c_real; st.H = 64 >> level; st.W = 64 >> level;
    st.fuse.base = base; st.fuse.out = out; st.fuse.n_up = static_cast<int>(ups.size());
    for (size_t k = 0; k < ups.size(); ++k) { st.fuse.up[k] = ups[k].first; st.fuse.shift[k] = ups[k].second; }
    st.fuse.H = 64 >> level; st.fuse.W = 64 >> level; st.fuse.C = cp; st.fuse.relu = relu ? 1 : 0;
    st.ins = std::move(ins);
    st.outs.push_back({out, name, c_real, 64 >> level, 64 >> level});
    h->backbone.push_back(st);
    return 0;
}

static int build_hrnet(hmv_handle* h) {
    const int R0 = h->img / 4;                        // 64
    HMV_CHECK(R0 == 64, "the HRNet plan is built for 256 x 256 inputs");
    auto cpad = [&](int c) { return !h->bf16 ? c : (c <= 64 ? 64 : c <= 128 ? 128 : c <= 192 ? 192 : (c + 63) / 64 * 64); };
    // ---- stem (hrnet.py:244-247, 380-385): conv3x3/2 3->64 (CUDA cores), conv3x3/2 64->64 ----
    void *s1 = nullptr, *s2 = nullptr;
    if (dev_alloc(h, &s1, static_cast<size_t>(h->mb_img) * 128 * 128 * 64 * h->esz) || hr_alloc(h, &s2, 0, 64)) return 1;
    {
        std::vector<float> wf, bf;
        if (fold_conv(h, "backbone.conv1", "backbone.bn1", false, 64, 3, 3, wf, bf)) return 1;
        if (upload_f32(h, &h->hr_stem_w, wf) || upload_f32(h, &h->stem_b, bf)) return 1;
        Step st; st.kind = SK_HR_STEM; st.name = "hr.conv1"; st.out = s1; st.C = 64; st.H = 128; st.W = 128;
        st.outs.push_back({s1, "hr.conv1", 64, 128, 128});
        h->backbone.push_back(st);
    }
    int idx;
    if (add_conv(h, "hr.conv2", "backbone.conv2", "backbone.bn2", false, 64, 64, 3, 2, 128, 128, s1, s2, ACT_RELU, nullptr, &idx)) return 1;
    add_gemm_step(h, "hr.conv2", idx, s2, 64, 64, 64, {{s1, "hr.conv1", 64, 128, 128}});
    // ---- stage 1: 4 Bottlenecks of 64 planes at 64 x 64 (hrnet.py:249-253) ----
    void *bx = nullptr, *by = nullptr, *bt1 = nullptr, *bt2 = nullptr, *bds = nullptr;
    if (hr_alloc(h, &bx, 0, 256) || hr_alloc(h, &by, 0, 256) || hr_alloc(h, &bt1, 0, 64) || hr_alloc(h, &bt2, 0, 64) || hr_alloc(h, &bds, 0, 256)) return 1;
    void* cur = s2; std::string cur_name = "hr.conv2"; int C = 64;
    for (int b = 0; b < 4; ++b) {
        const std::string p = "backbone.layer1." + std::to_string(b), sp = "hr.layer1." + std::to_string(b);
        void* nxt = (cur == bx) ? by : bx;
        if (add_conv(h, sp + ".conv1", p + ".conv1", p + ".bn1", false, C, 64, 1, 1, 64, 64, cur, bt1, ACT_RELU, nullptr, &idx)) return 1;
        add_gemm_step(h, sp + ".conv1", idx, bt1, 64, 64, 64, {{cur, cur_name, C, 64, 64}});
        if (add_conv(h, sp + ".conv2", p + ".conv2", p + ".bn2", false, 64, 64, 3, 1, 64, 64, bt1, bt2, ACT_RELU, nullptr, &idx)) return 1;
        add_gemm_step(h, sp + ".conv2", idx, bt2, 64, 64, 64, {{bt1, sp + ".conv1", 64, 64, 64}});
        const void* res = cur; StepIO res_io{cur, cur_name, C, 64, 64};
        if (b == 0) {
            if (add_conv(h, sp + ".downsample", p + ".downsample.0", p + ".downsample.1", false, C, 256, 1, 1, 64, 64, cur, bds, ACT_NONE, nullptr, &idx)) return 1;
            add_gemm_step(h, sp + ".downsample", idx, bds, 256, 64, 64, {{cur, cur_name, C, 64, 64}});
            res = bds; res_io = StepIO{bds, sp + ".downsample", 256, 64, 64};
        }
        if (add_conv(h, sp + ".conv3", p + ".conv3", p + ".bn3", false, 64, 256, 1, 1, 64, 64, bt2, nxt, ACT_RELU, res, &idx)) return 1;
        add_gemm_step(h, sp + ".conv3", idx, nxt, 256, 64, 64, {{bt2, sp + ".conv2", 64, 64, 64}, res_io});
        cur = nxt; cur_name = sp + ".conv3"; C = 256;
    }
    // ---- stages 2..4 ----
    const int stages[3][2] = {{1, 2}, {4, 3}, {3, 4}};   // (modules, branches): hrnet.py:453-485
    struct Br { void *x = nullptr, *y = nullptr, *t = nullptr, *z = nullptr, *r = nullptr, *d = nullptr; int c = 0, cp = 0; std::string name; };
    Br br[4];
    for (int l = 0; l < 4; ++l) {
        br[l].c = h->hr_c[l]; br[l].cp = cpad(br[l].c);
        HMV_CHECK(br[l].cp == h->hr_cp[l], "HRNet channel padding mismatch");
        // x / y: block ping-pong, t: BasicBlock temp, z: fused output, r: running sum of the down chains, d: chain intermediate (<= widest source)
        if (hr_alloc(h, &br[l].x, l, br[l].cp) || hr_alloc(h, &br[l].y, l, br[l].cp) || hr_alloc(h, &br[l].t, l, br[l].cp) ||
            hr_alloc(h, &br[l].z, l, br[l].cp) || hr_alloc(h, &br[l].r, l, br[l].cp) || hr_alloc(h, &br[l].d, l, 512))
            return 1;
    }
    void* up[4][4] = {};                                  // up[i][j]: conv1x1_ij(x_j) at resolution j with cp_i channels
    for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) if (hr_alloc(h, &up[i][j], j, br[i].cp)) return 1;
    int nprev = 1;                                        // branches alive before the stage
    void* prev_last = cur; std::string prev_last_name = cur_name; int prev_last_c = 256, prev_last_cp = 256, prev_last_level = 0;
    for (int si = 0; si < 3; ++si) {
        const int nmod = stages[si][0], nbr = stages[si][1];
        const std::string t = "backbone.transition" + std::to_string(si + 1);
        // transitions (hrnet.py:318-345, 387-408)
        for (int i = 0; i < nbr; ++i) {
            const std::string tn = "hr.stage" + std::to_string(si + 2) + ".in" + std::to_string(i);
            const int res_i = 64 >> i;
            if (i < nprev) {
                if (find_w(h, t + "." + std::to_string(i) + ".0.weight")) {     // only transition1.0 in w40 / w64 (256 -> c0)
                    HMV_CHECK(si == 0 && i == 0, "unexpected transition layer on an existing branch");
                    if (add_conv(h, tn, t + ".0.0", t + ".0.1", false, prev_last_c, br[0].c, 3, 1, 64, 64, prev_last, br[0].x, ACT_RELU, nullptr, &idx,
                                 prev_last_cp, br[0].cp)) return 1;
                    add_gemm_step(h, tn, idx, br[0].x, br[0].c, 64, 64, {{prev_last, prev_last_name, prev_last_c, 64, 64}});
                    br[0].name = tn;
                }                                          // else: the branch carries over unchanged (its x buffer already holds it)
            } else {                                       // a new branch: one stride-2 3x3 conv from the previous stage's last branch
                HMV_CHECK(i == nprev, "more than one new branch per stage");
                const std::string k = t + "." + std::to_string(i) + ".0";
                const int src_res = 64 >> prev_last_level;
                HMV_CHECK(src_res == 2 * res_i, "transition must halve the resolution");
                if (add_conv(h, tn, k + ".0", k + ".1", false, prev_last_c, br[i].c, 3, 2, src_res, src_res, prev_last, br[i].x, ACT_RELU, nullptr, &idx,
                             prev_last_cp, br[i].cp)) return 1;
                add_gemm_step(h, tn, idx, br[i].x, br[i].c, res_i, res_i, {{prev_last, prev_last_name, prev_last_c, src_res, src_res}});
                br[i].name = tn;
            }
        }
        for (int m = 0; m < nmod; ++m) {
            const std::string p = "backbone.stage" + std::to_string(si + 2) + "." + std::to_string(m);
            const std::string sp = "hr.stage" + std::to_string(si + 2) + "." + std::to_string(m);
            // branches: 4 BasicBlocks each (hrnet.py:38-55)
            for (int i = 0; i < nbr; ++i) {
                const int res = 64 >> i;
                for (int blk = 0; blk < 4; ++blk) {
                    const std::string q = p + ".branches." + std::to_string(i) + "." + std::to_string(blk);
                    const std::string qn = sp + ".b" + std::to_string(i) + "." + std::to_string(blk);
                    if (add_conv(h, qn + ".conv1", q + ".conv1", q + ".bn1", false, br[i].c, br[i].c, 3, 1, res, res, br[i].x, br[i].t, ACT_RELU, nullptr, &idx,
                                 br[i].cp, br[i].cp)) return 1;
                    add_gemm_step(h, qn + ".conv1", idx, br[i].t, br[i].c, res, res, {{br[i].x, br[i].name, br[i].c, res, res}});
                    const std::string out_name = blk == 3 ? sp + ".branch" + std::to_string(i) : qn + ".conv2";
                    if (add_conv(h, out_name, q + ".conv2", q + ".bn2", false, br[i].c, br[i].c, 3, 1, res, res, br[i].t, br[i].y, ACT_RELU, br[i].x, &idx,
                                 br[i].cp, br[i].cp)) return 1;
                    add_gemm_step(h, out_name, idx, br[i].y, br[i].c, res, res, {{br[i].t, qn + ".conv1", br[i].c, res, res}, {br[i].x, br[i].name, br[i].c, res, res}});
                    std::swap(br[i].x, br[i].y);
                    br[i].name = out_name;
                }
            }
            // fusion (hrnet.py:177-211, 222-231)
            for (int i = 0; i < nbr; ++i) {
                const int res_i = 64 >> i;
                const std::string on = sp + ".out" + std::to_string(i);
                const bool has_up = i + 1 < nbr;
                // up terms first (their temps are independent): conv1x1 + BN at the source resolution
                std::vector<std::pair<const void*, int>> ups;
                std::vector<StepIO> fuse_ins;
                for (int j = i + 1; j < nbr; ++j) {
                    const std::string f = p + ".fuse_layers." + std::to_string(i) + "." + std::to_string(j);
                    const std::string fn = sp + ".up" + std::to_string(i) + std::to_string(j);
                    const int res_j = 64 >> j;
                    if (add_conv(h, fn, f + ".0", f + ".1", false, br[j].c, br[i].c, 1, 1, res_j, res_j, br[j].x, up[i][j], ACT_NONE, nullptr, &idx,
                                 br[j].cp, br[i].cp)) return 1;
                    add_gemm_step(h, fn, idx, up[i][j], br[i].c, res_j, res_j, {{br[j].x, br[j].name, br[j].c, res_j, res_j}});
                    ups.push_back({up[i][j], j - i});
                    fuse_ins.push_back({up[i][j], fn, br[i].c, res_j, res_j});
                }
                // down chains: the last conv of every chain adds the running sum (starting from x_i) as its residual
                const void* run = br[i].x; std::string run_name = br[i].name;
                for (int j = 0; j < i; ++j) {
                    const std::string f = p + ".fuse_layers." + std::to_string(i) + "." + std::to_string(j);
                    const void* src = br[j].x; std::string src_name = br[j].name;
                    for (int k = 0; k < i - j; ++k) {
                        const bool last = k == i - j - 1;
                        const int rin = 64 >> (j + k), rout = rin / 2;
                        const std::string fk = f + "." + std::to_string(k);
                        const std::string fn = sp + ".down" + std::to_string(i) + std::to_string(j) + "." + std::to_string(k);
                        if (!last) {
                            if (add_conv(h, fn, fk + ".0", fk + ".1", false, br[j].c, br[j].c, 3, 2, rin, rin, src, br[j + k + 1].d, ACT_RELU, nullptr, &idx,
                                         br[j].cp, br[j].cp)) return 1;
                            add_gemm_step(h, fn, idx, br[j + k + 1].d, br[j].c, rout, rout, {{const_cast<void*>(src), src_name, br[j].c, rin, rin}});
                            src = br[j + k + 1].d; src_name = fn;
                        } else {
                            // final term of the whole sum and nothing to upsample: ReLU here, straight into the fused output
                            const bool finish = j == i - 1 && !has_up;
                            void* dst = finish ? br[i].z : br[i].r;
                            if (add_conv(h, finish ? on : fn, fk + ".0", fk + ".1", false, br[j].c, br[i].c, 3, 2, rin, rin, src, dst, finish ? ACT_RELU : ACT_NONE, run,
                                         &idx, br[j].cp, br[i].cp)) return 1;
                            add_gemm_step(h, finish ? on : fn, idx, dst, br[i].c, res_i, res_i,
                                          {{const_cast<void*>(src), src_name, br[j].c, rin, rin}, {const_cast<void*>(run), run_name, br[i].c, res_i, res_i}});
                            run = dst; run_name = finish ? on : fn;
                        }
                    }
                }
                if (has_up) {
                    fuse_ins.insert(fuse_ins.begin(), StepIO{const_cast<void*>(run), run_name, br[i].c, res_i, res_i});
                    add_fuse_step(h, on, run, ups, br[i].z, i, br[i].cp, br[i].c, true, fuse_ins);
                }
            }
            for (int i = 0; i < nbr; ++i) { std::swap(br[i].x, br[i].z); br[i].name = sp + ".out" + std::to_string(i); }
        }
        nprev = nbr;
        prev_last = br[nbr - 1].x; prev_last_name = br[nbr - 1].name; prev_last_c = br[nbr - 1].c; prev_last_cp = br[nbr - 1].cp; prev_last_level = nbr - 1;
    }
    for (int l = 0; l < 4; ++l) h->hr_lvl[l] = br[l].x;
    h->featbuf = br[0].x;
    for (auto& st : h->backbone) {                    // every activation with C channels lives in a buffer of cpad(C) channels
        for (auto& io : st.ins) io.ld = cpad(io.C);
        for (auto& io : st.outs) io.ld = cpad(io.C);
    }
    return 0;
}

static int build_heads(hmv_handle* h) {
    const int hw = h->hm * h->hm;
    const int rows_max = h->fcap * h->S;
    const size_t e = h->esz;
    if (dev_alloc_t(h, &h->hm_int, static_cast<size_t>(h->mb_img) * kJoints * hw * 4)) return 1;
    if (dev_alloc_t(h, &h->xy, static_cast<size_t>(h->mb_img) * kJoints * 2 * 4)) return 1;
    if (dev_alloc_t(h, &h->xy_scaled, static_cast<size_t>(h->mb_img) * kJoints * 2 * 4)) return 1;
    if (dev_alloc_t(h, &h->wts, static_cast<size_t>(h->mb_img) * kJoints * 4 * 4)) return 1;
    if (dev_alloc_t(h, &h->bbox_int, static_cast<size_t>(h->mb_img) * 4 * 4)) return 1;
    if (dev_alloc_t(h, &h->intr_int, static_cast<size_t>(h->mb_img) * 4 * 4)) return 1;
    if (dev_alloc_t(h, &h->joints_int, static_cast<size_t>(h->fcap) * kJoints * 3 * 4)) return 1;
    if (dev_alloc_t(h, &h->gcn_h1, static_cast<size_t>(h->fcap) * kJoints * 256 * 4)) return 1;
    const size_t tokb = static_cast<size_t>(rows_max) * h->pitch;
    if (dev_alloc_t(h, &h->tok0_f32, tokb * 4) || dev_alloc_t(h, &h->tokA_f32, tokb * 4) || dev_alloc_t(h, &h->tokB_f32, tokb * 4) || dev_alloc_t(h, &h->ybuf, tokb * 4) ||
        dev_alloc_t(h, &h->hbuf, tokb * 4) || dev_alloc_t(h, &h->y2buf, tokb * 4))
        return 1;
    if (dev_alloc(h, &h->tok0_lp, tokb * e) || dev_alloc(h, &h->tokA_lp, tokb * e) || dev_alloc(h, &h->tokB_lp, tokb * e) || dev_alloc(h, &h->hnbuf, tokb * e) ||
        dev_alloc(h, &h->qkvbuf, static_cast<size_t>(rows_max) * 3072 * e) ||
        dev_alloc(h, &h->attbuf, static_cast<size_t>(rows_max) * 1024 * e) ||
        dev_alloc(h, &h->f1buf, static_cast<size_t>(rows_max) * 128 * e))
        return 1;

    const int lp_out = h->bf16 ? OUT_BF16_ROWMAJOR : OUT_F32_ROWMAJOR;
    if (h->hr) {
        // ---- pose_net of the HRNet configs: Conv2d(c0, 21, 3, stride 2, padding 1) on level 0 (handmvnet.py:50-56) ----
        std::vector<float> wf, bf;
        if (fold_conv(h, "pose_net", "", true, kJoints, h->hr_c[0], 3, wf, bf)) return 1;
        if (add_conv_raw(h, "pose_net", wf, bf, h->hr_c[0], kJoints, 3, 2, 64, 64, h->hr_lvl[0], h->hm_int, ACT_NONE, nullptr, &h->pose3, hw, h->hr_cp[0], 0)) return 1;
        // ---- one SampleNet per level on gathered rows (nets.py:55-63; gather-then-conv identity of SURVEY appendix D) ----
        for (int l = 0; l < 4; ++l) {
            const std::string q = "sample_nets." + std::to_string(l);
            const int c = h->hr_c[l], cp = h->hr_cp[l], co = c / 2;
            const size_t rows = static_cast<size_t>(h->mb_img) * kJoints * 4;
            if (dev_alloc(h, &h->hr_rows[l], rows * cp * e) || dev_alloc_t(h, &h->hr_wts[l], rows * sizeof(float))) return 1;
            std::vector<float> wf2, bf2;
            if (fold_conv(h, q + ".conv.0", q + ".conv.1", true, co, c, 1, wf2, bf2)) return 1;
            std::vector<float> wpad(static_cast<size_t>(co) * cp, 0.f);
            for (int r = 0; r < co; ++r) memcpy(&wpad[static_cast<size_t>(r) * cp], &wf2[static_cast<size_t>(r) * c], sizeof(float) * c);
            Layer L;
            L.name = q; L.kind = LK_FLAT; L.cin = cp; L.cout = co; L.K = cp;
            L.max_units = static_cast<int>(rows); L.in = h->hr_rows[l];
            L.bn = tc_pick_bn(co);
            if (L.bn == 0) L.bn = tc_pick_bn((co + 191) / 192 * 192);
            HMV_CHECK(L.bn > 0, "no tile width for " + q);
            L.n_alloc = (co + L.bn - 1) / L.bn * L.bn;
            if (!h->bf16) L.n_alloc = (co + 3) / 4 * 4;
            if (dev_alloc_t(h, &h->hr_g[l], rows * L.n_alloc * sizeof(float))) return 1;
            L.ep = make_ep(h->hr_g[l], L.n_alloc, OUT_F32_ROWMAJOR, ACT_RELU);
            if (finish_layer(h, L, wpad, bf2)) return 1;
            if (h->bf16 && build_tc(h, L)) return 1;
            h->hr_samp[l] = static_cast<int>(h->layers.size());
            h->layers.push_back(L);
        }
    } else {
    // ---- pose_net (layers.py:318-334 via handmvnet.py:71) ----
    if (add_conv(h, "pose_net.0", "pose_net.0", "pose_net.1", true, 1024, 512, 1, 1, h->hm, h->hm, h->featbuf, h->bufT1, ACT_RELU, nullptr, &h->pose0)) return 1;
    {
        std::vector<float> wf, bf;
        if (fold_conv(h, "pose_net.3", "", true, kJoints, 512, 1, wf, bf)) return 1;
        if (add_conv_raw(h, "pose_net.3", wf, bf, 512, kJoints, 1, 1, h->hm, h->hm, h->bufT1, h->hm_int, ACT_NONE, nullptr, &h->pose3, hw)) return 1;
    }
    // ---- SampleNet conv on gathered rows (nets.py:55-63; SURVEY appendix D identity) ----
    {
        // gathered rows live in bufT2 as [mb_img*84, 1024]; output G fp32 [mb_img*84, 512] in bufDS
        std::vector<float> wf, bf;
        if (fold_conv(h, "sample_nets.0.conv.0", "sample_nets.0.conv.1", true, 512, 1024, 1, wf, bf)) return 1;
        Layer L;
        L.name = "sample_nets.0"; L.kind = LK_FLAT; L.cin = 1024; L.cout = 512; L.K = 1024;
        L.max_units = h->mb_img * kJoints * 4; L.in = h->bufT2;
        L.bn = 256; L.n_alloc = 512;
        L.ep = make_ep(h->bufDS, 512, OUT_F32_ROWMAJOR, ACT_RELU);
        if (finish_layer(h, L, wf, bf)) return 1;
        if (h->bf16 && build_tc(h, L)) return 1;
        h->samp = static_cast<int>(h->layers.size());
        h->layers.push_back(L);
    }
    }
    // ---- constants: positional table (layers.py:136-150) and Chebyshev basis (layers.py:405-445) ----
    {
        std::vector<float> pe;
        if (const HostTensor* t = find_w(h, "pe")) {
            HMV_CHECK(static_cast<int>(t->data.size()) == h->S * h->d, "pe must be [21*V, feat_dim]");
            pe = t->data;
        } else {
            pe.assign(static_cast<size_t>(h->S) * h->d, 0.f);
            for (int p = 0; p < h->S; ++p)
                for (int i = 0; 2 * i < h->d; ++i) {
                    const float div = expf(static_cast<float>(2 * i) * static_cast<float>(-log(10000.0) / h->d));
                    const float a = static_cast<float>(p) * div;
                    pe[static_cast<size_t>(p) * h->d + 2 * i] = sinf(a);
                    if (2 * i + 1 < h->d) pe[static_cast<size_t>(p) * h->d + 2 * i + 1] = cosf(a);
                }
        }
        if (upload_f32(h, &h->pe, pe)) return 1;
        std::vector<float> basis;
        if (const HostTensor* t = find_w(h, "cheb_basis")) {
            HMV_CHECK(t->data.size() == 3 * 21 * 21, "cheb_basis must be [3,21,21]");
            basis = t->data;
        } else {
            static const int edges[20][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {0, 5}, {5, 6}, {6, 7}, {7, 8}, {0, 9}, {9, 10},
                                             {10, 11}, {11, 12}, {0, 13}, {13, 14}, {14, 15}, {15, 16}, {0, 17}, {17, 18}, {18, 19}, {19, 20}};
            float A[21][21] = {}, Lm[21][21], T2[21][21];
            for (auto& ed : edges) { A[ed[0]][ed[1]] = 1.f; A[ed[1]][ed[0]] = 1.f; }
            for (int i = 0; i < 21; ++i) A[i][i] += 1.f;
            for (int i = 0; i < 21; ++i) {            // row-normalise (utils.py:89-96)
                float rs = 0.f;
                for (int j = 0; j < 21; ++j) rs += A[i][j];
                for (int j = 0; j < 21; ++j) A[i][j] *= 1.f / rs;
            }
            float dg[21];
            for (int i = 0; i < 21; ++i) {            // D^-1/2 of the (unit) row sums (layers.py:439-441)
                float rs = 0.f;
                for (int j = 0; j < 21; ++j) rs += A[i][j];
                dg[i] = powf(rs, -0.5f);
            }
            for (int i = 0; i < 21; ++i)
                for (int j = 0; j < 21; ++j) Lm[i][j] = (i == j ? 1.f : 0.f) - dg[i] * A[i][j] * dg[j];
            for (int i = 0; i < 21; ++i)
                for (int j = 0; j < 21; ++j) {
                    float acc = 0.f;
                    for (int k = 0; k < 21; ++k) acc += Lm[i][k] * Lm[k][j];
                    T2[i][j] = 2.f * acc - (i == j ? 1.f : 0.f);
                }
            basis.assign(3 * 441, 0.f);
            for (int i = 0; i < 21; ++i)
                for (int j = 0; j < 21; ++j) {
                    basis[i * 21 + j] = i == j ? 1.f : 0.f;
                    basis[441 + i * 21 + j] = Lm[i][j];
                    basis[882 + i * 21 + j] = T2[i][j];
                }
        }
        if (upload_f32(h, &h->basis, basis)) return 1;
    }
    // ---- fusion transformer (fusion.py:7-30, layers.py:177-237) ----
    const int nl = h->cfg.fusion_layers;
    const int half = (nl - 1) / 2;
    // tokens (+PE) stay in tok0 so the stage API can read them back; layers ping-pong between A and B
    float* cur_f = h->tok0_f32; void* cur_l = h->tok0_lp;
    float* nxt_f = h->tokA_f32; void* nxt_l = h->tokA_lp;
    int s_in = h->S;
    for (int i = 0; i < nl; ++i) {
        const std::string p = "joints_late_fusion.attn_fusion." + std::to_string(i);
        FusionLayerPlan fp{};
        fp.s_in = s_in;
        if (i == half) { fp.nq = kJoints; fp.nk = s_in - kJoints; fp.kv_row0 = kJoints; }
        else { fp.nq = s_in; fp.nk = s_in; fp.kv_row0 = 0; }
        HMV_CHECK(fp.nk > 0, "cross-attention layer needs at least 2 views");
        NEED(wq, p + ".to_q.weight"); NEED(wk, p + ".to_k.weight"); NEED(wv, p + ".to_v.weight");
        NEED(wo, p + ".to_out.weight"); NEED(bo, p + ".to_out.bias");
        NEED(w1, p + ".ff.net.1.weight"); NEED(b1, p + ".ff.net.1.bias");
        NEED(w2, p + ".ff.net.4.weight"); NEED(b2, p + ".ff.net.4.bias");
        const int rows_in = h->fcap * s_in, rows_q = h->fcap * fp.nq;
        if (add_linear(h, p + ".qkv", {wq, wk, wv}, nullptr, h->d, h->pitch, cur_l, rows_in,
                       make_ep(h->qkvbuf, 3072, lp_out, ACT_NONE), &fp.qkv)) return 1;
        Epilogue eo = make_ep(h->ybuf, h->pitch, OUT_F32_ROWMAJOR, ACT_NONE);
        eo.residual = cur_f; eo.res_mode = RES_F32; eo.res_ld = h->pitch; eo.res_group = fp.nq; eo.res_stride = s_in;
        if (add_linear(h, p + ".to_out", {wo}, bo, 1024, 1024, h->attbuf, rows_q, eo, &fp.outp)) return 1;
        if (add_linear(h, p + ".ff1", {w1}, b1, h->d, h->pitch, h->hnbuf, rows_q, make_ep(h->f1buf, 128, lp_out, ACT_GELU), &fp.ff1)) return 1;
        Epilogue e2 = make_ep(h->y2buf, h->pitch, OUT_F32_ROWMAJOR, ACT_NONE);
        e2.residual = h->hbuf; e2.res_mode = RES_F32; e2.res_ld = h->pitch;
        if (add_linear(h, p + ".ff2", {w2}, b2, 128, 128, h->f1buf, rows_q, e2, &fp.ff2)) return 1;
        const char* lnn[3] = {".norm1", ".ff.net.0", ".norm2"};
        const float** gp[3] = {&fp.g1, &fp.gff, &fp.g2};
        const float** bp[3] = {&fp.b1, &fp.bff, &fp.b2};
        for (int k = 0; k < 3; ++k) {
            NEED(g, p + lnn[k] + ".weight"); NEED(b, p + lnn[k] + ".bias");
            HMV_CHECK(static_cast<int>(g->data.size()) == h->d, "unexpected LayerNorm shape in " + p);
            float *dg, *db;
            if (upload_f32(h, &dg, g->data) || upload_f32(h, &db, b->data)) return 1;
            *gp[k] = dg; *bp[k] = db;
        }
        if (h->bf16) {
            const int P = h->pitch;
            std::vector<float> wop(static_cast<size_t>(P) * 1024, 0.f), w2p(static_cast<size_t>(P) * 128, 0.f), bop(P, 0.f), b2p(P, 0.f);
            HMV_CHECK(static_cast<int64_t>(wo->data.size()) == static_cast<int64_t>(h->d) * 1024 && static_cast<int64_t>(w2->data.size()) == static_cast<int64_t>(h->d) * 128,
                      "unexpected to_out / ff.net.4 shape in " + p);
            memcpy(wop.data(), wo->data.data(), wo->data.size() * sizeof(float));
            memcpy(w2p.data(), w2->data.data(), w2->data.size() * sizeof(float));
            memcpy(bop.data(), bo->data.data(), h->d * sizeof(float));
            memcpy(b2p.data(), b2->data.data(), h->d * sizeof(float));
            if (upload_weights(h, &fp.wo_p, wop) || upload_weights(h, &fp.w2_p, w2p) || upload_f32(h, &fp.bo_p, bop) || upload_f32(h, &fp.b2_p, b2p)) return 1;
            for (int k = 0; k < 3; ++k) {
                NEED(g, p + lnn[k] + ".weight"); NEED(b, p + lnn[k] + ".bias");
                std::vector<float> gp2(P, 0.f), bp2(P, 0.f);
                memcpy(gp2.data(), g->data.data(), h->d * sizeof(float));
                memcpy(bp2.data(), b->data.data(), h->d * sizeof(float));
                if (upload_f32(h, &fp.ln_p[2 * k], gp2) || upload_f32(h, &fp.ln_p[2 * k + 1], bp2)) return 1;
            }
        }
        fp.res_in = cur_f; fp.in_lp = cur_l; fp.out_f32 = nxt_f; fp.out_lp = nxt_l;
        h->fusion.push_back(fp);
        cur_f = nxt_f; cur_l = nxt_l;
        if (nxt_f == h->tokA_f32) { nxt_f = h->tokB_f32; nxt_l = h->tokB_lp; }
        else { nxt_f = h->tokA_f32; nxt_l = h->tokA_lp; }
        s_in = fp.nq;
    }
    h->fused_f32 = cur_f;
    // ---- graph head (nets.py:119-139) ----
    const int gc[4] = {h->d, 256, 64, 3};
    for (int i = 0; i < 3; ++i) {
        const std::string p = "joints_decoder.joints_gcn" + std::to_string(i + 1);
        NEED(w, p + ".weight"); NEED(b, p + ".bias");
        HMV_CHECK(static_cast<int64_t>(w->data.size()) == 3LL * gc[i] * gc[i + 1] && static_cast<int>(b->data.size()) == gc[i + 1],
                  "unexpected ChebConv shape in " + p);
        if (upload_f32(h, &h->gcn_w[i], w->data) || upload_f32(h, &h->gcn_b[i], b->data)) return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------------
template <typename T>
static int run_backbone_t(hmv_handle* h, const float* x, int n_img, int num_steps, cudaStream_t s) {
    const int total = static_cast<int>(h->backbone.size());
    if (num_steps < 0 || num_steps > total) num_steps = total;
    for (int i = 0; i < num_steps; ++i) {
        Step& st = h->backbone[i];
        if (st.kind == SK_PACK) {
            ++h->launches;
            const float* xf = x;
            if (h->x_u8) {                            // fp32 check mode of the uint8 entry points: normalise first
                const size_t total = static_cast<size_t>(n_img) * 3 * h->img * h->img;
                if (!h->u8_f32 && dev_alloc_t(h, &h->u8_f32, static_cast<size_t>(h->mb_img) * 3 * h->img * h->img * sizeof(float), false)) return 1;
                ++h->launches;
                if (u8_to_f32_norm_launch(reinterpret_cast<const uint8_t*>(x), h->u8_f32, total, h->img * h->img, h->norm, s)) return 1;
                xf = h->u8_f32;
            }
            if (pack_input_launch<T>(xf, static_cast<T*>(st.out), n_img, h->img, h->img, st.H, st.W, 3, s)) return 1;
        } else if (st.kind == SK_STEM_POOL) {
            ++h->launches;
            if (stem_pool_launch(x, h->x_u8, h->norm, static_cast<const bf16*>(h->stem_w), h->stem_b, static_cast<bf16*>(st.out), n_img,
                                 h->num_sms, h->err_flag_dev, s)) return 1;
        } else if (st.kind == SK_HR_STEM) {
            ++h->launches;
            if (hr_stem_launch<T>(x, h->x_u8, h->norm, h->hr_stem_w, h->stem_b, static_cast<T*>(st.out), n_img, h->img, s)) return 1;
        } else if (st.kind == SK_FUSE_SUM) {
            ++h->launches;
            FuseSumParams fp = st.fuse;
            fp.n_img = n_img;
            if (fuse_sum_launch<T>(fp, s)) return 1;
        } else if (st.kind == SK_TAIL) {
            if (run_tail(h, st.layer, n_img, s)) return 1;
        } else if (st.kind == SK_SEAM) {
            if (run_seam(h, st.layer, n_img, s)) return 1;
        } else if (st.kind == SK_MAXPOOL) {
            ++h->launches;
            if (maxpool_launch<T>(static_cast<const T*>(st.in), static_cast<T*>(st.out), n_img, st.H * 2, st.W * 2, st.C, s)) return 1;
        } else {
            if (run_layer(h, h->layers[st.layer], n_img, s)) return 1;
        }
    }
    return 0;
}
static int run_backbone(hmv_handle* h, const float* x, int n_img, int num_steps, cudaStream_t s) {
    return h->bf16 ? run_backbone_t<bf16>(h, x, n_img, num_steps, s) : run_backbone_t<float>(h, x, n_img, num_steps, s);
}

static int run_pose(hmv_handle* h, int n_img, float* heatmap_out, float* xy_scaled_out, cudaStream_t s) {
    if (h->pose0 >= 0 && run_layer(h, h->layers[h->pose0], n_img, s)) return 1;     // (HRNet: pose_net is one 3x3 / stride-2 conv, handmvnet.py:50-56)
    float* hm = heatmap_out ? heatmap_out : h->hm_int;
    if (run_layer(h, h->layers[h->pose3], n_img, s, hm)) return 1;
    ++h->launches;
    return softargmax_launch(hm, h->xy, xy_scaled_out ? xy_scaled_out : h->xy_scaled, n_img * kJoints, h->hm, h->hm,
                             1000.f, static_cast<float>(h->img) / static_cast<float>(h->hm), s);
}

// tokens of this micro-batch are written at sample offset `off` of the (fusion-pass wide) token stream
template <typename T>
static int run_sample_t(hmv_handle* h, int n_img, const float* bbox, const float* intr, cudaStream_t s, int off = 0) {
    if (h->hr) {
        // one SampleNet per level (handmvnet.py:97,185): the joint coordinates address every level with the SAME un-rescaled
        // heat-map pixel numbers (nets.py:46-53 normalises by the level's own size) - reproduced by sample_gather_kernel
        TokenParams tp{};
        tp.n_src = 4;
        for (int l = 0; l < 4; ++l) {
            const int res = 64 >> l;
            h->launches += 1;
            if (sample_gather_launch<T>(static_cast<const T*>(h->hr_lvl[l]), h->xy, static_cast<T*>(h->hr_rows[l]), h->hr_wts[l], n_img, res, res, h->hr_cp[l], s)) return 1;
            const Layer& L = h->layers[h->hr_samp[l]];
            if (run_layer(h, h->layers[h->hr_samp[l]], n_img * kJoints * 4, s)) return 1;
            tp.src[l] = TokenSource{h->hr_g[l], h->hr_wts[l], L.ep.ldc, h->hr_c[l] / 2};
        }
        tp.xy = h->xy;
        tp.bbox = bbox; tp.intr = intr; tp.pe = h->cfg.use_sin ? h->pe : nullptr;
        const size_t tok_off = static_cast<size_t>(off) * h->S * h->pitch;
        tp.tok_f32 = h->tok0_f32 + tok_off; tp.tok_lp = static_cast<T*>(h->tok0_lp) + tok_off;
        tp.n_img = n_img; tp.feat = h->feat; tp.d = h->d; tp.pitch = h->pitch; tp.tokens_per_sample = h->S;
        tp.use_pos2d = h->cfg.use_pos2d; tp.use_crop = h->cfg.use_crop;
        ++h->launches;
        return tokens_launch<T>(tp, s);
    }
    h->launches += 2;
    if (sample_gather_launch<T>(static_cast<const T*>(h->featbuf), h->xy, static_cast<T*>(h->bufT2), h->wts, n_img, h->hm, h->hm, 1024, s)) return 1;
    if (run_layer(h, h->layers[h->samp], n_img * kJoints * 4, s)) return 1;
    TokenParams tp{};
    tp.n_src = 1;
    tp.src[0] = TokenSource{static_cast<const float*>(h->bufDS), h->wts, 512, h->feat};
    tp.xy = h->xy;
    tp.bbox = bbox; tp.intr = intr; tp.pe = h->cfg.use_sin ? h->pe : nullptr;
    const size_t tok_off = static_cast<size_t>(off) * h->S * h->pitch;
    tp.tok_f32 = h->tok0_f32 + tok_off; tp.tok_lp = static_cast<T*>(h->tok0_lp) + tok_off;
    tp.n_img = n_img; tp.feat = h->feat; tp.d = h->d; tp.pitch = h->pitch; tp.tokens_per_sample = h->S;
    tp.use_pos2d = h->cfg.use_pos2d; tp.use_crop = h->cfg.use_crop;
    return tokens_launch<T>(tp, s);
}

template <typename T>
static int run_fusion_t(hmv_handle* h, int n, cudaStream_t s) {
    for (auto& fp : h->fusion) {
        const int rows_in = n * fp.s_in, rows_q = n * fp.nq;
        if (run_layer(h, h->layers[fp.qkv], rows_in, s)) return 1;
        if constexpr (sizeof(T) == 2) {
            if (h->fuse_block && h->pitch == 576) {      // attention + to_out + residual + 3 LayerNorms + feed-forward in one kernel
                FusionBlockParams q{};
                q.qkv = static_cast<const bf16*>(h->qkvbuf); q.ld_qkv = 3072;
                q.res_in = fp.res_in; q.s_in = fp.s_in; q.nq = fp.nq; q.nk = fp.nk; q.kv_row0 = fp.kv_row0;
                q.pitch = h->pitch; q.d = h->d; q.scale_log2e = 0.08838834764831845f * 1.4426950408889634f;
                q.wo = static_cast<const bf16*>(fp.wo_p); q.bo = fp.bo_p;
                q.g1 = fp.ln_p[0]; q.b1 = fp.ln_p[1]; q.gff = fp.ln_p[2]; q.bff = fp.ln_p[3]; q.g2 = fp.ln_p[4]; q.b2 = fp.ln_p[5];
                q.w1 = static_cast<const bf16*>(h->layers[fp.ff1].w); q.bf1 = h->layers[fp.ff1].bias;
                q.w2 = static_cast<const bf16*>(fp.w2_p); q.bf2 = fp.b2_p;
                q.out_f32 = fp.out_f32; q.out_lp = static_cast<bf16*>(fp.out_lp);
                ++h->launches;
                if (fusion_block_launch(q, n, s)) return 1;
                continue;
            }
        }
        h->launches += 3;
        if constexpr (sizeof(T) == 2) {
            if (attention_mma_launch(static_cast<const bf16*>(h->qkvbuf), 3072, static_cast<bf16*>(h->attbuf), 1024, n, fp.s_in, 0,
                                     fp.nq, fp.kv_row0, fp.nk, 8, 128, 0.08838834764831845f /* 128^-0.5 */, s)) return 1;
        } else {
            if (attention_launch<T>(static_cast<const T*>(h->qkvbuf), 3072, static_cast<T*>(h->attbuf), 1024, n, fp.s_in, 0, fp.nq,
                                    fp.kv_row0, fp.nk, 8, 128, 0.08838834764831845f /* 128^-0.5 */, s)) return 1;
        }
        if (run_layer(h, h->layers[fp.outp], rows_q, s)) return 1;
        if (layernorm_launch<T>(h->ybuf, h->pitch, fp.g1, fp.b1, h->hbuf, h->pitch, fp.gff, fp.bff, static_cast<T*>(h->hnbuf),
                                h->pitch, rows_q, h->d, 1e-5f, s)) return 1;
        if (run_layer(h, h->layers[fp.ff1], rows_q, s)) return 1;
        if (run_layer(h, h->layers[fp.ff2], rows_q, s)) return 1;
        if (layernorm_launch<T>(h->y2buf, h->pitch, fp.g2, fp.b2, fp.out_f32, h->pitch, nullptr, nullptr, static_cast<T*>(fp.out_lp),
                                h->pitch, rows_q, h->d, 1e-5f, s)) return 1;
    }
    return 0;
}

static int run_gcn(hmv_handle* h, int n, float* out, cudaStream_t s) {
    GcnParams g{};
    g.x = h->fused_f32; g.ld = h->pitch; g.d_in = h->d;
    for (int i = 0; i < 3; ++i) { g.w[i] = h->gcn_w[i]; g.b[i] = h->gcn_b[i]; }
    g.basis = h->basis; g.out = out ? out : h->joints_int; g.batch = n;
    h->launches += 2;
    return gcn_launch(g, h->gcn_h1, s);
}

static int mark_phase(hmv_handle* h, int id, cudaStream_t s) {
    if (!h->profiling) return 0;
    cudaEvent_t e;
    if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); }
    else HMV_CUDA(cudaEventCreate(&e));
    HMV_CUDA(cudaEventRecord(e, s));
    h->phase_marks.push_back({id, e});
    return 0;
}

// front half for one micro-batch of n <= mb samples (pointers already offset): backbone, pose_net, soft-argmax,
// sampling and token assembly; its tokens land at sample offset `off` of the current fusion pass
static int run_front(hmv_handle* h, const float* x, const float* bbox, const float* intr, int n, int off, float* heatmap,
                     float* xy_scaled, cudaStream_t s) {
    const int n_img = n * h->V;
    if (mark_phase(h, 0, s)) return 1;
    if (run_backbone(h, x, n_img, -1, s)) return 1;
    if (mark_phase(h, 1, s)) return 1;
    if (run_pose(h, n_img, heatmap, xy_scaled, s)) return 1;
    if (h->bf16 ? run_sample_t<bf16>(h, n_img, bbox, intr, s, off) : run_sample_t<float>(h, n_img, bbox, intr, s, off)) return 1;
    return mark_phase(h, 2, s);
}
// back half for n <= fcap samples: fusion transformer + graph head (small, latency-bound kernels: run once per pass)
static int run_back(hmv_handle* h, int n, float* joints, cudaStream_t s) {
    if (mark_phase(h, 3, s)) return 1;
    if (h->bf16 ? run_fusion_t<bf16>(h, n, s) : run_fusion_t<float>(h, n, s)) return 1;
    if (mark_phase(h, 4, s)) return 1;
    if (run_gcn(h, n, joints, s)) return 1;
    return mark_phase(h, 5, s);
}

static int check_flag(hmv_handle* h) {
    if (h->err_flag_host && *h->err_flag_host != 0) {
        const int code = *h->err_flag_host;
        *h->err_flag_host = 0;
        set_error("device pipeline timeout (role code " + std::to_string(code) +
                  ": conv_gemm_tc 1=TMA producer, 2=MMA/tmem-empty, 3=MMA/smem-full, 4=epilogue, 5/6=residual ring; stem_pool 11-14; "
                  "bottleneck_tail 21-32; 41 = hmv_preprocess box empty or more than 9x the output size)");
        return 1;
    }
    return 0;
}

}  // namespace hmv

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* hmv_last_error(void) { return hmv::get_error(); }
const char* hmv_version(void) { return "handmvnet_b200 0.1 (sm_100a)"; }

int hmv_create(const hmv_config* cfg, hmv_handle** out) {
    HMV_CHECK(cfg && out, "hmv_create: null argument");
    HMV_CHECK(cfg->num_views >= 1 && cfg->num_views <= 16, "num_views must be in [1,16]");
    HMV_CHECK(cfg->fusion_layers >= 1 && cfg->fusion_layers % 2 == 1, "num_layers must be an odd number");   // fusion.py:11
    HMV_CHECK(cfg->image_size == 256 && cfg->heatmap_size == 32, "only image_size 256 / heatmap_size 32 (release configs) are supported");
    HMV_CHECK(cfg->micro_batch >= 1, "micro_batch must be >= 1");
    HMV_CHECK(cfg->precision == HMV_PRECISION_BF16 || cfg->precision == HMV_PRECISION_FP32, "unknown precision");
    HMV_CHECK(cfg->backbone == HMV_BACKBONE_RESNET50_PAPER || cfg->backbone == HMV_BACKBONE_HRNET, "unknown backbone");
    int ndev = 0;
    HMV_CUDA(cudaGetDeviceCount(&ndev));
    HMV_CHECK(ndev > 0, "no CUDA device: handmvnet_b200 has no CPU fallback");
    HMV_CHECK(cfg->device >= 0 && cfg->device < ndev, "hmv_create: no such CUDA device");
    hmv::DeviceGuard guard(cfg->device);
    cudaDeviceProp prop;
    HMV_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    HMV_CHECK(prop.major == 10, "handmvnet_b200 is built for sm_100a (Blackwell B200) only");
    hmv_handle* h = new hmv_handle();
    h->cfg = *cfg;
    h->bf16 = cfg->precision == HMV_PRECISION_BF16;
    h->esz = h->bf16 ? 2 : 4;
    h->V = cfg->num_views; h->S = 21 * cfg->num_views;
    h->img = cfg->image_size; h->hm = cfg->heatmap_size;
    h->hr = cfg->backbone == HMV_BACKBONE_HRNET;
    if (h->hr) {                                     // feat_dim = sum(backbone_channels) / 2 (handmvnet.py:88)
        int sum = 0;
        for (int l = 0; l < 4; ++l) {
            h->hr_c[l] = cfg->hr_channels[l];
            if (h->hr_c[l] < 8 || h->hr_c[l] % 8 != 0 || h->hr_c[l] > 512) { hmv::set_error("hr_channels must be multiples of 8 in [8, 512]"); delete h; return 1; }
            const int c = h->hr_c[l];
            h->hr_cp[l] = !h->bf16 ? c : (c <= 64 ? 64 : c <= 128 ? 128 : c <= 192 ? 192 : (c + 63) / 64 * 64);
            sum += c;
        }
        h->feat = sum / 2;
    }
    h->d = h->feat + (cfg->use_pos2d ? 2 : 0) + (cfg->use_crop ? 10 : 0);
    h->pitch = (h->d + 63) / 64 * 64;
    if (h->pitch > 256 && h->pitch % 128 != 0 && h->pitch < (h->d + 175) / 176 * 176) h->pitch = ((h->d + 175) / 176 * 176 + 63) / 64 * 64;   // to_out tile width (176) must fit the token pitch
    h->mb = cfg->micro_batch; h->mb_img = h->mb * h->V;
    h->fcap = h->mb >= 256 ? h->mb : (256 / h->mb) * h->mb;          // a multiple of the micro-batch
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("HMV_NUM_SMS")) {         // cap the persistent grids (two handles sharing one GPU on two streams)
        const int v = atoi(e);
        if (v >= 1 && v < h->num_sms) h->num_sms = v;
    }
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->err_flag_host), sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->err_flag_dev), h->err_flag_host, 0) != cudaSuccess) {
        hmv::set_error("cannot allocate the mapped error flag");
        delete h;
        return 1;
    }
    *h->err_flag_host = 0;
    if (h->bf16 && (hmv::tc_init() || hmv::bt_init() || hmv::bn_init())) { delete h; return 1; }
    {
        const char* e = getenv("HMV_FUSE_NEXT");
        h->fuse_next = !(e && e[0] == '0');
    }
    {
        const char* u = getenv("HMV_FUSION_UNFUSED");
        h->fuse_block = !(u && u[0] == '1');
        const char* e = getenv("HMV_FUSE_TAIL");
        if (e && e[0] >= '0' && e[0] <= '7') h->fuse_mask = e[0] - '0';
    }
    *out = h;
    return 0;
}

int hmv_destroy(hmv_handle* h) {
    if (!h) return 0;
    hmv::DeviceGuard guard(h->cfg.device);
    cudaDeviceSynchronize();
    if (h->ws_event) cudaEventDestroy(h->ws_event);
    for (void* p : h->allocs) cudaFree(p);
    for (int i = 0; i < 2; ++i) {
        if (h->xstage[i]) cudaFree(h->xstage[i]);
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
        if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
    }
    if (h->d_bbox) cudaFree(h->d_bbox);
    if (h->d_intr) cudaFree(h->d_intr);
    if (h->d_hm) cudaFree(h->d_hm);
    if (h->d_xy) cudaFree(h->d_xy);
    if (h->d_j) cudaFree(h->d_j);
    for (auto e : h->ev_done) if (e) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
    for (auto& kv : h->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (auto& g : h->ptr_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
    if (h->graph_in) cudaEventDestroy(h->graph_in);
    if (h->graph_out) cudaEventDestroy(h->graph_out);
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->err_flag_host) cudaFreeHost(h->err_flag_host);
    delete h;
    return 0;
}

int hmv_set_weight(hmv_handle* h, const char* name, const float* data, const int64_t* dims, int32_t ndim) {
    HMV_CHECK(h && name && data, "hmv_set_weight: null argument");
    HMV_CHECK(!h->prepared, "hmv_set_weight after hmv_prepare");
    hmv::HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) { t.dims.push_back(dims[i]); n *= dims[i]; }
    t.data.assign(data, data + n);
    h->weights[name] = std::move(t);
    return 0;
}

int hmv_prepare(hmv_handle* h) {
    HMV_CHECK(h, "hmv_prepare: null handle");
    HMV_CHECK(!h->prepared, "hmv_prepare called twice");
    hmv::DeviceGuard guard(h->cfg.device);
    if (h->hr ? hmv::build_hrnet(h) : hmv::build_backbone(h)) return 1;
    if (hmv::build_heads(h)) return 1;
    {   // small-batch CUDA-graph path (HMV_NO_GRAPH=1 disables it)
        const char* e = getenv("HMV_NO_GRAPH");
        h->graph_max_batch = (e && e[0] == '1') ? 0 : (h->mb < 8 ? h->mb : 8);
        if (h->graph_max_batch > 0) {
            const size_t nimg = static_cast<size_t>(h->graph_max_batch) * h->V;
            if (hmv::dev_alloc_t(h, &h->g_x, nimg * 3 * h->img * h->img * sizeof(float), false) ||
                hmv::dev_alloc_t(h, &h->g_bbox, nimg * 4 * sizeof(float)) || hmv::dev_alloc_t(h, &h->g_intr, nimg * 4 * sizeof(float)) ||
                hmv::dev_alloc_t(h, &h->g_hm, nimg * 21 * h->hm * h->hm * sizeof(float)) ||
                hmv::dev_alloc_t(h, &h->g_xy, nimg * 21 * 2 * sizeof(float)) ||
                hmv::dev_alloc_t(h, &h->g_j, static_cast<size_t>(h->graph_max_batch) * 21 * 3 * sizeof(float)))
                return 1;
            HMV_CUDA(cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking));
            HMV_CUDA(cudaEventCreateWithFlags(&h->graph_in, cudaEventDisableTiming));
            HMV_CUDA(cudaEventCreateWithFlags(&h->graph_out, cudaEventDisableTiming));
        }
    }
    HMV_CUDA(cudaDeviceSynchronize());
    h->weights.clear();
    h->prepared = true;
    return 0;
}

// x advanced by `samples` samples (x is fp32, or uint8 when the call came through a *_u8 entry point)
static inline const float* x_at(const hmv_handle* h, const float* x, size_t samples) {
    const size_t per_sample = static_cast<size_t>(h->V) * 3 * h->img * h->img * (h->x_u8 ? 1 : sizeof(float));
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(x) + samples * per_sample);
}

static int forward_eager(hmv_handle* h, const float* x, const float* bbox, const float* intr, int batch, float* heatmap,
                         float* joints_crop_img, float* joints_cam, cudaStream_t s) {
    for (int p0 = 0; p0 < batch; p0 += h->fcap) {                     // fusion passes
        const int np = batch - p0 < h->fcap ? batch - p0 : h->fcap;
        for (int s0 = p0; s0 < p0 + np; s0 += h->mb) {                // backbone micro-batches
            const int n = p0 + np - s0 < h->mb ? p0 + np - s0 : h->mb;
            if (hmv::run_front(h, x_at(h, x, s0), bbox ? bbox + static_cast<size_t>(s0) * h->V * 4 : nullptr,
                               intr ? intr + static_cast<size_t>(s0) * h->V * 4 : nullptr, n, s0 - p0,
                               heatmap ? heatmap + static_cast<size_t>(s0) * h->V * 21 * h->hm * h->hm : nullptr,
                               joints_crop_img ? joints_crop_img + static_cast<size_t>(s0) * h->V * 21 * 2 : nullptr, s))
                return 1;
        }
        if (hmv::run_back(h, np, joints_cam ? joints_cam + static_cast<size_t>(p0) * 21 * 3 : nullptr, s)) return 1;
    }
    return 0;
}

// Small batches are launch-bound (88 launches of a few microseconds each): replay them as one CUDA graph over the
// handle's internal I/O buffers; inputs / outputs are staged with device-to-device copies around the launch.
static int forward_graph(hmv_handle* h, const float* x, const float* bbox, const float* intr, int batch, float* heatmap,
                         float* joints_crop_img, float* joints_cam, cudaStream_t s) {
    hmv_handle::GraphSlot& g = h->graphs[batch];
    if (g.exec == nullptr && g.calls == 0) {          // first call: eager (lazy cudaFuncSetAttribute calls happen here)
        ++g.calls;
        return forward_eager(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, s);
    }
    const size_t nimg = static_cast<size_t>(batch) * h->V;
    const size_t xb = nimg * 3 * h->img * h->img * sizeof(float);
    cudaStream_t user = s;
    s = h->graph_stream;                               // everything below runs on the handle's own stream
    HMV_CUDA(cudaEventRecord(h->graph_in, user));
    HMV_CUDA(cudaStreamWaitEvent(s, h->graph_in, 0));
    HMV_CUDA(cudaMemcpyAsync(h->g_x, x, xb, cudaMemcpyDeviceToDevice, s));
    if (h->cfg.use_crop) {
        HMV_CUDA(cudaMemcpyAsync(h->g_bbox, bbox, nimg * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        HMV_CUDA(cudaMemcpyAsync(h->g_intr, intr, nimg * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (g.exec == nullptr) {
        cudaGraph_t graph = nullptr;
        HMV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int64_t launches0 = h->launches;
        int rc = forward_eager(h, h->g_x, h->cfg.use_crop ? h->g_bbox : nullptr, h->cfg.use_crop ? h->g_intr : nullptr, batch,
                               h->g_hm, h->g_xy, h->g_j, s);
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        g.kernels = static_cast<int>(h->launches - launches0);
        h->launches = launches0;
        if (rc == 0 && ce == cudaSuccess && graph != nullptr) ce = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != 0 || ce != cudaSuccess || g.exec == nullptr) {     // capture unsupported here: eager from now on
            cudaGetLastError();
            g.exec = nullptr;
            h->graph_max_batch = 0;
            if (rc != 0) return rc;
            if (forward_eager(h, h->g_x, h->cfg.use_crop ? h->g_bbox : nullptr, h->cfg.use_crop ? h->g_intr : nullptr, batch,
                              h->g_hm, h->g_xy, h->g_j, s)) return 1;
        }
        g.calls = 2;
    }
    if (g.exec != nullptr) {
        HMV_CUDA(cudaGraphLaunch(g.exec, s));
        h->launches += g.kernels;                      // kernels inside the replayed graph
    }
    if (heatmap) HMV_CUDA(cudaMemcpyAsync(heatmap, h->g_hm, nimg * 21 * h->hm * h->hm * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (joints_crop_img) HMV_CUDA(cudaMemcpyAsync(joints_crop_img, h->g_xy, nimg * 21 * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (joints_cam) HMV_CUDA(cudaMemcpyAsync(joints_cam, h->g_j, static_cast<size_t>(batch) * 21 * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    HMV_CUDA(cudaEventRecord(h->graph_out, s));
    HMV_CUDA(cudaStreamWaitEvent(user, h->graph_out, 0));
    return 0;
}

// Pointer-keyed graphs for batches above graph_max_batch (see hmv_handle::PtrGraph).
static int forward_graph_ptr(hmv_handle* h, const float* x, const float* bbox, const float* intr, int batch, float* heatmap,
                             float* joints_crop_img, float* joints_cam, cudaStream_t user) {
    constexpr size_t kMaxEntries = 8;
    hmv_handle::PtrGraph* g = nullptr;
    for (auto& e : h->ptr_graphs)
        if (e.x == x && e.bbox == bbox && e.intr == intr && e.hm == heatmap && e.j2d == joints_crop_img && e.j3d == joints_cam && e.batch == batch) { g = &e; break; }
    ++h->ptr_graph_clock;
    if (!g) {                                          // first sight: remember the pointer set, run eagerly
        if (h->ptr_graphs.size() >= kMaxEntries) {
            size_t lru = 0;
            for (size_t i = 1; i < h->ptr_graphs.size(); ++i) if (h->ptr_graphs[i].last_use < h->ptr_graphs[lru].last_use) lru = i;
            if (h->ptr_graphs[lru].exec) cudaGraphExecDestroy(h->ptr_graphs[lru].exec);
            h->ptr_graphs.erase(h->ptr_graphs.begin() + lru);
        }
        h->ptr_graphs.push_back({x, bbox, intr, heatmap, joints_crop_img, joints_cam, batch, 1, 0, nullptr, h->ptr_graph_clock});
        return forward_eager(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, user);
    }
    g->last_use = h->ptr_graph_clock;
    cudaStream_t s = h->graph_stream;
    if (g->exec == nullptr) {                          // second sight: capture
        cudaGraph_t graph = nullptr;
        HMV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int64_t launches0 = h->launches;
        const int rc = forward_eager(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, s);
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        g->kernels = static_cast<int>(h->launches - launches0);
        h->launches = launches0;
        if (rc == 0 && ce == cudaSuccess && graph != nullptr) ce = cudaGraphInstantiate(&g->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != 0 || ce != cudaSuccess || g->exec == nullptr) {      // capture unsupported here: eager from now on
            cudaGetLastError();
            g->exec = nullptr;
            h->ptr_graphs_ok = false;
            if (rc != 0) return rc;
            return forward_eager(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, user);
        }
    }
    HMV_CUDA(cudaEventRecord(h->graph_in, user));
    HMV_CUDA(cudaStreamWaitEvent(s, h->graph_in, 0));
    HMV_CUDA(cudaGraphLaunch(g->exec, s));
    h->launches += g->kernels;
    HMV_CUDA(cudaEventRecord(h->graph_out, s));
    HMV_CUDA(cudaStreamWaitEvent(user, h->graph_out, 0));
    return 0;
}

int hmv_forward(hmv_handle* h, const float* x, const float* bbox, const float* intr, int32_t batch, float* heatmap,
                float* joints_crop_img, float* joints_cam, void* stream) {
    HMV_CHECK(h && h->prepared, "hmv_forward: handle not prepared");
    HMV_CHECK(batch >= 0, "negative batch");
    HMV_CHECK(x || batch == 0, "hmv_forward: x is null");
    HMV_CHECK(!h->cfg.use_crop || (bbox && intr) || batch == 0, "'crop' positional encoding needs bbox and cam_params[\"intrinsic\"]");
    if (hmv::check_flag(h)) return 1;
    if (batch == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (hmv::ws_acquire(h, s)) return 1;
    int rc;
    if (batch <= h->graph_max_batch && !h->profiling) rc = forward_graph(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, s);
    else if (h->graph_max_batch > 0 && h->ptr_graphs_ok && !h->profiling)
        rc = forward_graph_ptr(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, s);
    else rc = forward_eager(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, s);
    if (rc) return rc;
    return hmv::ws_release(h, s);          // the graph paths end with `s` waiting for the replay, so `s` orders after it
}

int hmv_synchronize(hmv_handle* h) {
    HMV_CHECK(h, "null handle");
    hmv::DeviceGuard guard(h->cfg.device);
    HMV_CUDA(cudaDeviceSynchronize());
    return hmv::check_flag(h);
}

static int ensure_host_pipeline(hmv_handle* h, int batch) {
    const size_t per_sample_x = static_cast<size_t>(h->V) * 3 * h->img * h->img;
    if (!h->copy_stream) {
        HMV_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        HMV_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->xstage[i]), per_sample_x * h->mb * sizeof(float)));
            HMV_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
            HMV_CUDA(cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
        }
    }
    if (batch > h->host_cap) {
        float** bufs[5] = {&h->d_bbox, &h->d_intr, &h->d_hm, &h->d_xy, &h->d_j};
        for (float** b : bufs) { if (*b) cudaFree(*b); *b = nullptr; }     // a failed cudaMalloc below leaves nulls, not freed addresses
        h->host_cap = 0;
        const size_t nimg = static_cast<size_t>(batch) * h->V;
        HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_bbox), nimg * 4 * sizeof(float)));
        HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_intr), nimg * 4 * sizeof(float)));
        HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_hm), nimg * 21 * h->hm * h->hm * sizeof(float)));
        HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_xy), nimg * 21 * 2 * sizeof(float)));
        HMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_j), static_cast<size_t>(batch) * 21 * 3 * sizeof(float)));
        h->host_cap = batch;
    }
    return 0;
}

// Enqueue one host-buffer forward: x is staged through two device buffers (copy stream) while the compute stream
// works on the previous chunk; bbox / intrinsics / outputs travel on the compute stream.  `ramp` uses two smaller
// first chunks so that compute starts after a short copy (synchronous calls); asynchronous calls keep full chunks
// because the copy of their first chunk already overlaps the previous call.
static int enqueue_host(hmv_handle* h, const float* x, const float* bbox, const float* intr, int batch, float* heatmap,
                        float* joints_crop_img, float* joints_cam, bool ramp) {
    const size_t per_sample_bytes = static_cast<size_t>(h->V) * 3 * h->img * h->img * (h->x_u8 ? 1 : sizeof(float));
    const size_t nimg = static_cast<size_t>(batch) * h->V;
    cudaStream_t cs = h->copy_stream, ks = h->compute_stream;
    if (hmv::ws_acquire(h, ks)) return 1;             // e.g. an hmv_forward still running on the caller's stream
    if (h->cfg.use_crop) {
        HMV_CUDA(cudaMemcpyAsync(h->d_bbox, bbox, nimg * 4 * sizeof(float), cudaMemcpyHostToDevice, ks));
        HMV_CUDA(cudaMemcpyAsync(h->d_intr, intr, nimg * 4 * sizeof(float), cudaMemcpyHostToDevice, ks));
    }
    int chunk = 0;
    for (int p0 = 0; p0 < batch; p0 += h->fcap) {
        const int np = batch - p0 < h->fcap ? batch - p0 : h->fcap;
        for (int s0 = p0; s0 < p0 + np; ++chunk) {
            int cap = h->mb;
            if (ramp && h->mb >= 8 && chunk == 0) cap = h->mb / 4;
            else if (ramp && h->mb >= 8 && chunk == 1) cap = h->mb - h->mb / 4;
            const int n = p0 + np - s0 < cap ? p0 + np - s0 : cap;
            const int b = static_cast<int>(h->chunk_seq & 1);
            if (h->chunk_seq >= 2) HMV_CUDA(cudaStreamWaitEvent(cs, h->ev_consumed[b], 0));
            HMV_CUDA(cudaMemcpyAsync(h->xstage[b], x_at(h, x, s0), per_sample_bytes * n, cudaMemcpyHostToDevice, cs));
            HMV_CUDA(cudaEventRecord(h->ev_copied[b], cs));
            HMV_CUDA(cudaStreamWaitEvent(ks, h->ev_copied[b], 0));
            if (hmv::run_front(h, h->xstage[b], h->cfg.use_crop ? h->d_bbox + static_cast<size_t>(s0) * h->V * 4 : nullptr,
                               h->cfg.use_crop ? h->d_intr + static_cast<size_t>(s0) * h->V * 4 : nullptr, n, s0 - p0,
                               h->d_hm + static_cast<size_t>(s0) * h->V * 21 * h->hm * h->hm,
                               h->d_xy + static_cast<size_t>(s0) * h->V * 21 * 2, ks))
                return 1;
            HMV_CUDA(cudaEventRecord(h->ev_consumed[b], ks));
            ++h->chunk_seq;
            s0 += n;
        }
        if (hmv::run_back(h, np, h->d_j + static_cast<size_t>(p0) * 21 * 3, ks)) return 1;
    }
    if (heatmap) HMV_CUDA(cudaMemcpyAsync(heatmap, h->d_hm, nimg * 21 * h->hm * h->hm * sizeof(float), cudaMemcpyDeviceToHost, ks));
    if (joints_crop_img) HMV_CUDA(cudaMemcpyAsync(joints_crop_img, h->d_xy, nimg * 21 * 2 * sizeof(float), cudaMemcpyDeviceToHost, ks));
    if (joints_cam) HMV_CUDA(cudaMemcpyAsync(joints_cam, h->d_j, static_cast<size_t>(batch) * 21 * 3 * sizeof(float), cudaMemcpyDeviceToHost, ks));
    return hmv::ws_release(h, ks);
}

static int host_call_checks(hmv_handle* h, const float* x, const float* bbox, const float* intr, int batch, const char* who) {
    HMV_CHECK(h && h->prepared, std::string(who) + ": handle not prepared");
    HMV_CHECK(batch >= 0, "negative batch");
    HMV_CHECK(x || batch == 0, std::string(who) + ": x is null");
    HMV_CHECK(!h->cfg.use_crop || (bbox && intr) || batch == 0, "'crop' positional encoding needs bbox and cam_params[\"intrinsic\"]");
    return hmv::check_flag(h);
}

int hmv_forward_host(hmv_handle* h, const float* x, const float* bbox, const float* intr, int32_t batch, float* heatmap,
                     float* joints_crop_img, float* joints_cam) {
    if (host_call_checks(h, x, bbox, intr, batch, "hmv_forward_host")) return 1;
    if (batch == 0) return 0;
    hmv::DeviceGuard guard(h->cfg.device);
    HMV_CHECK(h->tickets_issued == h->tickets_waited, "hmv_forward_host: asynchronous calls are still in flight (hmv_host_wait them first)");
    if (ensure_host_pipeline(h, batch)) return 1;
    if (enqueue_host(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, /*ramp=*/true)) return 1;
    HMV_CUDA(cudaStreamSynchronize(h->compute_stream));
    HMV_CUDA(cudaStreamSynchronize(h->copy_stream));
    return hmv::check_flag(h);
}

int hmv_forward_host_async(hmv_handle* h, const float* x, const float* bbox, const float* intr, int32_t batch, float* heatmap,
                           float* joints_crop_img, float* joints_cam, int64_t* ticket) {
    if (host_call_checks(h, x, bbox, intr, batch, "hmv_forward_host_async")) return 1;
    HMV_CHECK(ticket, "hmv_forward_host_async: ticket is null");
    HMV_CHECK(batch > 0, "hmv_forward_host_async: empty batch");
    hmv::DeviceGuard guard(h->cfg.device);
    while (h->tickets_issued - h->tickets_waited >= hmv_handle::kMaxInflight) {      // ring full: block on the oldest call
        HMV_CUDA(cudaEventSynchronize(h->ev_done[h->tickets_waited % hmv_handle::kMaxInflight]));
        ++h->tickets_waited;
    }
    if (batch > h->host_cap && h->tickets_issued != h->tickets_waited) {      // growing the output buffers frees the old ones
        hmv::set_error("hmv_forward_host_async: batch larger than any earlier call while calls are in flight");
        return 1;
    }
    if (ensure_host_pipeline(h, batch)) return 1;
    if (enqueue_host(h, x, bbox, intr, batch, heatmap, joints_crop_img, joints_cam, /*ramp=*/h->tickets_issued == h->tickets_waited)) return 1;
    const int slot = static_cast<int>(h->tickets_issued % hmv_handle::kMaxInflight);
    if (!h->ev_done[slot]) HMV_CUDA(cudaEventCreateWithFlags(&h->ev_done[slot], cudaEventDisableTiming));
    HMV_CUDA(cudaEventRecord(h->ev_done[slot], h->compute_stream));
    *ticket = h->tickets_issued++;
    return 0;
}

int hmv_host_wait(hmv_handle* h, int64_t ticket) {
    HMV_CHECK(h && h->prepared, "hmv_host_wait: handle not prepared");
    HMV_CHECK(ticket >= 0 && ticket < h->tickets_issued, "hmv_host_wait: unknown ticket");
    if (ticket < h->tickets_waited) return hmv::check_flag(h);      // tickets complete in issue order: this one already has
    hmv::DeviceGuard guard(h->cfg.device);
    HMV_CUDA(cudaEventSynchronize(h->ev_done[ticket % hmv_handle::kMaxInflight]));   // the stream is in-order: earlier tickets are done too
    h->tickets_waited = ticket + 1;
    return hmv::check_flag(h);
}

int hmv_set_input_norm(hmv_handle* h, const float* mean3, const float* std3) {
    HMV_CHECK(h && mean3 && std3, "hmv_set_input_norm: null argument");
    for (int c = 0; c < 3; ++c) {
        HMV_CHECK(std3[c] > 0.f, "hmv_set_input_norm: std must be positive");
        h->norm.mean[c] = mean3[c]; h->norm.std[c] = std3[c];
    }
    return 0;
}

int hmv_preprocess(hmv_handle* h, const uint8_t* frames, const int32_t* bbox, int32_t n_img, int32_t frame_h, int32_t frame_w,
                   float* x_out, void* stream) {
    HMV_CHECK(h && h->prepared, "hmv_preprocess: handle not prepared");
    HMV_CHECK(n_img >= 0 && (n_img == 0 || (frames && bbox && x_out)), "hmv_preprocess: bad argument");
    if (hmv::check_flag(h)) return 1;
    ++h->launches;
    return hmv::preprocess_launch(frames, bbox, x_out, n_img, frame_h, frame_w, h->img, h->norm, h->err_flag_dev,
                                  static_cast<cudaStream_t>(stream));
}

int hmv_forward_u8(hmv_handle* h, const uint8_t* x, const float* bbox, const float* intr, int32_t batch, float* heatmap,
                   float* joints_crop_img, float* joints_cam, void* stream) {
    HMV_CHECK(h && h->prepared, "hmv_forward_u8: handle not prepared");
    HMV_CHECK(batch >= 0, "negative batch");
    HMV_CHECK(x || batch == 0, "hmv_forward_u8: x is null");
    HMV_CHECK(!h->cfg.use_crop || (bbox && intr) || batch == 0, "'crop' positional encoding needs bbox and cam_params[\"intrinsic\"]");
    if (hmv::check_flag(h)) return 1;
    if (batch == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (hmv::ws_acquire(h, s)) return 1;
    h->x_u8 = true;
    const int rc = forward_eager(h, reinterpret_cast<const float*>(x), bbox, intr, batch, heatmap, joints_crop_img, joints_cam, s);
    h->x_u8 = false;
    if (rc) return rc;
    return hmv::ws_release(h, s);
}

int hmv_forward_host_u8_async(hmv_handle* h, const uint8_t* x, const float* bbox, const float* intr, int32_t batch, float* heatmap,
                              float* joints_crop_img, float* joints_cam, int64_t* ticket) {
    HMV_CHECK(h, "hmv_forward_host_u8_async: null handle");
    h->x_u8 = true;
    const int rc = hmv_forward_host_async(h, reinterpret_cast<const float*>(x), bbox, intr, batch, heatmap, joints_crop_img, joints_cam, ticket);
    h->x_u8 = false;
    return rc;
}

int hmv_stage_run(hmv_handle* h, int32_t stage, const float* x, const float* bbox, const float* intr, int32_t batch,
                  void* stream) {
    HMV_CHECK(h && h->prepared, "hmv_stage_run: handle not prepared");
    HMV_CHECK(batch >= 0 && batch <= h->mb, "hmv_stage_run: batch must be <= micro_batch");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_img = batch * h->V;
    if (hmv::ws_acquire(h, s)) return 1;
    int rc = 1;
    switch (stage) {
        case HMV_STAGE_BACKBONE:
            HMV_CHECK(x, "backbone stage needs x");
            rc = hmv::run_backbone(h, x, n_img, -1, s);
            break;
        case HMV_STAGE_POSE:
            rc = hmv::run_pose(h, n_img, nullptr, nullptr, s);
            break;
        case HMV_STAGE_SAMPLE:
            HMV_CHECK(!h->cfg.use_crop || (bbox && intr), "'crop' positional encoding needs bbox and intrinsics");
            rc = h->bf16 ? hmv::run_sample_t<hmv::bf16>(h, n_img, bbox, intr, s) : hmv::run_sample_t<float>(h, n_img, bbox, intr, s);
            break;
        case HMV_STAGE_FUSION:
            rc = h->bf16 ? hmv::run_fusion_t<hmv::bf16>(h, batch, s) : hmv::run_fusion_t<float>(h, batch, s);
            break;
        case HMV_STAGE_GCN:
            rc = hmv::run_gcn(h, batch, nullptr, s);
            break;
        case HMV_STAGE_SOFTARGMAX:
            ++h->launches;
            rc = hmv::softargmax_launch(h->hm_int, h->xy, h->xy_scaled, n_img * hmv::kJoints, h->hm, h->hm, 1000.f,
                                        static_cast<float>(h->img) / static_cast<float>(h->hm), s);
            break;
        default:
            hmv::set_error("unknown stage");
            return 1;
    }
    if (rc) return rc;
    return hmv::ws_release(h, s);
}

int hmv_tensor_get(hmv_handle* h, int32_t tensor, float* dst, int32_t batch, void* stream) {
    HMV_CHECK(h && h->prepared && dst, "hmv_tensor_get: bad argument");
    HMV_CHECK(batch >= 0 && batch <= h->mb, "hmv_tensor_get: batch must be <= micro_batch");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_img = batch * h->V, hw = h->hm * h->hm;
    if (batch == 0) return 0;
    switch (tensor) {
        case HMV_T_FEAT: case HMV_T_FEAT1: case HMV_T_FEAT2: case HMV_T_FEAT3: {
            const int l = tensor == HMV_T_FEAT ? 0 : tensor - HMV_T_FEAT1 + 1;
            HMV_CHECK(h->hr || l == 0, "feature levels 1..3 exist for the HRNet backbone only");
            const int C = h->hr ? h->hr_c[l] : 1024, ld = h->hr ? h->hr_cp[l] : 1024, px = h->hr ? (64 >> l) * (64 >> l) : hw;
            const void* src = h->hr ? h->hr_lvl[l] : h->featbuf;
            const size_t total = static_cast<size_t>(n_img) * C * px;
            if (h->bf16) hmv::nhwc_to_nchw_kernel<hmv::bf16><<<hmv::nblk(total), 256, 0, s>>>(static_cast<const hmv::bf16*>(src), dst, C, px, total, ld);
            else hmv::nhwc_to_nchw_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(static_cast<const float*>(src), dst, C, px, total, ld);
            break;
        }
        case HMV_T_HEATMAP:
            HMV_CUDA(cudaMemcpyAsync(dst, h->hm_int, static_cast<size_t>(n_img) * 21 * hw * 4, cudaMemcpyDeviceToDevice, s));
            break;
        case HMV_T_XY:
            HMV_CUDA(cudaMemcpyAsync(dst, h->xy, static_cast<size_t>(n_img) * 21 * 2 * 4, cudaMemcpyDeviceToDevice, s));
            break;
        case HMV_T_TOKENS: {
            const size_t total = static_cast<size_t>(batch) * h->S * h->d;
            hmv::rows_export_kernel<<<hmv::nblk(total), 256, 0, s>>>(h->tok0_f32, dst, h->d, h->pitch, total);
            break;
        }
        case HMV_T_FUSED: {
            const size_t total = static_cast<size_t>(batch) * 21 * h->d;
            hmv::rows_export_kernel<<<hmv::nblk(total), 256, 0, s>>>(h->fused_f32, dst, h->d, h->pitch, total);
            break;
        }
        case HMV_T_JOINTS:
            HMV_CUDA(cudaMemcpyAsync(dst, h->joints_int, static_cast<size_t>(batch) * 63 * 4, cudaMemcpyDeviceToDevice, s));
            break;
        default:
            hmv::set_error("unknown tensor id");
            return 1;
    }
    HMV_CUDA(cudaGetLastError());
    return 0;
}

int hmv_tensor_set(hmv_handle* h, int32_t tensor, const float* src, int32_t batch, void* stream) {
    HMV_CHECK(h && h->prepared && src, "hmv_tensor_set: bad argument");
    HMV_CHECK(batch >= 0 && batch <= h->mb, "hmv_tensor_set: batch must be <= micro_batch");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_img = batch * h->V, hw = h->hm * h->hm;
    if (batch == 0) return 0;
    switch (tensor) {
        case HMV_T_FEAT: case HMV_T_FEAT1: case HMV_T_FEAT2: case HMV_T_FEAT3: {
            const int l = tensor == HMV_T_FEAT ? 0 : tensor - HMV_T_FEAT1 + 1;
            HMV_CHECK(h->hr || l == 0, "feature levels 1..3 exist for the HRNet backbone only");
            const int C = h->hr ? h->hr_c[l] : 1024, ld = h->hr ? h->hr_cp[l] : 1024, px = h->hr ? (64 >> l) * (64 >> l) : hw;
            void* dstb = h->hr ? h->hr_lvl[l] : h->featbuf;
            const size_t total = static_cast<size_t>(n_img) * ld * px;
            if (h->bf16) hmv::nchw_to_nhwc_kernel<hmv::bf16><<<hmv::nblk(total), 256, 0, s>>>(src, static_cast<hmv::bf16*>(dstb), C, px, total, ld);
            else hmv::nchw_to_nhwc_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(src, static_cast<float*>(dstb), C, px, total, ld);
            break;
        }
        case HMV_T_HEATMAP:
            HMV_CUDA(cudaMemcpyAsync(h->hm_int, src, static_cast<size_t>(n_img) * 21 * hw * 4, cudaMemcpyDeviceToDevice, s));
            break;
        case HMV_T_XY:
            HMV_CUDA(cudaMemcpyAsync(h->xy, src, static_cast<size_t>(n_img) * 21 * 2 * 4, cudaMemcpyDeviceToDevice, s));
            break;
        case HMV_T_TOKENS: {
            const size_t total = static_cast<size_t>(batch) * h->S * h->d;
            if (h->bf16) hmv::rows_import_kernel<hmv::bf16><<<hmv::nblk(total), 256, 0, s>>>(src, h->tok0_f32, static_cast<hmv::bf16*>(h->tok0_lp), h->d, h->pitch, total);
            else hmv::rows_import_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(src, h->tok0_f32, static_cast<float*>(h->tok0_lp), h->d, h->pitch, total);
            break;
        }
        case HMV_T_FUSED: {
            const size_t total = static_cast<size_t>(batch) * 21 * h->d;
            hmv::rows_import_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(src, h->fused_f32, static_cast<float*>(nullptr), h->d, h->pitch, total);
            break;
        }
        default:
            hmv::set_error("tensor id cannot be set");
            return 1;
    }
    HMV_CUDA(cudaGetLastError());
    return 0;
}

int hmv_debug_num_steps(hmv_handle* h) { return h ? static_cast<int>(h->backbone.size()) : 0; }
const char* hmv_debug_step_name(hmv_handle* h, int32_t step) {
    if (!h || step < 0 || step >= static_cast<int>(h->backbone.size())) return "";
    return h->backbone[step].name.c_str();
}

int hmv_debug_backbone(hmv_handle* h, const float* x, int32_t n_img, int32_t num_steps, float* out, int32_t* chw,
                       void* stream) {
    HMV_CHECK(h && h->prepared && x && out && chw, "hmv_debug_backbone: bad argument");
    HMV_CHECK(n_img >= 1 && n_img <= h->mb_img, "hmv_debug_backbone: n_img exceeds the workspace");
    HMV_CHECK(num_steps >= 1 && num_steps <= static_cast<int>(h->backbone.size()), "hmv_debug_backbone: bad step count");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (hmv::run_backbone(h, x, n_img, num_steps, s)) return 1;
    const hmv::Step& st = h->backbone[num_steps - 1];
    chw[0] = st.C; chw[1] = st.H; chw[2] = st.W;
    const size_t total = static_cast<size_t>(n_img) * st.C * st.H * st.W;
    const int ld = st.outs.empty() ? 0 : st.outs[0].ld;
    if (h->bf16) hmv::nhwc_to_nchw_kernel<hmv::bf16><<<hmv::nblk(total), 256, 0, s>>>(static_cast<const hmv::bf16*>(st.out), out, st.C, st.H * st.W, total, ld);
    else hmv::nhwc_to_nchw_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(static_cast<const float*>(st.out), out, st.C, st.H * st.W, total, ld);
    HMV_CUDA(cudaGetLastError());
    return 0;
}

int hmv_debug_step_io(hmv_handle* h, int32_t step, char* buf, int32_t buflen) {
    HMV_CHECK(h && h->prepared && buf && buflen > 0, "hmv_debug_step_io: bad argument");
    HMV_CHECK(step >= 0 && step < static_cast<int>(h->backbone.size()), "hmv_debug_step_io: bad step");
    const hmv::Step& st = h->backbone[step];
    std::string d;
    for (const auto& io : st.ins) d += "in " + io.tap + " " + std::to_string(io.C) + " " + std::to_string(io.H) + " " + std::to_string(io.W) + ";";
    for (const auto& io : st.outs) d += "out " + io.tap + " " + std::to_string(io.C) + " " + std::to_string(io.H) + " " + std::to_string(io.W) + ";";
    HMV_CHECK(static_cast<int>(d.size()) < buflen, "hmv_debug_step_io: buffer too small");
    memcpy(buf, d.c_str(), d.size() + 1);
    return 0;
}

int hmv_debug_step_run(hmv_handle* h, int32_t step, int32_t n_img, const float* const* inputs, float* const* outputs,
                       void* stream) {
    HMV_CHECK(h && h->prepared && inputs && outputs, "hmv_debug_step_run: bad argument");
    HMV_CHECK(step >= 0 && step < static_cast<int>(h->backbone.size()), "hmv_debug_step_run: bad step");
    HMV_CHECK(n_img >= 1 && n_img <= h->mb_img, "hmv_debug_step_run: n_img exceeds the workspace");
    hmv::Step& st = h->backbone[step];
    HMV_CHECK(!st.ins.empty() && !st.outs.empty(), "hmv_debug_step_run: this step reads the network input (run it with hmv_debug_backbone)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (hmv::ws_acquire(h, s)) return 1;
    for (size_t i = 0; i < st.ins.size(); ++i) {
        const hmv::StepIO& io = st.ins[i];
        HMV_CHECK(inputs[i], "hmv_debug_step_run: null input");
        const int ld = io.ld ? io.ld : io.C;
        const size_t total = static_cast<size_t>(n_img) * ld * io.H * io.W;
        if (h->bf16) hmv::nchw_to_nhwc_kernel<hmv::bf16><<<hmv::nblk(total), 256, 0, s>>>(inputs[i], static_cast<hmv::bf16*>(io.ptr), io.C, io.H * io.W, total, ld);
        else hmv::nchw_to_nhwc_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(inputs[i], static_cast<float*>(io.ptr), io.C, io.H * io.W, total, ld);
    }
    HMV_CUDA(cudaGetLastError());
    int rc = 1;
    if (st.kind == hmv::SK_TAIL) rc = hmv::run_tail(h, st.layer, n_img, s);
    else if (st.kind == hmv::SK_SEAM) rc = hmv::run_seam(h, st.layer, n_img, s);
    else if (st.kind == hmv::SK_GEMM) rc = hmv::run_layer(h, h->layers[st.layer], n_img, s);
    else if (st.kind == hmv::SK_MAXPOOL) {
        ++h->launches;
        rc = h->bf16 ? hmv::maxpool_launch<hmv::bf16>(static_cast<const hmv::bf16*>(st.in), static_cast<hmv::bf16*>(st.out), n_img, st.H * 2, st.W * 2, st.C, s)
                     : hmv::maxpool_launch<float>(static_cast<const float*>(st.in), static_cast<float*>(st.out), n_img, st.H * 2, st.W * 2, st.C, s);
    } else if (st.kind == hmv::SK_FUSE_SUM) {
        ++h->launches;
        hmv::FuseSumParams fp = st.fuse;
        fp.n_img = n_img;
        rc = h->bf16 ? hmv::fuse_sum_launch<hmv::bf16>(fp, s) : hmv::fuse_sum_launch<float>(fp, s);
    } else {
        hmv::set_error("hmv_debug_step_run: unsupported step kind");
    }
    if (rc) return rc;
    for (size_t i = 0; i < st.outs.size(); ++i) {
        const hmv::StepIO& io = st.outs[i];
        HMV_CHECK(outputs[i], "hmv_debug_step_run: null output");
        const size_t total = static_cast<size_t>(n_img) * io.C * io.H * io.W;
        if (h->bf16) hmv::nhwc_to_nchw_kernel<hmv::bf16><<<hmv::nblk(total), 256, 0, s>>>(static_cast<const hmv::bf16*>(io.ptr), outputs[i], io.C, io.H * io.W, total, io.ld);
        else hmv::nhwc_to_nchw_kernel<float><<<hmv::nblk(total), 256, 0, s>>>(static_cast<const float*>(io.ptr), outputs[i], io.C, io.H * io.W, total, io.ld);
    }
    HMV_CUDA(cudaGetLastError());
    return hmv::ws_release(h, s);
}

int hmv_conv_bn_act(int32_t precision, const float* in, const float* w, const float* scale, const float* shift,
                    const float* residual, float* out, int32_t n_img, int32_t cin, int32_t hin, int32_t win,
                    int32_t cout, int32_t ksize, int32_t stride, int32_t relu, float* elapsed_ms, int32_t iters,
                    void* stream) {
    HMV_CHECK(in && w && out && n_img > 0, "hmv_conv_bn_act: bad argument");
    HMV_CHECK(precision == HMV_PRECISION_BF16 || precision == HMV_PRECISION_FP32, "unknown precision");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    hmv_handle tmp;
    hmv_handle* h = &tmp;
    h->bf16 = precision == HMV_PRECISION_BF16;
    h->esz = h->bf16 ? 2 : 4;
    h->mb_img = n_img;
    int dev = 0;
    HMV_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    HMV_CUDA(cudaGetDeviceProperties(&prop, dev));
    HMV_CHECK(prop.major == 10, "handmvnet_b200 is built for sm_100a (Blackwell B200) only");
    h->num_sms = prop.multiProcessorCount;
    if (h->bf16 && hmv::tc_init()) return 1;
    int rc = 1;
    int* flag_host = nullptr;
    do {
        if (cudaHostAlloc(reinterpret_cast<void**>(&flag_host), sizeof(int), cudaHostAllocMapped) != cudaSuccess) { hmv::set_error("flag alloc failed"); break; }
        *flag_host = 0;
        h->err_flag_host = flag_host;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->err_flag_dev), flag_host, 0) != cudaSuccess) { hmv::set_error("flag map failed"); break; }
        const int hout = hin / stride, wout = win / stride;
        const size_t in_elems = static_cast<size_t>(n_img) * cin * hin * win, out_elems = static_cast<size_t>(n_img) * cout * hout * wout;
        const int kk = ksize * ksize;
        // host copies of the small tensors (weights / scale / shift) -> folded layout
        std::vector<float> wh(static_cast<size_t>(cout) * cin * kk), sc(cout, 1.f), sh(cout, 0.f);
        if (cudaMemcpy(wh.data(), w, wh.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { hmv::set_error("weight copy failed"); break; }
        if (scale && cudaMemcpy(sc.data(), scale, cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { hmv::set_error("scale copy failed"); break; }
        if (shift && cudaMemcpy(sh.data(), shift, cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { hmv::set_error("shift copy failed"); break; }
        std::vector<float> wf(wh.size());
        for (int co = 0; co < cout; ++co)
            for (int ci = 0; ci < cin; ++ci)
                for (int t = 0; t < kk; ++t)
                    wf[(static_cast<size_t>(co) * kk + t) * cin + ci] = wh[(static_cast<size_t>(co) * cin + ci) * kk + t] * sc[co];
        void *din = nullptr, *dout = nullptr, *dres = nullptr;
        if (hmv::dev_alloc(h, &din, in_elems * h->esz) || hmv::dev_alloc(h, &dout, out_elems * h->esz)) break;
        if (residual && hmv::dev_alloc(h, &dres, out_elems * h->esz)) break;
        if (h->bf16) {
            hmv::nchw_to_nhwc_kernel<hmv::bf16><<<hmv::nblk(in_elems), 256, 0, s>>>(in, static_cast<hmv::bf16*>(din), cin, hin * win, in_elems);
            if (residual) hmv::nchw_to_nhwc_kernel<hmv::bf16><<<hmv::nblk(out_elems), 256, 0, s>>>(residual, static_cast<hmv::bf16*>(dres), cout, hout * wout, out_elems);
        } else {
            hmv::nchw_to_nhwc_kernel<float><<<hmv::nblk(in_elems), 256, 0, s>>>(in, static_cast<float*>(din), cin, hin * win, in_elems);
            if (residual) hmv::nchw_to_nhwc_kernel<float><<<hmv::nblk(out_elems), 256, 0, s>>>(residual, static_cast<float*>(dres), cout, hout * wout, out_elems);
        }
        int idx = -1;
        const bool direct_nchw = cout % 16 != 0;       // narrow heads (pose_net.3): fp32 NCHW epilogue straight into `out`
        if (direct_nchw && residual) { hmv::set_error("hmv_conv_bn_act: residual needs cout % 16 == 0"); break; }
        if (direct_nchw) {
            if (hmv::add_conv_raw(h, "unit_conv", wf, sh, cin, cout, ksize, stride, hin, win, din, out, relu ? hmv::ACT_RELU : hmv::ACT_NONE, nullptr, &idx, /*nchw_hw=*/hout * wout)) break;
        } else {
            if (hmv::add_conv_raw(h, "unit_conv", wf, sh, cin, cout, ksize, stride, hin, win, din, dout, relu ? hmv::ACT_RELU : hmv::ACT_NONE, dres, &idx)) break;
        }
        if (hmv::run_layer(h, h->layers[idx], n_img, s)) break;
        if (elapsed_ms && iters > 0) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, s);
            bool ok = true;
            for (int i = 0; i < iters && ok; ++i) ok = hmv::run_layer(h, h->layers[idx], n_img, s) == 0;
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            *elapsed_ms = ms / iters;
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            if (!ok) break;
        }
        if (direct_nchw) { /* already written in the caller's layout */ }
        else if (h->bf16) hmv::nhwc_to_nchw_kernel<hmv::bf16><<<hmv::nblk(out_elems), 256, 0, s>>>(static_cast<const hmv::bf16*>(dout), out, cout, hout * wout, out_elems);
        else hmv::nhwc_to_nchw_kernel<float><<<hmv::nblk(out_elems), 256, 0, s>>>(static_cast<const float*>(dout), out, cout, hout * wout, out_elems);
        if (cudaStreamSynchronize(s) != cudaSuccess) { hmv::set_error(std::string("hmv_conv_bn_act: ") + cudaGetErrorString(cudaGetLastError())); break; }
        if (hmv::check_flag(h)) break;
        rc = 0;
    } while (0);
    cudaStreamSynchronize(s);
    for (void* p : h->allocs) cudaFree(p);
    if (flag_host) cudaFreeHost(flag_host);
    return rc;
}

int hmv_profile_enable(hmv_handle* h, int32_t enable) {
    HMV_CHECK(h, "null handle");
    h->profiling = enable != 0;
    return 0;
}

/* Sums the recorded tensor-core GEMM launches since the last read: device time (ms), algorithmic FLOPs
 * (2*M*N*K over the real, unpadded extents) and launch count; optionally appends one CSV line per launch. */
int hmv_profile_read(hmv_handle* h, double* tc_ms, double* tc_flops, int64_t* tc_launches, const char* csv_path) {
    HMV_CHECK(h, "null handle");
    HMV_CUDA(cudaDeviceSynchronize());
    double ms = 0.0, fl = 0.0;
    FILE* f = csv_path ? fopen(csv_path, "w") : nullptr;
    if (f) fprintf(f, "layer,M,N,K_real,bn,ms,gflop,tflops,mbytes,mmas,mma_cycles\n");
    for (auto& r : h->prof) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        const hmv::Layer& L = h->layers[r.layer];
        const double M = static_cast<double>(r.units) * L.rows_per_unit();
        const int cin_alg = L.cin_real > 0 ? L.cin_real : L.cin;       // algorithmic work: the layer's own channels, not the padded buffer's
        const double kreal = L.kind == hmv::LK_STEM ? 147.0 : (L.kind == hmv::LK_FLAT ? static_cast<double>(cin_alg) : static_cast<double>(cin_alg) * L.ksize * L.ksize);
        double flop = 2.0 * M * L.cout * kreal;
        // algorithmic HBM bytes of the launch: every input / residual / output element and every weight once
        auto out_bytes = [&](const hmv::Layer& X, double rows) { return rows * X.cout * (X.ep.out_mode == hmv::OUT_BF16_ROWMAJOR ? 2.0 : 4.0); };
        auto res_bytes = [&](const hmv::Layer& X, double rows) { return X.ep.res_mode == hmv::RES_NONE ? 0.0 : rows * X.cout * (X.ep.res_mode == hmv::RES_BF16 ? 2.0 : 4.0); };
        const double in_rows = L.kind == hmv::LK_FLAT || L.ksize == 1 ? M : static_cast<double>(r.units) * L.hin * L.win;
        double bytes = in_rows * cin_alg * 2.0 + static_cast<double>(L.cout) * kreal * 2.0;
        std::string name = L.name;
        int ncol = L.cout;
        double kcol = kreal;
        // tcgen05.mma instructions (128 rows per SM, K = 16) the launch issues and the tensor-pipe cycles they take at the
        // measured rate: max(N / 2, 48) cycles for an N-column MMA (tools/mma_issue_bench.cu: 48 / 64 / 96 / 128 cycles for
        // N = 64 / 128 / 192 / 256, one CTA or a cta_group::2 pair alike) - the shape-limited tensor floor of the launch
        auto mma_cyc = [](double n) { return n / 2.0 > 48.0 ? n / 2.0 : 48.0; };
        const double m_tiles = L.kind == hmv::LK_FLAT ? ceil(M / 128.0) : static_cast<double>(r.units) * L.tc.p.tpi;
        double mmas = m_tiles * (L.n_alloc / (L.bn > 0 ? L.bn : 1)) * (L.K / 16.0);
        double mma_cycles = mmas * mma_cyc(L.bn);
        if (r.seam >= 0) {
            mmas = m_tiles * (8.0 * 16.0 + 64.0);     // 8 conv3 chunks x K 256 (N = 128) + next conv1 K 1024 (N = 256)
            mma_cycles = m_tiles * (8.0 * 16.0 * mma_cyc(128) + 64.0 * mma_cyc(256));
            // fused conv3(b) + conv1(b+1): conv2 output in, residual in, block output + next conv1 output out
            const hmv::Layer& L1 = h->layers[h->seams[r.seam].l1];
            flop += 2.0 * M * L1.cout * L1.cin;
            bytes += out_bytes(L, M) + res_bytes(L, M) + out_bytes(L1, M) + static_cast<double>(L1.cout) * L1.K * 2.0;
            name = L.name + "+next.conv1";
        } else if (r.tail >= 0) {                     // fused conv2 + conv3: both GEMMs' work, N / K columns of conv3 (the conv2 output stays on chip / in L2)
            const hmv::Layer& L3 = h->layers[h->tails[r.tail].l3];
            const int l_ds = h->tails[r.tail].l_ds;
            const double k_ds = l_ds >= 0 ? h->layers[l_ds].K : 0.0;          // folded downsample: extra K of conv3, block input instead of a residual
            mmas = m_tiles * (L.K / 16.0 + (L3.cout / 128.0) * ((L3.K + k_ds) / 16.0));
            mma_cycles = m_tiles * (L.K / 16.0 * mma_cyc(L.cout) + (L3.cout / 128.0) * ((L3.K + k_ds) / 16.0) * mma_cyc(128));
            flop += 2.0 * M * L3.cout * (L3.cin + k_ds);
            bytes += out_bytes(L3, M) + static_cast<double>(L3.cout) * (L3.K + k_ds) * 2.0 + (l_ds >= 0 ? M * k_ds * 2.0 : res_bytes(L3, M));
            name = L.name + (l_ds >= 0 ? "+conv3+downsample" : "+conv3");
            ncol = L3.cout; kcol = kreal + L3.cin + k_ds;
        } else {
            bytes += out_bytes(L, M) + res_bytes(L, M);
        }
        ms += t; fl += flop;
        if (f) fprintf(f, "%s,%.0f,%d,%.0f,%d,%.6f,%.4f,%.2f,%.3f,%.0f,%.0f\n", name.c_str(), M, ncol, kcol, L.bn, t, flop * 1e-9, flop / (t * 1e-3) * 1e-12, bytes * 1e-6, mmas, mma_cycles);
        h->ev_pool.push_back(r.e0); h->ev_pool.push_back(r.e1);
    }
    if (f) fclose(f);
    if (tc_ms) *tc_ms = ms;
    if (tc_flops) *tc_flops = fl;
    if (tc_launches) *tc_launches = static_cast<int64_t>(h->prof.size());
    h->prof.clear();
    return 0;
}

/* Stream time (ms, including launch gaps) the profiled passes spent in: [0] backbone, [1] pose_net + soft-argmax +
 * sampling + token assembly, [2] fusion transformer, [3] graph head.  Resets the marks. */
int hmv_profile_phases(hmv_handle* h, double* out4) {
    HMV_CHECK(h && out4, "null argument");
    HMV_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 4; ++i) out4[i] = 0.0;
    for (size_t i = 0; i + 1 < h->phase_marks.size(); ++i) {
        const int a = h->phase_marks[i].first, b = h->phase_marks[i + 1].first;
        int slot = -1;
        if (a == 0 && b == 1) slot = 0;
        else if (a == 1 && b == 2) slot = 1;
        else if (a == 3 && b == 4) slot = 2;
        else if (a == 4 && b == 5) slot = 3;
        if (slot < 0) continue;
        float t = 0.f;
        cudaEventElapsedTime(&t, h->phase_marks[i].second, h->phase_marks[i + 1].second);
        out4[slot] += t;
    }
    for (auto& m : h->phase_marks) h->ev_pool.push_back(m.second);
    h->phase_marks.clear();
    return 0;
}

int64_t hmv_launch_count(hmv_handle* h) { return h ? h->launches : 0; }
int hmv_num_sms(hmv_handle* h) { return h ? h->num_sms : 0; }

}  // extern "C"
