// Fused bottleneck seam for layer3 (P = 256): conv3 1x1 (+BN) + residual + ReLU of block b  ->  conv1 1x1 (+BN) + ReLU
// of block b+1, in ONE persistent tcgen05 kernel (reference backbones/resnet.py:135-143 followed by :127-129 of the
// next Bottleneck).
//
// Why: both layers are HBM-bound as separate kernels (conv3 writes the 1024-channel block output and re-reads the
// 1024-channel residual, conv1 reads the block output again).  Here the block output tile is produced in shared
// memory, TMA-stored once, and consumed from shared memory as the A operand of the next conv1 - the 671 MB re-read of a
// B=64 pass disappears and conv1's tensor work hides under conv3's residual / store traffic.
//
//   per M tile (128 pixels), MMA issue order (one thread):
//     T3(0) T3(1) T1(0) T3(2) T1(1) ... T3(7) T1(6) T1(7)
//     T3(c): acc3[c & 1][128 x 128] = Y2[m] * W3[c]^T            (K = 256 = 4 K blocks: A and W3 chunk through the ring)
//     T1(c): acc1[128 x 256]      += OUT[m, chunk c] * W1[:, chunk c]^T   (2 K blocks; A = the two staging slots of
//                                                                   chunk c, B = W1 K block through the ring)
//   epilogue (8 warps), in the same order: E3(c) = TMEM -> +bias3 +residual -> ReLU -> bf16 IN PLACE in the slot the
//   residual was prefetched into (128B swizzle = the canonical K-major A operand layout) -> TMA store of the block output
//   + "A ready" to the MMA issuer;  E1 = acc1 -> +bias1 -> ReLU -> bf16 -> TMA store of the next block's conv1 output.
//   A slot is recycled when the TMA store has read it AND the T1 MMAs that use it have retired.
//   TMEM: acc1 at columns [0, 256), acc3 double-buffered at 256 / 384.   Shared memory: 5 operand stages of 32 KiB
//   and 4 chunk slots of 16 KiB (measured: ring depth matters more than slot slack - 4 stages + 6 slots 0.457 ms,
//   3 stages + 8 slots with a two-chunk lag 0.50 ms, 5 stages + 4 slots 0.40 ms per B=64 launch; the two separate
//   kernels take 0.269 + 0.166 ms).
#include "conv_gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <map>
#include <utility>

namespace hmv {

namespace {

constexpr int kBnP = 256;                                       // planes: K of conv3, N of the next conv1
constexpr int kBnN3 = 4 * kBnP;                                 // block output channels
constexpr int kBnChunk = 128;                                   // conv3 output columns per T3
constexpr int kBnNch = kBnN3 / kBnChunk;                        // 8 chunks per tile
constexpr int kBnKb3 = kBnP / kTcBlockK;                        // 4 K blocks per T3
constexpr int kBnSlotCols = 64;
constexpr int kBnSlotBytes = kTcBlockM * kBnSlotCols * 2;       // 16 KiB
// Shared memory of a CTA: the tile's conv2 output Y2 (the A operand of all 8 conv3 chunks, 4 K blocks x 16 KiB, loaded once
// per tile), a ring of weight stages, and the staging slots.  A stage holds two K blocks of one conv3 chunk's weights
// (T3: 8 MMAs) or one K block of the next conv1's weights (T1: 4 MMAs) - 512 tensor cycles either way.
//   one CTA per tile:  stage 32 KiB (2 x [128 x 64] of W3 / [256 x 64] of W1), 3 stages, 4 slots
//   CTA pair (cta_group::2, PAIR): a CTA holds HALF of every weight tile: stage 16 KiB, 4 stages, 6 slots
// (Until round 2 the A operand travelled through the ring with every chunk's weights - 24-32 KiB stages, 5-6 of them and
//  only 4 slots: the slots, each used 5 times per tile with a residual prefetch, the epilogue, the next conv1's MMAs and a
//  store in every cycle, were what bounded the kernel once the MMA issue path was lean.)
#ifndef HMV_BN_SLOTS
#define HMV_BN_SLOTS 4
#endif
#ifndef HMV_BN_STAGES
#define HMV_BN_STAGES 3
#endif
#ifndef HMV_BN_PAIR_SLOTS
#define HMV_BN_PAIR_SLOTS 6
#endif
#ifndef HMV_BN_PAIR_STAGES
#define HMV_BN_PAIR_STAGES 4
#endif
#ifndef HMV_BN_LAG
#define HMV_BN_LAG 1
#endif
#ifndef HMV_BN_PAIR_LAG
#define HMV_BN_PAIR_LAG 2                                       // (4 stages + 6 slots: lag 1 0.355 ms, lag 2 0.344; 6 + 4, lag 1: 0.348)
#endif
#ifndef HMV_BN_STORE_DEPTH
#define HMV_BN_STORE_DEPTH 1                                    // slot stores the store warp keeps in flight beyond the newest
#endif
constexpr int kBnY2Bytes = kTcBlockM * kBnP * 2;                // 64 KiB: 4 K blocks of [128 rows x 64] bf16, 128B swizzle
template <bool PAIR> struct BnGeo {
    static constexpr int kHalfBytes = PAIR ? 8 * 1024 : 16 * 1024;          // one K block of a conv3 chunk's weights
    static constexpr int kStageBytes = 2 * kHalfBytes;
    static constexpr int kStages = PAIR ? HMV_BN_PAIR_STAGES : HMV_BN_STAGES;
    static constexpr int kSlots = PAIR ? HMV_BN_PAIR_SLOTS : HMV_BN_SLOTS;
    static constexpr int kLag = PAIR ? HMV_BN_PAIR_LAG : HMV_BN_LAG;      // T1(c - kLag) follows T3(c): gives the epilogue of chunk c - kLag time to finish
    static constexpr int kSmemBytes = kBnY2Bytes + kStages * kStageBytes + kSlots * kBnSlotBytes + 1024 /*align*/ + 512 /*barriers*/;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(kStages >= 2 && kSlots >= 4 && kSlots % 2 == 0, "two epilogue teams alternate over the slots");
};
constexpr int kBnSlotsPerTile = 2 * kBnNch + kBnP / kBnSlotCols; // 16 conv3 slots + 4 conv1 slots

template <int kBnLag, typename F3, typename F1>
__device__ __forceinline__ bool bn_tile_schedule(F3&& t3, F1&& t1) {
    for (int c = 0; c < kBnNch; ++c) {
        if (!t3(c)) return false;
        if (c >= kBnLag && !t1(c - kBnLag)) return false;
    }
    for (int c = kBnNch - kBnLag; c < kBnNch; ++c)
        if (!t1(c)) return false;
    return true;
}

// mbar_wait that also accumulates the stall time (clock cycles) into `acc` when PROF (HMV_BN_PROF=1 runs)
template <bool PROF>
__device__ __forceinline__ bool bn_wait(uint32_t bar, uint32_t parity, int* err_flag, int code, long long& acc) {
    if (!PROF) return mbar_wait(bar, parity, err_flag, code);
    const long long t0 = clock64();                  // (try_wait may suspend inside the instruction: time it as well)
    const bool ok = mbar_wait(bar, parity, err_flag, code);
    acc += clock64() - t0;
    return ok;
}

// (A 2-CTA cluster variant with multicast weight tiles was measured at -1 % in round 1 - the kernel is not bound by the
// L2 -> SM fill bandwidth - and has been removed.)
// PAIR: launched as clusters of two CTAs on neighbouring M tiles; the leader (rank 0) issues tcgen05.mma.cta_group::2 of
// M = 256 for both (A rows from both CTAs' shared memory at the same offsets - operand stages and the epilogue's slots alike -
// and half of each weight tile from each).  Both CTAs' TMA loads count on the leader's `full` barriers, the leader's commits
// arrive on `empty` / `t3full` / `t1full` / `sfree` in both CTAs, and the peer's epilogue warps arrive on the leader's
// `t3empty` / `t1empty` / `aready` remotely.  The weights cross L2 -> SM once per pair and 6 operand stages fit where 5 did.
template <bool PROF, bool PAIR>
__global__ void __launch_bounds__(kTcThreads, 1)
bottleneck_next_kernel(const __grid_constant__ CUtensorMap tmY2,   // conv2 output [rows, 256], load box {64, 128}
                       const __grid_constant__ CUtensorMap tmW3,   // [1024, 256], box {64, 128}  
                       const __grid_constant__ CUtensorMap tmRes,  // residual [rows, 1024], load box {64, 128}
                       const __grid_constant__ CUtensorMap tmOut,  // block output [rows, 1024], store box {64, 128}
                       const __grid_constant__ CUtensorMap tmW1,   // next conv1 weights [256, 1024], box {64, 256}  
                       const __grid_constant__ CUtensorMap tmY1,   // next conv1 output [rows, 256], store box {64, 128}
                       const __grid_constant__ BnParams p,
                       const __grid_constant__ BiasBank bank) {            // conv3 biases [0, 1024), next conv1 biases [1024, 1280)
    constexpr int kBnStages = BnGeo<PAIR>::kStages;
    constexpr int kBnStageBytes = BnGeo<PAIR>::kStageBytes;
    constexpr int kBnHalfBytes = BnGeo<PAIR>::kHalfBytes;
    constexpr int kBnSlots = BnGeo<PAIR>::kSlots;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* ring = smem + kBnY2Bytes;                        // [0, 64 KiB): the tile's Y2
    uint8_t* slots = ring + kBnStages * kBnStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(slots + kBnSlots * kBnSlotBytes);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * kBnStages;
    const uint32_t t3full0 = empty0 + 8 * kBnStages;
    const uint32_t t3empty0 = t3full0 + 16;
    const uint32_t t1full = t3empty0 + 16;
    const uint32_t t1empty = t1full + 8;
    const uint32_t y2full = t1empty + 8;                      // the tile's Y2 has landed (in both CTAs of a pair)
    const uint32_t y2empty = y2full + 8;                      // the tile's last conv3 MMA has retired: Y2 may be overwritten
    const uint32_t sres0 = y2empty + 8;                       // [slots] residual landed / slot handed to the epilogue
    const uint32_t aready0 = sres0 + 8 * kBnSlots;            // [slots] finished block-output chunk is in the slot
    const uint32_t sfree0 = aready0 + 8 * kBnSlots;           // [slots] store has read the slot and its MMAs retired
    const uint32_t stready0 = sfree0 + 8 * kBnSlots;          // [slots] this CTA's four slabs of the slot are written: the store warp may store it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBnStages + 8 + 4 * kBnSlots);
    const uint32_t y2_base = smem_u32(smem);
    const uint32_t smem_base = smem_u32(ring);                // weight ring
    const uint32_t slots_base = smem_u32(slots);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    constexpr uint32_t kCtas = PAIR ? 2u : 1u;
    // arrive on a barrier the MMA issuer waits on: in a pair that barrier lives in the leader CTA
    auto arrive_mma = [&](uint32_t bar) {
        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(bar, 0));
        else mbar_arrive(bar);
    };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmY2); prefetch_tmap(&tmW3); prefetch_tmap(&tmRes); prefetch_tmap(&tmOut); prefetch_tmap(&tmW1); prefetch_tmap(&tmY1);
        for (int i = 0; i < kBnStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(t3full0 + 8 * i, 1);
            mbar_init(t3empty0 + 8 * i, 8 * kCtas);         // one arrival per epilogue warp (of both CTAs of a pair)
        }
        mbar_init(t1full, 1);
        mbar_init(t1empty, 8 * kCtas);
        mbar_init(y2full, 1);
        mbar_init(y2empty, 1);
        for (int i = 0; i < kBnSlots; ++i) {
            mbar_init(sres0 + 8 * i, 1);
            mbar_init(aready0 + 8 * i, 4 * kCtas);           // the four slab warps of the team that handles the slot (in both CTAs of a pair)
            mbar_init(sfree0 + 8 * i, 2);                   // the store warp + the MMA commit (conv3 slots) / the slot producer (conv1 slots)
            mbar_init(stready0 + 8 * i, 4);                 // the four slab warps of the owning team (this CTA's)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer's barriers exist before anything remote lands on them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();
    pdl_launch_dependents();

    // one CTA per tile: this CTA's tiles are blockIdx.x, + gridDim.x, ...   pair: the cluster walks pairs of M tiles (the tile
    // count is even: 8 tiles per image) and this CTA takes tile 2 * pair + rank
    const int units = PAIR ? p.num_m_tiles / 2 : p.num_m_tiles;
    const int first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int n_i = (units - first + step - 1) / step;
    auto tile_of = [&](int i) { return PAIR ? 2 * (first + i * step) + static_cast<int>(crank) : first + i * step; };

    if (warp == 0) {
        // ===================== TMA producer of the operand ring (same order as the MMA issuer consumes it) =====================
        {   // the whole warp walks the schedule and waits; one elected lane issues the copies (see elect_one() in tc_ptx.cuh)
            int stage = 0;
            uint32_t phase = 0;
            long long w_e3 = 0, w_e1 = 0, w_y2e = 0;
            for (int i = 0; i < n_i; ++i) {
                const int m = tile_of(i);
                // this tile's Y2: once the previous tile's last conv3 MMA has retired
                if (!bn_wait<PROF>(y2empty, (static_cast<uint32_t>(i) & 1u) ^ 1u, p.err_flag, 63, w_y2e)) break;
                if (elect_one()) {
                    if (PAIR) {                    // both CTAs' bytes are counted on the LEADER's barrier, which only the leader arms
                        const uint32_t lfb = crank == 0 ? y2full : mapa_u32(y2full, 0);
                        if (crank == 0) mbar_arrive_expect_tx(y2full, 2u * kBnY2Bytes);
#pragma unroll
                        for (int kb = 0; kb < kBnKb3; ++kb) tma_load_2d_2sm(y2_base + kb * 16384, &tmY2, lfb, kb * kTcBlockK, m * kTcBlockM);
                    } else {
                        mbar_arrive_expect_tx(y2full, kBnY2Bytes);
#pragma unroll
                        for (int kb = 0; kb < kBnKb3; ++kb) tma_load_2d(y2_base + kb * 16384, &tmY2, y2full, kb * kTcBlockK, m * kTcBlockM);
                    }
                }
                const bool ok = bn_tile_schedule<BnGeo<PAIR>::kLag>(
                    [&](int c) {
                        for (int s2 = 0; s2 < kBnKb3 / 2; ++s2) {          // a stage = two K blocks of this chunk's weights
                            if (!bn_wait<PROF>(empty0 + 8 * stage, phase ^ 1, p.err_flag, 51, w_e3)) return false;
                            if (elect_one()) {
                                const uint32_t fb = full0 + 8 * stage;
                                const uint32_t sb = smem_base + stage * kBnStageBytes;
                                if (PAIR) {        // this CTA's 64 of the chunk's 128 weight rows; bytes counted on the leader
                                    const uint32_t lfb = crank == 0 ? fb : mapa_u32(fb, 0);
                                    if (crank == 0) mbar_arrive_expect_tx(fb, 2u * kBnStageBytes);
#pragma unroll
                                    for (int hb = 0; hb < 2; ++hb)
                                        tma_load_2d_2sm(sb + hb * kBnHalfBytes, &tmW3, lfb, (2 * s2 + hb) * kTcBlockK, c * kBnChunk + static_cast<int>(crank) * (kBnChunk / 2));
                                } else {
                                    mbar_arrive_expect_tx(fb, kBnStageBytes);
#pragma unroll
                                    for (int hb = 0; hb < 2; ++hb) tma_load_2d(sb + hb * kBnHalfBytes, &tmW3, fb, (2 * s2 + hb) * kTcBlockK, c * kBnChunk);
                                }
                            }
                            if (++stage == kBnStages) { stage = 0; phase ^= 1; }
                        }
                        return true;
                    },
                    [&](int c) {
                        for (int j = 0; j < 2; ++j) {
                            if (!bn_wait<PROF>(empty0 + 8 * stage, phase ^ 1, p.err_flag, 52, w_e1)) return false;
                            if (elect_one()) {
                                const uint32_t fb = full0 + 8 * stage;
                                if (PAIR) {        // this CTA's 128 of the 256 weight rows of the K block
                                    const uint32_t lfb = crank == 0 ? fb : mapa_u32(fb, 0);
                                    if (crank == 0) mbar_arrive_expect_tx(fb, 2u * kBnStageBytes);
                                    tma_load_2d_2sm(smem_base + stage * kBnStageBytes, &tmW1, lfb, (2 * c + j) * kTcBlockK, static_cast<int>(crank) * (kBnP / 2));
                                } else {
                                    mbar_arrive_expect_tx(fb, kBnStageBytes);
                                    tma_load_2d(smem_base + stage * kBnStageBytes, &tmW1, fb, (2 * c + j) * kTcBlockK, 0);
                                }
                            }
                            if (++stage == kBnStages) { stage = 0; phase ^= 1; }
                        }
                        return true;
                    });
                if (!ok) break;
            }
            if (PROF && p.prof && lane == 0) { p.prof[blockIdx.x * 24 + 6] = w_e3; p.prof[blockIdx.x * 24 + 7] = w_e1; p.prof[blockIdx.x * 24 + 22] = w_y2e; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (crank == 0) {                                    // pair: only the leader issues, for both CTAs
            // the whole warp walks the schedule and waits; one elected lane issues (see elect_one() in tc_ptx.cuh for why)
            constexpr uint32_t idesc3 = PAIR ? make_idesc_mn(2 * kTcBlockM, kBnChunk) : make_idesc(kBnChunk);
            constexpr uint32_t idesc1 = PAIR ? make_idesc_mn(2 * kTcBlockM, kBnP) : make_idesc(kBnP);
            auto wait_epi = [&](uint32_t bar, uint32_t parity, int code, long long& acc) {      // barriers the epilogue warps arrive on
                if (!PAIR) return bn_wait<PROF>(bar, parity, p.err_flag, code, acc);
                const long long t0 = PROF ? clock64() : 0;
                const bool ok = mbar_wait_cluster(bar, parity, p.err_flag, code);
                if (PROF) acc += clock64() - t0;
                return ok;
            };
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
                if (PAIR) umma_f16_2sm(d, a, b, idesc, accum);
                else umma_f16(d, a, b, idesc, accum);
            };
            auto commit = [&](uint32_t bar) {
                if (PAIR) umma_commit_2sm_mc(bar, static_cast<uint16_t>(3));
                else umma_commit(bar);
            };
            int stage = 0;
            uint32_t phase = 0;
            uint32_t q3 = 0;                                 // running conv3 chunk counter (TMEM slot = q3 & 1)
            long long w_t3e = 0, w_f3 = 0, w_t1e = 0, w_ar = 0, w_f1 = 0, w_y2f = 0;
            const long long t_start = clock64();
            for (int i = 0; i < n_i; ++i) {
                const uint32_t gbase = static_cast<uint32_t>(i) * kBnSlotsPerTile;
                const bool ok = bn_tile_schedule<BnGeo<PAIR>::kLag>(
                    [&](int c) {
                        const uint32_t s = q3 & 1u, use = q3 >> 1;
                        if (!wait_epi(t3empty0 + 8 * s, (use & 1u) ^ 1u, 53, w_t3e)) return false;
                        if (c == 0 && !bn_wait<PROF>(y2full, static_cast<uint32_t>(i) & 1u, p.err_flag, 64, w_y2f)) return false;
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + kBnP + s * kBnChunk;
                        for (int s2 = 0; s2 < kBnKb3 / 2; ++s2) {
                            if (!bn_wait<PROF>(full0 + 8 * stage, phase, p.err_flag, 54, w_f3)) return false;
                            tc_fence_after();
                            if (elect_one()) {
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {
                                    const uint64_t adesc = make_sw128_desc(y2_base + (2 * s2 + hb) * 16384);          // resident Y2, K block 2*s2+hb
                                    const uint64_t bdesc = make_sw128_desc(smem_base + stage * kBnStageBytes + hb * kBnHalfBytes);
#pragma unroll
                                    for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k)   // +32 bytes of K per MMA = +2 in the address field
                                        mma(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc3, (s2 | hb | k) != 0 ? 1u : 0u);
                                }
                                commit(empty0 + 8 * stage);
                                if (s2 == kBnKb3 / 2 - 1) {
                                    commit(t3full0 + 8 * s);
                                    if (c == kBnNch - 1) commit(y2empty);                  // the tile's Y2 is no longer needed (both CTAs of a pair)
                                }
                            }
                            if (++stage == kBnStages) { stage = 0; phase ^= 1; }
                        }
                        ++q3;
                        return true;
                    },
                    [&](int c) {
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t g = gbase + 2 * c + j, slot = g % kBnSlots, use = g / kBnSlots;
                            if (c == 0 && j == 0) {          // acc1 drained by the epilogue of the previous tile
                                if (!wait_epi(t1empty, (static_cast<uint32_t>(i) & 1u) ^ 1u, 55, w_t1e)) return false;
                            }
                            if (!wait_epi(aready0 + 8 * slot, use & 1u, 56, w_ar)) return false;
                            if (!bn_wait<PROF>(full0 + 8 * stage, phase, p.err_flag, 57, w_f1)) return false;
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t adesc = make_sw128_desc(slots_base + slot * kBnSlotBytes);   // finished block-output chunk = A operand
                                const uint64_t bdesc = make_sw128_desc(smem_base + stage * kBnStageBytes);
#pragma unroll
                                for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k)
                                    mma(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc1, (c | j | k) != 0 ? 1u : 0u);
                                commit(empty0 + 8 * stage);
                                commit(sfree0 + 8 * slot);                                 // the slot's MMAs have retired (in both CTAs of a pair)
                                if (c == kBnNch - 1 && j == 1) commit(t1full);
                            }
                            if (++stage == kBnStages) { stage = 0; phase ^= 1; }
                        }
                        return true;
                    });
                if (!ok) break;
            }
            if (PROF && p.prof && lane == 0) {
                long long* o = p.prof + blockIdx.x * 24;
                o[0] = clock64() - t_start; o[1] = w_t3e; o[2] = w_f3; o[3] = w_t1e; o[4] = w_ar; o[5] = w_f1; o[15] = n_i; o[21] = w_y2f;
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== slot producer: residual prefetch (conv3 slots) / plain hand-over (conv1 slots) =====================
        {   // whole warp + one elected lane per copy, as above
            uint32_t g = 0;
            bool alive = true;
            long long w_sf = 0;
            for (int i = 0; i < n_i && alive; ++i) {
                const int m = tile_of(i);
                for (int c = 0; c < kBnSlotsPerTile && alive; ++c, ++g) {
                    const uint32_t slot = g % kBnSlots, use = g / kBnSlots;
                    if (!bn_wait<PROF>(sfree0 + 8 * slot, (use & 1u) ^ 1u, p.err_flag, 58, w_sf)) { alive = false; break; }
                    if (elect_one()) {
                        if (c < 2 * kBnNch) {
                            mbar_arrive_expect_tx(sres0 + 8 * slot, kBnSlotBytes);
                            tma_load_2d(slots_base + slot * kBnSlotBytes, &tmRes, sres0 + 8 * slot, c * kBnSlotCols, m * kTcBlockM);
                        } else {
                            mbar_arrive(sres0 + 8 * slot);
                            mbar_arrive(sfree0 + 8 * slot);  // stands in for the MMA commit: conv1 slots feed no MMA
                        }
                    }
                }
            }
            if (PROF && p.prof && lane == 0) p.prof[blockIdx.x * 24 + 8] = w_sf;
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== store warp: finished slots -> global memory =====================
        // One TMA store per 16 KiB slot, issued here instead of by the epilogue warps: waiting until a store has read its
        // slot (before the slot goes back to the residual prefetcher) cost every epilogue warp ~750 cycles per slot -
        // a quarter of its time - and the epilogue is what bounds this kernel once the MMAs are issued leanly.
        {
            uint32_t g = 0;
            bool alive = true;
            long long w_st = 0, w_rd = 0;
            for (int i = 0; i < n_i && alive; ++i) {
                const int m = tile_of(i);
                for (int j = 0; j < kBnSlotsPerTile && alive; ++j, ++g) {
                    const uint32_t slot = g % kBnSlots, use = g / kBnSlots;
                    if (!bn_wait<PROF>(stready0 + 8 * slot, use & 1u, p.err_flag, 62, w_st)) { alive = false; break; }
                    if (elect_one()) {                     // (elect.sync picks the same lane every time: the bulk groups are its own)
                        if (j < 2 * kBnNch) tma_store_2d(&tmOut, slots_base + slot * kBnSlotBytes, j * kBnSlotCols, m * kTcBlockM);
                        else tma_store_2d(&tmY1, slots_base + slot * kBnSlotBytes, (j - 2 * kBnNch) * kBnSlotCols, m * kTcBlockM);
                        bulk_commit();
                        if (g >= HMV_BN_STORE_DEPTH) {     // the store issued HMV_BN_STORE_DEPTH slots ago has read its slot
                            if (PROF) { const long long t0 = clock64(); bulk_wait_read<HMV_BN_STORE_DEPTH>(); w_rd += clock64() - t0; }
                            else bulk_wait_read<HMV_BN_STORE_DEPTH>();
                            mbar_arrive(sfree0 + 8 * ((g - HMV_BN_STORE_DEPTH) % kBnSlots));
                        }
                    }
                }
            }
            if (elect_one()) bulk_wait_read<0>();            // staged data must stay valid until every store has read it
            __syncwarp();
            if (PROF && p.prof && lane == 0) { p.prof[blockIdx.x * 24 + 13] = w_st; p.prof[blockIdx.x * 24 + 14] = w_rd; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 2 teams x 4 warps; a warp owns one 32-row slab (TMEM lane quarter) of a whole 64-column slot ==========
        // Team t takes the slots with running index g = t (mod 2): the two 64-column halves of a conv3 chunk / alternate
        // conv1 slots are in flight at the same time, and a warp needs no other warp to finish its slab (no named barrier):
        // TMEM -> +bias (+residual already in the slab) -> ReLU -> bf16 in place -> its own TMA store.
        const int quarter = warp & 3;
        const int team = (warp - 4) >> 2;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t slab_off = quarter * (32 * 128) + lane * 128;
        uint32_t g = static_cast<uint32_t>(team);            // this warp's running slot index (advances by 2)
        bool alive = true;
        long long w_sres = 0, w_t3f = 0, w_t1f = 0, w_ldt = 0, w_math = 0, w_fence = 0, w_issue = 0, w_head = 0;
        const long long e_start = clock64();

        auto do_slot = [&](uint32_t tcol, int bias_off, bool is_conv3, uint32_t release_bar) {
            const uint32_t slot = g % kBnSlots, use = g / kBnSlots;
            uint8_t* srow = slots + slot * kBnSlotBytes + slab_off;
            long long tp = PROF ? clock64() : 0;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                 // two 32-column halves of the slab row
                uint32_t r[32];
                tmem_ld32(lane_base + tcol + hf * 32, r);
                float4 bq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) bq[j] = *reinterpret_cast<const float4*>(&bank.v[bias_off + hf * 32 + 4 * j]);   // constant cache
                if (hf == 0 && !bn_wait<PROF>(sres0 + 8 * slot, use & 1u, p.err_flag, 59, w_sres)) alive = false;
                uint4 rq[4];
                if (is_conv3) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rq[j] = *reinterpret_cast<const uint4*>(srow + ((static_cast<uint32_t>(hf * 4 + j) ^ (lane & 7)) << 4));
                }
                if (PROF) { const long long t0 = clock64(); w_head += t0 - tp; tmem_ld_wait(); tp = clock64(); w_ldt += tp - t0; } else tmem_ld_wait();
                if (hf == 1 && release_bar != 0) {           // this warp's part of the accumulator is read: hand TMEM back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_mma(release_bar);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 v0 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])), make_float2(bq[2 * j].x, bq[2 * j].y));
                    float2 v1 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), make_float2(bq[2 * j].z, bq[2 * j].w));
                    float2 v2 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), make_float2(bq[2 * j + 1].x, bq[2 * j + 1].y));
                    float2 v3 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])), make_float2(bq[2 * j + 1].z, bq[2 * j + 1].w));
                    if (is_conv3) {
                        v0 = __fadd2_rn(v0, bf16x2_to_f2(rq[j].x)); v1 = __fadd2_rn(v1, bf16x2_to_f2(rq[j].y));
                        v2 = __fadd2_rn(v2, bf16x2_to_f2(rq[j].z)); v3 = __fadd2_rn(v3, bf16x2_to_f2(rq[j].w));
                    }
                    uint4 o;
                    o.x = cvt_bf16x2(v0.x, v0.y, true); o.y = cvt_bf16x2(v1.x, v1.y, true);
                    o.z = cvt_bf16x2(v2.x, v2.y, true); o.w = cvt_bf16x2(v3.x, v3.y, true);
                    *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(hf * 4 + j) ^ (lane & 7)) << 4)) = o;
                }
                if (PROF) { const long long t0 = clock64(); w_math += t0 - tp; tp = t0; }
            }
            fence_async_smem();                              // generic-proxy writes -> visible to the TMA store and to the tensor core
            __syncwarp();
            if (PROF) { const long long t0 = clock64(); w_fence += t0 - tp; tp = t0; }
            if (lane == 0) {
                arrive_mma(aready0 + 8 * slot);              // every slot use (conv1 slots too: keeps the barrier's phase == use count)
                mbar_arrive(stready0 + 8 * slot);            // this warp's slab is in the slot: the store warp stores it once all four are
            }
            __syncwarp();
            if (PROF) { w_issue += clock64() - tp; }
            g += 2;
        };

        uint32_t q3 = 0;
        for (int i = 0; i < n_i && alive; ++i) {
            const int m = tile_of(i);
            for (int c = 0; c < kBnNch && alive; ++c, ++q3) {
                const uint32_t s = q3 & 1u;
                if (!bn_wait<PROF>(t3full0 + 8 * s, (q3 >> 1) & 1u, p.err_flag, 60, w_t3f)) { alive = false; break; }
                tc_fence_after();
                do_slot(kBnP + s * kBnChunk + team * kBnSlotCols, c * kBnChunk + team * kBnSlotCols, true, t3empty0 + 8 * s);
            }
            if (!alive) break;
            if (!bn_wait<PROF>(t1full, static_cast<uint32_t>(i) & 1u, p.err_flag, 61, w_t1f)) { alive = false; break; }
            tc_fence_after();
#pragma unroll 1
            for (int k = 0; k < kBnP / kBnSlotCols / 2 && alive; ++k) {
                const int cc = 2 * k + team;
                do_slot(cc * kBnSlotCols, kBnBias1Off + cc * kBnSlotCols, false, k == kBnP / kBnSlotCols / 2 - 1 ? t1empty : 0u);
            }
        }
        if (PROF && p.prof && warp == 4 && lane == 0) {
            long long* o = p.prof + blockIdx.x * 24;
            o[9] = clock64() - e_start; o[10] = w_t3f; o[11] = w_sres; o[12] = w_t1f; o[16] = w_ldt;
            o[17] = w_head; o[18] = w_math; o[19] = w_fence; o[20] = w_issue;
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer may still arrive on this CTA's barriers / the leader's MMAs read this CTA's shared memory
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

int bn_init() {
    HMV_CUDA(cudaFuncSetAttribute(bottleneck_next_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BnGeo<false>::kSmemBytes));
    HMV_CUDA(cudaFuncSetAttribute(bottleneck_next_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BnGeo<false>::kSmemBytes));
    HMV_CUDA(cudaFuncSetAttribute(bottleneck_next_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BnGeo<true>::kSmemBytes));
    HMV_CUDA(cudaFuncSetAttribute(bottleneck_next_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BnGeo<true>::kSmemBytes));
    return 0;
}

int bn_launch(const BnLaunch& l, int num_sms, cudaStream_t stream) {
    if (l.p.num_m_tiles <= 0) return 0;
    if (l.pair) {                      // CTA pairs (tmW3 / tmW1 carry half-height boxes)
        HMV_CHECK(l.p.num_m_tiles % 2 == 0, "seam kernel pairs need an even tile count");
        static std::map<std::pair<int, int>, int> cache;                   // per (device, persistent-grid cap)
        int dev = 0;
        HMV_CUDA(cudaGetDevice(&dev));
        int& max_clusters = cache.emplace(std::make_pair(dev, num_sms), -1).first->second;
        if (max_clusters < 0) {        // the persistent grid must be co-resident; clusters are placed inside one GPC
            cudaLaunchConfig_t qc{};
            qc.gridDim = dim3(2 * (num_sms / 2)); qc.blockDim = dim3(kTcThreads); qc.dynamicSmemBytes = BnGeo<true>::kSmemBytes;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int n = 0;
            HMV_CUDA(cudaOccupancyMaxActiveClusters(&n, bottleneck_next_kernel<false, true>, &qc));
            HMV_CHECK(n > 0, "no CTA pair of the seam kernel fits on this device");
            max_clusters = n < num_sms / 2 ? n : num_sms / 2;
        }
        const int pairs = l.p.num_m_tiles / 2;
        const int clusters = pairs < max_clusters ? pairs : max_clusters;
        if (l.p.prof)
            HMV_CUDA(launch_kernel_cluster(bottleneck_next_kernel<true, true>, dim3(2 * clusters), dim3(kTcThreads), 2, BnGeo<true>::kSmemBytes, stream,
                                           l.tmY2, l.tmW3, l.tmRes, l.tmOut, l.tmW1, l.tmY1, l.p, l.bank));
        else
            HMV_CUDA(launch_kernel_cluster(bottleneck_next_kernel<false, true>, dim3(2 * clusters), dim3(kTcThreads), 2, BnGeo<true>::kSmemBytes, stream,
                                           l.tmY2, l.tmW3, l.tmRes, l.tmOut, l.tmW1, l.tmY1, l.p, l.bank));
        return 0;
    }
    const int grid = l.p.num_m_tiles < num_sms ? l.p.num_m_tiles : num_sms;
    if (l.p.prof)
        HMV_CUDA(launch_kernel(bottleneck_next_kernel<true, false>, dim3(grid), dim3(kTcThreads), BnGeo<false>::kSmemBytes, stream, l.tmY2, l.tmW3, l.tmRes,
                               l.tmOut, l.tmW1, l.tmY1, l.p, l.bank));
    else
        HMV_CUDA(launch_kernel(bottleneck_next_kernel<false, false>, dim3(grid), dim3(kTcThreads), BnGeo<false>::kSmemBytes, stream, l.tmY2, l.tmW3, l.tmRes,
                               l.tmOut, l.tmW1, l.tmY1, l.p, l.bank));
    return 0;
}

}  // namespace hmv
