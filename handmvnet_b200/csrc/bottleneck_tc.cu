// Fused bottleneck tail for sm_100a: conv3x3(+BN+ReLU) -> conv1x1(+BN) + residual + ReLU in ONE persistent
// tcgen05 kernel (reference: backbones/resnet.py:131-143, `conv2/bn2/relu/conv3/bn3/+=identity/relu`).
//
// Why: conv2 (K = 9*P) is tensor-bound and conv3 (K = P, 4*P outputs + a 4*P-wide residual) is HBM-bound; as two
// kernels their times add.  Here one CTA owns an M tile (128 pixels) for BOTH GEMMs and interleaves them, so the
// residual reads / output stores of conv3 stream while the tensor pipe works on the next tile's conv2:
//
//   MMA issue order of a CTA (tiles m_0, m_1, ... strided by the grid):
//     T2(m_0) | T2(m_1) | T2(m_2) with T3(m_0, chunk c) slotted in after every 4th K block | T2(m_3) with T3(m_1, .) | ... | T3 tails
//   (conv3 lags conv2 by two tiles so that the Y2 round trip below is never on the critical path)
//   T2(m):    acc1[128 x P]    = A_3x3[m] * W2^T          (K = 9*P, same TMA tap addressing as conv_gemm_tc.cu)
//   T3(m, c): acc2[128 x 128]  = Y2[m]    * W3[c]^T       (K = P; c = one of P/32 column chunks of the 4*P outputs)
//   Y2[m] (conv2 output, bf16, bias+ReLU applied) is written to global by this CTA's epilogue with TMA stores and read
//   back through TMA as the A operand of T3 (it never leaves L2); an mbarrier orders the store completion before
//   the reload.  TMEM: acc1 at column 0 (double-buffered at 0 / P when P <= 128), acc2 double-buffered at columns
//   256/384.  Epilogue order mirrors the MMA order: E3(m_{i-2}, all chunks) then E2(m_i).
//
// Shared memory: kStages operand stages (A 16 KiB + B max(P*128, 16 KiB)) and ONE ring of 4 chunk slots
// (128 rows x 64 bf16, 128B swizzle) that serves both as residual landing zone (TMA load by warp 2) and as store
// staging (the epilogue adds bias/residual in place, then TMA-stores the slot).
#include <map>
#include "conv_gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace hmv {

namespace {

#ifndef HMV_BT_PAIR_SLOTS
#define HMV_BT_PAIR_SLOTS 5
#endif
constexpr int kBtSlots1 = 4;                                    // chunk slots of the one-CTA-per-tile variant (pair: HMV_BT_PAIR_SLOTS)
constexpr int kBtChunkCols = 64;
constexpr int kBtChunkBytes = kTcBlockM * kBtChunkCols * 2;     // 16 KiB
constexpr int kBtAcc2Col = 256;                                 // TMEM column of the first conv3 accumulator
constexpr int kBtN3 = 128;                                      // conv3 chunk width
constexpr int kBtLag = 2;                                       // conv3 of tile i-kBtLag runs inside the K loop of tile i

template <int P, bool PAIR = false>
struct BtCfg {
    static constexpr int kABytes = kTcBlockM * kTcBlockK * 2;                  // 16 KiB
    // CTA pair (cta_group::2): a CTA holds its own A rows and HALF of every weight tile
    static constexpr int kB2Bytes = (PAIR ? P / 2 : P) * kTcBlockK * 2;
    static constexpr int kB3Bytes = (PAIR ? kBtN3 / 2 : kBtN3) * kTcBlockK * 2;
    static constexpr int kBSlot = kB2Bytes > kB3Bytes ? kB2Bytes : kB3Bytes;
    static constexpr int kStageBytes = kABytes + kBSlot;
    static constexpr int kSlots = PAIR ? HMV_BT_PAIR_SLOTS : kBtSlots1;         // the pair's half weight tiles leave room for a fifth slot
    static constexpr int kStagesRaw = (226 * 1024 - 4 * kBtChunkBytes - 1024) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kKB2 = 9 * P / kTcBlockK;                             // K blocks of conv2
    static constexpr int kKB3 = P / kTcBlockK;                                 // K blocks of one conv3 chunk
    static constexpr int kNCH = 4 * P / kBtN3;                                 // conv3 chunks per tile
    static constexpr int kNA = P <= 128 ? 2 : 1;                               // conv2 accumulators in TMEM (columns 0 / P)
    static constexpr int kSmemBytes = kStages * kStageBytes + kSlots * kBtChunkBytes + 1024 /*align*/ + 512 /*barriers*/;
    static_assert(P == 64 || P == 128 || P == 256, "bottleneck widths of ResNet-50");
    static_assert(kStages >= 3 && kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(kKB2 / 4 >= kNCH, "every conv3 chunk of the previous tile fits between the K blocks of conv2");
};

// Segment i of a CTA: the K loop of conv2 on its i-th tile (if any) with the conv3 chunks of tile i-1 (if any) slotted in
// after every 4th K block; whatever is left (the tail segment has no K loop) follows.  The TMA producer and the MMA
// issuer walk the same schedule.
template <int P, typename F2, typename F3>
__device__ __forceinline__ bool bt_segment(bool has_t2, bool has_t3, F2&& kblock2, F3&& chunk3) {
    const int nkb = has_t2 ? BtCfg<P>::kKB2 : 0;
    const int nch = has_t3 ? BtCfg<P>::kNCH : 0;
    int c = 0;
    for (int kb = 0; kb < nkb; ++kb) {
        if (!kblock2(kb)) return false;
        if ((kb & 3) == 3 && c < nch) {
            if (!chunk3(c)) return false;
            ++c;
        }
    }
    for (; c < nch; ++c)
        if (!chunk3(c)) return false;
    return true;
}

// mbar_wait that also accumulates the stall time (clock cycles) into `acc` -- read back by HMV_BT_PROF=1 runs
__device__ __forceinline__ bool timed_wait(uint32_t bar, uint32_t parity, int* err_flag, int code, long long& acc, bool prof) {
    if (!prof) return mbar_wait(bar, parity, err_flag, code);
    const long long t0 = clock64();                  // (try_wait may suspend inside the instruction: time it as well)
    const bool ok = mbar_wait(bar, parity, err_flag, code);
    acc += clock64() - t0;
    return ok;
}

__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// PAIR: launched as clusters of two CTAs on neighbouring M tiles; the leader (rank 0) issues tcgen05.mma.cta_group::2 of M = 256 for
// both (A rows from both CTAs' shared memory at the same offsets, half of each weight tile from each).  Both CTAs' TMA loads
// count on the leader's `full` barriers, the leader's commits arrive on `empty` / `t1full` / `t2full` in both CTAs, and the peer's
// epilogue warps arrive on the leader's `t1empty` / `t2empty` remotely.  The weights cross L2 -> SM once per pair.
template <int P, bool PROF, bool PAIR>
__global__ void __launch_bounds__(kTcThreads, 1)
bottleneck_tail_kernel(const __grid_constant__ CUtensorMap tmA,    // conv2 input, 5-D activation map
                       const __grid_constant__ CUtensorMap tmW2,   // [P, 9P] box {64, P}
                       const __grid_constant__ CUtensorMap tmY2s,  // conv2 output [rows, P], store box {64, 32}
                       const __grid_constant__ CUtensorMap tmY2l,  // conv2 output [rows, P], load box {64, 128}
                       const __grid_constant__ CUtensorMap tmW3,   // [4P, P] box {64, 128}
                       const __grid_constant__ CUtensorMap tmOut,  // block output [rows, 4P], store box {64, 32}
                       const __grid_constant__ CUtensorMap tmRes,  // residual [rows, 4P], load box {64, 128} (ds_kb > 0: the block input [rows, 64 * ds_kb])
                       const __grid_constant__ CUtensorMap tmWd,   // ds_kb > 0: downsample weights [4P, 64 * ds_kb], box {64, 128}
                       const __grid_constant__ BtParams p,
                       const __grid_constant__ BiasBank bank) {            // conv2 biases [0, P), conv3 biases [256, 256 + 4P)
    using Cfg = BtCfg<P, PAIR>;
    constexpr int kBtSlots = Cfg::kSlots;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* slots = smem + Cfg::kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(slots + kBtSlots * kBtChunkBytes);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * Cfg::kStages;
    const uint32_t t1full0 = empty0 + 8 * Cfg::kStages;
    const uint32_t t1empty0 = t1full0 + 16;
    const uint32_t t2full0 = t1empty0 + 16;
    const uint32_t t2empty0 = t2full0 + 16;
    const uint32_t cfull0 = t2empty0 + 16;
    const uint32_t cempty0 = cfull0 + 8 * kBtSlots;
    const uint32_t y2ready0 = cempty0 + 8 * kBtSlots;       // two barriers: even / odd tiles of this CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 8 + 2 * kBtSlots + 2);
    const uint32_t smem_base = smem_u32(smem);
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
        prefetch_tmap(&tmA); prefetch_tmap(&tmW2); prefetch_tmap(&tmY2s); prefetch_tmap(&tmY2l);
        prefetch_tmap(&tmW3); prefetch_tmap(&tmOut); prefetch_tmap(&tmRes);
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(t1full0 + 8 * i, 1);
            mbar_init(t1empty0 + 8 * i, (P == 64 ? 4 : 8) * kCtas);   // one arrival per epilogue warp that reads the conv2 accumulator (P = 64: one chunk, one team), of both CTAs of a pair
            mbar_init(t2full0 + 8 * i, 1);
            mbar_init(t2empty0 + 8 * i, 8 * kCtas);
            mbar_init(y2ready0 + 8 * i, P == 64 ? 4 : 8);   // every warp that stored a slab of Y2[m]
        }
        for (int i = 0; i < kBtSlots; ++i) {
            mbar_init(cfull0 + 8 * i, 1);
            mbar_init(cempty0 + 8 * i, 4);                  // the four slab-store issuers
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
    // count is even) and this CTA takes tile 2 * pair + rank
    const int units = PAIR ? p.num_m_tiles / 2 : p.num_m_tiles;
    const int first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int n_i = (units - first + step - 1) / step;
    auto tile_of = [&](int i) { return PAIR ? 2 * (first + i * step) + static_cast<int>(crank) : first + i * step; };
    constexpr bool prof = PROF;                               // stall counters compiled in only for HMV_BT_PROF=1 launches

    if (warp == 0) {
        // ===================== TMA producer of the operand ring =====================
        {   // the whole warp walks the schedule and waits; one elected lane issues the copies (see elect_one() in tc_ptx.cuh)
            int stage = 0;
            uint32_t phase = 0;
            long long w_empty = 0, w_y2 = 0;
            for (int i = 0; i < n_i + kBtLag; ++i) {
                const int j3 = i - kBtLag;                   // index of the tile whose conv3 runs in this segment
                const int m2 = tile_of(i), m3 = tile_of(j3);
                const int h0 = (m2 % p.tpi) * p.hbox, img = m2 / p.tpi;
                const bool ok = bt_segment<P>(i < n_i, j3 >= 0,
                    [&](int kb) {
                        if (!timed_wait(empty0 + 8 * stage, phase ^ 1, p.err_flag, 21, w_empty, prof)) return false;
                        if (elect_one()) {
                            const TcTap tap = p.taps[kb / (P / kTcBlockK)];
                            const int cb = kb % (P / kTcBlockK);
                            const uint32_t fb = full0 + 8 * stage;
                            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                            if (PAIR) {        // own A rows + this CTA's half of the weight rows; both CTAs' bytes are counted on the LEADER's barrier
                                const uint32_t lfb = crank == 0 ? fb : mapa_u32(fb, 0);
                                if (crank == 0) mbar_arrive_expect_tx(fb, 2u * (Cfg::kABytes + Cfg::kB2Bytes));
                                tma_load_5d_2sm(sa, &tmA, lfb, tap.c_off + cb * kTcBlockK, tap.dw, tap.a, h0 + tap.dh, img);
                                tma_load_2d_2sm(sa + Cfg::kABytes, &tmW2, lfb, kb * kTcBlockK, static_cast<int>(crank) * (P / 2));
                            } else {
                                mbar_arrive_expect_tx(fb, Cfg::kABytes + Cfg::kB2Bytes);
                                tma_load_5d(sa, &tmA, fb, tap.c_off + cb * kTcBlockK, tap.dw, tap.a, h0 + tap.dh, img);
                                tma_load_2d(sa + Cfg::kABytes, &tmW2, fb, kb * kTcBlockK, 0);
                            }
                        }
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        return true;
                    },
                    [&](int c) {
                        // Y2[m3] was TMA-stored by this CTA's epilogue; wait until those stores have completed
                        if (c == 0 && !timed_wait(y2ready0 + 8 * (j3 & 1), static_cast<uint32_t>(j3 >> 1) & 1u, p.err_flag, 22, w_y2, prof)) return false;
                        if (c == 0) fence_proxy_async_all();      // (every lane: the elected one is not known in advance)
                        for (int kb3 = 0; kb3 < Cfg::kKB3 + p.ds_kb; ++kb3) {     // (+ the folded downsample's K blocks: block input x its weights)
                            if (!timed_wait(empty0 + 8 * stage, phase ^ 1, p.err_flag, 23, w_empty, prof)) return false;
                            if (elect_one()) {
                                const uint32_t fb = full0 + 8 * stage;
                                const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                                const CUtensorMap* ma = kb3 < Cfg::kKB3 ? &tmY2l : &tmRes;      // conv2 output / (folded downsample) block input
                                const CUtensorMap* mb = kb3 < Cfg::kKB3 ? &tmW3 : &tmWd;
                                const int kc = (kb3 < Cfg::kKB3 ? kb3 : kb3 - Cfg::kKB3) * kTcBlockK;
                                if (PAIR) {
                                    const uint32_t lfb = crank == 0 ? fb : mapa_u32(fb, 0);
                                    if (crank == 0) mbar_arrive_expect_tx(fb, 2u * (Cfg::kABytes + Cfg::kB3Bytes));
                                    tma_load_2d_2sm(sa, ma, lfb, kc, m3 * kTcBlockM);
                                    tma_load_2d_2sm(sa + Cfg::kABytes, mb, lfb, kc, c * kBtN3 + static_cast<int>(crank) * (kBtN3 / 2));
                                } else {
                                    mbar_arrive_expect_tx(fb, Cfg::kABytes + Cfg::kB3Bytes);
                                    tma_load_2d(sa, ma, fb, kc, m3 * kTcBlockM);
                                    tma_load_2d(sa + Cfg::kABytes, mb, fb, kc, c * kBtN3);
                                }
                            }
                            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        }
                        return true;
                    });
                if (!ok) break;
            }
            if (PROF && p.prof && lane == 0) { p.prof[blockIdx.x * 24 + 5] = w_empty; p.prof[blockIdx.x * 24 + 6] = w_y2; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (crank == 0) {                                    // pair: only the leader issues, for both CTAs
            // the whole warp walks the schedule and waits; one elected lane issues (see elect_one() in tc_ptx.cuh for why)
            constexpr uint32_t idesc2 = PAIR ? make_idesc_mn(2 * kTcBlockM, P) : make_idesc(P);
            constexpr uint32_t idesc3 = PAIR ? make_idesc_mn(2 * kTcBlockM, kBtN3) : make_idesc(kBtN3);
            auto wait_epi = [&](uint32_t bar, uint32_t parity, int code, long long& acc) {      // barriers the epilogue warps arrive on
                if (!PAIR) return timed_wait(bar, parity, p.err_flag, code, acc, prof);
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
            uint32_t q = 0;                                  // running conv3 chunk counter (TMEM slot = q & 1)
            long long w_t1e = 0, w_f2 = 0, w_t2e = 0, w_f3 = 0, t_issue = 0, t_setup = 0, t_mma = 0, t_commit = 0;
            const long long t_start = clock64();
            for (int i = 0; i < n_i + kBtLag; ++i) {
                const uint32_t a = static_cast<uint32_t>(i) % Cfg::kNA, ause = static_cast<uint32_t>(i) / Cfg::kNA;
                const uint32_t acc1 = tmem_base + a * P;
                const bool ok = bt_segment<P>(i < n_i, i >= kBtLag,
                    [&](int kb) {
                        if (kb == 0) {                       // this conv2 accumulator was drained by the epilogue of its previous tile
                            if (!wait_epi(t1empty0 + 8 * a, (ause & 1u) ^ 1u, 24, w_t1e)) return false;
                            tc_fence_after();
                        }
                        if (!timed_wait(full0 + 8 * stage, phase, p.err_flag, 25, w_f2, prof)) return false;
                        tc_fence_after();
                        const long long c0 = PROF ? clock64() : 0;
                        if (elect_one()) {
                            const uint64_t adesc = make_sw128_desc(smem_base + stage * Cfg::kStageBytes);
                            const uint64_t bdesc = make_sw128_desc(smem_base + stage * Cfg::kStageBytes + Cfg::kABytes);
                            const long long c1 = PROF ? clock64() : 0;
#pragma unroll
                            for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k)       // +32 bytes of K per MMA = +2 in the address field
                                mma(acc1, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0 ? 1u : 0u);
                            const long long c2 = PROF ? clock64() : 0;
                            commit(empty0 + 8 * stage);
                            if (kb == Cfg::kKB2 - 1) commit(t1full0 + 8 * a);
                            if (PROF) { t_setup += c1 - c0; t_mma += c2 - c1; t_commit += clock64() - c2; }
                        }
                        if (PROF) t_issue += clock64() - c0;
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        return true;
                    },
                    [&](int) {
                        const uint32_t s = q & 1u, use = q >> 1;
                        if (!wait_epi(t2empty0 + 8 * s, (use & 1u) ^ 1u, 26, w_t2e)) return false;
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + kBtAcc2Col + s * kBtN3;
                        const int nk3 = Cfg::kKB3 + p.ds_kb;
                        for (int kb3 = 0; kb3 < nk3; ++kb3) {
                            if (!timed_wait(full0 + 8 * stage, phase, p.err_flag, 27, w_f3, prof)) return false;
                            tc_fence_after();
                            const long long c0 = PROF ? clock64() : 0;
                            if (elect_one()) {
                                const uint64_t adesc = make_sw128_desc(smem_base + stage * Cfg::kStageBytes);
                                const uint64_t bdesc = make_sw128_desc(smem_base + stage * Cfg::kStageBytes + Cfg::kABytes);
#pragma unroll
                                for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k)
                                    mma(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc3, (kb3 | k) != 0 ? 1u : 0u);
                                commit(empty0 + 8 * stage);
                                if (kb3 == nk3 - 1) commit(t2full0 + 8 * s);
                            }
                            if (PROF) t_issue += clock64() - c0;
                            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        }
                        ++q;
                        return true;
                    });
                if (!ok) break;
            }
            if (PROF && p.prof && lane == 0) {
                long long* o = p.prof + blockIdx.x * 24;
                o[0] = clock64() - t_start; o[1] = w_t1e; o[2] = w_f2; o[3] = w_t2e; o[4] = w_f3; o[13] = n_i; o[14] = t_issue; o[15] = 0; o[16] = t_setup; o[17] = t_mma; o[18] = t_commit;
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== chunk-slot producer: residual prefetch (conv3 chunks) / plain hand-over (conv2 chunks) ============
        {   // whole warp + one elected lane per copy, as above
            uint32_t g = 0;
            bool alive = true;
            long long w_ce = 0;
            for (int i = 0; i < n_i + kBtLag && alive; ++i) {
                if (p.prefetch && !p.ds_kb && i + 1 >= kBtLag && i + 1 - kBtLag < n_i && lane == 0)      // next segment's residual tile: HBM -> L2 ahead of its slot loads
                    for (int c = 0; c < Cfg::kNCH * 2; ++c) tma_prefetch_l2_2d(&tmRes, c * kBtChunkCols, tile_of(i + 1 - kBtLag) * kTcBlockM);
                __syncwarp();
                if (i >= kBtLag) {
                    const int m3 = tile_of(i - kBtLag);
                    for (int c = 0; c < Cfg::kNCH * 2 && alive; ++c, ++g) {
                        const uint32_t slot = g % kBtSlots, use = g / kBtSlots;
                        if (!timed_wait(cempty0 + 8 * slot, (use & 1u) ^ 1u, p.err_flag, 28, w_ce, prof)) { alive = false; break; }
                        if (elect_one()) {
                            if (p.ds_kb) {                  // folded downsample: no residual to fetch, plain hand-over
                                mbar_arrive(cfull0 + 8 * slot);
                            } else {
                                mbar_arrive_expect_tx(cfull0 + 8 * slot, kBtChunkBytes);
                                tma_load_2d(slots_base + slot * kBtChunkBytes, &tmRes, cfull0 + 8 * slot, c * kBtChunkCols, m3 * kTcBlockM);
                            }
                        }
                    }
                }
                if (i < n_i) {
                    for (int c = 0; c < P / kBtChunkCols && alive; ++c, ++g) {
                        const uint32_t slot = g % kBtSlots, use = g / kBtSlots;
                        if (!timed_wait(cempty0 + 8 * slot, (use & 1u) ^ 1u, p.err_flag, 29, w_ce, prof)) { alive = false; break; }
                        if (lane == 0) mbar_arrive(cfull0 + 8 * slot);
                    }
                }
            }
            if (PROF && p.prof && lane == 0) p.prof[blockIdx.x * 24 + 12] = w_ce;
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: 2 teams x 4 warps; a warp owns one 32-row slab (TMEM lane quarter) of a whole 64-column chunk =========
        // Every warp walks the chunk sequence; team t processes the chunks whose running index g is t (mod 2), so two
        // chunks are in flight at a time and a warp needs no other warp to finish its slab (no named barrier):
        // TMEM -> +bias (+residual already in the slab) -> ReLU -> bf16 in place -> its own TMA store.
        const int quarter = warp & 3;
        const uint32_t team = static_cast<uint32_t>((warp - 4) >> 2);
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t slab_off = quarter * (32 * 128) + lane * 128;
        uint32_t g = 0;                                      // running chunk counter (slot = g % kBtSlots)
        int pending = -1;                                    // lane 0: slot whose TMA store may still be reading smem
        bool alive = true;
        long long w_t2f = 0, w_cf = 0, w_t1f = 0, w_bulk = 0;

        auto do_chunk = [&](uint32_t tcol, int bias_off, bool has_res, uint32_t release_bar, const CUtensorMap* tm_out, int col,
                            int row) {
            const uint32_t slot = g % kBtSlots, use = g / kBtSlots;
            uint8_t* srow = slots + slot * kBtChunkBytes + slab_off;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                 // two 32-column halves of the slab row
                uint32_t r[32];
                tmem_ld32(lane_base + tcol + hf * 32, r);
                float4 bq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) bq[j] = *reinterpret_cast<const float4*>(&bank.v[bias_off + hf * 32 + 4 * j]);   // constant cache
                if (hf == 0 && !timed_wait(cfull0 + 8 * slot, use & 1u, p.err_flag, 30, w_cf, prof)) alive = false;
                uint4 rq[4];
                if (has_res) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rq[j] = *reinterpret_cast<const uint4*>(srow + ((static_cast<uint32_t>(hf * 4 + j) ^ (lane & 7)) << 4));
                }
                tmem_ld_wait();
                if (hf == 1 && release_bar != 0) {           // this warp's part of the accumulator is read: hand TMEM back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_mma(release_bar);        // (the leader's barrier in a pair)
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 v0 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])), make_float2(bq[2 * j].x, bq[2 * j].y));
                    float2 v1 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), make_float2(bq[2 * j].z, bq[2 * j].w));
                    float2 v2 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), make_float2(bq[2 * j + 1].x, bq[2 * j + 1].y));
                    float2 v3 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])), make_float2(bq[2 * j + 1].z, bq[2 * j + 1].w));
                    if (has_res) {
                        v0 = __fadd2_rn(v0, bf16x2_to_f2(rq[j].x)); v1 = __fadd2_rn(v1, bf16x2_to_f2(rq[j].y));
                        v2 = __fadd2_rn(v2, bf16x2_to_f2(rq[j].z)); v3 = __fadd2_rn(v3, bf16x2_to_f2(rq[j].w));
                    }
                    uint4 o;
                    o.x = cvt_bf16x2(v0.x, v0.y, true); o.y = cvt_bf16x2(v1.x, v1.y, true);
                    o.z = cvt_bf16x2(v2.x, v2.y, true); o.w = cvt_bf16x2(v3.x, v3.y, true);
                    *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(hf * 4 + j) ^ (lane & 7)) << 4)) = o;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(tm_out, slots_base + slot * kBtChunkBytes + quarter * (32 * 128), col, row + quarter * 32);
                bulk_commit();
                pending = static_cast<int>(slot);
            }
            __syncwarp();
        };
        // Before the warp waits for its next accumulator: the slab store issued a moment ago has read shared memory by
        // now, and the slot goes back to the residual prefetcher ahead of its next use.
        auto release_pending = [&]() {
            if (lane == 0 && pending >= 0) {
                if (PROF) { const long long tb = clock64(); bulk_wait_read<0>(); w_bulk += clock64() - tb; } else bulk_wait_read<0>();
                mbar_arrive(cempty0 + 8 * pending);
                pending = -1;
            }
        };

        uint32_t q = 0;
        for (int i = 0; i < n_i + kBtLag && alive; ++i) {
            if (i >= kBtLag) {
                const int m3 = tile_of(i - kBtLag);
                for (int c = 0; c < Cfg::kNCH && alive; ++c, ++q) {
                    const uint32_t s = q & 1u;
                    release_pending();
                    if (!timed_wait(t2full0 + 8 * s, (q >> 1) & 1u, p.err_flag, 31, w_t2f, prof)) { alive = false; break; }
                    tc_fence_after();
#pragma unroll 1
                    for (int cc = 0; cc < kBtN3 / kBtChunkCols && alive; ++cc, ++g) {     // two chunks: one per team
                        if ((g & 1u) != team) continue;
                        do_chunk(kBtAcc2Col + s * kBtN3 + cc * kBtChunkCols, kBtBias3Off + c * kBtN3 + cc * kBtChunkCols, p.ds_kb == 0, t2empty0 + 8 * s, &tmOut,
                                 c * kBtN3 + cc * kBtChunkCols, m3 * kTcBlockM);
                    }
                }
            }
            if (i < n_i && alive) {
                const int m2 = tile_of(i);
                const uint32_t a = static_cast<uint32_t>(i) % Cfg::kNA, ause = static_cast<uint32_t>(i) / Cfg::kNA;
                constexpr int kN2 = P / kBtChunkCols;        // conv2 output chunks (1, 2 or 4)
                bool mine = false;
                for (int cc = 0; cc < kN2; ++cc) mine = mine || (((g + cc) & 1u) == team);
                if (mine) {
                    release_pending();
                    if (!timed_wait(t1full0 + 8 * a, ause & 1u, p.err_flag, 32, w_t1f, prof)) { alive = false; break; }
                    tc_fence_after();
                }
#pragma unroll 1
                for (int cc = 0; cc < kN2 && alive; ++cc, ++g) {
                    if ((g & 1u) != team) continue;
                    if (cc >= 2) release_pending();
                    do_chunk(a * P + cc * kBtChunkCols, cc * kBtChunkCols, false, cc + 2 >= kN2 ? t1empty0 + 8 * a : 0u, &tmY2s,
                             cc * kBtChunkCols, m2 * kTcBlockM);
                }
                if (mine && lane == 0) {                     // Y2[m2] must be complete in global memory before it is reloaded
                    if (PROF) { const long long tb = clock64(); bulk_wait_all(); w_bulk += clock64() - tb; } else bulk_wait_all();
                    if (pending >= 0) mbar_arrive(cempty0 + 8 * pending);
                    pending = -1;
                    fence_proxy_async_all();
                    mbar_arrive(y2ready0 + 8 * (i & 1));
                }
            }
        }
        if (lane == 0 && pending >= 0) bulk_wait_read<0>();  // staged data must stay valid until every store has read it
        if (PROF && p.prof && warp == 4 && lane == 0) {
            long long* o = p.prof + blockIdx.x * 24;
            o[7] = w_t2f; o[8] = w_cf; o[9] = w_t1f; o[10] = w_bulk; o[11] = 0;
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

template <int P, bool PAIR>
int bt_launch_p(const BtLaunch& l, int num_sms, cudaStream_t stream) {
    using Cfg = BtCfg<P, PAIR>;
    if (PAIR) {                        // CTA pairs (tmW2 / tmW3 / tmWd carry half-height boxes)
        HMV_CHECK(l.p.num_m_tiles % 2 == 0, "tail kernel pairs need an even tile count");
        static std::map<std::pair<int, int>, int> cache;                   // per (device, persistent-grid cap)
        int dev = 0;
        HMV_CUDA(cudaGetDevice(&dev));
        int& max_clusters = cache.emplace(std::make_pair(dev, num_sms), -1).first->second;
        if (max_clusters < 0) {        // the persistent grid must be co-resident; clusters are placed inside one GPC
            cudaLaunchConfig_t qc{};
            qc.gridDim = dim3(2 * (num_sms / 2)); qc.blockDim = dim3(kTcThreads); qc.dynamicSmemBytes = Cfg::kSmemBytes;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int n = 0;
            HMV_CUDA(cudaOccupancyMaxActiveClusters(&n, bottleneck_tail_kernel<P, false, PAIR>, &qc));
            HMV_CHECK(n > 0, "no CTA pair of the tail kernel fits on this device");
            max_clusters = n < num_sms / 2 ? n : num_sms / 2;
        }
        const int pairs = l.p.num_m_tiles / 2;
        const int clusters = pairs < max_clusters ? pairs : max_clusters;
        if (l.p.prof)
            HMV_CUDA(launch_kernel_cluster(bottleneck_tail_kernel<P, true, PAIR>, dim3(2 * clusters), dim3(kTcThreads), 2, Cfg::kSmemBytes, stream, l.tmA,
                                           l.tmW2, l.tmY2s, l.tmY2l, l.tmW3, l.tmOut, l.tmRes, l.tmWd, l.p, l.bank));
        else
            HMV_CUDA(launch_kernel_cluster(bottleneck_tail_kernel<P, false, PAIR>, dim3(2 * clusters), dim3(kTcThreads), 2, Cfg::kSmemBytes, stream, l.tmA,
                                           l.tmW2, l.tmY2s, l.tmY2l, l.tmW3, l.tmOut, l.tmRes, l.tmWd, l.p, l.bank));
        return 0;
    }
    const int grid = l.p.num_m_tiles < num_sms ? l.p.num_m_tiles : num_sms;
    if (l.p.prof)
        HMV_CUDA(launch_kernel(bottleneck_tail_kernel<P, true, PAIR>, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, l.tmA, l.tmW2, l.tmY2s,
                               l.tmY2l, l.tmW3, l.tmOut, l.tmRes, l.tmWd, l.p, l.bank));
    else
        HMV_CUDA(launch_kernel(bottleneck_tail_kernel<P, false, PAIR>, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, l.tmA, l.tmW2, l.tmY2s,
                               l.tmY2l, l.tmW3, l.tmOut, l.tmRes, l.tmWd, l.p, l.bank));
    return 0;
}

template <int P, bool PROF, bool PAIR>
int bt_attr() {
    HMV_CUDA(cudaFuncSetAttribute(bottleneck_tail_kernel<P, PROF, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, BtCfg<P, PAIR>::kSmemBytes));
    return 0;
}

}  // namespace

int bt_init() {
    if (bt_attr<64, false, false>() || bt_attr<128, false, false>() || bt_attr<256, false, false>() || bt_attr<64, true, false>() ||
        bt_attr<128, true, false>() || bt_attr<256, true, false>())
        return 1;
    if (bt_attr<64, false, true>() || bt_attr<128, false, true>() || bt_attr<64, true, true>() || bt_attr<128, true, true>()) return 1;
    return 0;
}

int bt_launch(const BtLaunch& l, int num_sms, cudaStream_t stream) {
    if (l.p.num_m_tiles <= 0) return 0;
    switch (l.planes) {
        case 64: return l.pair ? bt_launch_p<64, true>(l, num_sms, stream) : bt_launch_p<64, false>(l, num_sms, stream);
        case 128: return l.pair ? bt_launch_p<128, true>(l, num_sms, stream) : bt_launch_p<128, false>(l, num_sms, stream);
        case 256: return bt_launch_p<256, false>(l, num_sms, stream);
    }
    set_error("bt_launch: unsupported bottleneck width " + std::to_string(l.planes));
    return 1;
}

}  // namespace hmv
