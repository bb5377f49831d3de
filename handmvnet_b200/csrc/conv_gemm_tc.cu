// See conv_gemm_tc.cuh for the design.  sm_100a only: tcgen05.mma / tcgen05.ld / TMEM alloc,
// cp.async.bulk.tensor (TMA) and mbarrier pipelines are written as inline PTX.
#include "conv_gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <map>
#include <utility>

namespace hmv {

// ------------------------------------------------------------------------------------------------
// epilogue: 16 consecutive output columns of one row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_16(const Epilogue& ep, const BiasBank& bank, bool bias_in_params, float (&v)[16], int row, int col0) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 b = bias_in_params ? *reinterpret_cast<const float4*>(&bank.v[col0 + i])
                                        : __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
    if (ep.res_mode != RES_NONE) {
        const size_t rrow = static_cast<size_t>(row / ep.res_group) * ep.res_stride + row % ep.res_group;
        if (ep.res_mode == RES_BF16) {
            const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const bf16*>(ep.residual) + rrow * ep.res_ld + col0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint4 q = __ldg(rp + i);
                const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
                v[8 * i + 0] += a.x; v[8 * i + 1] += a.y; v[8 * i + 2] += b.x; v[8 * i + 3] += b.y;
                v[8 * i + 4] += c.x; v[8 * i + 5] += c.y; v[8 * i + 6] += d.x; v[8 * i + 7] += d.y;
            }
        } else {
            const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(ep.residual) + rrow * ep.res_ld + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 q = __ldg(rp + i);
                v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
        }
    }
    if (ep.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (ep.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
    }
    if (ep.out_mode == OUT_BF16_ROWMAJOR) {
        uint4* op = reinterpret_cast<uint4*>(static_cast<bf16*>(ep.out) + static_cast<size_t>(row) * ep.ldc + col0);
        uint4 q0, q1;
        q0.x = pack_bf16x2(v[0], v[1]);   q0.y = pack_bf16x2(v[2], v[3]);
        q0.z = pack_bf16x2(v[4], v[5]);   q0.w = pack_bf16x2(v[6], v[7]);
        q1.x = pack_bf16x2(v[8], v[9]);   q1.y = pack_bf16x2(v[10], v[11]);
        q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
        op[0] = q0;
        op[1] = q1;
    } else if (ep.out_mode == OUT_F32_ROWMAJOR) {
        float4* op = reinterpret_cast<float4*>(static_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldc + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {  // OUT_F32_NCHW: lanes are consecutive pixels -> coalesced per channel
        const int img = row / ep.hw, pix = row % ep.hw;
        float* op = static_cast<float*>(ep.out) + (static_cast<size_t>(img) * ep.N) * ep.hw + pix;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (col0 + i < ep.N) op[static_cast<size_t>(col0 + i) * ep.hw] = v[i];
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
constexpr int kChunkCols = 64;                         // staged output chunk: 128 rows x 64 bf16 = 16 KiB
constexpr int kChunkBytes = kTcBlockM * kChunkCols * 2;

template <int BN, int MODE, bool PAIRED = false>
struct TcCfg {
    static constexpr int kABytes = kTcBlockM * kTcBlockK * 2;       // 16 KiB
    static constexpr int kBBytes = (PAIRED ? BN / 2 : BN) * kTcBlockK * 2;       // a CTA of a cta_group::2 pair holds half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    // staging rings of the TMA-store epilogues (see TcMode)
    static constexpr int kNumS = MODE == TC_DIRECT ? 0 : (MODE == TC_STORE_DEEP ? 4 : 2);     // store slots
    static constexpr int kNumR = MODE == TC_STORE_RES ? 3 : 0;                                 // residual slots
    static constexpr int kCBytes = (kNumS + kNumR) * kChunkBytes;
    static constexpr int kStagesRaw = (224 * 1024 - kCBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                     : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kCBytes + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
    static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment");
    static_assert(MODE == TC_DIRECT || BN % kChunkCols == 0, "TMA-store epilogue works on 64-column chunks");
    static_assert(kSmemBytes <= 227 * 1024 && kStages >= 2, "shared memory budget");
};

// CL = 4: CTA PAIR (tcgen05 cta_group::2).  The two CTAs of a cluster again own two neighbouring M tiles of one N tile, but
// the leader (rank 0) issues ONE tcgen05.mma of M = 256 per K step for both: each CTA's shared memory holds its own 128 A rows
// and only HALF of the B tile (its BN/2 rows), each CTA's TMEM receives its own 128 accumulator rows.  Both CTAs' TMA loads
// count their bytes on the leader's `full` barrier; the leader's tcgen05.commit arrives on `empty` / `tfull` in both CTAs; the
// peer's epilogue warps arrive on the leader's `tempty` remotely.  Per CTA a stage is 16 + BN/4 KiB instead of 16 + BN/2:
// for BN = 256 six operand stages fit where four did, and the B operand crosses L2 -> SM once per pair.
// CL = 2: launched as clusters of two CTAs that work on two neighbouring M tiles of the same N tile at the same time.
// Each CTA loads its own A tile and ONE HALF of the shared weight tile, multicast into both CTAs' shared memory
// (tmB then has a BN/2-row box), which halves the L2 -> SM traffic of the B operand.  A stage may only be refilled when
// BOTH CTAs have consumed it, so the MMA issuer's commit arrives on the `empty` barrier of both CTAs (count 2).
template <int BN, int MODE, int CL>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const __grid_constant__ TcParams p, const __grid_constant__ BiasBank bank) {
    constexpr bool PAIR = CL == 4;
    using Cfg = TcCfg<BN, MODE, PAIR>;
    static_assert(CL == 1 || ((CL == 2 || CL == 4) && BN % 32 == 0), "cluster width");
    const uint32_t crank = CL >= 2 ? cluster_ctarank() : 0u;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kCBytes);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * Cfg::kStages;
    const uint32_t tfull0 = empty0 + 8 * Cfg::kStages;
    const uint32_t tempty0 = tfull0 + 16;
    const uint32_t cfull0 = tempty0 + 16;                  // residual chunk landed (4 slots)
    const uint32_t cempty0 = cfull0 + 32;                  // staging slot free again (4 slots)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 12);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sbuf0 = smem_base + Cfg::kStages * Cfg::kStageBytes;      // store staging ring
    const uint32_t rbuf0 = sbuf0 + Cfg::kNumS * kChunkBytes;                  // residual ring

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool has_res = MODE == TC_STORE_RES && p.ep.res_mode == RES_BF16;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (MODE != TC_DIRECT) prefetch_tmap(&tmC);
        if (has_res) prefetch_tmap(&tmR);
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, CL == 2 ? 2 : 1);                 // CL = 2: one commit per CTA of the cluster; pair: the leader's commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, (MODE == TC_DIRECT ? 4 : 8) * (PAIR ? 2 : 1));     // one arrival per epilogue warp (pair: of both CTAs, on the leader)
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(cfull0 + 8 * i, 1);
            mbar_init(cempty0 + 8 * i, 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL >= 2) cluster_sync_all();   // the peer's barriers are initialised before anything remote can land on them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();                        // everything above overlapped the previous kernel; its results are visible from here
    pdl_launch_dependents();

    // Work items.  CL == 1: tile = (m_tile, n_tile), strided by the grid.  CL == 2: the cluster walks (pair of M tiles,
    // n_tile) items and this CTA takes M tile 2 * pair + rank; an M tile past the end is a dummy (zero-filled loads,
    // clipped stores) that still takes part in the stage hand-shake.
    const int num_tiles = CL >= 2 ? ((p.num_m_tiles + 1) / 2) * p.num_n_tiles : p.num_m_tiles * p.num_n_tiles;
    const int tile0 = CL >= 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_step = CL >= 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto m_of = [&](int tile) { return CL >= 2 ? 2 * (tile / p.num_n_tiles) + static_cast<int>(crank) : tile / p.num_n_tiles; };
    const int num_kb = p.num_taps * p.cblks;

    if (warp == 0) {
        // ===================== TMA producer (A / B operand tiles) =====================
        {   // the whole warp walks the schedule and waits; one elected lane issues the copies (see elect_one() in tc_ptx.cuh)
            int stage = 0;
            uint32_t phase = 0;
            bool alive = true;
            for (int tile = tile0; tile < num_tiles && alive; tile += tile_step) {
                const int n_tile = tile % p.num_n_tiles;
                const int m_tile = m_of(tile);
                const int w0 = p.flat ? m_tile * kTcBlockM : 0;
                const int h0 = p.flat ? 0 : (m_tile % p.tpi) * p.hbox;
                const int img = p.flat ? 0 : m_tile / p.tpi;
                int kb = 0;
                for (int t = 0; t < p.num_taps && alive; ++t) {
                    const TcTap tap = p.taps[t];
                    for (int cb = 0; cb < p.cblks; ++cb, ++kb) {
                        if (!mbar_wait(empty0 + 8 * stage, phase ^ 1, p.err_flag, 1)) { alive = false; break; }
                        if (elect_one()) {
                            const uint32_t fb = full0 + 8 * stage;
                            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                            if (PAIR) {    // both CTAs' bytes are counted on the LEADER's barrier, which only the leader arms
                                const uint32_t lfb = crank == 0 ? fb : mapa_u32(fb, 0);
                                if (crank == 0) mbar_arrive_expect_tx(fb, 2u * (static_cast<uint32_t>(p.a_bytes) + Cfg::kBBytes));
                                tma_load_5d_2sm(sa, &tmA, lfb, tap.c_off + cb * kTcBlockK, w0 + tap.dw, tap.a, h0 + tap.dh, img);
                                tma_load_2d_2sm(sa + Cfg::kABytes, &tmB, lfb, kb * kTcBlockK, n_tile * BN + static_cast<int>(crank) * (BN / 2));
                            } else {
                                mbar_arrive_expect_tx(fb, static_cast<uint32_t>(p.a_bytes) + Cfg::kBBytes);
                                tma_load_5d(sa, &tmA, fb, tap.c_off + cb * kTcBlockK, w0 + tap.dw, tap.a, h0 + tap.dh, img);
                                if (CL == 2)   // this CTA's half of the weight tile, delivered to both CTAs
                                    tma_load_2d_mc(sa + Cfg::kABytes + crank * (Cfg::kBBytes / 2), &tmB, fb, kb * kTcBlockK,
                                                   n_tile * BN + static_cast<int>(crank) * (BN / 2), static_cast<uint16_t>(3));
                                else
                                    tma_load_2d(sa + Cfg::kABytes, &tmB, fb, kb * kTcBlockK, n_tile * BN);
                            }
                        }
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (!(PAIR && crank != 0)) {                        // pair: only the leader issues (for both CTAs)
            // The whole warp walks the schedule and waits; one elected lane issues (see elect_one() for why).
            constexpr uint32_t idesc = PAIR ? make_idesc_mn(2 * kTcBlockM, BN) : make_idesc(BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            bool alive = true;
            for (int tile = tile0; tile < num_tiles && alive; tile += tile_step) {
                if (PAIR ? !mbar_wait_cluster(tempty0 + 8 * acc, acc_phase ^ 1, p.err_flag, 2)
                         : !mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, p.err_flag, 2)) { alive = false; break; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (!mbar_wait(full0 + 8 * stage, phase, p.err_flag, 3)) { alive = false; break; }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = make_sw128_desc(smem_base + stage * Cfg::kStageBytes);
                        const uint64_t bdesc = make_sw128_desc(smem_base + stage * Cfg::kStageBytes + Cfg::kABytes);
#pragma unroll
                        for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k) {         // +32 bytes of K per MMA = +2 in the address field
                            if (PAIR) umma_f16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            else umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        if (PAIR) umma_commit_2sm_mc(empty0 + 8 * stage, static_cast<uint16_t>(3));  // frees the stage in both CTAs
                        else if (CL == 2) umma_commit_mc(empty0 + 8 * stage, static_cast<uint16_t>(3));   // both CTAs' producers wait for both consumers
                        else umma_commit(empty0 + 8 * stage);     // smem slot free once these MMAs retire
                        if (kb == num_kb - 1) {
                            if (PAIR) umma_commit_2sm_mc(tfull0 + 8 * acc, static_cast<uint16_t>(3));       // both CTAs' epilogues
                            else umma_commit(tfull0 + 8 * acc);      // accumulator ready for the epilogue
                        }
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (!alive) break;
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== residual producer (TC_STORE_RES with a bf16 residual) =====================
        if (MODE == TC_STORE_RES && has_res) {                 // whole warp + one elected lane per copy, as above
            uint32_t g = 0;
            bool alive = true;
            for (int tile = tile0; tile < num_tiles && alive; tile += tile_step) {
                const int n_tile = tile % p.num_n_tiles;
                const int m_tile = m_of(tile);
                for (int c = 0; c < BN / kChunkCols; ++c, ++g) {
                    constexpr uint32_t R = Cfg::kNumR > 0 ? Cfg::kNumR : 1;
                    const uint32_t slot = g % R, use = g / R;
                    if (!mbar_wait(cempty0 + 8 * slot, (use & 1) ^ 1, p.err_flag, 5)) { alive = false; break; }
                    if (elect_one()) {
                        mbar_arrive_expect_tx(cfull0 + 8 * slot, kChunkBytes);
                        tma_load_2d(rbuf0 + slot * kChunkBytes, &tmR, cfull0 + 8 * slot, n_tile * BN + c * kChunkCols,
                                    m_tile * p.tile_rows);
                    }
                }
            }
        }
        __syncwarp();
    } else if (MODE == TC_DIRECT && warp >= 4 && warp < 8) {
        // ===================== direct epilogue (4 warps, one TMEM lane quarter each) =====================
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            const int n_tile = tile % p.num_n_tiles;
            const int m_tile = m_of(tile);
            if (!mbar_wait(tfull0 + 8 * acc, acc_phase, p.err_flag, 4)) break;
            tc_fence_after();
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            const int row = m_tile * p.tile_rows + quarter * 32 + lane;
            const bool row_ok = row < p.ep.M && quarter * 32 < p.tile_rows;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r0[16], r1[16];
                tmem_ld16(t0 + c, r0);
                if (c + 16 < BN) tmem_ld16(t0 + c + 16, r1);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r0[i]);
                if (row_ok) epilogue_16(p.ep, bank, p.bias_in_params != 0, v, row, n_tile * BN + c);
                if (c + 16 < BN) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r1[i]);
                    if (row_ok) epilogue_16(p.ep, bank, p.bias_in_params != 0, v, row, n_tile * BN + c + 16);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * acc, 0));
                else mbar_arrive(tempty0 + 8 * acc);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (MODE != TC_DIRECT && warp >= 4) {
        // ===================== TMA-store epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
        // Each 64-column chunk of the tile is staged as [128 rows][128 B] (128B swizzle); a quarter's 32-row slab is
        // written by its two warps (32 columns each) and stored by lane 0 of the first one.
        const int quarter = warp & 3;
        const int half = (warp - 4) >> 2;
        const bool issuer = half == 0 && lane == 0;
        const bool relu = p.ep.act == ACT_RELU;
        const bool gelu = p.ep.act == ACT_GELU;
        int acc = 0;
        uint32_t acc_phase = 0;
        bool alive = true;
        uint32_t g = 0;                                    // running chunk counter
        for (int tile = tile0; tile < num_tiles && alive; tile += tile_step) {
            const int n_tile = tile % p.num_n_tiles;
            const int m_tile = m_of(tile);
            if (!mbar_wait(tfull0 + 8 * acc, acc_phase, p.err_flag, 4)) { alive = false; break; }
            tc_fence_after();
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * 32;
#pragma unroll 1
            for (int c = 0; c < BN / kChunkCols; ++c, ++g) {
                constexpr uint32_t S = Cfg::kNumS > 0 ? Cfg::kNumS : 1, R = Cfg::kNumR > 0 ? Cfg::kNumR : 1;
                const uint32_t sslot = g % S, rslot = g % R;
                if (issuer) bulk_wait_read<(Cfg::kNumS > 0 ? Cfg::kNumS - 1 : 0)>();   // the slab's previous store left smem
                named_bar_sync(1 + quarter, 64);
                uint32_t r[32];
                tmem_ld32(t0 + c * kChunkCols, r);
                const int col0 = n_tile * BN + c * kChunkCols + half * 32;
                float4 bq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    bq[j] = p.bias_in_params ? *reinterpret_cast<const float4*>(&bank.v[col0 + 4 * j])
                                             : __ldg(reinterpret_cast<const float4*>(p.ep.bias + col0) + j);
                const uint32_t slab_off = Cfg::kStages * Cfg::kStageBytes + quarter * (32 * 128) + lane * 128;
                uint4 rq[4];
                if (has_res) {
                    if (!mbar_wait(cfull0 + 8 * rslot, (g / R) & 1, p.err_flag, 6)) { alive = false; }
                    const uint8_t* rrow = smem + slab_off + (Cfg::kNumS + rslot) * kChunkBytes;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rq[j] = *reinterpret_cast<const uint4*>(rrow + ((static_cast<uint32_t>(half * 4 + j) ^ (lane & 7)) << 4));
                }
                tmem_ld_wait();
                if (c == BN / kChunkCols - 1) {                // accumulator drained: hand TMEM back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * acc, 0));
                        else mbar_arrive(tempty0 + 8 * acc);
                    }
                }
                uint8_t* srow = smem + slab_off + sslot * kChunkBytes;
#pragma unroll
                for (int j = 0; j < 4; ++j) {                  // 4 x (8 columns = 16 bytes)
                    float2 v0 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])), make_float2(bq[2 * j].x, bq[2 * j].y));
                    float2 v1 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), make_float2(bq[2 * j].z, bq[2 * j].w));
                    float2 v2 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), make_float2(bq[2 * j + 1].x, bq[2 * j + 1].y));
                    float2 v3 = __fadd2_rn(make_float2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])), make_float2(bq[2 * j + 1].z, bq[2 * j + 1].w));
                    if (has_res) {
                        v0 = __fadd2_rn(v0, bf16x2_to_f2(rq[j].x)); v1 = __fadd2_rn(v1, bf16x2_to_f2(rq[j].y));
                        v2 = __fadd2_rn(v2, bf16x2_to_f2(rq[j].z)); v3 = __fadd2_rn(v3, bf16x2_to_f2(rq[j].w));
                    }
                    if (gelu) {
                        v0.x = gelu_erf(v0.x); v0.y = gelu_erf(v0.y); v1.x = gelu_erf(v1.x); v1.y = gelu_erf(v1.y);
                        v2.x = gelu_erf(v2.x); v2.y = gelu_erf(v2.y); v3.x = gelu_erf(v3.x); v3.y = gelu_erf(v3.y);
                    }
                    uint4 o;
                    o.x = cvt_bf16x2(v0.x, v0.y, relu); o.y = cvt_bf16x2(v1.x, v1.y, relu);
                    o.z = cvt_bf16x2(v2.x, v2.y, relu); o.w = cvt_bf16x2(v3.x, v3.y, relu);
                    *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(half * 4 + j) ^ (lane & 7)) << 4)) = o;   // 128B swizzle
                }
                fence_async_smem();                        // generic-proxy smem writes -> visible to the TMA engine
                if (has_res) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(cempty0 + 8 * rslot);          // residual slot consumed by this warp
                }
                named_bar_sync(1 + quarter, 64);           // both column halves of the slab are written
                if (issuer) {
                    if (quarter * 32 < p.tile_rows)        // (8 x 8 maps: rows 64..127 of the MMA tile hold nothing)
                        tma_store_2d(&tmC, sbuf0 + sslot * kChunkBytes + quarter * (32 * 128), n_tile * BN + c * kChunkCols,
                                     m_tile * p.tile_rows + quarter * 32);
                    bulk_commit();
                }
                if (!alive) break;
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (issuer) bulk_wait_read<0>();                   // staged data must stay valid until every store has read it
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (CL >= 2) cluster_sync_all();   // the peer may still multicast into this CTA's shared memory / arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                         : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

template <int BN, int MODE>
static int set_attr() {
    HMV_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TcCfg<BN, MODE>::kSmemBytes));
    if (BN == 256) {
        HMV_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, MODE, BN == 256 ? 2 : 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<BN, MODE>::kSmemBytes));
        HMV_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, MODE, BN == 256 ? 4 : 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<BN, MODE, BN == 256>::kSmemBytes));
    }
    return 0;
}

int tc_init() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HMV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    HMV_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    if (set_attr<32, TC_DIRECT>() || set_attr<64, TC_DIRECT>() || set_attr<128, TC_DIRECT>() || set_attr<176, TC_DIRECT>() ||
        set_attr<192, TC_DIRECT>() || set_attr<256, TC_DIRECT>() ||
        set_attr<64, TC_STORE>() || set_attr<128, TC_STORE>() || set_attr<192, TC_STORE>() || set_attr<256, TC_STORE>() ||
        set_attr<64, TC_STORE_DEEP>() || set_attr<128, TC_STORE_DEEP>() || set_attr<192, TC_STORE_DEEP>() || set_attr<256, TC_STORE_DEEP>() ||
        set_attr<64, TC_STORE_RES>() || set_attr<128, TC_STORE_RES>() || set_attr<192, TC_STORE_RES>() || set_attr<256, TC_STORE_RES>())
        return 1;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

static int encode(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box) {
    HMV_CHECK(g_encode != nullptr, "tc_init() was not called");
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) s[i] = strides[i];
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, s, b, e,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::string m = "cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ") rank " +
                        std::to_string(rank) + " dims";
        for (int i = 0; i < rank; ++i) m += " " + std::to_string(dims[i]);
        m += " strides";
        for (int i = 0; i < rank - 1; ++i) m += " " + std::to_string(strides[i]);
        m += " box";
        for (int i = 0; i < rank; ++i) m += " " + std::to_string(box[i]);
        set_error(m);
        return 1;
    }
    return 0;
}

int tc_make_tmap_act(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                     const uint32_t box[5]) {
    return encode(out, base, 5, dims, strides_bytes, box);
}

int tc_make_tmap_wgt(CUtensorMap* out, const void* base, uint64_t k_total, uint64_t n_alloc, int bn) {
    const uint64_t dims[2] = {k_total, n_alloc};
    const uint64_t strides[1] = {k_total * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kTcBlockK), static_cast<uint32_t>(bn)};
    return encode(out, base, 2, dims, strides, box);
}

int tc_make_tmap_out(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, int box_rows) {
    const uint64_t dims[2] = {cols, rows};
    const uint64_t strides[1] = {cols * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kChunkCols), static_cast<uint32_t>(box_rows)};
    return encode(out, base, 2, dims, strides, box);
}

int tc_pick_bn(int n) {
    if (n <= 32) return 32;
    if (n <= 64) return 64;
    if (n <= 128) return 128;
    if (n % 256 == 0) return 256;
    if (n % 192 == 0) return 192;   // HRNet-w40 widths padded: 160 -> 192, 320 -> 2 x 192 (the store clips at the tensor's 320 columns)
    if (n % 176 == 0) return 176;   // 528 = 3 x 176 (d_model 524 padded)
    if (n % 128 == 0) return 128;
    return 0;
}

template <int BN, int MODE>
static int launch_bn(const TcLaunch& l, int num_sms, cudaStream_t stream) {
    if constexpr (BN == 256) {
        if (l.cluster == 4) {          // CTA pairs issuing cta_group::2 MMAs (l.tmB has a BN/2-row box)
            static std::map<std::pair<int, int>, int> cache4;
            int dev = 0;
            HMV_CUDA(cudaGetDevice(&dev));
            int& max_clusters = cache4.emplace(std::make_pair(dev, num_sms), -1).first->second;
            if (max_clusters < 0) {
                cudaLaunchConfig_t qc{};
                qc.gridDim = dim3(2 * (num_sms / 2)); qc.blockDim = dim3(kTcThreads); qc.dynamicSmemBytes = TcCfg<BN, MODE, true>::kSmemBytes;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                qc.attrs = qa; qc.numAttrs = 1;
                int n = 0;
                HMV_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_gemm_tc_kernel<BN, MODE, 4>, &qc));
                HMV_CHECK(n > 0, "no CTA pair of the conv kernel fits on this device");
                max_clusters = n < num_sms / 2 ? n : num_sms / 2;
            }
            const int items = ((l.p.num_m_tiles + 1) / 2) * l.p.num_n_tiles;
            const int clusters = items < max_clusters ? items : max_clusters;
            HMV_CUDA(launch_kernel_cluster(conv_gemm_tc_kernel<BN, MODE, 4>, dim3(2 * clusters), dim3(kTcThreads), 2,
                                           TcCfg<BN, MODE, true>::kSmemBytes, stream, l.tmA, l.tmB, l.tmC, l.tmR, l.p, l.bank));
            return 0;
        }
        if (l.cluster == 2) {          // pairs of CTAs sharing multicast weight tiles (l.tmB has a BN/2-row box)
            // the persistent grid must be co-resident: clusters are placed inside one GPC, so fewer than num_sms / 2 may fit
            static std::map<std::pair<int, int>, int> cache;               // per (device, persistent-grid cap)
            int dev = 0;
            HMV_CUDA(cudaGetDevice(&dev));
            int& max_clusters = cache.emplace(std::make_pair(dev, num_sms), -1).first->second;
            if (max_clusters < 0) {
                cudaLaunchConfig_t qc{};
                qc.gridDim = dim3(2 * (num_sms / 2)); qc.blockDim = dim3(kTcThreads); qc.dynamicSmemBytes = TcCfg<BN, MODE>::kSmemBytes;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                qc.attrs = qa; qc.numAttrs = 1;
                int n = 0;
                HMV_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_gemm_tc_kernel<BN, MODE, 2>, &qc));
                HMV_CHECK(n > 0, "no 2-CTA cluster of the conv kernel fits on this device");
                max_clusters = n < num_sms / 2 ? n : num_sms / 2;
            }
            const int items = ((l.p.num_m_tiles + 1) / 2) * l.p.num_n_tiles;
            const int clusters = items < max_clusters ? items : max_clusters;
            HMV_CUDA(launch_kernel_cluster(conv_gemm_tc_kernel<BN, MODE, 2>, dim3(2 * clusters), dim3(kTcThreads), 2,
                                           TcCfg<BN, MODE>::kSmemBytes, stream, l.tmA, l.tmB, l.tmC, l.tmR, l.p, l.bank));
            return 0;
        }
    }
    const int tiles = l.p.num_m_tiles * l.p.num_n_tiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    HMV_CUDA(launch_kernel(conv_gemm_tc_kernel<BN, MODE, 1>, dim3(grid), dim3(kTcThreads), TcCfg<BN, MODE>::kSmemBytes, stream, l.tmA, l.tmB,
                           l.tmC, l.tmR, l.p, l.bank));
    return 0;
}

template <int BN>
static int launch_mode(const TcLaunch& l, int num_sms, cudaStream_t stream) {
    if constexpr (BN % kChunkCols == 0) {
        if (l.mode == TC_STORE) return launch_bn<BN, TC_STORE>(l, num_sms, stream);
        if (l.mode == TC_STORE_DEEP) return launch_bn<BN, TC_STORE_DEEP>(l, num_sms, stream);
        if (l.mode == TC_STORE_RES) return launch_bn<BN, TC_STORE_RES>(l, num_sms, stream);
    }
    if (l.mode != TC_DIRECT) {
        set_error("tc_launch: TMA-store epilogue needs a tile width that is a multiple of 64");
        return 1;
    }
    return launch_bn<BN, TC_DIRECT>(l, num_sms, stream);
}

int tc_launch(const TcLaunch& l, int num_sms, cudaStream_t stream) {
    if (l.p.num_m_tiles <= 0 || l.p.num_n_tiles <= 0) return 0;
    switch (l.bn) {
        case 32: return launch_mode<32>(l, num_sms, stream);
        case 64: return launch_mode<64>(l, num_sms, stream);
        case 128: return launch_mode<128>(l, num_sms, stream);
        case 176: return launch_mode<176>(l, num_sms, stream);
        case 192: return launch_mode<192>(l, num_sms, stream);
        case 256: return launch_mode<256>(l, num_sms, stream);
    }
    set_error("tc_launch: unsupported BN " + std::to_string(l.bn));
    return 1;
}

}  // namespace hmv
