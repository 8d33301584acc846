// conv3x3_tc2 — the CTA-pair (tcgen05 cta_group::2) form of conv3x3_tc: same math, same canvas / C8 layout, same
// epilogue semantics, but two SMs of a TPC cooperate on one 256-row work item:
//   * UMMA M = 256 (128 rows from each CTA), N = BN; each CTA stages only ITS HALF of every weight block
//     ([BN/2][64], 8 KB): L2 -> SM weight traffic and the B-operand shared-memory reads per SM are halved (the
//     single-CTA kernel was L2-bandwidth bound on weight streaming: ~25 B/clk/SM of a ~24 B/clk/SM budget);
//   * each CTA owns ONE 128-row tile, so its 512 TMEM columns hold main + lo accumulators DOUBLE-buffered:
//     the drain of item j overlaps the MMAs of item j+1 also on the split-precision rung, and the epilogue per
//     item is half as long (4 warps x 128 rows), which shrinks the exposed tail;
//   * 256 threads per CTA: every epilogue thread can hold its 128-column accumulator row without setmaxnreg.
// Protocol (leader = cluster rank 0 issues every MMA): `full` barriers live in the leader and receive the TMA
// transaction bytes of BOTH CTAs (the peer's loads signal the leader's barrier through its shared::cluster
// address); `empty` / `tmem_full` barriers are local to each CTA and are signalled by multicast tcgen05.commit;
// `tmem_empty` lives in the leader and collects the epilogue warps of both CTAs (remote mbarrier.arrive).
// Reference lines replaced: see conv3x3_tc.cuh.
#pragma once
#include "common.cuh"
#include "conv3x3_tc.cuh"
#include "ptx.cuh"

namespace sb {

// Split rung: 2 activation-slab buffers and a 12-stage weight ring measured best (profiles/r01s2_ring_depth.log: +2.4 % on
// 10bx128, +3.5 % on 20bx256 over 3 slabs + 8 stages; 16 stages: +2 % / +3 %) — the weight stream is what the MMA
// issuer waits for, a third slab buffer is not.  Overridable for experiments.
#ifndef SB_TC2_NA
#define SB_TC2_NA 2      // activation-slab buffers (split rung)
#endif
#ifndef SB_TC2_NB
#define SB_TC2_NB 12     // weight-stage ring depth (split rung, <= 18)
#endif
constexpr int kTileRows2 = 128;                              // rows per CTA per item
constexpr int kSlabRows2 = kTileRows2 + 2 * kSlabMargin;     // 176

template <bool SPLIT>
struct Conv2Cfg {
    static constexpr int kParts = SPLIT ? 2 : 1;
    static constexpr int kSlabPartBytes = kSlabRows2 * 128;          // [8 chunks][176 rows][16 B]
    static constexpr int kSlabBytes = kParts * kSlabPartBytes;
    static constexpr int kNumSlabs = SPLIT ? SB_TC2_NA : 3;
    static constexpr int kBStageBytes = 64 * 128;                    // this CTA's half: up to 64 rows x 64 fp16
    // fp16 rung: 18 stages x 8 KB hold ALL of this CTA's weights of a C <= 128 layer (9 taps x 2 k-halves): they are
    // then loaded once per launch and stay resident (ConvParams::resident), instead of being re-streamed for every
    // work item (147 KB per item per CTA, the L2 -> SM bound of this rung).  The split rung has no room for that.
    static constexpr int kNumBStages = SPLIT ? SB_TC2_NB : 18;
    static constexpr int kOffB = kNumSlabs * kSlabBytes;
    static constexpr int kOffBar = kOffB + kNumBStages * kBStageBytes;
    static constexpr int kOffBias = kOffBar + 512;
    static constexpr int kSmemBytes = kOffBias + kMaxConvWidth * 4 + 1024;
    static constexpr int kTmemCols = 512;
    // Epilogue column parts per TMEM lane quadrant = epilogue warps per scheduler.  2 on both rungs: 4 parts (16 epilogue
    // warps, 640 threads) measured 2 % SLOWER on the fp16 rung (265 k vs 271 k evals/s) and 13 % slower on 1x1-conv
    // heavy towers — after the resident-weight change that rung is bound by the tensor core's operand reads from
    // shared memory (~95 cycles per N=128 MMA instead of 64), not by the epilogue.
    static constexpr int kEpiParts = 2;
    static constexpr int kThreads = 128 + kEpiParts * 128;
    static constexpr int kMaxGroups = 128 / kEpiParts / 16;   // 16-column groups per epilogue thread (BN <= 128)
    static_assert(kNumBStages <= 18 && kNumSlabs <= 4, "barrier block layout");
    static_assert(kSlabPartBytes % 1024 == 0, "slab parts must keep the weight stages 1024-byte aligned");
    static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
};

// ---- cluster / cta_group::2 PTX ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of THIS kernel's layout) inside CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // Default (.release.cta) semantics, as CUTLASS' ClusterBarrier::arrive(cta_id): the explicit .release.cluster form
    // compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. every epilogue warp waited once per work item for ALL its outstanding
    // global loads / stores (8.7 % of the kernel's stall samples, profiles/r01s2_ncu_source_*.txt).  The only data
    // the MMA issuer must see ordered before this arrive are our tcgen05.ld reads of the accumulator, which
    // tcgen05.wait::ld + tcgen05.fence::before_thread_sync already order.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion bytes are credited to a barrier given by its shared::cluster address
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}

// Work units.  An item is (256-row super tile st, N tile nt) at N = BN.  To cut the wave-quantisation loss of a
// persistent grid (401 items over 74 CTA pairs = 5.42 -> 6 waves at batch 256), the host may turn the items of
// the last, partial wave into TWO half units each (N = BN/2, same rows): units [0, n_full) are whole items, units
// n_full + 2k, n_full + 2k + 1 are the N-halves of item n_full + k, so that the last wave costs half a wave.
struct ConvUnit {
    int st, n0, bn;
};
__device__ __forceinline__ ConvUnit conv_unit(int u, const ConvParams& p) {
    ConvUnit w;
    if (u < p.n_full) {
        w.st = u / p.n_ntiles;
        w.bn = p.bn;
        w.n0 = (u % p.n_ntiles) * p.bn;
    } else {
        const int v = u - p.n_full, item = p.n_full + (v >> 1);
        w.st = item / p.n_ntiles;
        w.bn = p.bn >> 1;
        w.n0 = (item % p.n_ntiles) * p.bn + (v & 1) * w.bn;
    }
    return w;
}

template <bool SPLIT, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Conv2Cfg<SPLIT>::kThreads, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                   const __grid_constant__ CUtensorMap tmWq_hi, const __grid_constant__ CUtensorMap tmWq_lo,
                   const ConvParams p) {
    using Cfg = Conv2Cfg<SPLIT>;
    const int KH = p.kh, BN = p.bn;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t slab_addr = smem_base;
    const uint32_t bst_addr = smem_base + Cfg::kOffB;
    const uint32_t bar_addr = smem_base + Cfg::kOffBar;
    const uint32_t a_full = bar_addr + 0, a_empty = bar_addr + 32;                // [3] each
    const uint32_t tmem_full = bar_addr + 64, tmem_empty = bar_addr + 80;         // [2] each
    const uint32_t b_full = bar_addr + 128, b_empty = bar_addr + 288;             // [<= 18] each
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + Cfg::kOffBar + 448);
    float* sbias = reinterpret_cast<float*>(smem_gen + Cfg::kOffBias);
    constexpr uint32_t kNA = Cfg::kNumSlabs;
    // resident weights: the ring is exactly one item's worth of stages, filled once, never recycled
    const bool resident = p.resident != 0;
    // a weight stage holds this CTA's N-half of one [BN][64] block: 8 KB slots for BN <= 128, 16 KB (two slots) for the
    // N = 256 tiles of the fp16 rung, which therefore has half as many stages in the same ring
    const uint32_t stage_stride = BN > 128 ? 2u * Cfg::kBStageBytes : (uint32_t)Cfg::kBStageBytes;
    const uint32_t kNB = resident ? (uint32_t)(p.kh * p.ntaps * Cfg::kParts)
                                  : (uint32_t)(Cfg::kNumBStages * Cfg::kBStageBytes) / stage_stride;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_items = p.n_units;   // whole items + half units of the tail wave (conv_unit)

    for (int i = threadIdx.x; i < p.cout && i < kMaxConvWidth; i += blockDim.x) sbias[i] = p.bias[i];

    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)kNA; ++i) {
            mbar_init(a_full + 8 * i, 1);    // leader's own arrive.expect_tx (bytes of both CTAs)
            mbar_init(a_empty + 8 * i, 1);   // one multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full + 8 * i, 1);
            mbar_init(tmem_empty + 8 * i, 2 * 4 * Cfg::kEpiParts);  // every epilogue warp of both CTAs
        }
        for (int i = 0; i < Cfg::kNumBStages; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmW_hi);
        if (SPLIT) {
            tma_prefetch_desc(&tmA_lo);
            tma_prefetch_desc(&tmW_lo);
        }
        if (p.n_units > p.n_full) {
            tma_prefetch_desc(&tmWq_hi);
            if (SPLIT) tma_prefetch_desc(&tmWq_lo);
        }
    }
    if (warp == 2) {
        tmem_alloc2(smem_u32(tmem_ptr_smem), Cfg::kTmemCols);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // barriers of both CTAs are initialised before any remote arrive / TMA credit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) pdl_launch_dependents();   // the next layer may start its prologue whenever SMs free up

    if (warp == 3) {
        // ===================== activation-slab producer (own 128-row tile, +-24 rows) =====================
        pdl_wait();   // the slabs are the previous layer's output
        uint32_t it = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int st = conv_unit(item, p).st;
            const int row_lo = kGuardRows + st * kSuperRows + (int)rank * kTileRows2 - kSlabMargin;
            for (int h = 0; h < KH; ++h, ++it) {
                const uint32_t s = it % kNA, ph = (it / kNA) & 1u;
                if ((p.dbg & 64) && it >= kNA) continue;   // ablation: no slab traffic after the first fills
                mbar_wait(a_empty + 8 * s, ph ^ 1u, p.err, 1);
                if (elect_one()) {
                    const uint32_t full0 = mapa_u32(a_full + 8 * s, 0);
                    if (leader) mbar_arrive_expect_tx(a_full + 8 * s, 2u * Cfg::kSlabBytes);
#pragma unroll
                    for (int part = 0; part < Cfg::kParts; ++part) {
                        tma2_load_3d(slab_addr + s * Cfg::kSlabBytes + part * Cfg::kSlabPartBytes, part ? &tmA_lo : &tmA_hi, 0,
                                     row_lo, h * 8, full0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 0) {
        // ===================== weight-stage producer (own N-half of every block) =====================
        uint32_t s = 0, ph = 0;   // ring position and phase, advanced incrementally (kNB is a run-time value)
        for (int item = cluster_id; item < n_items && !(p.dbg & 16); item += n_clusters) {
            if (resident && item != cluster_id) break;     // everything is already in shared memory
            const ConvUnit w = conv_unit(item, p);
            const int HB = w.bn >> 1;                      // weight rows staged by this CTA
            const bool whole = w.bn == BN;
            for (int h = 0; h < KH; ++h) {
                for (int tap = 0; tap < p.ntaps; ++tap) {
#pragma unroll
                    for (int part = 0; part < Cfg::kParts; ++part) {
                        mbar_wait(b_empty + 8 * s, ph ^ 1u, p.err, 2);
                        if (elect_one()) {
                            const uint32_t full0 = mapa_u32(b_full + 8 * s, 0);
                            if (leader) mbar_arrive_expect_tx(b_full + 8 * s, 2u * (uint32_t)HB * 128u);
                            tma2_load_2d(bst_addr + s * stage_stride,
                                         whole ? (part ? &tmW_lo : &tmW_hi) : (part ? &tmWq_lo : &tmWq_hi),
                                         tap * (KH * 64) + h * 64, w.n0 + (int)rank * HB, full0);
                        }
                        __syncwarp();
                        if (++s == kNB) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===================== MMA issuer (leader CTA only) =====================
            // TMEM columns per CTA: stage as -> main at as*2*BN, lo at as*2*BN + BN (fp16 rung: main at as*BN).
            const uint32_t idesc_whole = umma_idesc_f16(256, BN), idesc_half = umma_idesc_f16(256, BN >> 1);
            constexpr uint64_t kAStep = 2 * kSlabRows2 * 16 / 16;   // one K=16 step = two channel chunks
            const bool stats = p.stats != nullptr;
            uint32_t a_it = 0, j = 0;
            uint32_t bs = 0, bph = 0;   // weight ring position and phase (incremental: kNB is a run-time value)
            long long t_wait_tmem = 0, t_wait_slab = 0, t_wait_b = 0, t0 = 0;
            const long long t_begin = stats ? clock64() : 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, ++j) {
                const uint32_t as = j & 1u, aph = (j >> 1) & 1u;
                const uint32_t idesc = item < p.n_full ? idesc_whole : idesc_half;
                if (stats) t0 = clock64();
                mbar_wait(tmem_empty + 8 * as, aph ^ 1u, p.err, 3);
                if (stats) t_wait_tmem += clock64() - t0;
                tc_fence_after();
                const uint32_t d_main = tmem_base + (SPLIT ? as * 2 * BN : as * BN);
                const uint32_t d_lo = d_main + BN;
                for (int h = 0; h < KH; ++h, ++a_it) {
                    const uint32_t s = a_it % kNA, sph = (a_it / kNA) & 1u;
                    if (stats) t0 = clock64();
                    if (!((p.dbg & 64) && a_it >= kNA)) mbar_wait(a_full + 8 * s, sph, p.err, 4);
                    if (stats) t_wait_slab += clock64() - t0;
                    tc_fence_after();
                    const uint32_t a_hi = slab_addr + s * Cfg::kSlabBytes;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int shift = (p.ntaps == 1 || (p.dbg & 8)) ? 0 : (tap / 3 - 1) * p.pitch + (tap % 3 - 1);   // 1 tap = 1x1 convolution
                        const uint32_t first = (h | tap) == 0 ? 0u : 1u;
                        const uint64_t ad0 = umma_desc_nosw(a_hi + (uint32_t)(kSlabMargin + shift) * 16u, kSlabRows2 * 16u, 128u);
                        {   // weights hi x activations hi -> main ; x activations lo -> lo accumulator
                            if (stats) t0 = clock64();
                            if (!(p.dbg & 16) && !(resident && j > 0)) mbar_wait(b_full + 8 * bs, bph, p.err, 5);
                            if (stats) t_wait_b += clock64() - t0;
                            tc_fence_after();
                            const uint64_t bd0 = umma_desc_sw128(bst_addr + bs * stage_stride);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma2_f16(d_main, ad0 + kAStep * k, bd0 + 2 * k, idesc, (k == 0) ? first : 1u);
                                    if (SPLIT)
                                        umma2_f16(d_lo, ad0 + (Cfg::kSlabPartBytes >> 4) + kAStep * k, bd0 + 2 * k, idesc,
                                                  (k == 0) ? first : 1u);
                                }
                                if (!resident) umma2_commit_mc(b_empty + 8 * bs, 3);   // resident stages are never recycled
                            }
                            __syncwarp();
                            if (++bs == kNB) {
                                bs = 0;
                                if (!resident) bph ^= 1u;
                            }
                        }
                        if (SPLIT) {   // weights lo x activations hi -> lo accumulator
                            if (stats) t0 = clock64();
                            if (!(p.dbg & 16) && !(resident && j > 0)) mbar_wait(b_full + 8 * bs, bph, p.err, 6);
                            if (stats) t_wait_b += clock64() - t0;
                            tc_fence_after();
                            const uint64_t bd0 = umma_desc_sw128(bst_addr + bs * stage_stride);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma2_f16(d_lo, ad0 + kAStep * k, bd0 + 2 * k, idesc, 1u);
                                if (!resident) umma2_commit_mc(b_empty + 8 * bs, 3);   // resident stages are never recycled
                            }
                            __syncwarp();
                            if (++bs == kNB) {
                                bs = 0;
                                if (!resident) bph ^= 1u;
                            }
                        }
                    }
                    if (elect_one()) umma2_commit_mc(a_empty + 8 * s, 3);
                    __syncwarp();
                }
                if (elect_one()) umma2_commit_mc(tmem_full + 8 * as, 3);
                __syncwarp();
            }
            if (stats && lane == 0) {
                long long* st = p.stats + (size_t)cluster_id * 8;
                st[0] = clock64() - t_begin;
                st[1] = t_wait_tmem;
                st[2] = t_wait_slab;
                st[3] = t_wait_b;
                st[6] = j;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 4 * kEpiParts warps; warp%4 = TMEM lane quadrant, (warp-4)/4 = column part =====================
        // kEpiParts warps per scheduler: each thread owns 1/kEpiParts of an accumulator row (<= 64 / 32 registers), so
        // the dependent ALU/MUFU chains of one warp are hidden behind the other.
        const int q = warp & 3;
        const int part = (warp - 4) >> 2;                      // which column part of the accumulator row
        const bool stats = p.stats != nullptr;
        uint32_t j = 0;
        long long t_wait_full = 0, t_drain = 0;
        const long long t_begin = stats ? clock64() : 0;
        const uint32_t empty0 = mapa_u32(tmem_empty, 0);   // leader's tmem_empty[0]; [1] is +8
        pdl_wait();   // residual reads, and our stores may overwrite a buffer the previous layer still reads
        for (int item = cluster_id; item < n_items; item += n_clusters, ++j) {
            const ConvUnit w = conv_unit(item, p);
            const int st = w.st;
            // columns of this thread: w.bn split into kEpiParts runs rounded to the 16-column ld granule (may be 0);
            // a run longer than 64 columns (N = 256 tiles of the fp16 rung) is processed in passes of 64
            const int c_run = ((w.bn + Cfg::kEpiParts - 1) / Cfg::kEpiParts + 15) & ~15;
            const int cbase = min(part * c_run, w.bn);
            const int HC = min(c_run, w.bn - cbase);
            const int n_pass = max(1, (HC + 63) >> 6);
            const uint32_t as = j & 1u, aph = (j >> 1) & 1u;

            const int row = kGuardRows + st * kSuperRows + (int)rank * kTileRows2 + q * 32 + lane;
            const bool live = p.mask[row] != 0;
            const bool has_res = live && p.res_hi != nullptr && !(p.dbg & 1);
            const size_t chunk_stride = (size_t)p.rows * 8;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);

            for (int pass = 0; pass < n_pass; ++pass) {
                const int pbase = cbase + pass * 64;            // first column of this pass inside the N tile
                const int PC = min(64, HC - pass * 64);         // columns of this pass (multiple of 16, may be 0)
                // Everything the epilogue needs from global memory is requested BEFORE waiting for the accumulators:
                // the mask byte and the residual pieces of the first 16-column group; the pieces of group g+1 are
                // requested while group g is computed (a load issued at its point of use stalled ~1 us per group).
                const size_t off = act_index(row, w.n0 + pbase, p.rows);
                uint4 rh[2][2], rl[2][2];
                rh[0][0] = rh[0][1] = rh[1][0] = rh[1][1] = make_uint4(0u, 0u, 0u, 0u);
                rl[0][0] = rl[0][1] = rl[1][0] = rl[1][1] = make_uint4(0u, 0u, 0u, 0u);
                if (has_res && PC > 0) {
                    rh[0][0] = *reinterpret_cast<const uint4*>(p.res_hi + off);
                    rh[0][1] = *reinterpret_cast<const uint4*>(p.res_hi + off + chunk_stride);
                    if (SPLIT) {
                        rl[0][0] = *reinterpret_cast<const uint4*>(p.res_lo + off);
                        rl[0][1] = *reinterpret_cast<const uint4*>(p.res_lo + off + chunk_stride);
                    }
                }
                if (pass == 0) {
                    const long long t0 = stats ? clock64() : 0;
                    mbar_wait(tmem_full + 8 * as, aph, p.err, 7);
                    if (stats) t_wait_full += clock64() - t0;
                    tc_fence_after();
                }
                const uint32_t t_main = lane_base + (SPLIT ? as * 2 * BN : as * BN) + pbase;
                const uint32_t t_lo = t_main + BN;
                float acc[Cfg::kMaxGroups * 16];
                const long long t_d0 = stats ? clock64() : 0;
#pragma unroll
                for (int g = 0; g < Cfg::kMaxGroups; ++g) {
                    if (g * 16 < PC) {
                        uint32_t r[16];
                        tmem_ld16(t_main + g * 16, r);
                        if (SPLIT) {
                            uint32_t r2[16];
                            tmem_ld16(t_lo + g * 16, r2);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[g * 16 + i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]);
                        } else {
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[g * 16 + i] = __uint_as_float(r[i]);
                        }
                    }
                }
                if (pass == n_pass - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(empty0 + 8 * as);   // accumulator stage is free again
                }
                if (stats) t_drain += clock64() - t_d0;

#pragma unroll
                for (int g = 0; g < Cfg::kMaxGroups; ++g) {
                    if (g * 16 < PC) {
                        const int c0 = g * 16;
                        const size_t o0 = off + (size_t)(c0 >> 3) * chunk_stride, o1 = o0 + chunk_stride;
                        if (has_res && (g + 1) * 16 < PC) {   // next group's residual pieces
                            const size_t n0 = o0 + 2 * chunk_stride, n1 = n0 + chunk_stride;
                            rh[(g + 1) & 1][0] = *reinterpret_cast<const uint4*>(p.res_hi + n0);
                            rh[(g + 1) & 1][1] = *reinterpret_cast<const uint4*>(p.res_hi + n1);
                            if (SPLIT) {
                                rl[(g + 1) & 1][0] = *reinterpret_cast<const uint4*>(p.res_lo + n0);
                                rl[(g + 1) & 1][1] = *reinterpret_cast<const uint4*>(p.res_lo + n1);
                            }
                        }
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = acc[c0 + i] + sbias[w.n0 + pbase + c0 + i];
                        if (has_res) {
                            const __half* hh0 = reinterpret_cast<const __half*>(&rh[g & 1][0]);
                            const __half* hh1 = reinterpret_cast<const __half*>(&rh[g & 1][1]);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                v[i] += __half2float(hh0[i]);
                                v[8 + i] += __half2float(hh1[i]);
                            }
                            if (SPLIT) {
                                const __half* ll0 = reinterpret_cast<const __half*>(&rl[g & 1][0]);
                                const __half* ll1 = reinterpret_cast<const __half*>(&rl[g & 1][1]);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    v[i] += __half2float(ll0[i]);
                                    v[8 + i] += __half2float(ll1[i]);
                                }
                            }
                        }
                        uint32_t oh[8], ol[8];
                        if (!(p.dbg & 2)) activate16<ACT>(v);
                        if (!live) {   // select, not multiply: garbage rows may hold NaN
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = 0.f;
                        }
                        split16(v, oh, ol, SPLIT);
                        if ((p.dbg & 1) && oh[0] != 0x12345678u) continue;
                        // cout may be a multiple of 8 only (weight rows are zero-padded to 16): channels >= cout belong to
                        // somebody else (the value half of the head buffer behind a RepLK 1x1) and are not written
                        const int ch0 = w.n0 + pbase + c0;
                        if (ch0 < p.cout) {
                            *reinterpret_cast<uint4*>(p.out_hi + o0) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                            if (SPLIT) *reinterpret_cast<uint4*>(p.out_lo + o0) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                        }
                        if (ch0 + 8 < p.cout) {
                            *reinterpret_cast<uint4*>(p.out_hi + o1) = make_uint4(oh[4], oh[5], oh[6], oh[7]);
                            if (SPLIT) *reinterpret_cast<uint4*>(p.out_lo + o1) = make_uint4(ol[4], ol[5], ol[6], ol[7]);
                        }
                    }
                }
            }
        }
        if (stats && leader && warp == 4 && lane == 0) {
            long long* st = p.stats + (size_t)cluster_id * 8;
            st[4] = t_wait_full;
            st[5] = clock64() - t_begin;
            st[7] = t_drain;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // nobody leaves (or frees TMEM) while the peer may still read this CTA's memory
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

}  // namespace sb
