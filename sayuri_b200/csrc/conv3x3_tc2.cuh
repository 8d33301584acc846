// conv3x3_tc2 — 3x3 same-pad convolution (+bias, +residual, *mask, activation) as an implicit GEMM on the 5th-gen tensor
// cores, CTA-pair (tcgen05 cta_group::2) form: TMA-staged channel-blocked NHWC tiles in shared memory -> tcgen05.mma
// (UMMA M = 256 = 128 rows from each CTA of the pair, N = BN, K = 16 per instruction, fp16 operands, fp32 accumulation in
// TMEM) -> tcgen05.ld epilogue.
//
// Replaces, for the tower and input convolutions, the reference's
//   Convolution<3>::Forward + Im2col          /root/reference/src/neural/blas/convolution.h:41-125
//   (or WinogradConvolution3::Forward         /root/reference/src/neural/blas/winograd_convolution3.cc:280-291)
//   followed by AddSpatialBiases::Forward     /root/reference/src/neural/blas/biases.cc:14-45
// and on the reference GPU path im2col/Winograd kernels + cuBLAS/cuDNN + add_spatial
//   (/root/reference/src/neural/cuda/cuda_kernels.cu:37-79,182-239,522-667); with POOL also the pooling pass of the SE unit
//   (GlobalPooling, /root/reference/src/neural/blas/se_unit.cc:9-37; GPU twin cuda_kernels.cu:241-321).
//
// GEMM view: M = canvas rows (pixels of all samples, halo cells included), N = Cout, K = taps * Cin.  Because of the
// canvas layout (common.cuh) the A operand of tap (ky,kx) is the SAME row-major tile shifted by (ky-1)*P + (kx-1) rows,
// so one "slab" of 128 + 2*24 rows x 64 channels per CTA is loaded once per work item and k-half and re-used by all 9
// taps through row-shifted UMMA descriptors (9x less activation traffic than per-tap loads).
//   * each CTA stages only ITS HALF of every weight block ([BN/2][64], 8 KB): L2 -> SM weight traffic and the B-operand
//     shared-memory reads per SM are halved against a single-CTA kernel;
//   * 128 + kEpiParts*128 threads per CTA: every epilogue thread holds its part of an accumulator row in registers.
// Protocol (leader = cluster rank 0 issues every MMA): `full` barriers live in the leader and receive the TMA
// transaction bytes of BOTH CTAs (the peer's loads signal the leader's barrier through its shared::cluster address);
// `empty` / `tmem_full` / `lo_full` barriers are local to each CTA and are signalled by multicast tcgen05.commit;
// `tmem_empty` / `lo_empty` live in the leader and collect the epilogue warps of both CTAs (remote mbarrier.arrive).
//
// Precision.  SPLIT = true evaluates x*w as hi*hi + lo*hi + hi*lo with x = x_hi + x_lo, w = w_hi + w_lo (fp16 pairs, ~22
// significant bits): the fp32-faithful rung.  The tensor core adds into its fp32 accumulator with TRUNCATION: every MMA
// into a large accumulator loses a fraction of an ulp TOWARDS ZERO, a systematic shrink of ~1e-8 per MMA that adds up
// coherently over K (144 main MMAs at C = 256) and over the layers of a deep tower (measured, profiles/r02_precision_probe*:
// 20bx256 trunk 5.7e-5 relative = 1.2e-3 absolute, 12x the fp32 noise floor).  Therefore:
//   * the low-order products (lo*hi, hi*lo) have their own accumulator (small values: their truncation is harmless);
//   * the MAIN product is accumulated in CHUNKS of one k-half (64 input channels x 9 taps = 36 MMAs); each chunk starts
//     from zero in one of two TMEM stages and is drained by the epilogue and added in fp32 round-to-nearest in registers,
//     scaled by 1 + the expected truncation loss of a chunk, while the next chunk runs in the other stage.  (Chunks of 3
//     taps or 1 tap are 2x / 4x more accurate still but stall the issuer behind the epilogue: -12 % / -31 %, measured.)
// SPLIT = false uses the hi parts only, one chunk per item (the reference's own --fp16 trade).
//
// Warp roles:  warp 0 : TMA producer for weight stages      warp 1 : tcgen05.mma issuer (leader CTA; converged, elected lane)
//              warp 2 : TMEM allocator, then publisher of    warp 3 : TMA producer for activation slabs (acquires the
//                       the per-tile completion counters              completion counters of its input tiles first)
//              warps 4.. : epilogue (warp%4 = TMEM lane quadrant, (warp-4)/4 = column part of the accumulator row)
// One launch runs a CHAIN of convolutions ("Chains of convolutions in one launch" below); a plain launch is a chain of one.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

struct ConvParams {
    __half* out_hi;
    __half* out_lo;          // unused when !SPLIT
    const __half* res_hi;    // optional residual (same layout as out), nullptr if none
    const __half* res_lo;
    const float* bias;       // [cout padded to the N tile]
    int bias_off;            // first float of this layer's biases in the kernel's shared-memory staging area (all layers of a chain)
    const uint8_t* mask;     // [rows]: 1 = real board cell of its sample, 0 = halo / off-board / padding
    int cout;                // real output channels (bias length)
    int rows;                // R: rows per channel chunk of the C8 activation tensors (out/res)
    int kh;                  // number of 64-channel K blocks per tap = padded Cin / 64
    int bn;                  // UMMA N = output channels per work item (multiple of 16)
    int n_super;             // number of 256-row work items along M
    int n_ntiles;            // cout / bn
    int resident;            // fp16 rung: this CTA's weights fit the stage ring and are loaded once per launch
    int n_full, n_units;     // whole items, and whole items + half units of the tail wave
    int pitch;               // P = N + 1
    int ntaps;               // 9 = 3x3 convolution, 1 = 1x1 convolution (centre tap only)
    int dbg;                 // ablation bits for profiling only: 1 skip stores, 2 skip activation+split math, 8 no tap shifts, 16 no weight stream, 64 no slab stream,
                             // 128 general MMA issue loop also for 3x3 convolutions (A/B against the unrolled one; same MMAs in the same order)
    int chunk_kh;            // SPLIT: k-halves (64 input channels x all taps) per main-accumulator chunk; kh = one chunk per item
    float chunk_scale;       // SPLIT: a drained chunk is multiplied by this (1 + expected truncation loss of a chunk)
    float* pool_part;        // POOL: [group][2][pool_c] sums and maxima of the output over groups of 2^pool_log2 canvas rows
    int pool_log2;           // 2..4; groups never straddle samples (the engine checks kGuardRows and SS are multiples)
    int pool_groups;         // groups covered by the batch; groups beyond are not written
    int pool_c;              // channel stride of pool_part
    // Layer overlap (see "Cross-layer dependencies" below): completion counters per 256-row super tile, in units of
    // (output columns) x (epilogue quadrant warps); a super tile of a launch is complete at 8 * n_ntiles * bn.
    int* done_out;           // counters of THIS launch's output, nullptr = not signalled
    const int* done_in;      // counters of the launch that wrote the input tensor; nullptr = whole-grid dependency (griddepcontrol.wait)
    int done_in_full;        // value of a complete super tile in done_in
    int rot;                 // items are dealt round-robin to the CTA pairs starting at pair (n_pairs - rot) % n_pairs: rotates from layer to layer of a chain
    int* err;                // device int, receives a site code if a barrier wait times out
    long long* stats;        // optional [grid][8] cycle counters (see sb_conv_stats), nullptr = off
};

// Split rung: 2 activation-slab buffers and a 12-stage weight ring measured best, in round 1 (profiles/r01s2_ring_depth.log)
// and again after the MMA issue loop was rewritten (profiles/r02_epilogue_parts.log: 16 stages -4 %, 3 slabs + 6 stages
// -8 % on 10bx128).  Overridable for experiments.
// Epilogue warps per TMEM lane quadrant (= column parts of an accumulator row), a template parameter of the kernel.
// Measured after the MMA issue loop stopped being the limiter (profiles/r02_epilogue_parts.log, 10bx128, batch 256): the
// epilogue of 2 warps per scheduler is latency-bound (tcgen05.ld waits, ex2/rcp chains, residual loads) and takes as long
// per item as the MMAs; 4 parts: split 151 k -> 159 k evals/s, fp16 335 k -> 362 k.  The N = 256 tiles of the fp16 rung
// (C > 128) are tensor-bound and lose 3 % to the extra warps: they keep 2.
#ifndef SB_TC2_EPI_PARTS_SPLIT
#define SB_TC2_EPI_PARTS_SPLIT 4
#endif
#ifndef SB_TC2_EPI_PARTS_FP16
#define SB_TC2_EPI_PARTS_FP16 4    // N <= 128
#endif
#ifndef SB_TC2_EPI_PARTS_FP16_WIDE
#define SB_TC2_EPI_PARTS_FP16_WIDE 2   // N > 128
#endif
#ifndef SB_TC2_NA
#define SB_TC2_NA 2      // activation-slab buffers (split rung)
#endif
#ifndef SB_TC2_NB
#define SB_TC2_NB 12     // weight-stage ring depth (split rung, <= 18)
#endif
constexpr int kChainBiasFloats = 2048;                       // padded output channels of all layers of a chain, at most
constexpr int kTileRows2 = 128;                              // rows per CTA per item
constexpr int kSlabRows2 = kTileRows2 + 2 * kSlabMargin;     // 176

template <bool SPLIT, int PARTS = 2>
struct Conv2Cfg {
    static constexpr int kParts = SPLIT ? 2 : 1;
    static constexpr int kSlabPartBytes = kSlabRows2 * 128;          // [8 chunks][176 rows][16 B]
    static constexpr int kSlabBytes = kParts * kSlabPartBytes;
    static constexpr int kNumSlabs = SPLIT ? SB_TC2_NA : 3;
    static constexpr int kBStageBytes = 64 * 128;                    // this CTA's half: up to 64 rows x 64 fp16
    // fp16 rung: 18 stages x 8 KB hold ALL of this CTA's weights of a C <= 128 layer (9 taps x 2 k-halves): they are
    // then loaded once per layer and stay resident (ConvParams::resident) instead of being re-streamed for every work
    // item (147 KB per item per CTA), and the MMA issuer waits for no weight barrier after the first item of a layer.
    // The split rung has no room for that (288 KB per CTA and layer) and, with the unrolled issuer, no need.
    static constexpr int kNumBStages = SPLIT ? SB_TC2_NB : 18;
    static constexpr int kOffB = kNumSlabs * kSlabBytes;
    static constexpr int kOffBar = kOffB + kNumBStages * kBStageBytes;
    static constexpr int kOffBias = kOffBar + 512;            // biases of all layers of a chain, staged once per launch
    static constexpr int kSmemBytes = kOffBias + kChainBiasFloats * 4 + 1024;   // + slack for the 1024-byte alignment of the base
    static constexpr int kTmemCols = 512;
    // Epilogue column parts per TMEM lane quadrant = epilogue warps per scheduler (see SB_TC2_EPI_PARTS_* above).
    static constexpr int kEpiParts = PARTS;
    static constexpr int kThreads = 128 + kEpiParts * 128;
    static constexpr int kMaxGroups = 128 / kEpiParts / 16;   // 16-column groups per epilogue thread and pass (BN <= 128: one pass)
    static constexpr int kPassCols = kMaxGroups * 16;         // a longer run of columns (N = 256 tiles of the fp16 rung) takes several passes
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

// Reduce-scatter of 16 per-row values over the 2^L lanes of an aligned lane group (recursive halving): step `bit` pairs
// lane and lane ^ bit, each keeps one half of its current index range and receives the partner's values of that half.
// The summation tree of every channel depends only on the lane bits: results do not depend on where in the batch the
// rows lie.  Afterwards lane l holds 16 >> L consecutive channels starting at pool_first(l); v[0 .. 16>>L) are valid.
template <int L>
__device__ __forceinline__ void pool_reduce16(float (&s)[16], float (&m)[16], int lane) {
    int n = 16;
#pragma unroll
    for (int step = 0; step < L; ++step) {
        const int bit = 1 << (L - 1 - step);
        const int half = n >> 1;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < half) {
                const float send_s = upper ? s[i] : s[i + half], keep_s = upper ? s[i + half] : s[i];
                const float send_m = upper ? m[i] : m[i + half], keep_m = upper ? m[i + half] : m[i];
                s[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                m[i] = fmaxf(keep_m, __shfl_xor_sync(0xffffffffu, send_m, bit));
            }
        }
        n = half;
    }
}
__device__ __forceinline__ int pool_first(int lane, int L) {   // first channel (of 16) a lane holds after pool_reduce16<L>
    int first = 0, n = 16;
    for (int step = 0; step < L; ++step) {
        n >>= 1;
        if (lane & (1 << (L - 1 - step))) first += n;
    }
    return first;
}

// ---- Cross-layer dependencies ------------------------------------------------------------------------------------
// A convolution launch whose input was written by the previous convolution launch does not wait for that whole grid
// (griddepcontrol.wait): item st needs rows of the input's super tiles st-1, st, st+1 only (its slab reaches 24 rows into
// the neighbours).  When the epilogue warps of a CTA have stored their parts of an item, the CTA's share is added to
// done[st] with a gpu-scope release (by its publisher warp, or by the epilogue warps themselves: SB_TC2_PUBLISHER_* below);
// the slab producer of the consuming launch acquires done[st-1 .. st+1] before the first TMA load of an item (plus a
// generic->async proxy fence) and then tells the epilogue warps of its CTA, through a counter in shared memory, that the
// residual pieces of the item may be prefetched (the residual tensor is the input of the producing launch: complete for
// these rows by transitivity, the engine enables the mode only then).  With programmatic dependent launch the CTAs of layer l+1 become
// resident as the CTA pairs of layer l run out of items and start on the tiles that are ready: the partial last wave
// (400 items over 74 pairs = 5.4 waves) and the launch ramp of one layer are filled with the next layer's work.
// No deadlock: a dependent grid starts only after EVERY CTA of the primary has started (each triggers at its top), so a
// counter that is waited for is always owned by a resident or finished CTA.  Buffers are re-used across layers
// (x -> t -> u -> t ...): a tile is overwritten by layer l+2 only after layer l+1's tiles st-1 .. st+1 — the only readers
// of those rows — have completed, which is exactly what the slab producer waited for.
__device__ __forceinline__ int ld_acquire_gpu(const int* ptr) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(int* ptr, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// The three super tiles a slab of item st reads (st - 1 and st + 1 where they exist), polled together: one round trip.
// Bounded like mbar_wait: a protocol bug surfaces as a launch failure with a site code, never as a hung GPU.
__device__ __forceinline__ void wait_tiles_done3(const int* done, int st, int n_super, int full, int* err, int site) {
    const int* c0 = done + (st > 0 ? st - 1 : st);
    const int* c1 = done + st;
    const int* c2 = done + (st + 1 < n_super ? st + 1 : st);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        const int v0 = ld_acquire_gpu(c0), v1 = ld_acquire_gpu(c1), v2 = ld_acquire_gpu(c2);
        if (v0 >= full && v1 >= full && v2 >= full) return;
        __nanosleep(64);
    }
    if (err) atomicExch(err, site);
    __threadfence_system();
    __trap();
}

__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void wait_deps_ctr(uint32_t addr, uint32_t want, int* err, int site) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        if (ld_acquire_cta_shared(addr) >= want) return;
    }
    if (err) atomicExch(err, site);
    __threadfence_system();
    __trap();
}

// Barrier wait of the MMA issuer: one try_wait on the fast path, the bounded spin behind it.
__device__ __forceinline__ void mbar_wait_issuer(uint32_t bar, uint32_t parity, int* err, int site, bool stats, long long& t_acc) {
    if (!stats) {
        if (mbar_try_wait(bar, parity)) return;
        mbar_wait(bar, parity, err, site);
    } else {   // counters on: the whole wait is timed (try_wait itself may block for a while before it reports failure)
        const long long t0 = clock64();
        mbar_wait(bar, parity, err, site);
        t_acc += clock64() - t0;
    }
}

// ---- Chains of convolutions in one launch --------------------------------------------------------------------------
// The kernel runs a CHAIN of up to kMaxChain convolutions: every warp role loops over the layers, its ring positions,
// barrier phases and TMEM stages carried from one layer into the next.  Layer l+1 depends on layer l tile by tile (the
// completion counters above), so a CTA pair that has finished its items of layer l starts on layer l+1 at once, while its
// own epilogue warps still work on layer l's last item and other pairs are still inside layer l: the launch ramp, the
// drain of the last epilogue and the partial last wave of EVERY layer boundary inside a chain disappear (they cost ~25 %
// of a 10bx128 tower launch at batch 256: 109 k cycles per launch against 5.4 items x 15.8 k).  Items are dealt
// round-robin from a start that rotates by the layer's remainder (`rot`), so the pairs that took the extra item of one
// layer are not the ones that take it in the next.  The engine chains launches that share the kernel instantiation, the
// grid, the weight-ring geometry and the resident-weights mode; a launch that is not chained is a chain of one.
// Co-residency: the grid never exceeds one CTA per SM, and forwards of one GPU are serialised (chain_forwards), so every
// CTA a counter is waited on becomes resident without needing any of the waiters to finish.
// Who releases the completion counters.  fp16 rung: a publisher warp per CTA (the epilogue warps are its critical path and
// only arrive on a CTA-local barrier).  Split rung: every epilogue warp releases its own share, deferred until the
// accumulators of its next item are drained — by then the stores it covers are long done and the release is cheap; the
// epilogue has slack there, and the publisher form measured 4 % slower (profiles/r02_conv_chain.md).
#ifndef SB_TC2_PUBLISHER_SPLIT
#define SB_TC2_PUBLISHER_SPLIT 0
#endif
#ifndef SB_TC2_PUBLISHER_FP16
#define SB_TC2_PUBLISHER_FP16 1
#endif
constexpr int kMaxChain = 8;
struct ConvLayer {
    CUtensorMap tmA_hi, tmA_lo;     // input activations (lo unused when !SPLIT)
    CUtensorMap tmW_hi, tmW_lo;     // weights at N = bn
    CUtensorMap tmWq_hi, tmWq_lo;   // weights at N = bn / 2 (half units of a tail wave)
    ConvParams p;
};
struct ConvChain {
    int n_layers;
    ConvLayer layer[kMaxChain];
};

// Barrier addresses and ring geometry the MMA issuer needs (shared::cta addresses of the leader CTA).
struct Conv2Issue {
    uint32_t tmem_base, slab_addr, bst_addr;
    uint32_t a_full, a_empty, tmem_full, tmem_empty, lo_full, lo_empty, b_full, b_empty;
    uint32_t kNB, stage_stride;
    int chunk_kh, first_item, n_clusters, n_items;
    bool resident;
};
// Ring positions and phases of the MMA issuer, carried across the layers of a chain.
struct Conv2IssueState {
    uint32_t j = 0;     // items issued so far (low-order accumulator stage = j & 1)
    uint32_t cc = 0;    // main-accumulator chunks issued so far (stage = cc & 1)
    uint32_t a_it = 0;  // slabs consumed so far
    uint32_t as = 0, aph = 0;   // slab ring position and phase
    uint32_t bs = 0, bph = 0;   // weight ring position and phase
};

// MMA issue for 3x3 convolutions: the nine taps of a k-half are straight-line code, the loop state is warp-uniform.
// The general loop below (per tap: barrier wait, elect, re-derive both descriptors from loop counters, 4-8 MMAs, commit)
// spent ~60 SASS instructions of a single warp — R2UR moves, uniform-datapath chains, reconvergence brackets — on every
// 4 MMAs: measured 93 cycles per N = 128 MMA on the fp16 rung with RESIDENT weights and no barrier waits at all
// (tools/conv_stats.py, total - waits), against 64 cycles of tensor-pipe time: the issuing warp, not the tensor core, set
// the pace, and every wait on its critical path (a successful try_wait costs ~11 cycles, 432 of them per split item) added
// to it.  Here the tap shifts and stage addresses are formed once per k-half, the per-MMA work is one 64-bit add per
// descriptor, and a wait that succeeds at once costs one instruction and one branch.  Same MMAs, same order, same
// accumulators as the general loop: results are bit-identical (test_scheduling_knobs_do_not_change_a_single_bit).
// Resident weights (fp16 rung): the stages of a layer are waited for on the layer's first item only and handed back to
// the producer (for the next layer of a chain) behind the MMAs of its last item.
template <bool SPLIT>
__device__ __forceinline__ void conv2_issue_taps9(const ConvParams& p, const Conv2Issue& q, Conv2IssueState& S, long long* stats_out) {
    using Cfg = Conv2Cfg<SPLIT>;
    constexpr uint32_t kNA = Cfg::kNumSlabs;
    constexpr uint64_t kAStep = 2 * kSlabRows2 * 16 / 16;   // one K=16 step = two channel chunks
    constexpr uint64_t kALo = Cfg::kSlabPartBytes >> 4;     // hi slab part -> lo slab part
    const int KH = p.kh, BN = p.bn, pitch = p.pitch;
    const bool stats = stats_out != nullptr;
    const bool elected = elect_one();   // the whole warp runs the loops (uniform control flow), this lane issues
    const bool wait_w = !(p.dbg & 16), no_slab = (p.dbg & 64) != 0, resident = q.resident;
    const uint32_t idesc_whole = umma_idesc_f16(256, BN), idesc_half = umma_idesc_f16(256, BN >> 1);
    long long t_wait_tmem = 0, t_wait_slab = 0, t_wait_b = 0;
    const long long t_begin = stats ? clock64() : 0;
    uint32_t jl = 0;   // items of THIS layer issued so far
    for (int item = q.first_item; item < q.n_items; item += q.n_clusters, ++S.j, ++jl) {
        const uint32_t ls = S.j & 1u, lph = (S.j >> 1) & 1u;
        const uint32_t idesc = item < p.n_full ? idesc_whole : idesc_half;
        const uint32_t d_lo = q.tmem_base + (2u * ls + 1u) * BN;
        const bool fill = !resident || jl == 0;                                   // the stages are (re)filled for this item
        const bool hand_back = !resident || item + q.n_clusters >= q.n_items;     // ... and free again behind its MMAs
        if (SPLIT) {
            mbar_wait_issuer(q.lo_empty + 8 * ls, lph ^ 1u, p.err, 8, stats, t_wait_tmem);
        }
        uint32_t d_main = 0;
        int h_in_chunk = 0;
        for (int h = 0; h < KH; ++h, ++S.a_it) {
            if (h_in_chunk == 0) {
                const uint32_t cs = S.cc & 1u, cph = (S.cc >> 1) & 1u;
                mbar_wait_issuer(q.tmem_empty + 8 * cs, cph ^ 1u, p.err, 3, stats, t_wait_tmem);
                d_main = q.tmem_base + (SPLIT ? 2u * cs : cs) * BN;
            }
            if (!(no_slab && S.a_it >= kNA)) mbar_wait_issuer(q.a_full + 8 * S.as, S.aph, p.err, 4, stats, t_wait_slab);
            tc_fence_after();
            // descriptor of the un-shifted tile; a tap shift of s rows is +s in the 16-byte start-address field
            const uint64_t a_base = umma_desc_nosw(q.slab_addr + S.as * Cfg::kSlabBytes + (uint32_t)kSlabMargin * 16u, kSlabRows2 * 16u, 128u);
            const uint32_t first_main = h_in_chunk == 0 ? 0u : 1u;
            const uint32_t first_lo = h == 0 ? 0u : 1u;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int shift = (tap / 3 - 1) * pitch + (tap % 3 - 1);
                const uint64_t ad0 = a_base + (uint64_t)(int64_t)shift;
                {   // weights hi x activations hi -> main ; x activations lo -> low-order accumulator
                    if (wait_w && fill) mbar_wait_issuer(q.b_full + 8 * S.bs, S.bph, p.err, 5, stats, t_wait_b);
                    tc_fence_after();
                    const uint64_t bd0 = umma_desc_sw128(q.bst_addr + S.bs * q.stage_stride);
                    if (elected) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma2_f16(d_main, ad0 + kAStep * k, bd0 + 2 * k, idesc, (tap == 0 && k == 0) ? first_main : 1u);
                            if (SPLIT) umma2_f16(d_lo, ad0 + kALo + kAStep * k, bd0 + 2 * k, idesc, (tap == 0 && k == 0) ? first_lo : 1u);
                        }
                        if (hand_back) umma2_commit_mc(q.b_empty + 8 * S.bs, 3);
                    }
                    __syncwarp();
                    if (++S.bs == q.kNB) {
                        S.bs = 0;
                        if (!resident) S.bph ^= 1u;
                    }
                }
                if (SPLIT) {   // weights lo x activations hi -> low-order accumulator
                    if (wait_w && fill) mbar_wait_issuer(q.b_full + 8 * S.bs, S.bph, p.err, 6, stats, t_wait_b);
                    tc_fence_after();
                    const uint64_t bd0 = umma_desc_sw128(q.bst_addr + S.bs * q.stage_stride);
                    if (elected) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma2_f16(d_lo, ad0 + kAStep * k, bd0 + 2 * k, idesc, 1u);
                        if (hand_back) umma2_commit_mc(q.b_empty + 8 * S.bs, 3);
                    }
                    __syncwarp();
                    if (++S.bs == q.kNB) {
                        S.bs = 0;
                        if (!resident) S.bph ^= 1u;
                    }
                }
            }
            const bool last_h = h == KH - 1;
            const bool chunk_done = ++h_in_chunk == q.chunk_kh || last_h;
            if (elected) {
                umma2_commit_mc(q.a_empty + 8 * S.as, 3);
                if (chunk_done) {   // hand the chunk's accumulator stage (and, behind the last one, the low-order sums) to the epilogue
                    if (SPLIT && last_h) umma2_commit_mc(q.lo_full + 8 * ls, 3);
                    umma2_commit_mc(q.tmem_full + 8 * (S.cc & 1u), 3);
                }
            }
            __syncwarp();
            if (chunk_done) {
                h_in_chunk = 0;
                ++S.cc;
            }
            if (++S.as == kNA) {
                S.as = 0;
                S.aph ^= 1u;
            }
        }
    }
    if (resident) S.bph ^= 1u;   // every stage barrier of a resident layer completes exactly once
    if (stats && elected) {
        stats_out[0] = clock64() - t_begin;
        stats_out[1] = t_wait_tmem;
        stats_out[2] = t_wait_slab;
        stats_out[3] = t_wait_b;
        stats_out[6] = jl;
    }
}

// General MMA issue loop: any tap count (1 = 1x1 convolution), the ablation bits of ConvParams::dbg, per-wait counters.
template <bool SPLIT>
__device__ __forceinline__ void conv2_issue_general(const ConvParams& p, const Conv2Issue& q, Conv2IssueState& S, long long* stats_out) {
    using Cfg = Conv2Cfg<SPLIT>;
    constexpr uint32_t kNA = Cfg::kNumSlabs;
    constexpr uint64_t kAStep = 2 * kSlabRows2 * 16 / 16;   // one K=16 step = two channel chunks
    const int KH = p.kh, BN = p.bn;
    const bool stats = stats_out != nullptr, resident = q.resident;
    const uint32_t idesc_whole = umma_idesc_f16(256, BN), idesc_half = umma_idesc_f16(256, BN >> 1);
    long long t_wait_tmem = 0, t_wait_slab = 0, t_wait_b = 0, t0 = 0;
    const long long t_begin = stats ? clock64() : 0;
    uint32_t jl = 0;
    for (int item = q.first_item; item < q.n_items; item += q.n_clusters, ++S.j, ++jl) {
        const uint32_t ls = S.j & 1u, lph = (S.j >> 1) & 1u;
        const uint32_t idesc = item < p.n_full ? idesc_whole : idesc_half;
        const uint32_t d_lo = q.tmem_base + (2u * ls + 1u) * BN;
        const bool fill = !resident || jl == 0;
        const bool hand_back = !resident || item + q.n_clusters >= q.n_items;
        if (SPLIT) {
            if (stats) t0 = clock64();
            mbar_wait(q.lo_empty + 8 * ls, lph ^ 1u, p.err, 8);
            if (stats) t_wait_tmem += clock64() - t0;
        }
        uint32_t d_main = 0;
        int h_in_chunk = 0;      // k-halves issued into the current chunk
        for (int h = 0; h < KH; ++h, ++S.a_it) {
            if (h_in_chunk == 0) {   // a new chunk: its accumulator stage must have been drained
                const uint32_t cs = S.cc & 1u, cph = (S.cc >> 1) & 1u;
                if (stats) t0 = clock64();
                mbar_wait(q.tmem_empty + 8 * cs, cph ^ 1u, p.err, 3);
                if (stats) t_wait_tmem += clock64() - t0;
                d_main = q.tmem_base + (SPLIT ? 2u * cs : cs) * BN;
            }
            if (stats) t0 = clock64();
            if (!((p.dbg & 64) && S.a_it >= kNA)) mbar_wait(q.a_full + 8 * S.as, S.aph, p.err, 4);
            if (stats) t_wait_slab += clock64() - t0;
            tc_fence_after();
            const uint32_t a_hi = q.slab_addr + S.as * Cfg::kSlabBytes;
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int shift = (p.ntaps == 1 || (p.dbg & 8)) ? 0 : (tap / 3 - 1) * p.pitch + (tap % 3 - 1);   // 1 tap = 1x1 convolution
                const uint32_t first_main = (h_in_chunk | tap) == 0 ? 0u : 1u;
                const uint32_t first_lo = (h | tap) == 0 ? 0u : 1u;
                const uint64_t ad0 = umma_desc_nosw(a_hi + (uint32_t)(kSlabMargin + shift) * 16u, kSlabRows2 * 16u, 128u);
                {   // weights hi x activations hi -> main ; x activations lo -> lo accumulator
                    if (stats) t0 = clock64();
                    if (!(p.dbg & 16) && fill) mbar_wait(q.b_full + 8 * S.bs, S.bph, p.err, 5);
                    if (stats) t_wait_b += clock64() - t0;
                    tc_fence_after();
                    const uint64_t bd0 = umma_desc_sw128(q.bst_addr + S.bs * q.stage_stride);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma2_f16(d_main, ad0 + kAStep * k, bd0 + 2 * k, idesc, (k == 0) ? first_main : 1u);
                            if (SPLIT)
                                umma2_f16(d_lo, ad0 + (Cfg::kSlabPartBytes >> 4) + kAStep * k, bd0 + 2 * k, idesc,
                                          (k == 0) ? first_lo : 1u);
                        }
                        if (hand_back) umma2_commit_mc(q.b_empty + 8 * S.bs, 3);
                    }
                    __syncwarp();
                    if (++S.bs == q.kNB) {
                        S.bs = 0;
                        if (!resident) S.bph ^= 1u;
                    }
                }
                if (SPLIT) {   // weights lo x activations hi -> lo accumulator
                    if (stats) t0 = clock64();
                    if (!(p.dbg & 16) && fill) mbar_wait(q.b_full + 8 * S.bs, S.bph, p.err, 6);
                    if (stats) t_wait_b += clock64() - t0;
                    tc_fence_after();
                    const uint64_t bd0 = umma_desc_sw128(q.bst_addr + S.bs * q.stage_stride);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma2_f16(d_lo, ad0 + kAStep * k, bd0 + 2 * k, idesc, 1u);
                        if (hand_back) umma2_commit_mc(q.b_empty + 8 * S.bs, 3);
                    }
                    __syncwarp();
                    if (++S.bs == q.kNB) {
                        S.bs = 0;
                        if (!resident) S.bph ^= 1u;
                    }
                }
            }
            const bool last_h = h == KH - 1;
            const bool chunk_done = ++h_in_chunk == q.chunk_kh || last_h;
            if (elect_one()) {
                umma2_commit_mc(q.a_empty + 8 * S.as, 3);
                if (chunk_done) {   // hand the chunk's accumulator stage (and, behind the last one, the low-order sums) to the epilogue
                    if (SPLIT && last_h) umma2_commit_mc(q.lo_full + 8 * ls, 3);
                    umma2_commit_mc(q.tmem_full + 8 * (S.cc & 1u), 3);
                }
            }
            __syncwarp();
            if (chunk_done) {
                h_in_chunk = 0;
                ++S.cc;
            }
            if (++S.as == kNA) {
                S.as = 0;
                S.aph ^= 1u;
            }
        }
    }
    if (resident) S.bph ^= 1u;
    if (stats && (threadIdx.x & 31) == 0) {
        stats_out[0] = clock64() - t_begin;
        stats_out[1] = t_wait_tmem;
        stats_out[2] = t_wait_slab;
        stats_out[3] = t_wait_b;
        stats_out[6] = jl;
    }
}

// first item of a cluster in a layer: round-robin from a start that rotates with the layer (ConvParams::rot)
__device__ __forceinline__ int conv_first_item(int cluster_id, int n_clusters, const ConvParams& p) {
    const int v = cluster_id + p.rot;
    return v >= n_clusters ? v - n_clusters : v;
}

template <bool SPLIT, int ACT, bool POOL, int PARTS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Conv2Cfg<SPLIT, PARTS>::kThreads, 1)
conv3x3_tc2_kernel(const __grid_constant__ ConvChain chain) {
    using Cfg = Conv2Cfg<SPLIT, PARTS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t slab_addr = smem_base;
    const uint32_t bst_addr = smem_base + Cfg::kOffB;
    const uint32_t bar_addr = smem_base + Cfg::kOffBar;
    const uint32_t a_full = bar_addr + 0, a_empty = bar_addr + 32;                // [<= 4] each
    const uint32_t tmem_full = bar_addr + 64, tmem_empty = bar_addr + 80;         // [2] each: main accumulator stages (chunks)
    const uint32_t lo_full = bar_addr + 96, lo_empty = bar_addr + 112;            // [2] each: low-order accumulators (items)
    const uint32_t b_full = bar_addr + 128, b_empty = bar_addr + 288;             // [<= 18] each
    const uint32_t stored = bar_addr + 456;                                       // [4]: the epilogue warps of THIS CTA have stored an item
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + Cfg::kOffBar + 448);
    const uint32_t deps_ctr = bar_addr + 488;   // items of this CTA whose input tiles the slab producer has seen complete
    constexpr uint32_t kNA = Cfg::kNumSlabs;
    constexpr bool kPublisher = SPLIT ? (SB_TC2_PUBLISHER_SPLIT != 0) : (SB_TC2_PUBLISHER_FP16 != 0);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: role branches and their loop state stay in uniform registers
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_layers = chain.n_layers;

    float* sbias = reinterpret_cast<float*>(smem_gen + Cfg::kOffBias);
    for (int l = 0; l < n_layers; ++l) {   // weights are constants: staged before any dependency wait
        const ConvParams& pl = chain.layer[l].p;
        const int nb = pl.n_ntiles * pl.bn;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) sbias[pl.bias_off + i] = pl.bias[i];
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)kNA; ++i) {
            mbar_init(a_full + 8 * i, 1);    // leader's own arrive.expect_tx (bytes of both CTAs)
            mbar_init(a_empty + 8 * i, 1);   // one multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full + 8 * i, 1);
            mbar_init(tmem_empty + 8 * i, 2 * 4 * Cfg::kEpiParts);  // every epilogue warp of both CTAs
            mbar_init(lo_full + 8 * i, 1);
            mbar_init(lo_empty + 8 * i, 2 * 4 * Cfg::kEpiParts);
        }
        for (int i = 0; i < 4; ++i) mbar_init(stored + 8 * i, 4 * Cfg::kEpiParts);
        for (int i = 0; i < Cfg::kNumBStages; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        st_release_cta_shared(deps_ctr, 0u);
        fence_barrier_init();
        const ConvLayer& L0 = chain.layer[0];
        tma_prefetch_desc(&L0.tmA_hi);
        tma_prefetch_desc(&L0.tmW_hi);
        if (SPLIT) {
            tma_prefetch_desc(&L0.tmA_lo);
            tma_prefetch_desc(&L0.tmW_lo);
        }
        if (L0.p.n_units > L0.p.n_full) {
            tma_prefetch_desc(&L0.tmWq_hi);
            if (SPLIT) tma_prefetch_desc(&L0.tmWq_lo);
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
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    if (threadIdx.x == 0) pdl_launch_dependents();   // the next launch may start its prologue whenever SMs free up

    if (warp == 3) {
        // ===================== activation-slab producer (own 128-row tile, +-24 rows) =====================
        uint32_t it = 0, s = 0, ph = 0;   // slabs loaded so far, ring position and phase
        uint32_t n_dep = 0;               // items started so far (all layers)
        for (int l = 0; l < n_layers; ++l) {
            const ConvLayer& L = chain.layer[l];
            const ConvParams& p = L.p;
            const int KH = p.kh;
            const bool tile_deps = p.done_in != nullptr;
            if (!tile_deps) pdl_wait();   // the slabs are the previous launch's output (only the first layer of a chain)
            for (int item = conv_first_item(cluster_id, n_clusters, p); item < p.n_units; item += n_clusters) {
                const int st = conv_unit(item, p).st;
                const int row_lo = kGuardRows + st * kSuperRows + (int)rank * kTileRows2 - kSlabMargin;
                for (int h = 0; h < KH; ++h, ++it) {
                    if ((p.dbg & 64) && it >= kNA) continue;   // ablation: no slab traffic after the first fills
                    if (tile_deps && h == 0) {   // the producing layer has finished the super tiles this slab reads
                        if (lane == 0) {
                            wait_tiles_done3(p.done_in, st, p.n_super, p.done_in_full, p.err, 10);
                            st_release_cta_shared(deps_ctr, n_dep + 1u);   // ... which the epilogue warps of this CTA learn here
                        }
                        __syncwarp();
                    }
                    if (h == 0) ++n_dep;
                    mbar_wait(a_empty + 8 * s, ph ^ 1u, p.err, 1);
                    if (elect_one()) {
                        if (tile_deps) fence_proxy_async_global();   // rows written through the generic proxy, read by TMA
                        const uint32_t full0 = mapa_u32(a_full + 8 * s, 0);
                        if (leader) mbar_arrive_expect_tx(a_full + 8 * s, 2u * Cfg::kSlabBytes);
#pragma unroll
                        for (int part = 0; part < Cfg::kParts; ++part) {
                            tma2_load_3d(slab_addr + s * Cfg::kSlabBytes + part * Cfg::kSlabPartBytes, part ? &L.tmA_lo : &L.tmA_hi, 0,
                                         row_lo, h * 8, full0);
                        }
                    }
                    __syncwarp();
                    if (++s == kNA) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ===================== weight-stage producer (own N-half of every block) =====================
        uint32_t s = 0, ph = 0;   // ring position and phase, advanced incrementally (kNB is a run-time value)
        for (int l = 0; l < n_layers; ++l) {
            const ConvLayer& L = chain.layer[l];
            const ConvParams& p = L.p;
            const int KH = p.kh, BN = p.bn;
            const bool resident = p.resident != 0;
            const uint32_t stage_stride = BN > 128 ? 2u * Cfg::kBStageBytes : (uint32_t)Cfg::kBStageBytes;
            const uint32_t kNB = resident ? (uint32_t)(p.kh * p.ntaps * Cfg::kParts)
                                          : (uint32_t)(Cfg::kNumBStages * Cfg::kBStageBytes) / stage_stride;
            const int first = conv_first_item(cluster_id, n_clusters, p);
            for (int item = first; item < p.n_units && !(p.dbg & 16); item += n_clusters) {
                if (resident && item != first) break;     // everything is already in shared memory
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
                                             whole ? (part ? &L.tmW_lo : &L.tmW_hi) : (part ? &L.tmWq_lo : &L.tmWq_hi),
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
        }
    } else if (warp == 2) {
        // ===================== publisher: per-tile completion counters for the next layer / launch =====================
        // The epilogue warps only arrive on a CTA-local barrier when their part of an item is stored; the gpu-scope
        // release (MEMBAR.ALL.GPU: wait for the stores to be visible at L2) is paid here, off the epilogue's path
        // (as a release per epilogue warp it cost the fp16 rung 8 %).  Cumulativity makes the one release cover the
        // stores of all arriving warps, as in a grid-wide barrier built on bar.sync + one fencing thread.
        // Four barriers in rotation: an epilogue warp cannot run four items ahead of another one (they share two TMEM
        // stages), so an arrival never lands in the phase of an item that is still being stored.
        uint32_t jp = 0;   // published items so far
        for (int l = 0; l < n_layers && kPublisher; ++l) {
            const ConvParams& p = chain.layer[l].p;
            if (p.done_out == nullptr) continue;
            for (int item = conv_first_item(cluster_id, n_clusters, p); item < p.n_units; item += n_clusters, ++jp) {
                const ConvUnit w = conv_unit(item, p);
                mbar_wait(stored + 8 * (jp & 3u), (jp >> 2) & 1u, p.err, 14);
                if (lane == 0) {
                    fence_proxy_async_global();   // the consumer reads these rows with TMA (async proxy)
                    red_release_gpu(p.done_out + w.st, 4 * w.bn);   // 4 lane quadrants x the unit's columns, from each CTA of the pair
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===================== MMA issuer (leader CTA only) =====================
            // TMEM columns per CTA, split rung: [main 0 | lo 0 | main 1 | lo 1] x BN (main stage = chunk counter & 1, low-order
            // accumulator = item counter & 1); fp16 rung: main stage s at s * BN.
            Conv2IssueState S;
            for (int l = 0; l < n_layers; ++l) {
                const ConvParams& p = chain.layer[l].p;
                Conv2Issue q;
                q.tmem_base = tmem_base; q.slab_addr = slab_addr; q.bst_addr = bst_addr;
                q.a_full = a_full; q.a_empty = a_empty; q.tmem_full = tmem_full; q.tmem_empty = tmem_empty;
                q.lo_full = lo_full; q.lo_empty = lo_empty; q.b_full = b_full; q.b_empty = b_empty;
                q.resident = p.resident != 0;
                // a weight stage holds this CTA's N-half of one [BN][64] block: 8 KB slots for BN <= 128, 16 KB (two slots) for
                // the N = 256 tiles of the fp16 rung, which therefore has half as many stages in the same ring; resident
                // weights: the ring is exactly one item's worth of stages, filled once per layer
                q.stage_stride = p.bn > 128 ? 2u * Cfg::kBStageBytes : (uint32_t)Cfg::kBStageBytes;
                q.kNB = q.resident ? (uint32_t)(p.kh * p.ntaps * Cfg::kParts) : (uint32_t)(Cfg::kNumBStages * Cfg::kBStageBytes) / q.stage_stride;
                // main-accumulator chunks: the k-halves of an item are cut every chunk_kh k-halves (fp16 rung: one chunk per item)
                q.chunk_kh = SPLIT ? p.chunk_kh : p.kh;
                q.first_item = conv_first_item(cluster_id, n_clusters, p);
                q.n_clusters = n_clusters;
                q.n_items = p.n_units;   // whole items + half units of the tail wave (conv_unit)
                long long* stats_out = p.stats ? p.stats + (size_t)cluster_id * 8 : nullptr;
                if (p.ntaps == 9 && !(p.dbg & (8 | 128))) conv2_issue_taps9<SPLIT>(p, q, S, stats_out);
                else conv2_issue_general<SPLIT>(p, q, S, stats_out);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 4 * kEpiParts warps; warp%4 = TMEM lane quadrant, (warp-4)/4 = column part =====================
        // kEpiParts warps per scheduler: each thread owns 1/kEpiParts of an accumulator row, so the dependent ALU/MUFU
        // chains, tcgen05.ld waits and residual loads of one warp are hidden behind the others.
        const int q = warp & 3;
        const int part = (warp - 4) >> 2;                      // which column part of the accumulator row
        uint32_t j = 0, cc = 0;                                // items / main chunks drained so far (all layers)
        uint32_t jp = 0;                                       // items handed to the publisher warp so far
        int* pend_ptr = nullptr;                               // !SB_TC2_PUBLISHER: completion counter of the previous item, not yet released
        int pend_val = 0;
        const uint32_t empty0 = mapa_u32(tmem_empty, 0);   // leader's tmem_empty[0]; [1] is +8
        const uint32_t lo_empty0 = mapa_u32(lo_empty, 0);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int l = 0; l < n_layers; ++l) {
        const ConvParams& p = chain.layer[l].p;
        const int KH = p.kh, BN = p.bn;
        const int chunk_kh = SPLIT ? p.chunk_kh : KH;
        const int n_chunks = (KH + chunk_kh - 1) / chunk_kh;
        const bool stats = p.stats != nullptr;
        long long t_wait_full = 0, t_drain = 0;
        const long long t_begin = stats ? clock64() : 0;
        const float chunk_scale = p.chunk_scale;
        const float* sb = sbias + p.bias_off;
        const bool tile_deps = p.done_in != nullptr;
        if (!tile_deps) pdl_wait();   // residual reads, and our stores may overwrite a buffer the previous launch still reads
        for (int item = conv_first_item(cluster_id, n_clusters, p); item < p.n_units; item += n_clusters, ++j) {
            const ConvUnit w = conv_unit(item, p);
            const int st = w.st;
            if (tile_deps && p.res_hi != nullptr) {
                // The residual rows of this tile are final once this CTA's slab producer has seen the input tiles of the item
                // complete (see "Cross-layer dependencies"): it says so through a counter in shared memory, normally long
                // before we get here (it works one item ahead).  Its gpu-scope acquire also invalidated this SM's L1
                // (LDG.STRONG.GPU + CCTL.IVALL) after the rows were published, so the plain loads below cannot hit a line
                // cached before they were written (a polling acquire per epilogue warp and item did the same at the price of
                // an L2 round trip in front of every residual prefetch and 16 L1 invalidations per item).
                wait_deps_ctr(deps_ctr, j + 1u, p.err, 13);
            }
            // columns of this thread: w.bn split into kEpiParts runs rounded to the 16-column ld granule (may be 0);
            // a run longer than kPassCols columns (N = 256 tiles of the fp16 rung) is processed in passes
            const int c_run = ((w.bn + Cfg::kEpiParts - 1) / Cfg::kEpiParts + 15) & ~15;
            const int cbase = min(part * c_run, w.bn);
            const int HC = min(c_run, w.bn - cbase);
            const int n_pass = max(1, (HC + Cfg::kPassCols - 1) / Cfg::kPassCols);     // 1 on the split rung (BN <= 128)
            const uint32_t ls = j & 1u, lph = (j >> 1) & 1u;

            const int row = kGuardRows + st * kSuperRows + (int)rank * kTileRows2 + q * 32 + lane;
            const bool live = p.mask[row] != 0;
            const bool has_res = live && p.res_hi != nullptr && !(p.dbg & 1);
            const size_t chunk_stride = (size_t)p.rows * 8;

            for (int pass = 0; pass < n_pass; ++pass) {
                const int pbase = cbase + pass * Cfg::kPassCols;            // first column of this pass inside the N tile
                const int PC = min(Cfg::kPassCols, HC - pass * Cfg::kPassCols);   // columns of this pass (multiple of 16, may be 0)
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
                float acc[Cfg::kMaxGroups * 16];
                if (SPLIT) {
                    // ---- split rung: drain every main chunk as it completes and add it in fp32 round-to-nearest ----
                    for (int ch = 0; ch < n_chunks; ++ch, ++cc) {
                        const uint32_t cs = cc & 1u, cph = (cc >> 1) & 1u;
                        const long long t0 = stats ? clock64() : 0;
                        mbar_wait(tmem_full + 8 * cs, cph, p.err, 7);
                        if (stats) t_wait_full += clock64() - t0;
                        tc_fence_after();
                        const long long t_d0 = stats ? clock64() : 0;
                        const uint32_t t_main = lane_base + 2u * cs * BN + pbase;
#pragma unroll
                        for (int g = 0; g < Cfg::kMaxGroups; ++g) {
                            if (g * 16 < PC) {
                                uint32_t r[16];
                                tmem_ld16(t_main + g * 16, r);
                                tmem_ld_wait();
                                if (ch == 0) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) acc[g * 16 + i] = __uint_as_float(r[i]) * chunk_scale;
                                } else {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) acc[g * 16 + i] = fmaf(__uint_as_float(r[i]), chunk_scale, acc[g * 16 + i]);
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(empty0 + 8 * cs);   // main stage is free for the chunk after next
                        if (stats) t_drain += clock64() - t_d0;
                    }
                    {   // the low-order accumulator of this item (committed together with the last chunk)
                        mbar_wait(lo_full + 8 * ls, lph, p.err, 9);
                        tc_fence_after();
                        const uint32_t t_lo = lane_base + (2u * ls + 1u) * BN + pbase;
#pragma unroll
                        for (int g = 0; g < Cfg::kMaxGroups; ++g) {
                            if (g * 16 < PC) {
                                uint32_t r2[16];
                                tmem_ld16(t_lo + g * 16, r2);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; ++i) acc[g * 16 + i] += __uint_as_float(r2[i]);
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(lo_empty0 + 8 * ls);
                    }
                } else {
                    // ---- fp16 rung: one chunk per item ----
                    const uint32_t cs = cc & 1u, cph = (cc >> 1) & 1u;
                    if (pass == 0) {
                        const long long t0 = stats ? clock64() : 0;
                        mbar_wait(tmem_full + 8 * cs, cph, p.err, 7);
                        if (stats) t_wait_full += clock64() - t0;
                        tc_fence_after();
                    }
                    const uint32_t t_main = lane_base + cs * BN + pbase;
                    const long long t_d0 = stats ? clock64() : 0;
#pragma unroll
                    for (int g = 0; g < Cfg::kMaxGroups; ++g) {
                        if (g * 16 < PC) {
                            uint32_t r[16];
                            tmem_ld16(t_main + g * 16, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[g * 16 + i] = __uint_as_float(r[i]);
                        }
                    }
                    if (pass == n_pass - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(empty0 + 8 * cs);   // accumulator stage is free again
                        ++cc;
                    }
                    if (stats) t_drain += clock64() - t_d0;
                }

                if (!kPublisher && pend_ptr != nullptr) {   // the previous item's stores were issued an MMA phase ago: the release finds them done
                    __syncwarp();
                    if (lane == 0) {
                        fence_proxy_async_global();
                        red_release_gpu(pend_ptr, pend_val);
                    }
                    pend_ptr = nullptr;
                }
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
                        for (int i = 0; i < 16; ++i) v[i] = acc[c0 + i] + sb[w.n0 + pbase + c0 + i];
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
                        if (POOL) {
                            // SE pooling partials of the (pre-rounding) outputs: sum and max over the board cells of every
                            // aligned group of 2^pool_log2 rows, per channel; off-board rows count 0 / -5000 (se_unit.cc:22)
                            float ps[16], pm[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                ps[i] = v[i];                       // already 0 on dead rows
                                pm[i] = live ? v[i] : -5000.f;
                            }
                            const int PL = p.pool_log2;
                            if (PL == 4) pool_reduce16<4>(ps, pm, lane);
                            else if (PL == 3) pool_reduce16<3>(ps, pm, lane);
                            else pool_reduce16<2>(ps, pm, lane);
                            const int gi = (row - kGuardRows) >> PL;
                            const int cnt = 16 >> PL, first = pool_first(lane, PL);
                            if (gi < p.pool_groups && ch0 + first < p.cout) {
                                float* dst = p.pool_part + (size_t)gi * 2 * p.pool_c + ch0 + first;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    if (i < cnt) {
                                        dst[i] = ps[i];
                                        dst[p.pool_c + i] = pm[i];
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (p.done_out != nullptr) {   // this warp's part of the tile is stored: tell the publisher warp
                if (kPublisher) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(stored + 8 * (jp & 3u));
                    ++jp;
                } else if (HC > 0) {
                    if (item + n_clusters >= p.n_units) {   // last item of the layer: nothing to hide the release behind
                        __syncwarp();
                        if (lane == 0) {
                            fence_proxy_async_global();
                            red_release_gpu(p.done_out + st, HC);
                        }
                    } else {
                        pend_ptr = p.done_out + st;
                        pend_val = HC;
                    }
                }
            }
        }
        if (stats && leader && warp == 4 && lane == 0) {
            long long* so = p.stats + (size_t)cluster_id * 8;
            so[4] = t_wait_full;
            so[5] = clock64() - t_begin;
            so[7] = t_drain;
        }
        }   // layers
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // nobody leaves (or frees TMEM) while the peer may still read this CTA's memory
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

}  // namespace sb
