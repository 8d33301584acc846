// Engine + C ABI (include/sayuri_b200.h).  One Replica per GPU (weights blob, tensor maps), a few
// independent Slots per replica (stream, pinned staging, activations) so that H2D / compute / D2H of
// consecutive batches overlap.  Mirrors the lifetime and semantics of the reference GPU pipe
//   CudaForwardPipe / NNGraph      /root/reference/src/neural/cuda/cuda_forward_pipe.cc:14-131,133-613,684-1136
// but none of its structure: NHWC canvas activations, one fused tensor-core conv kernel per layer,
// device-side canvas placement / crop / policy-channel gather.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <dlfcn.h>
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sayuri_b200.h"
#include "aux_kernels.cuh"
#include "common.cuh"
#include "conv3x3_tc2.cuh"
#include "host_net.h"

namespace sb {

// ------------------------------------------------------------------------------------------------
struct CudaError {
    std::string msg;
};
#define SB_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            throw CudaError{std::string("CUDA Error: ") + cudaGetErrorString(e__) + " at " + #call + " (" + \
                            __FILE__ + ":" + std::to_string(__LINE__) + ")"};                             \
        }                                                                                                 \
    } while (0)

static inline int RoundUp(int v, int m) { return (v + m - 1) / m * m; }

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency,
// so the library also loads on a box without a driver (the "symbols exported" CPU test).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn GetEncodeTiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    if (!fn) throw CudaError{"cuTensorMapEncodeTiled is unavailable (driver too old?)"};
    return fn;
}

// 2-D fp16 row-major tensor [rows][cols], box = [box_rows][64 cols], 128-byte swizzle, zero OOB fill.
static CUtensorMap MakeMap2D(const void* base, int rows, int cols, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = GetEncodeTiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)};
    return m;
}

// C8 activation tensor [chunks][R rows][8 ch] as a 3-D {8, R, chunks} view with a 176-row box: one box {8, 176, 8} lands in
// shared memory as [chunk][176 rows][16 B], the per-CTA slab of the conv kernel.
static CUtensorMap MakeActMap3(const void* base, int R, int chunks) {
    CUtensorMap m;
    cuuint64_t dims[3] = {8u, (cuuint64_t)R, (cuuint64_t)chunks};
    cuuint64_t strides[2] = {16u, (cuuint64_t)R * 16u};
    cuuint32_t box[3] = {8u, (cuuint32_t)kSlabRows2, 8u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = GetEncodeTiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled (activation, pair kernel) failed with code " + std::to_string((int)r)};
    return m;
}

// ------------------------------------------------------------------------------------------------
// Weight blob: every device-resident parameter of one replica, packed once on the host.
struct ConvLayout {
    int cin = 0, cinp = 0, cout = 0, coutp = 0, kh = 0, bn = 0, ntiles = 0, taps = 9;
    size_t w_hi = 0, w_lo = 0, bias = 0;  // byte offsets into the blob
};
struct DwLayout {   // depthwise k x k filters, fp32 [ch][k*k], + bias [ch]
    int ch = 0, k = 0;
    size_t w = 0, b = 0;
};
struct FcLayout {
    int in = 0, out = 0;
    size_t w = 0, b = 0;
};
struct BlobLayout {
    ConvLayout input;
    std::vector<std::vector<ConvLayout>> bconv;   // per block, loader order (HostBlock::convs; a Mixer's depthwise conv is in bdw)
    std::vector<DwLayout> bdw;                    // per block: Mixer depthwise filters (k == 0 otherwise)
    DwLayout p_dw;                                // RepLK policy head (k == 0 for the Normal head)
    ConvLayout p_pt;
    std::vector<FcLayout> squeeze, excite;  // per block (unused entries when se_size == 0)
    ConvLayout head;                        // policy + value head-entry 1x1 convs as one single-tap conv, cout = P + V
    FcLayout p_inter, pass, v_inter, misc;
    size_t prob_w = 0, prob_b = 0, own_w = 0, own_b = 0;
    size_t bytes = 0;
};

static size_t Take(size_t& cursor, size_t bytes) {
    const size_t at = cursor;
    cursor = (cursor + bytes + 255) & ~(size_t)255;
    return at;
}

static ConvLayout LayConv(size_t& cur, int cin, int cout, int taps = 9) {
    ConvLayout c;
    c.taps = taps;
    c.cin = cin;
    c.cinp = RoundUp(cin, 64);
    c.cout = cout;
    c.coutp = RoundUp(cout, 16);   // weight rows are zero-padded to the UMMA N granule; the epilogue stores only cout
    c.kh = c.cinp / 64;
    c.bn = 16;
    for (int bn = 128; bn >= 16; bn -= 16) {   // widest N tile (multiple of 16, <= 128) that divides the width
        if (c.coutp % bn == 0) {
            c.bn = bn;
            break;
        }
    }
    c.ntiles = c.coutp / c.bn;
    const size_t mat = (size_t)c.coutp * taps * c.cinp * sizeof(__half);
    c.w_hi = Take(cur, mat);
    c.w_lo = Take(cur, mat);
    c.bias = Take(cur, (size_t)c.coutp * sizeof(float));
    return c;
}
static DwLayout LayDw(size_t& cur, int ch, int k) {
    DwLayout d;
    d.ch = ch;
    d.k = k;
    d.w = Take(cur, (size_t)ch * k * k * sizeof(float));
    d.b = Take(cur, (size_t)ch * sizeof(float));
    return d;
}
static FcLayout LayFc(size_t& cur, int in, int out) {
    FcLayout f;
    f.in = in;
    f.out = out;
    f.w = Take(cur, (size_t)in * out * sizeof(float));
    f.b = Take(cur, (size_t)out * sizeof(float));
    return f;
}

static BlobLayout ComputeLayout(int blocks, int C, int P, int V, const std::vector<int>& se, const std::vector<int>& types,
                                const std::vector<int>& inner, const std::vector<int>& dwk, int replk_kernel) {
    BlobLayout L;
    size_t cur = 0;
    L.input = LayConv(cur, SB_INPUT_CHANNELS, C);
    L.bconv.resize(blocks);
    L.bdw.resize(blocks);
    L.squeeze.resize(blocks);
    L.excite.resize(blocks);
    for (int b = 0; b < blocks; ++b) {
        if (types[b] == SB_BLOCK_RESIDUAL) {
            L.bconv[b] = {LayConv(cur, C, C), LayConv(cur, C, C)};
        } else if (types[b] == SB_BLOCK_MIXER) {   // depthwise k x k, 1x1 C -> F, 1x1 F -> C
            L.bdw[b] = LayDw(cur, C, dwk[b]);
            L.bconv[b] = {LayConv(cur, C, inner[b], 1), LayConv(cur, inner[b], C, 1)};
        } else {   // pre 1x1, 2 or 4 inner 3x3, post 1x1
            const int I = inner[b], n3 = types[b] == SB_BLOCK_BOTTLENECK ? 2 : 4;
            L.bconv[b].push_back(LayConv(cur, C, I, 1));
            for (int q = 0; q < n3; ++q) L.bconv[b].push_back(LayConv(cur, I, I));
            L.bconv[b].push_back(LayConv(cur, I, C, 1));
        }
        if (se[b] > 0) {
            L.squeeze[b] = LayFc(cur, 3 * C, se[b]);
            L.excite[b] = LayFc(cur, se[b], 2 * C);
        }
    }
    L.head = LayConv(cur, C, P + V, 1);
    if (replk_kernel > 0) {
        L.p_dw = LayDw(cur, P, replk_kernel);
        L.p_pt = LayConv(cur, P, P, 1);
    }
    L.p_inter = LayFc(cur, 3 * P, P);
    L.pass = LayFc(cur, P, 5);
    L.v_inter = LayFc(cur, 3 * V, 3 * V);
    L.misc = LayFc(cur, 3 * V, 15);
    L.prob_w = Take(cur, (size_t)5 * P * sizeof(float));
    L.prob_b = Take(cur, 5 * sizeof(float));
    L.own_w = Take(cur, (size_t)V * sizeof(float));
    L.own_b = Take(cur, sizeof(float));
    L.bytes = cur;
    return L;
}

// OIHW fp32 -> K-major [cout][tap * cinp + c] fp16 hi / lo (the B operand of the implicit GEMM).
static void PackConv(const HostConv& hc, const ConvLayout& L, uint8_t* blob) {
    __half* hi = reinterpret_cast<__half*>(blob + L.w_hi);
    __half* lo = reinterpret_cast<__half*>(blob + L.w_lo);
    const size_t K = (size_t)L.taps * L.cinp;
    for (int o = 0; o < L.cout; ++o) {
        for (int tap = 0; tap < L.taps; ++tap) {
            for (int c = 0; c < L.cinp; ++c) {
                float w = 0.f;
                if (c < L.cin) w = hc.w[((size_t)o * L.cin + c) * L.taps + tap];
                const __half h = __float2half_rn(w);
                const __half l = __float2half_rn(w - __half2float(h));
                hi[(size_t)o * K + (size_t)tap * L.cinp + c] = h;
                lo[(size_t)o * K + (size_t)tap * L.cinp + c] = l;
            }
        }
    }
    std::memcpy(blob + L.bias, hc.b.data(), (size_t)L.cout * sizeof(float));
}
static void PackDw(const HostConv& hc, const DwLayout& L, uint8_t* blob) {
    std::memcpy(blob + L.w, hc.w.data(), hc.w.size() * sizeof(float));
    std::memcpy(blob + L.b, hc.b.data(), hc.b.size() * sizeof(float));
}
static void PackFc(const HostFC& f, const FcLayout& L, uint8_t* blob) {
    std::memcpy(blob + L.w, f.w.data(), f.w.size() * sizeof(float));
    std::memcpy(blob + L.b, f.b.data(), f.b.size() * sizeof(float));
}

static std::vector<uint8_t> PackBlob(const HostNet& n, const BlobLayout& L) {
    std::vector<uint8_t> blob(L.bytes, 0);
    uint8_t* p = blob.data();
    PackConv(n.input_conv, L.input, p);
    for (int b = 0; b < n.blocks; ++b) {
        const bool mixer = n.tower[b].type == SB_BLOCK_MIXER;
        if (mixer) PackDw(n.tower[b].convs[0], L.bdw[b], p);
        for (size_t q = mixer ? 1 : 0; q < n.tower[b].convs.size(); ++q) PackConv(n.tower[b].convs[q], L.bconv[b][q - (mixer ? 1 : 0)], p);
        if (n.tower[b].se_size > 0) {
            PackFc(n.tower[b].squeeze, L.squeeze[b], p);
            PackFc(n.tower[b].excite, L.excite[b], p);
        }
    }
    const int C = n.channels, P = n.P, V = n.V;
    {   // policy and value head-entry filters concatenated along the output dimension
        HostConv head;
        head.in = C;
        head.out = P + V;
        head.k = 1;
        head.w = n.p_hd_conv.w;
        head.w.insert(head.w.end(), n.v_hd_conv.w.begin(), n.v_hd_conv.w.end());
        head.b = n.p_hd_conv.b;
        head.b.insert(head.b.end(), n.v_hd_conv.b.begin(), n.v_hd_conv.b.end());
        PackConv(head, L.head, p);
    }
    if (n.replk) {
        PackDw(n.p_dw_conv, L.p_dw, p);
        PackConv(n.p_pt_conv, L.p_pt, p);
    }
    PackFc(n.p_inter_fc, L.p_inter, p);
    PackFc(n.pass_fc, L.pass, p);
    PackFc(n.v_inter_fc, L.v_inter, p);
    PackFc(n.v_misc, L.misc, p);
    std::memcpy(p + L.prob_w, n.prob_conv.w.data(), (size_t)5 * P * sizeof(float));
    std::memcpy(p + L.prob_b, n.prob_conv.b.data(), 5 * sizeof(float));
    std::memcpy(p + L.own_w, n.v_ownership.w.data(), (size_t)V * sizeof(float));
    std::memcpy(p + L.own_b, n.v_ownership.b.data(), sizeof(float));
    return blob;
}

// ------------------------------------------------------------------------------------------------
struct ActBuf {
    __half* hi = nullptr;
    __half* lo = nullptr;
    int channels = 0;   // padded to a multiple of 64
    int rows = 0;       // R
    CUtensorMap tm3_hi, tm3_lo;     // 176-row slabs
    // layer overlap (conv3x3_tc2.cuh, "Cross-layer dependencies"): set while the tensor holds the output of a convolution
    // launch of the forward being enqueued that publishes per-tile completion counters
    const int* done = nullptr;      // the counters, nullptr = written by something else (whole-grid dependency)
    int done_full = 0;              // value of a complete super tile
    const ActBuf* done_src = nullptr;   // the input tensor of that launch
};

// A chain of convolution launches waiting to be issued as ONE kernel (conv3x3_tc2.cuh, "Chains of convolutions in one launch").
struct PendingChain {
    ConvChain chain{};
    bool split = false, wide = false, pool = false;
    int act = 0, grid2 = 0, resident = 0, ring_key = 0, last_rem = 0, bn = 0, bias_floats = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_done = nullptr;   // forward of this slot has finished computing (before its D2H)
    float* h_in = nullptr;     // pinned [max_batch][SB_PLANE_FLOATS]
    float* d_in = nullptr;
    int* h_meta = nullptr;     // pinned [2][max_batch]: board sizes, policy offsets
    int* d_meta = nullptr;
    sb_packed_position* h_packed = nullptr;   // pinned [max_batch]: compact records (host_pack.cc)
    sb_packed_position* d_packed = nullptr;
    bool packed = false;       // the batch in flight came as compact records (unpack_packed_kernel)
    float* d_out = nullptr;    // [max_batch][kOutFloats]
    float* h_out = nullptr;    // pinned
    ActBuf in, x, t, u;
    ActBuf ia, ib, ic;         // bottleneck / feed-forward width buffers (allocated only when the tower needs them)
    ActBuf pq;                 // RepLK policy head: depthwise output (P channels, padded to 64)
    ActBuf* trunk = nullptr;   // which buffer holds the tower output after the last forward
    uint8_t* mask = nullptr;
    float* gb = nullptr;       // [max_batch][2C]
    float* pool_part = nullptr;   // [row groups][2][C]: SE pooling partials written by the conv epilogue (PoolLog2)
    float* pooled = nullptr;   // [max_batch][2 * max(C, 64)]: per-channel sums and maxima between the pooling CTAs and the FC tail
    int* counters = nullptr;   // [max_batch]: pooling CTAs of a sample that have published (self-resetting)
    ActBuf pv;                 // head-entry conv output: P policy + V value channels (padded to 64)
    float* pint = nullptr;
    float* pass5 = nullptr;
    float* misc15 = nullptr;
    long long* d_stats = nullptr;  // [sm_count][8] cycle counters of the last conv launch (option "stats")
    int* h_err = nullptr;      // mapped pinned: barrier-timeout site code survives a trapped context
    int* d_err = nullptr;
    int n = 0;                 // samples of the batch in flight / last uploaded
    int conv_counter = 0;      // conv launches of the forward being enqueued on this slot (option "stats_launch")
    int* d_done = nullptr;     // [kMaxDoneLaunches][done_stride] per-tile completion counters of the conv launches of a forward
    int done_stride = 0;
    PendingChain* pending = nullptr;   // convolution launches collected into the next chained launch
    bool fwd_chain = false, fwd_overlap = false;   // ChainMode / LayerOverlap, decided ONCE for the forward being enqueued
    bool busy = false;
    std::vector<int> sizes, offsets;
};

struct DevConv {
    ConvLayout L;
    // box [(bn >> level) / 2][64] = one CTA's half of an N tile of width bn >> level.  Level 0 is the
    // throughput shape; levels 1-2 spread small batches over more CTA pairs and serve the tail-wave half units.
    CUtensorMap tm2_hi[4], tm2_lo[4];
    int levels = 1;               // usable N widths: bn0 >> 0 .. bn0 >> (levels - 1), each a multiple of 16
    int bn0 = 0;                  // widest N tile of the CTA-pair kernel for this layer (L.bn, or 256 on the fp16 rung)
    float* wT = nullptr;  // fp32 [9*cinp][cout], SIMT debug only
};

struct Replica {
    int device = -1;
    uint8_t* blob = nullptr;
    DevConv input, head, p_pt;
    std::vector<std::vector<DevConv>> bconv;   // per block, loader order
    std::vector<Slot> slots;
    std::vector<Slot> bslots;     // the batcher's own slots (sb_eval), allocated when its workers start
    // Forwards of different slots are chained on the device: H2D / D2H copies of one slot overlap the other slot's
    // kernels, but two forwards never interleave (each conv launch fills the GPU anyway, and two interleaved batches
    // evict each other's activations from L2).
    std::unique_ptr<std::mutex> chain_mutex{new std::mutex};
    cudaEvent_t chain_tail = nullptr;   // ev_done of the forward enqueued last on this replica
    bool registered = false;            // counted in g_replicas_on_device (chained launches need the device to themselves)
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    int sm_count = 0;
};


// ---- batcher (sb_eval): host-side data structures ------------------------------------------------------------
// One LANE per replica (GPU): its own ring of host batches in pinned memory, its own mutex and condition variables, its
// own two worker threads (one per device slot / stream).  A calling thread is bound to a lane by a per-thread ticket
// (thread -> GPU affinity, SURVEY.md §8(e) "static game -> GPU affinity"), claims an index in the lane's FILLING batch
// under the lane's short mutex hold, packs its position into the batch record outside the lock and sleeps on the batch's
// futex word (or keeps the ticket and polls: sb_eval_submit / sb_eval_poll).  With N GPUs there are N independent
// mutexes, rings and wake-up domains: nothing on the per-evaluation path is shared between GPUs.
struct HostBatch {
    sb_packed_position* rec = nullptr;   // pinned [max_batch]
    std::atomic<float*> raw{nullptr};    // pinned [max_batch][SB_PLANE_FLOATS], allocated on first unpackable position
    float* out = nullptr;                // pinned [max_batch][kOutFloats]
    int count = 0;                       // claimed entries (under Lane::m)
    std::atomic<int> ready{0};           // entries whose record is complete
    std::atomic<int> readers{0};         // callers that still have to copy their result out
    std::atomic<uint32_t> done_seq{0};   // futex word: bumped when the results (or the error) are published
    std::atomic<int> any_raw{0};
    int rc = 0;
    std::string err;
    std::chrono::steady_clock::time_point first, last;   // arrival of the first / the latest position
};

struct Lane {
    int gpu = 0;
    std::mutex m;
    std::condition_variable cv_work;     // workers: a batch was closed / the first position of a batch arrived
    std::condition_variable cv_space;    // callers: a batch became free
    std::vector<std::unique_ptr<HostBatch>> ring;
    std::vector<int> free_list;
    std::deque<int> closed;
    int fill = -1;
    int cur_wait_us = 200;               // adaptive (cf. batch_forward_pipe.cc:120-147): halved when a timer close saw no
                                         // arrival during the second half of the wait, doubled when batches fill up
    int in_flight = 0;                   // batches currently running on this GPU (under m)
    std::chrono::steady_clock::time_point last_finish;   // when the latest batch of this GPU was published
};

struct Batcher {
    std::vector<std::unique_ptr<Lane>> lanes;
    std::atomic<int> batch_size{0};
    std::atomic<int> wait_us{200};       // configured ceiling (reference default gpu_waittime = 2 ms, config.cc:59)
    std::atomic<bool> quit{false};
    std::atomic<int> outstanding{0};     // calls between claim and result copy (blocking callers and open tickets)
    std::vector<std::thread> workers;
    std::atomic<long long> n_batches{0}, n_positions{0}, n_full{0}, n_timer{0}, n_raw{0};
    int max_batch = 0;
    bool buffers_freed = false;
    void FreeBuffers() {
        if (buffers_freed) return;
        for (auto& ln : lanes)
            for (auto& hb : ln->ring) {
                cudaFreeHost(hb->rec);
                cudaFreeHost(hb->raw.load());
                cudaFreeHost(hb->out);
                hb->rec = nullptr;
                hb->raw.store(nullptr);
                hb->out = nullptr;
            }
        buffers_freed = true;
    }
};

}  // namespace sb

using namespace sb;

struct sb_engine {
    HostNet net_shape;  // scalar fields + se sizes only (tensors dropped after packing)
    std::vector<int> se_sizes;
    std::vector<int> block_types;      // SB_BLOCK_* per block
    std::vector<int> inner_channels;   // bottleneck / feed-forward width per block (0 for plain residual blocks)
    std::vector<int> dw_kernels;       // Mixer depthwise filter size per block (0 otherwise)
    int replk_kernel = 0;              // RepLK policy head depthwise filter size (0 = Normal head)
    BlobLayout layout;
    std::vector<Replica> replicas;
    Geom geom;
    int max_batch = 0;
    int rows_alloc = 0;
    int precision = SB_PRECISION_FP32_SPLIT;
    int collect_stats = 0;
    int stats_launch = -1;   // conv launch index within a forward whose counters are kept (-1: every launch, last wins)
    int conv_dbg = 0;
    int chunk_accumulate = 1;    // split rung: the main accumulator of a 3x3 conv is re-accumulated in fp32 RN per k-half (0 = whole K in TMEM)
    int acc_comp_ppb = 12;       // split rung: a drained chunk is scaled by 1 + ppb * 1e-9 * (main MMAs of the chunk): the expected
                                 // loss of the tensor core's truncating fp32 accumulation (measured: the TC - SIMT slope crosses
                                 // zero at 9-13 ppb on 1..41-layer towers, profiles/r02_precision_probe.log)
    int fuse_se_pool = 1;        // SE pooling partials come from the conv epilogue (0 = separate pooling pass over the tensor)
    int wide_n = 1;              // fp16 rung: N = 256 tiles for 256-wide layers (read at engine creation / reload)
    int resident_weights = 1;    // fp16 rung: keep a C <= 128 layer's weights in shared memory for the whole launch
    int chain_forwards = 1;      // forwards of different slots of a replica run back to back, never interleaved
    int small_batch_split = 1;   // narrower N tiles when a batch does not fill one wave of CTA pairs
    int use_pdl = 1;     // programmatic dependent launch between consecutive conv3x3_tc2 launches
    int pdl_aux = 1;     // ... and for the unpack / SE / head kernels between them: 0 off, 1 for batches <= 64 (measured:
                         // -2.0..-2.5 % per forward at batch 32, +-1 % noise at batch 256; profiles/r01s2_pdl_aux_ab.md), 2 always
    int tail_split = 1;  // split the items of a partial last wave into N-halves (conv3x3_tc2)
    int conv_chain = 1;      // consecutive convolution launches of one kernel instantiation run as ONE launch, layer l+1
                             // starting on its tiles while layer l is still being finished elsewhere (conv3x3_tc2.cuh)
    std::atomic<long long> chained_layers{0};   // convolutions issued as part of a multi-layer launch (sb_launch_count counts launches)
    int layer_overlap = 1;   // convolution launches depend on their producer tile by tile instead of grid by grid
                             // (conv3x3_tc2.cuh, "Cross-layer dependencies"; needs use_pdl): 0 off, 1 split rung at batches
                             // <= 64, 2 always.  Measured (profiles/r02_layer_overlap.md): -2..-4 % per forward at batch
                             // 1..64 on the split rung; nothing at batch 256 (the SMs are held by the previous layer's CTAs
                             // anyway); -8 % THROUGHPUT on the fp16 rung, whose epilogue warps are the critical path and pay
                             // for the gpu-scope release per item
    int pack_inputs = 1; // pageable inputs travel as compact exact records (host_pack.cc); 0 = always fp32 staging
    int pack_threads = 4;
    int batcher_batch = 0;      // 0 = max_batch
    int batcher_wait_us = 200;
    std::atomic<sb::Batcher*> batcher{nullptr};             // sb_eval; owned by `batchers`
    std::vector<std::unique_ptr<sb::Batcher>> batchers;     // the live batcher and the retired ones (freed at sb_destroy: a
                                                            // caller that raced with a stop may still hold a pointer)
    std::mutex batcher_start_mutex;
    std::mutex error_mutex;
    int n_slots = 2;
    bool weights_ready = false;
    std::atomic<long long> launches{0};
    // weight distribution (DistributeBlob): ONE host->device upload per (re)load, then a device-side broadcast
    std::vector<void*> nccl_comms;     // ncclComm_t per replica (in-process communicator, created on first use)
    long long stat_h2d_uploads = 0;    // blob uploads from host memory since creation
    long long stat_d2d_fills = 0;      // replicas filled device-to-device since creation
    int bcast_method = 0;              // 0 none (single replica), 1 ncclBroadcast, 2 cudaMemcpyPeerAsync
    int nccl_version = 0;
    int verify_ok = 1;                 // device checksums of all replicas agreed after the last (re)load
    double bcast_ms = 0.0;
    int use_nccl = 1;                  // option "nccl": 0 forces the peer-copy broadcast
    std::string last_error;
    std::vector<float> host_wT_scratch;
};

static std::string g_create_error;

namespace sb {

static bool Split(const sb_engine* e) { return e->precision != SB_PRECISION_FP16; }
// Replicas (of any engine of this process) per device.  A chained launch waits, inside the kernel, for tiles owned by
// CTAs of the same grid: every CTA of the grid must become resident without any waiter having to finish.  That holds
// when the forwards on a device are serialised — one replica, chain_forwards — and not when two replicas (two engines, or
// one engine created with the same device twice) may each hold a part of the SMs with a part of their grid.
constexpr int kMaxDevices = 64;
static std::atomic<int> g_replicas_on_device[kMaxDevices];
// Convolutions are chained into one launch when option conv_chain is on, programmatic dependent launch is on, forwards of
// the replica are serialised and the replica has its device to itself (within this process: another PROCESS on the same
// device time-slices the whole GPU unless MPS is used; under MPS set conv_chain = 0).
static bool ChainMode(const sb_engine* e, const Replica& r) {
    return e->conv_chain && e->use_pdl && e->chain_forwards && e->precision != SB_PRECISION_SIMT_DEBUG && r.device >= 0 &&
           r.device < kMaxDevices && g_replicas_on_device[r.device].load(std::memory_order_relaxed) == 1;
}

static bool LayerOverlap(const sb_engine* e, int n) {
    if (!e->use_pdl || e->precision == SB_PRECISION_SIMT_DEBUG) return false;
    return e->layer_overlap == 2 || (e->layer_overlap == 1 && Split(e) && n <= 64);
}
// epilogue warps per TMEM lane quadrant of the conv kernel, per rung (conv3x3_tc2.cuh)
constexpr int kMaxDoneLaunches = 256;   // conv launches of a forward that can publish per-tile completion counters
constexpr int kPartsSplit = SB_TC2_EPI_PARTS_SPLIT, kPartsFp16 = SB_TC2_EPI_PARTS_FP16, kPartsFp16Wide = SB_TC2_EPI_PARTS_FP16_WIDE;

static void FreeSlot(Slot& s) {
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.ev_a) cudaEventDestroy(s.ev_a);
    if (s.ev_b) cudaEventDestroy(s.ev_b);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    cudaFreeHost(s.h_in);
    cudaFree(s.d_in);
    cudaFreeHost(s.h_meta);
    cudaFree(s.d_meta);
    cudaFreeHost(s.h_packed);
    cudaFree(s.d_packed);
    cudaFree(s.d_out);
    cudaFreeHost(s.h_out);
    for (ActBuf* a : {&s.in, &s.x, &s.t, &s.u, &s.pv, &s.ia, &s.ib, &s.ic, &s.pq}) {
        if (!a->hi) continue;
        if (a->lo != a->hi) cudaFree(a->lo);   // single-pass fp16 mode aliases lo to hi
        cudaFree(a->hi);
    }
    cudaFree(s.mask);
    cudaFree(s.gb);
    cudaFree(s.pooled);
    cudaFree(s.pool_part);
    cudaFree(s.counters);
    cudaFree(s.pint);
    cudaFree(s.pass5);
    cudaFree(s.misc15);
    cudaFree(s.d_stats);
    cudaFree(s.d_done);
    s.d_done = nullptr;
    delete s.pending;
    s.pending = nullptr;
    cudaFreeHost(s.h_err);
    s = Slot{};
}

// SE pooling partials from the conv epilogue need row groups (2^L rows) that never straddle two samples: the first sample
// starts at kGuardRows and every sample covers SS rows.  Returns L in {4, 3, 2}, or 0 when the canvas does not allow it
// (even board sizes: SS is odd) and the separate pooling kernel has to run.
static int PoolLog2(const Geom& g) {
    for (int L = 4; L >= 2; --L)
        if (g.SS % (1 << L) == 0 && kGuardRows % (1 << L) == 0) return L;
    return 0;
}

static void AllocAct(ActBuf& a, int rows, int channels, bool split) {
    a.channels = channels;
    a.rows = rows;
    const size_t bytes = (size_t)rows * channels * sizeof(__half);
    SB_CUDA(cudaMalloc(&a.hi, bytes));
    SB_CUDA(cudaMemset(a.hi, 0, bytes));
    a.tm3_hi = MakeActMap3(a.hi, rows, channels / 8);
    a.tm3_lo = a.tm3_hi;
    if (split) {
        SB_CUDA(cudaMalloc(&a.lo, bytes));
        SB_CUDA(cudaMemset(a.lo, 0, bytes));
        a.tm3_lo = MakeActMap3(a.lo, rows, channels / 8);
    } else {
        a.lo = a.hi;
    }
}

static void AllocSlotVec(sb_engine* e, Replica& r, std::vector<Slot>& slots, int count) {
    SB_CUDA(cudaSetDevice(r.device));
    const int C = e->net_shape.channels, P = e->net_shape.P;
    const int Cp = RoundUp(C, 64);
    const int rows = e->rows_alloc;
    const bool split = Split(e);
    slots.resize(count);
    for (Slot& s : slots) {
        SB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreate(&s.ev_a));
        SB_CUDA(cudaEventCreate(&s.ev_b));
        SB_CUDA(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        const size_t in_bytes = (size_t)e->max_batch * SB_PLANE_FLOATS * sizeof(float);
        SB_CUDA(cudaHostAlloc(&s.h_in, in_bytes, cudaHostAllocDefault));
        SB_CUDA(cudaMalloc(&s.d_in, in_bytes));
        SB_CUDA(cudaHostAlloc(&s.h_meta, (size_t)2 * e->max_batch * sizeof(int), cudaHostAllocDefault));
        SB_CUDA(cudaMalloc(&s.d_meta, (size_t)2 * e->max_batch * sizeof(int)));
        SB_CUDA(cudaHostAlloc(&s.h_packed, (size_t)e->max_batch * sizeof(sb_packed_position), cudaHostAllocDefault));
        SB_CUDA(cudaMalloc(&s.d_packed, (size_t)e->max_batch * sizeof(sb_packed_position)));
        const size_t out_bytes = (size_t)e->max_batch * kOutFloats * sizeof(float);
        SB_CUDA(cudaMalloc(&s.d_out, out_bytes));
        SB_CUDA(cudaHostAlloc(&s.h_out, out_bytes, cudaHostAllocDefault));
        AllocAct(s.in, rows, kInputChannelsPadded, split);
        AllocAct(s.x, rows, Cp, split);
        AllocAct(s.t, rows, Cp, split);
        AllocAct(s.u, rows, Cp, split);
        int inner_max = 0;
        for (int v : e->inner_channels) inner_max = std::max(inner_max, v);
        if (inner_max > 0) {
            const int Ip = RoundUp(inner_max, 64);
            AllocAct(s.ia, rows, Ip, split);
            AllocAct(s.ib, rows, Ip, split);
            AllocAct(s.ic, rows, Ip, split);
        }
        if (e->replk_kernel > 0) AllocAct(s.pq, rows, 64, split);
        SB_CUDA(cudaMalloc(&s.mask, (size_t)rows));
        SB_CUDA(cudaMemset(s.mask, 0, (size_t)rows));
        SB_CUDA(cudaMalloc(&s.gb, (size_t)e->max_batch * 2 * C * sizeof(float)));
        SB_CUDA(cudaMalloc(&s.pooled, (size_t)e->max_batch * 2 * std::max(C, 64) * sizeof(float)));
        if (PoolLog2(e->geom) > 0)
            SB_CUDA(cudaMalloc(&s.pool_part, ((size_t)rows >> PoolLog2(e->geom)) * 2 * C * sizeof(float)));
        SB_CUDA(cudaMalloc(&s.counters, (size_t)e->max_batch * sizeof(int)));
        SB_CUDA(cudaMemset(s.counters, 0, (size_t)e->max_batch * sizeof(int)));
        AllocAct(s.pv, rows, 64, split);
        SB_CUDA(cudaMalloc(&s.pint, (size_t)e->max_batch * P * sizeof(float)));
        SB_CUDA(cudaMalloc(&s.pass5, (size_t)e->max_batch * 5 * sizeof(float)));
        SB_CUDA(cudaMalloc(&s.misc15, (size_t)e->max_batch * 15 * sizeof(float)));
        SB_CUDA(cudaMalloc(&s.d_stats, (size_t)r.sm_count * 8 * sizeof(long long)));
        SB_CUDA(cudaMemset(s.d_stats, 0, (size_t)r.sm_count * 8 * sizeof(long long)));
        s.pending = new PendingChain();
        s.done_stride = e->geom.n_super(e->max_batch) + 1;
        SB_CUDA(cudaMalloc(&s.d_done, (size_t)kMaxDoneLaunches * s.done_stride * sizeof(int)));
        SB_CUDA(cudaMemset(s.d_done, 0, (size_t)kMaxDoneLaunches * s.done_stride * sizeof(int)));
        SB_CUDA(cudaHostAlloc(&s.h_err, sizeof(int), cudaHostAllocMapped));
        *s.h_err = 0;
        SB_CUDA(cudaHostGetDevicePointer(&s.d_err, s.h_err, 0));
        s.sizes.assign(e->max_batch, 0);
        s.offsets.assign(e->max_batch, 0);
    }
}

static void AllocSlots(sb_engine* e, Replica& r) { AllocSlotVec(e, r, r.slots, e->n_slots); }

static void MakeConvMaps(const Replica& r, DevConv& c, bool wide_n) {
    const int K = c.L.taps * c.L.cinp;
    // CTA-pair kernel, fp16 rung: a layer wider than 128 (up to 256) runs as ONE N tile (the accumulators of the fp16
    // rung leave room in TMEM: 2 stages x N <= 512 columns): the activation slab is loaded, and read from shared memory
    // by the tensor core, once instead of once per narrow tile (20bx256: 43 k -> 60 k evals/s)
    c.bn0 = (wide_n && c.L.coutp > 128 && c.L.coutp <= 256) ? c.L.coutp : c.L.bn;
    c.levels = 0;
    for (int lv = 0; lv < 4; ++lv) {
        const int bn = c.bn0 >> lv;
        if (bn < 16 || bn % 16 || (bn << lv) != c.bn0) break;    // UMMA N (M = 256) must be a multiple of 16
        c.tm2_hi[lv] = MakeMap2D(r.blob + c.L.w_hi, c.L.coutp, K, bn / 2);
        c.tm2_lo[lv] = MakeMap2D(r.blob + c.L.w_lo, c.L.coutp, K, bn / 2);
        c.levels = lv + 1;
    }
    for (int lv = c.levels; lv < 4; ++lv) {
        c.tm2_hi[lv] = c.tm2_hi[c.levels - 1];
        c.tm2_lo[lv] = c.tm2_lo[c.levels - 1];
    }
}

// fp32 transposed weights for the SIMT cross-check kernel, derived from the packed hi/lo matrices.
static void MakeSimtWeights(sb_engine* e, Replica& r, DevConv& c, const std::vector<uint8_t>& blob) {
    const size_t K = (size_t)c.L.taps * c.L.cinp;
    std::vector<float> wT(K * c.L.cout);
    const __half* hi = reinterpret_cast<const __half*>(blob.data() + c.L.w_hi);
    const __half* lo = reinterpret_cast<const __half*>(blob.data() + c.L.w_lo);
    for (int o = 0; o < c.L.cout; ++o)
        for (size_t k = 0; k < K; ++k) wT[k * c.L.cout + o] = __half2float(hi[o * K + k]) + __half2float(lo[o * K + k]);
    if (!c.wT) SB_CUDA(cudaMalloc(&c.wT, wT.size() * sizeof(float)));
    SB_CUDA(cudaMemcpy(c.wT, wT.data(), wT.size() * sizeof(float), cudaMemcpyHostToDevice));
    (void)e;
    (void)r;
}

static void BuildReplica(sb_engine* e, Replica& r, const std::vector<uint8_t>* blob) {
    SB_CUDA(cudaSetDevice(r.device));
    if (!r.registered && r.device >= 0 && r.device < kMaxDevices) {
        g_replicas_on_device[r.device].fetch_add(1);
        r.registered = true;
    }
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, r.device));
    if (prop.major != 10) {
        throw CudaError{"sayuri_b200 requires a Blackwell sm_100 device; device " + std::to_string(r.device) + " is sm_" +
                        std::to_string(prop.major) + std::to_string(prop.minor)};
    }
    r.sm_count = prop.multiProcessorCount;
    if (!r.blob) SB_CUDA(cudaMalloc(&r.blob, e->layout.bytes));   // filled by DistributeBlob (one upload + broadcast)
    const int blocks = e->net_shape.blocks;
    r.input.L = e->layout.input;
    const bool wide_n = !Split(e) && e->wide_n;
    MakeConvMaps(r, r.input, wide_n);
    r.head.L = e->layout.head;
    MakeConvMaps(r, r.head, wide_n);
    if (e->layout.p_dw.k > 0) {
        r.p_pt.L = e->layout.p_pt;
        MakeConvMaps(r, r.p_pt, wide_n);
    }
    r.bconv.resize(blocks);
    for (int b = 0; b < blocks; ++b) {
        r.bconv[b].resize(e->layout.bconv[b].size());
        for (size_t q = 0; q < r.bconv[b].size(); ++q) {
            r.bconv[b][q].L = e->layout.bconv[b][q];
            MakeConvMaps(r, r.bconv[b][q], wide_n);
        }
    }
    if (blob && e->precision == SB_PRECISION_SIMT_DEBUG) {
        MakeSimtWeights(e, r, r.input, *blob);
        MakeSimtWeights(e, r, r.head, *blob);
        if (e->layout.p_dw.k > 0) MakeSimtWeights(e, r, r.p_pt, *blob);
        for (int b = 0; b < blocks; ++b) {
            for (auto& c : r.bconv[b]) MakeSimtWeights(e, r, c, *blob);
        }
    }
    // dynamic shared memory opt-in, sized for the widest supported net (C = 256) so that engines of different
    // widths can coexist in one process
    for (int act = 0; act < 8; ++act) {
        SB_DISPATCH_ACT(act, ACT,
            SB_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<true, ACT, false, kPartsSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<true>::kSmemBytes));
            SB_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<false, ACT, false, kPartsFp16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<false>::kSmemBytes));
            SB_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<false, ACT, false, kPartsFp16Wide>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<false>::kSmemBytes)));
    }
    SB_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<true, kIdentity, true, kPartsSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<true>::kSmemBytes));
    SB_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<false, kIdentity, true, kPartsFp16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<false>::kSmemBytes));
    SB_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<false, kIdentity, true, kPartsFp16Wide>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<false>::kSmemBytes));
}


// ------------------------------------------------------------------------------------------------
// Weight distribution at (re)load.  The reference uploads every tensor to every GPU from host memory
// (CudaForwardPipe::Construct -> NNGraph::ConstructGraph per device, /root/reference/src/neural/cuda/cuda_forward_pipe.cc:85-116,
// 440-552 via MallocAndCopy, cuda_common.cc:228-243).  Here the packed blob crosses PCIe ONCE, into replica 0, and is
// then broadcast device-to-device over NVLink / NVSwitch: ncclBroadcast on an in-process communicator (libnccl.so.2 is
// opened at run time, so the library still loads on a box without it), or cudaMemcpyPeerAsync when NCCL cannot serve
// the device list (e.g. two replicas on one device).  A device-side checksum of every replica is compared afterwards.
typedef int (*NcclCommInitAllFn)(void**, int, const int*);
typedef int (*NcclCommDestroyFn)(void*);
typedef int (*NcclGroupFn)();
typedef int (*NcclBroadcastFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclGetVersionFn)(int*);
typedef const char* (*NcclGetErrorStringFn)(int);
struct NcclApi {
    void* lib = nullptr;
    NcclCommInitAllFn CommInitAll = nullptr;
    NcclCommDestroyFn CommDestroy = nullptr;
    NcclGroupFn GroupStart = nullptr, GroupEnd = nullptr;
    NcclBroadcastFn Broadcast = nullptr;
    NcclGetVersionFn GetVersion = nullptr;
    NcclGetErrorStringFn GetErrorString = nullptr;
    bool ok = false;
};
static NcclApi& Nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.CommInitAll = (NcclCommInitAllFn)dlsym(api.lib, "ncclCommInitAll");
        api.CommDestroy = (NcclCommDestroyFn)dlsym(api.lib, "ncclCommDestroy");
        api.GroupStart = (NcclGroupFn)dlsym(api.lib, "ncclGroupStart");
        api.GroupEnd = (NcclGroupFn)dlsym(api.lib, "ncclGroupEnd");
        api.Broadcast = (NcclBroadcastFn)dlsym(api.lib, "ncclBroadcast");
        api.GetVersion = (NcclGetVersionFn)dlsym(api.lib, "ncclGetVersion");
        api.GetErrorString = (NcclGetErrorStringFn)dlsym(api.lib, "ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast;
    });
    return api;
}
constexpr int kNcclUint8 = 1;   // ncclDataType_t::ncclUint8 (nccl.h)

__global__ void blob_checksum_kernel(const uint64_t* __restrict__ w, size_t n, unsigned long long* out) {
    uint64_t h = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = w[i] + (i + 1) * 0x9e3779b97f4a7c15ull;     // position-dependent splitmix64 finaliser, summed
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        h += z ^ (z >> 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)h);
}

static uint64_t DeviceChecksum(sb_engine* e, Replica& r) {
    SB_CUDA(cudaSetDevice(r.device));
    unsigned long long* d = nullptr;
    SB_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    SB_CUDA(cudaMemset(d, 0, sizeof(unsigned long long)));
    blob_checksum_kernel<<<r.sm_count * 4, 256>>>(reinterpret_cast<const uint64_t*>(r.blob), e->layout.bytes / 8, d);
    unsigned long long h = 0;
    cudaError_t err = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    SB_CUDA(err);
    e->launches++;
    return h;
}

static void DestroyComms(sb_engine* e) {
    if (!e->nccl_comms.empty() && Nccl().ok)
        for (void* c : e->nccl_comms)
            if (c) Nccl().CommDestroy(c);
    e->nccl_comms.clear();
}

// `host_blob` == nullptr: replica 0 already holds the blob (sb_weights_import), broadcast only.
static void DistributeBlob(sb_engine* e, const std::vector<uint8_t>* host_blob) {
    const size_t bytes = e->layout.bytes;
    const int n = (int)e->replicas.size();
    Replica& r0 = e->replicas[0];
    SB_CUDA(cudaSetDevice(r0.device));
    if (host_blob) {
        SB_CUDA(cudaMemcpy(r0.blob, host_blob->data(), bytes, cudaMemcpyHostToDevice));   // the ONLY PCIe crossing of the weights
        e->stat_h2d_uploads++;
    }
    e->bcast_method = 0;
    e->bcast_ms = 0.0;
    if (n > 1) {
        const auto t0 = std::chrono::steady_clock::now();
        bool done = false;
        std::vector<int> devs;
        for (Replica& r : e->replicas) devs.push_back(r.device);
        std::vector<int> uniq = devs;
        std::sort(uniq.begin(), uniq.end());
        const bool distinct = std::adjacent_find(uniq.begin(), uniq.end()) == uniq.end();
        if (e->use_nccl && distinct && Nccl().ok) {
            NcclApi& nc = Nccl();
            if (e->nccl_comms.empty()) {
                e->nccl_comms.assign(n, nullptr);
                if (nc.CommInitAll(e->nccl_comms.data(), n, devs.data()) != 0) e->nccl_comms.clear();
                if (nc.GetVersion) nc.GetVersion(&e->nccl_version);
            }
            if (!e->nccl_comms.empty()) {
                std::vector<cudaStream_t> st(n, nullptr);
                for (int i = 0; i < n; ++i) {
                    SB_CUDA(cudaSetDevice(devs[i]));
                    SB_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
                }
                int rc = nc.GroupStart();
                for (int i = 0; i < n && rc == 0; ++i) {
                    SB_CUDA(cudaSetDevice(devs[i]));
                    rc = nc.Broadcast(e->replicas[i].blob, e->replicas[i].blob, bytes, kNcclUint8, 0, e->nccl_comms[i], st[i]);
                }
                const int rc_end = nc.GroupEnd();
                if (rc == 0) rc = rc_end;
                for (int i = 0; i < n; ++i) {
                    SB_CUDA(cudaSetDevice(devs[i]));
                    cudaError_t se = cudaStreamSynchronize(st[i]);
                    cudaStreamDestroy(st[i]);
                    SB_CUDA(se);
                }
                if (rc != 0)
                    throw CudaError{std::string("ncclBroadcast of the weight blob failed: ") +
                                    (nc.GetErrorString ? nc.GetErrorString(rc) : "unknown NCCL error")};
                e->bcast_method = 1;
                done = true;
            }
        }
        if (!done) {   // peer copies replica 0 -> replica i (NVLink P2P when the devices differ; plain D2D on one device)
            SB_CUDA(cudaSetDevice(r0.device));
            for (int i = 1; i < n; ++i) {
                if (devs[i] != r0.device) {
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, r0.device, devs[i]);
                    if (can) {
                        cudaError_t pe = cudaDeviceEnablePeerAccess(devs[i], 0);
                        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) SB_CUDA(pe);
                        cudaGetLastError();
                    }
                }
                SB_CUDA(cudaMemcpyPeerAsync(e->replicas[i].blob, devs[i], r0.blob, r0.device, bytes, 0));
            }
            SB_CUDA(cudaDeviceSynchronize());
            e->bcast_method = 2;
        }
        e->stat_d2d_fills += n - 1;
        e->bcast_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    // every replica must hold the same bytes
    e->verify_ok = 1;
    if (n > 1) {
        const uint64_t h0 = DeviceChecksum(e, r0);
        for (int i = 1; i < n; ++i) {
            if (DeviceChecksum(e, e->replicas[i]) != h0) {
                e->verify_ok = 0;
                throw CudaError{"weight broadcast verification failed: replica " + std::to_string(i) + " differs from replica 0"};
            }
        }
    }
}

static void DestroyReplica(Replica& r) {
    if (r.registered && r.device >= 0 && r.device < kMaxDevices) g_replicas_on_device[r.device].fetch_sub(1);
    r.registered = false;
    if (r.device >= 0) cudaSetDevice(r.device);
    for (Slot& s : r.slots) FreeSlot(s);
    r.slots.clear();
    for (Slot& s : r.bslots) FreeSlot(s);
    r.bslots.clear();
    auto free_conv = [](DevConv& c) {
        cudaFree(c.wT);
        c.wT = nullptr;
    };
    free_conv(r.input);
    free_conv(r.head);
    free_conv(r.p_pt);
    for (auto& blk : r.bconv)
        for (auto& c : blk) free_conv(c);
    cudaFree(r.blob);
    r.blob = nullptr;
    cudaFree(r.flush_buf);
    r.flush_buf = nullptr;
}

// ------------------------------------------------------------------------------------------------
// Forward orchestration
// Profiling pass of sb_time_forward: event pairs around every group of NON-convolution kernels of one forward, so that
// the convolution launches keep their back-to-back programmatic-dependent-launch overlap exactly as in a normal step;
// time of the conv kernel = duration of the forward - sum of the bracketed groups.
struct ConvTimer {
    std::vector<cudaEvent_t> ev;  // pairs around the other kernels
    bool on = false;
    int conv_launches = 0;
    void Mark(cudaStream_t st) {
        if (!on) return;
        cudaEvent_t a;
        SB_CUDA(cudaEventCreate(&a));
        SB_CUDA(cudaEventRecord(a, st));
        ev.push_back(a);
    }
};

static void FlushChain(sb_engine* e, Slot& s) {
    PendingChain& pc = *s.pending;
    if (pc.chain.n_layers == 0) return;
    // launched with the programmatic-stream-serialization attribute: see pdl_wait() in conv3x3_tc2.cuh
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pc.grid2);
    cfg.blockDim = dim3(pc.split ? Conv2Cfg<true, kPartsSplit>::kThreads
                                 : pc.wide ? Conv2Cfg<false, kPartsFp16Wide>::kThreads : Conv2Cfg<false, kPartsFp16>::kThreads);
    cfg.stream = s.stream;
    cfg.dynamicSmemBytes = pc.split ? Conv2Cfg<true>::kSmemBytes : Conv2Cfg<false>::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const ConvChain& ch = pc.chain;
    if (pc.pool) {   // the last convolution of an SE block: identity activation, pooling partials from the epilogue
        if (pc.split) SB_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<true, kIdentity, true, kPartsSplit>, ch));
        else if (pc.wide) SB_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<false, kIdentity, true, kPartsFp16Wide>, ch));
        else SB_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<false, kIdentity, true, kPartsFp16>, ch));
    } else if (pc.split) {
        SB_DISPATCH_ACT(pc.act, ACT, SB_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<true, ACT, false, kPartsSplit>, ch)));
    } else if (pc.wide) {
        SB_DISPATCH_ACT(pc.act, ACT, SB_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<false, ACT, false, kPartsFp16Wide>, ch)));
    } else {
        SB_DISPATCH_ACT(pc.act, ACT, SB_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<false, ACT, false, kPartsFp16>, ch)));
    }
    SB_CUDA(cudaGetLastError());
    e->launches++;
    if (pc.chain.n_layers > 1) e->chained_layers += pc.chain.n_layers;
    pc.chain.n_layers = 0;
}

static void LaunchConv(sb_engine* e, Replica& r, Slot& s, const DevConv& c, const ActBuf& in, ActBuf& out,
                       const ActBuf* res, int act, int n, ConvTimer* tm, bool pool = false) {
    const int n_super = e->geom.n_super(n);
    if (tm && tm->on) tm->conv_launches++;
    if (e->precision == SB_PRECISION_SIMT_DEBUG) {
        dim3 grid((n_super * kSuperRows + 7) / 8, (c.L.cout + 31) / 32), block(32, 8);
        conv3x3_simt_kernel<<<grid, block, 0, s.stream>>>(in.hi, in.lo, true, c.L.cinp, c.wT,
                                                          reinterpret_cast<const float*>(r.blob + c.L.bias),
                                                          res ? res->hi : nullptr, res ? res->lo : nullptr, s.mask,
                                                          c.L.cout, n_super * kSuperRows, e->geom.P, c.L.taps, act, out.hi,
                                                          out.lo, out.rows);
        out.done = nullptr;
        SB_CUDA(cudaGetLastError());
        e->launches++;
        return;
    }
    if (pool && act != kIdentity) throw CudaError{"internal error: pooled convolution with an activation"};
    ConvLayer lay;
    ConvParams& p = lay.p;
    p.out_hi = out.hi;
    p.out_lo = out.lo;
    p.res_hi = res ? res->hi : nullptr;
    p.res_lo = res ? res->lo : nullptr;
    p.bias = reinterpret_cast<const float*>(r.blob + c.L.bias);
    p.mask = s.mask;
    p.cout = c.L.cout;
    p.rows = out.rows;
    p.kh = c.L.kh;
    p.n_super = n_super;
    p.pitch = e->geom.P;
    p.ntaps = c.L.taps;
    p.dbg = e->conv_dbg;
    // split rung: the main accumulator is drained and re-accumulated in fp32 RN after every k-half (conv3x3_tc2.cuh,
    // "Precision"); a 1x1 convolution (<= 24 main MMAs) is one chunk
    p.chunk_kh = (c.L.taps == 9 && e->chunk_accumulate) ? 1 : c.L.kh;
    p.chunk_scale = 1.0f + 1e-9f * (float)e->acc_comp_ppb * (float)(4 * p.chunk_kh * c.L.taps);
    p.pool_part = pool ? s.pool_part : nullptr;
    p.pool_log2 = PoolLog2(e->geom);
    p.pool_groups = (n * e->geom.SS) >> std::max(p.pool_log2, 1);
    p.pool_c = e->net_shape.channels;
    p.err = s.d_err;
    p.stats = (e->collect_stats && (e->stats_launch < 0 || e->stats_launch == s.conv_counter)) ? s.d_stats : nullptr;
    // Small batches: narrow the N tile (bn >> level) while all items still fit in one wave, so that a handful of
    // positions is spread over up to 4x more CTA pairs (latency of the single-position / GTP case).
    int level = 0;
    const int max_pairs = r.sm_count / 2;
    while (e->small_batch_split && level + 1 < c.levels && level < 2 &&
           n_super * (c.L.coutp / (c.bn0 >> (level + 1))) <= max_pairs)
        ++level;
    p.bn = c.bn0 >> level;
    p.n_ntiles = c.L.coutp / p.bn;
    const int items2 = n_super * p.n_ntiles;
    // fp16 rung, one N tile, <= 18 weight stages per item: weights stay resident in shared memory
    p.resident = (e->resident_weights && !Split(e) && p.n_ntiles == 1 && c.L.kh * c.L.taps <= Conv2Cfg<false>::kNumBStages &&
                  items2 > max_pairs) ? 1 : 0;
    const int pairs = std::min(items2, max_pairs);
    const int grid2 = 2 * pairs;
    const int rem = items2 % pairs;
    const bool split = Split(e);
    const bool wide = !split && p.bn > 128;   // N = 256 tiles of the fp16 rung: tensor-bound, fewer epilogue warps
    // weight-ring geometry (must agree along a chain): stage size class, and for resident weights the stages per layer
    const int ring_key = (p.bn > 128 ? 1 : 0) | (p.resident ? (c.L.kh * c.L.taps) << 1 : 0);

    // Dependencies.  The launch publishes per-tile completion counters when something may consume them; it depends on its
    // producer tile by tile when the input is the output of a launch that published them and the residual (if any) is that
    // launch's own input (conv3x3_tc2.cuh, "Cross-layer dependencies").
    const bool chain_mode = s.fwd_chain && !pool && p.stats == nullptr && items2 >= pairs;
    const bool overlap = s.fwd_overlap;
    const bool counters = (overlap || s.fwd_chain) && s.conv_counter < kMaxDoneLaunches;
    const bool deps_ok = in.done != nullptr && (res == nullptr || res == in.done_src);
    PendingChain& pc = *s.pending;
    const bool append = chain_mode && pc.chain.n_layers > 0 && pc.chain.n_layers < kMaxChain && deps_ok && !pc.pool &&
                        in.done == pc.chain.layer[pc.chain.n_layers - 1].p.done_out && pc.split == split && pc.wide == wide &&
                        pc.act == act && pc.grid2 == grid2 && pc.resident == p.resident && pc.ring_key == ring_key && pc.bn == p.bn &&
                        pc.bias_floats + c.L.coutp <= kChainBiasFloats;
    if (!append) FlushChain(e, s);
    if (c.L.coutp > kChainBiasFloats) throw CudaError{"internal error: convolution wider than the bias staging area"};
    p.bias_off = append ? pc.bias_floats : 0;
    p.done_out = counters ? s.d_done + (size_t)s.conv_counter * s.done_stride : nullptr;
    const bool tile_deps = deps_ok && (append || overlap);
    p.done_in = tile_deps ? in.done : nullptr;
    p.done_in_full = in.done_full;
    s.conv_counter++;
    // persistent CTA pairs; the items of a partial last wave are split into N-halves when that makes the wave half as long
    // (conv_unit in conv3x3_tc2.cuh).  Inside a chain the next layer fills the partial wave instead: no split, and the pair
    // that takes the first item rotates by the remainder from layer to layer.
    int n_tail = 0;
    if (e->tail_split && !chain_mode && !p.resident && items2 > pairs && rem > 0 && 2 * rem <= pairs && level + 1 < c.levels) n_tail = rem;
    p.n_full = items2 - n_tail;
    p.n_units = items2 + n_tail;
    p.rot = append ? (pc.chain.layer[pc.chain.n_layers - 1].p.rot + pc.last_rem) % pairs : 0;
    out.done = p.done_out;
    out.done_full = 8 * p.n_ntiles * p.bn;
    out.done_src = &in;
    lay.tmA_hi = in.tm3_hi;
    lay.tmA_lo = split ? in.tm3_lo : in.tm3_hi;
    lay.tmW_hi = c.tm2_hi[level];
    lay.tmW_lo = split ? c.tm2_lo[level] : c.tm2_hi[level];
    lay.tmWq_hi = c.tm2_hi[level + 1];
    lay.tmWq_lo = split ? c.tm2_lo[level + 1] : c.tm2_hi[level + 1];
    if (!append) {
        pc.split = split;
        pc.wide = wide;
        pc.pool = pool;
        pc.act = act;
        pc.grid2 = grid2;
        pc.resident = p.resident;
        pc.ring_key = ring_key;
        pc.bn = p.bn;   // the TMEM accumulator stages are laid out in multiples of bn: one layout along a chain
    }
    pc.last_rem = rem;
    pc.bias_floats = p.bias_off + c.L.coutp;
    pc.chain.layer[pc.chain.n_layers++] = lay;
    if (!chain_mode) FlushChain(e, s);   // not chainable: a chain of one, launched now
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait() in common.cuh).
template <typename... KArgs, typename... Args>
static void LaunchPdl(sb_engine* e, int n, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                      Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (e->use_pdl && (e->pdl_aux == 2 || (e->pdl_aux == 1 && n <= 64))) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SB_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

static void LaunchDw(sb_engine* e, Slot& s, const DwLayout& d, const uint8_t* blob, const ActBuf& in, ActBuf& out, int act,
                     bool add_input, int n, int n_rows) {
    const dim3 grid((n_rows + 127) / 128, d.ch / 8);
    const float* w = reinterpret_cast<const float*>(blob + d.w);
    const float* b = reinterpret_cast<const float*>(blob + d.b);
    const int halo = (d.k / 2) * (e->geom.P + 1);
    const size_t smem = ((size_t)d.k * d.k * 8 + 8 + (size_t)(128 + 2 * halo) * 8) * sizeof(float);
    SB_DISPATCH_ACT(act, ACT, (dwconv_kernel<ACT><<<grid, 128, smem, s.stream>>>(in.hi, in.lo, out.hi, out.lo, Split(e), w, b, s.d_meta,
                                                                              e->geom, n, n_rows, in.rows, out.rows, d.k, add_input)));
    out.done = nullptr;   // not a convolution launch: consumers depend on the whole grid
    SB_CUDA(cudaGetLastError());
    e->launches++;
}

// Everything between "inputs are in d_in / d_meta" and "outputs are in d_out", on the slot's stream.
static void EnqueueForward(sb_engine* e, Replica& r, Slot& s, int n, ConvTimer* tm = nullptr) {
    const HostNet& ns = e->net_shape;
    const Geom g = e->geom;
    const int C = ns.channels, P = ns.P, V = ns.V, act = ns.act;
    const bool split = Split(e);
    const int n_rows = g.n_super(n) * kSuperRows;
    const int* d_sizes = s.d_meta;
    const int* d_offsets = s.d_meta + e->max_batch;
    const BlobLayout& L = e->layout;
    auto F = [&](size_t off) { return reinterpret_cast<const float*>(r.blob + off); };

    s.conv_counter = 0;
    for (ActBuf* b : {&s.in, &s.x, &s.t, &s.u, &s.ia, &s.ib, &s.ic, &s.pv, &s.pq}) b->done = nullptr;
    s.pending->chain.n_layers = 0;
    s.fwd_chain = ChainMode(e, r);   // decided once: the counters zeroed below are the ones every launch of this forward uses
    s.fwd_overlap = LayerOverlap(e, n);
    if (s.fwd_overlap || s.fwd_chain) {   // per-tile completion counters of this forward's convolution launches
        size_t n_conv = 3;   // input, head entry, RepLK 1x1
        for (const auto& blk : r.bconv) n_conv += blk.size();
        SB_CUDA(cudaMemsetAsync(s.d_done, 0, std::min<size_t>(n_conv, kMaxDoneLaunches) * s.done_stride * sizeof(int), s.stream));
    }
    const int pool_log2 = PoolLog2(g);
    const bool pool_fused = e->fuse_se_pool && pool_log2 > 0 && e->precision != SB_PRECISION_SIMT_DEBUG;
    // brackets groups of non-convolution kernels (profiling pass); pending convolutions are issued first
    auto mark = [&]() {
        FlushChain(e, s);
        if (tm) tm->Mark(s.stream);
    };
    mark();
    {   // input planes -> canvas
        const int threads = n_rows * 8;
        if (s.packed) {
            LaunchPdl(e, n, unpack_packed_kernel, dim3((threads + 255) / 256), dim3(256), 0, s.stream, s.d_packed, s.d_in,
                      (size_t)SB_PLANE_FLOATS, g, n, n_rows, s.in.rows, e->max_batch, s.in.hi, s.in.lo, split, s.mask, s.d_meta);
        } else {
            LaunchPdl(e, n, unpack_planes_kernel, dim3((threads + 255) / 256), dim3(256), 0, s.stream, s.d_in,
                      (size_t)SB_PLANE_FLOATS, d_sizes, g, n, n_rows, s.in.rows, s.in.hi, s.in.lo, split, s.mask);
        }
        SB_CUDA(cudaGetLastError());
        e->launches++;
    }
    mark();
    ActBuf* x = &s.x;
    ActBuf* t = &s.t;
    ActBuf* u = &s.u;
    LaunchConv(e, r, s, r.input, s.in, *x, nullptr, act, n, tm);
    for (int b = 0; b < ns.blocks; ++b) {
        const int se = e->se_sizes[b];
        const std::vector<DevConv>& cv = r.bconv[b];
        // the block's last conv joins the skip connection; with an SE unit it is left linear and the skip is
        // added by se_apply (blas_forward_pipe.cc:73-87,147-161,249-262)
        const ActBuf* last_res = se > 0 ? nullptr : x;
        const int last_act = se > 0 ? kIdentity : act;
        const ActBuf* se_skip = x;
        const bool pool = se > 0 && pool_fused;
        if (e->block_types[b] == SB_BLOCK_MIXER) {
            // MixerBlockForward, blas_forward_pipe.cc:265-312: y = act(dw(x) + b) + x ; out = ffn2(act(ffn1(y))) (+ y)
            mark();
            LaunchDw(e, s, L.bdw[b], r.blob, *x, *t, act, true, n, n_rows);
            mark();
            LaunchConv(e, r, s, cv[0], *t, s.ia, nullptr, act, n, tm);
            LaunchConv(e, r, s, cv[1], s.ia, *u, se > 0 ? nullptr : t, last_act, n, tm, pool);
            se_skip = t;   // the skip of this block and of its SE unit is y
        } else if (e->block_types[b] == SB_BLOCK_RESIDUAL) {
            // ResidualBlockForward, blas_forward_pipe.cc:46-88
            LaunchConv(e, r, s, cv[0], *x, *t, nullptr, act, n, tm);
            LaunchConv(e, r, s, cv[1], *t, *u, last_res, last_act, n, tm, pool);
        } else if (e->block_types[b] == SB_BLOCK_BOTTLENECK) {
            // BottleneckBlockForward, blas_forward_pipe.cc:90-162: 1x1 down, 3x3, 3x3, 1x1 up (+ skip)
            LaunchConv(e, r, s, cv[0], *x, s.ia, nullptr, act, n, tm);
            LaunchConv(e, r, s, cv[1], s.ia, s.ib, nullptr, act, n, tm);
            LaunchConv(e, r, s, cv[2], s.ib, s.ic, nullptr, act, n, tm);
            LaunchConv(e, r, s, cv[3], s.ic, *u, last_res, last_act, n, tm, pool);
        } else {
            // NestedBottleneckBlockForward, blas_forward_pipe.cc:164-263: 1x1 down, two inner residual blocks, 1x1 up
            LaunchConv(e, r, s, cv[0], *x, s.ia, nullptr, act, n, tm);       // a
            LaunchConv(e, r, s, cv[1], s.ia, s.ib, nullptr, act, n, tm);     // b = act(conv1(a))
            LaunchConv(e, r, s, cv[2], s.ib, s.ic, &s.ia, act, n, tm);       // c = act(conv2(b) + a)
            LaunchConv(e, r, s, cv[3], s.ic, s.ib, nullptr, act, n, tm);     // d = act(conv3(c))
            LaunchConv(e, r, s, cv[4], s.ib, s.ia, &s.ic, act, n, tm);       // e = act(conv4(d) + c)
            LaunchConv(e, r, s, cv[5], s.ia, *u, last_res, last_act, n, tm, pool);
        }
        if (se > 0) {
            mark();
            const size_t smem = ((size_t)3 * C + se) * sizeof(float);
            if (pool_fused) {
                // the block's last convolution left per-(16-row group, channel) sums and maxima: fixed-order finalize + FCs
                SB_DISPATCH_ACT(act, ACT, LaunchPdl(e, n, se_fc_kernel<ACT>, dim3(n), dim3(256), smem, s.stream,
                                                    (const float*)s.pool_part, g.SS >> pool_log2, d_sizes, C, se, F(L.squeeze[b].w),
                                                    F(L.squeeze[b].b), F(L.excite[b].w), F(L.excite[b].b), s.gb));
            } else {
                SB_DISPATCH_ACT(act, ACT, LaunchPdl(e, n, se_pool_fc_kernel<ACT>, dim3((C + 63) / 64, n), dim3(256), smem, s.stream,
                                                    (const __half*)u->hi, (const __half*)u->lo, split, (const uint8_t*)s.mask, d_sizes, g, C,
                                                    u->rows, se, F(L.squeeze[b].w), F(L.squeeze[b].b), F(L.excite[b].w),
                                                    F(L.excite[b].b), s.pooled, s.counters, s.gb, e->conv_dbg));
            }
            SB_CUDA(cudaGetLastError());
            const size_t total = (size_t)n_rows * (C / 8);
            SB_DISPATCH_ACT(act, ACT, LaunchPdl(e, n, se_apply_kernel<ACT>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s.stream,
                                                u->hi, u->lo, (const __half*)se_skip->hi, (const __half*)se_skip->lo, split,
                                                (const uint8_t*)s.mask, (const float*)s.gb, g, C, u->rows, n_rows));
            SB_CUDA(cudaGetLastError());
            u->done = nullptr;   // rewritten in place by se_apply
            e->launches += 2;
            mark();
        }
        std::swap(x, u);
    }
    s.trunk = x;
    // heads: the two head-entry 1x1 convs as one single-tap tensor-core launch
    LaunchConv(e, r, s, r.head, *x, s.pv, nullptr, act, n, tm);
    if (e->replk_kernel > 0) {
        // RepLK policy head, blas_forward_pipe.cc:443-471: depthwise k x k (+bias, act) on the P policy channels, then a
        // 1x1 P -> P (+bias, act) written back over the policy channels of pv (the V value channels stay untouched)
        mark();
        LaunchDw(e, s, L.p_dw, r.blob, s.pv, s.pq, act, false, n, n_rows);
        mark();
        LaunchConv(e, r, s, r.p_pt, s.pq, s.pv, nullptr, act, n, tm);
    }
    HeadWeights hw;
    hw.p_inter_w = F(L.p_inter.w);
    hw.p_inter_b = F(L.p_inter.b);
    hw.pass_w = F(L.pass.w);
    hw.pass_b = F(L.pass.b);
    hw.v_inter_w = F(L.v_inter.w);
    hw.v_inter_b = F(L.v_inter.b);
    hw.misc_w = F(L.misc.w);
    hw.misc_b = F(L.misc.b);
    hw.prob_w = F(L.prob_w);
    hw.prob_b = F(L.prob_b);
    hw.own_w = F(L.own_w);
    hw.own_b = F(L.own_b);
    mark();
    {
        const int PV = P + V;
        const size_t smem = ((size_t)3 * P + 3 * V + P + 3 * V) * sizeof(float);
        LaunchPdl(e, n, head_fused_kernel, dim3((PV + 31) / 32, n), dim3(256), smem, s.stream, (const __half*)s.pv.hi,
                  (const __half*)s.pv.lo, split, s.pv.rows, (const uint8_t*)s.mask, d_sizes, d_offsets, g, P, V, hw, act, s.pooled,
                  s.counters, s.d_out);
        SB_CUDA(cudaGetLastError());
    }
    e->launches += 1;
    mark();
}

// Enqueue one forward on the slot's stream, ordered after the forward enqueued last on the same replica.
static void EnqueueChainedForward(sb_engine* e, Replica& r, Slot& s, int n) {
    if (!e->chain_forwards) {
        EnqueueForward(e, r, s, n);
        return;
    }
    std::lock_guard<std::mutex> lk(*r.chain_mutex);
    if (r.chain_tail && r.chain_tail != s.ev_done) SB_CUDA(cudaStreamWaitEvent(s.stream, r.chain_tail, 0));
    EnqueueForward(e, r, s, n);
    SB_CUDA(cudaEventRecord(s.ev_done, s.stream));
    r.chain_tail = s.ev_done;
}

static void CheckSlotError(Slot& s, cudaError_t err, const char* what) {
    if (err == cudaSuccess) return;
    std::string msg = std::string("CUDA Error: ") + cudaGetErrorString(err) + " in " + what;
    if (s.h_err && *s.h_err != 0) msg += " (pipeline barrier timeout at site " + std::to_string(*s.h_err) + ")";
    throw CudaError{msg};
}

static void Configure(sb_engine* e, int board, int max_batch) {
    e->geom = Geom(board);
    e->max_batch = max_batch;
    e->rows_alloc = e->geom.rows_alloc(max_batch);
    for (Replica& r : e->replicas) {
        SB_CUDA(cudaSetDevice(r.device));
        for (Slot& s : r.slots) FreeSlot(s);
        r.slots.clear();
        r.chain_tail = nullptr;
        AllocSlots(e, r);
    }
}

static int Fail(sb_engine* e, int code, const std::string& msg) {
    if (e) e->last_error = msg;
    else g_create_error = msg;
    return code;
}

static int CreateImpl(sb_engine** out, HostNet* net, const sb_net_desc* shape_only, const int* gpu_ids, int n_gpus,
                      int board, int max_batch, int precision) {
    if (!out) return Fail(nullptr, SB_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (board < 2 || board > SB_MAX_BOARD_SIZE) return Fail(nullptr, SB_ERR_INVALID, "NN board size should be in [2, 19]");
    if (max_batch < 1) return Fail(nullptr, SB_ERR_INVALID, "NN batch size should be larger than zero");
    if (precision < 0 || precision > 2) return Fail(nullptr, SB_ERR_INVALID, "unknown precision");
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
        cudaGetLastError();
        return Fail(nullptr, SB_ERR_CUDA, "No executable GPU device!");   // cuda_forward_pipe.cc:105-107
    }
    std::vector<int> gpus;
    for (int i = 0; i < n_gpus; ++i) {
        if (gpu_ids && gpu_ids[i] >= 0 && gpu_ids[i] < dev_count) gpus.push_back(gpu_ids[i]);
    }
    if (gpus.empty()) {   // reference: assign all devices automatically (cuda_forward_pipe.cc:98-104)
        if (n_gpus > 0 && gpu_ids) return Fail(nullptr, SB_ERR_INVALID, "Not found the requested GPU device(s)");
        for (int i = 0; i < dev_count; ++i) gpus.push_back(i);
    }
    std::unique_ptr<sb_engine> e(new sb_engine);
    e->precision = precision;
    if (net) {
        e->net_shape.version = net->version;
        e->net_shape.input_channels = net->input_channels;
        e->net_shape.blocks = net->blocks;
        e->net_shape.channels = net->channels;
        e->net_shape.P = net->P;
        e->net_shape.V = net->V;
        e->net_shape.act = net->act;
        e->se_sizes = net->se_sizes();
        e->block_types = net->block_types();
        e->inner_channels = net->inner_channels();
        e->dw_kernels = net->dw_kernels();
        e->replk_kernel = net->replk ? net->p_dw_conv.k : 0;
    } else {
        e->net_shape.version = shape_only->version;
        e->net_shape.input_channels = shape_only->input_channels;
        e->net_shape.blocks = shape_only->blocks;
        e->net_shape.channels = shape_only->channels;
        e->net_shape.P = shape_only->policy_channels;
        e->net_shape.V = shape_only->value_channels;
        e->net_shape.act = shape_only->activation;
        e->se_sizes.assign(shape_only->se_sizes, shape_only->se_sizes + shape_only->blocks);
        e->block_types.assign(shape_only->blocks, SB_BLOCK_RESIDUAL);
        e->inner_channels.assign(shape_only->blocks, 0);
        e->dw_kernels.assign(shape_only->blocks, 0);
        for (int b = 0; b < shape_only->blocks; ++b) {
            if (shape_only->block_types) e->block_types[b] = shape_only->block_types[b];
            const int ty = e->block_types[b];
            if (ty < SB_BLOCK_RESIDUAL || ty > SB_BLOCK_MIXER) return Fail(nullptr, SB_ERR_INVALID, "unsupported block type");
            if (ty != SB_BLOCK_RESIDUAL) {
                const int I = shape_only->inner_channels ? shape_only->inner_channels[b] : 0;
                if (I < 8 || I > 512 || I % 8) return Fail(nullptr, SB_ERR_INVALID, "unsupported bottleneck / feed-forward width");
                e->inner_channels[b] = I;
            }
            if (ty == SB_BLOCK_MIXER) {
                const int kk = shape_only->dw_kernels ? shape_only->dw_kernels[b] : 7;
                if (kk < 3 || kk > 15 || !(kk & 1)) return Fail(nullptr, SB_ERR_INVALID, "unsupported depthwise kernel size");
                e->dw_kernels[b] = kk;
            }
        }
        if (shape_only->policy_head_type == SB_POLICY_HEAD_REPLK) {
            e->replk_kernel = shape_only->policy_dw_kernel > 0 ? shape_only->policy_dw_kernel : 7;
            if (e->replk_kernel < 3 || e->replk_kernel > 15 || !(e->replk_kernel & 1)) return Fail(nullptr, SB_ERR_INVALID, "unsupported depthwise kernel size");
        } else if (shape_only->policy_head_type != SB_POLICY_HEAD_NORMAL) {
            return Fail(nullptr, SB_ERR_INVALID, "unsupported policy head type");
        }
        const int C = e->net_shape.channels, PV = e->net_shape.P + e->net_shape.V;
        if (C < 16 || C > 256 || C % 16 || (C > 128 && C % 32) || PV % 4 || PV > 64 || e->net_shape.input_channels != SB_INPUT_CHANNELS)
            return Fail(nullptr, SB_ERR_INVALID, "unsupported network shape");
    }
    {
        const int PV = e->net_shape.P + e->net_shape.V;
        if ((PV != 16 && PV != 32 && PV != 48 && PV != 64) || e->net_shape.P % 8 || e->net_shape.V % 8)
            return Fail(nullptr, SB_ERR_INVALID, "policy and value head channels must be multiples of 8 summing to 16, 32, 48 or 64");
        if (e->net_shape.channels % 8) return Fail(nullptr, SB_ERR_INVALID, "channels must be a multiple of 8");
    }
    e->layout = ComputeLayout(e->net_shape.blocks, e->net_shape.channels, e->net_shape.P, e->net_shape.V, e->se_sizes,
                              e->block_types, e->inner_channels, e->dw_kernels, e->replk_kernel);
    try {
        std::vector<uint8_t> blob;
        if (net) blob = PackBlob(*net, e->layout);
        e->replicas.resize(gpus.size());
        for (size_t i = 0; i < gpus.size(); ++i) {
            e->replicas[i].device = gpus[i];
            BuildReplica(e.get(), e->replicas[i], net ? &blob : nullptr);
        }
        if (net) DistributeBlob(e.get(), &blob);
        e->weights_ready = net != nullptr;
        Configure(e.get(), board, max_batch);
    } catch (const CudaError& ce) {
        DestroyComms(e.get());
        for (Replica& r : e->replicas) DestroyReplica(r);
        return Fail(nullptr, SB_ERR_CUDA, ce.msg);
    }
    *out = e.release();
    // SAYURI_B200_OPTIONS="key=value,key=value": sb_set_option knobs for callers that cannot reach the C ABI (A/B runs through the
    // unmodified reference front-end).  Unknown keys are reported on stderr and ignored.
    if (const char* env = std::getenv("SAYURI_B200_OPTIONS")) {
        std::string all(env);
        size_t at = 0;
        while (at < all.size()) {
            const size_t end = std::min(all.find(',', at), all.size());
            const std::string kv = all.substr(at, end - at);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos && sb_set_option(*out, kv.substr(0, eq).c_str(), std::atoi(kv.c_str() + eq + 1)) != SB_OK)
                std::fprintf(stderr, "sayuri_b200: SAYURI_B200_OPTIONS: %s ignored (%s)\n", kv.c_str(), sb_last_error(*out));
            at = end + 1;
        }
    }
    return SB_OK;
}

static int SubmitImpl(sb_engine* e, int gpu, int slot, int n, const float* planes, long long stride,
                      const float* const* plane_ptrs, const int* sizes, const int* offsets) {
    if (!e) return SB_ERR_INVALID;
    if (gpu < 0 || gpu >= (int)e->replicas.size()) return Fail(e, SB_ERR_INVALID, "gpu index out of range");
    if (slot < 0 || slot >= e->n_slots) return Fail(e, SB_ERR_INVALID, "slot index out of range");
    if (n < 1 || n > e->max_batch) return Fail(e, SB_ERR_INVALID, "batch size out of range [1, max_batch]");
    if ((!planes && !plane_ptrs) || !sizes || !offsets) return Fail(e, SB_ERR_INVALID, "null input pointer");
    if (!e->weights_ready) return Fail(e, SB_ERR_STATE, "weights have not been loaded");
    Replica& r = e->replicas[gpu];
    Slot& s = r.slots[slot];
    if (s.busy) return Fail(e, SB_ERR_STATE, "slot is busy: call sb_wait first");
    for (int i = 0; i < n; ++i) {
        if (sizes[i] < 2 || sizes[i] > e->geom.N) return Fail(e, SB_ERR_INVALID, "board size of a sample exceeds the NN canvas");
        if (offsets[i] < 0 || offsets[i] > 4) return Fail(e, SB_ERR_INVALID, "policy offset must be in [0, 4]");
    }
    try {
        SB_CUDA(cudaSetDevice(r.device));
        cudaGetLastError();   // drop any stale non-sticky error of an unrelated earlier call
        for (int i = 0; i < n; ++i) {
            s.h_meta[i] = sizes[i];
            s.h_meta[e->max_batch + i] = offsets[i];
            s.sizes[i] = sizes[i];
            s.offsets[i] = offsets[i];
        }
        SB_CUDA(cudaMemcpyAsync(s.d_meta, s.h_meta, (size_t)2 * e->max_batch * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        bool direct = false;
        if (planes) {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, planes) == cudaSuccess && attr.type == cudaMemoryTypeHost) direct = true;
            cudaGetLastError();
        }
        s.packed = false;
        if (direct) {
            const size_t width = (size_t)std::min<long long>(stride, SB_PLANE_FLOATS) * sizeof(float);
            SB_CUDA(cudaMemcpy2DAsync(s.d_in, (size_t)SB_PLANE_FLOATS * sizeof(float), planes, (size_t)stride * sizeof(float),
                                      width, n, cudaMemcpyHostToDevice, s.stream));
        } else if (e->pack_inputs) {
            // pageable caller memory: pack every position exactly into a 2.2 KB record in pinned memory (host_pack.cc)
            // instead of staging 62 KB of fp32; a position that is not two-valued per plane travels raw.
            auto pack_range = [&](int lo, int hi) {
                for (int i = lo; i < hi; ++i) {
                    const float* src = plane_ptrs ? plane_ptrs[i] : planes + (size_t)i * stride;
                    if (!sb_pack_position(src, sizes[i], offsets[i], &s.h_packed[i])) {
                        s.h_packed[i].board_size = sizes[i];
                        s.h_packed[i].offset = offsets[i];
                        s.h_packed[i].flags = SB_PACKED_RAW;
                        std::memcpy(s.h_in + (size_t)i * SB_PLANE_FLOATS, src,
                                    (size_t)SB_INPUT_CHANNELS * sizes[i] * sizes[i] * sizeof(float));
                    }
                }
            };
            const int n_thr = n >= 64 ? std::min(e->pack_threads, n / 32) : 1;
            if (n_thr > 1) {
                std::vector<std::thread> pool;
                for (int t = 1; t < n_thr; ++t) pool.emplace_back(pack_range, (int)((long long)n * t / n_thr), (int)((long long)n * (t + 1) / n_thr));
                pack_range(0, n / n_thr);
                for (auto& th : pool) th.join();
            } else {
                pack_range(0, n);
            }
            SB_CUDA(cudaMemcpyAsync(s.d_packed, s.h_packed, (size_t)n * sizeof(sb_packed_position), cudaMemcpyHostToDevice, s.stream));
            for (int i = 0; i < n; ++i) {
                if (s.h_packed[i].flags & SB_PACKED_RAW)
                    SB_CUDA(cudaMemcpyAsync(s.d_in + (size_t)i * SB_PLANE_FLOATS, s.h_in + (size_t)i * SB_PLANE_FLOATS,
                                            (size_t)SB_INPUT_CHANNELS * sizes[i] * sizes[i] * sizeof(float), cudaMemcpyHostToDevice, s.stream));
            }
            s.packed = true;
        } else {
            for (int i = 0; i < n; ++i) {
                const float* src = plane_ptrs ? plane_ptrs[i] : planes + (size_t)i * stride;
                std::memcpy(s.h_in + (size_t)i * SB_PLANE_FLOATS, src, (size_t)SB_INPUT_CHANNELS * sizes[i] * sizes[i] * sizeof(float));
            }
            SB_CUDA(cudaMemcpyAsync(s.d_in, s.h_in, (size_t)n * SB_PLANE_FLOATS * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        }
        s.n = n;
        EnqueueChainedForward(e, r, s, n);
        SB_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)n * kOutFloats * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        s.busy = true;
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

static int WaitImpl(sb_engine* e, int gpu, int slot, sb_output* out) {
    if (!e) return SB_ERR_INVALID;
    if (gpu < 0 || gpu >= (int)e->replicas.size()) return Fail(e, SB_ERR_INVALID, "gpu index out of range");
    if (slot < 0 || slot >= e->n_slots) return Fail(e, SB_ERR_INVALID, "slot index out of range");
    Replica& r = e->replicas[gpu];
    Slot& s = r.slots[slot];
    if (!s.busy) return Fail(e, SB_ERR_STATE, "slot is idle: nothing was submitted");
    try {
        SB_CUDA(cudaSetDevice(r.device));
        s.busy = false;
        CheckSlotError(s, cudaStreamSynchronize(s.stream), "forward");
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    if (out) {
        for (int i = 0; i < s.n; ++i) {
            const float* src = s.h_out + (size_t)i * kOutFloats;
            sb_output& o = out[i];
            std::memcpy(o.probabilities, src, sizeof(float) * SB_MAX_INTERSECTIONS);
            std::memcpy(o.ownership, src + SB_MAX_INTERSECTIONS, sizeof(float) * SB_MAX_INTERSECTIONS);
            const float* m = src + 2 * SB_MAX_INTERSECTIONS;
            o.pass_probability = m[0];
            o.wdl[0] = m[1];
            o.wdl[1] = m[2];
            o.wdl[2] = m[3];
            o.stm_winrate = m[4];
            o.final_score = m[5];
            o.q_error = m[6];
            o.score_error = m[7];
            o.board_size = s.sizes[i];
            o.offset = s.offsets[i];
            o.fp16 = e->precision == SB_PRECISION_FP16 ? 1 : 0;
        }
    }
    return SB_OK;
}

// ------------------------------------------------------------------------------------------------
// Batcher (sb_eval).  Replaces BatchForwardPipe::SendQueryAndWait / Worker / GatherBatches
// (/root/reference/src/neural/batch_forward_pipe.cc:7-193).
static void FutexWait(std::atomic<uint32_t>* a, uint32_t expected) {
    syscall(SYS_futex, reinterpret_cast<uint32_t*>(a), FUTEX_WAIT_PRIVATE, expected, nullptr, nullptr, 0);
}
static void FutexWakeAll(std::atomic<uint32_t>* a) {
    syscall(SYS_futex, reinterpret_cast<uint32_t*>(a), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
}

constexpr int kBatcherSlots = 2;   // worker threads (= device slots = streams) per GPU

// caller holds L.m: the FILLING batch stops accepting positions; the next free batch (if any) takes over
static int CloseFill(Lane& L) {
    const int idx = L.fill;
    HostBatch& hb = *L.ring[idx];
    hb.readers.store(hb.count, std::memory_order_relaxed);
    if (L.free_list.empty()) {
        L.fill = -1;
    } else {
        L.fill = L.free_list.back();
        L.free_list.pop_back();
    }
    return idx;
}

static void RunHostBatch(sb_engine* e, Replica& r, Slot& s, HostBatch& hb) {
    const int n = hb.count;
    try {
        cudaGetLastError();
        SB_CUDA(cudaMemcpyAsync(s.d_packed, hb.rec, (size_t)n * sizeof(sb_packed_position), cudaMemcpyHostToDevice, s.stream));
        if (hb.any_raw.load(std::memory_order_relaxed)) {
            const float* raw = hb.raw.load(std::memory_order_acquire);
            for (int i = 0; i < n; ++i) {
                if (hb.rec[i].flags & SB_PACKED_RAW) {
                    const int bs = hb.rec[i].board_size;
                    SB_CUDA(cudaMemcpyAsync(s.d_in + (size_t)i * SB_PLANE_FLOATS, raw + (size_t)i * SB_PLANE_FLOATS,
                                            (size_t)SB_INPUT_CHANNELS * bs * bs * sizeof(float), cudaMemcpyHostToDevice, s.stream));
                }
            }
        }
        for (int i = 0; i < n; ++i) {
            s.sizes[i] = hb.rec[i].board_size;
            s.offsets[i] = hb.rec[i].offset;
        }
        s.n = n;
        s.packed = true;
        EnqueueChainedForward(e, r, s, n);
        SB_CUDA(cudaMemcpyAsync(hb.out, s.d_out, (size_t)n * kOutFloats * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        CheckSlotError(s, cudaStreamSynchronize(s.stream), "batched forward");
        hb.rc = SB_OK;
    } catch (const CudaError& ce) {
        hb.rc = SB_ERR_CUDA;
        hb.err = ce.msg;
    }
}

static void PublishBatch(HostBatch& hb) {
    hb.done_seq.fetch_add(1, std::memory_order_release);
    FutexWakeAll(&hb.done_seq);
}

static void BatchWorker(sb_engine* e, Batcher* Bp, int gpu, int k) {
    Batcher& B = *Bp;
    Lane& L = *B.lanes[gpu];
    Replica& r = e->replicas[gpu];
    Slot& s = r.bslots[k];
    cudaSetDevice(r.device);
    for (;;) {
        int idx = -1;
        {
            std::unique_lock<std::mutex> lk(L.m);
            for (;;) {
                if (B.quit.load(std::memory_order_relaxed)) return;
                if (!L.closed.empty()) {
                    idx = L.closed.front();
                    L.closed.pop_front();
                    break;
                }
                if (L.fill >= 0 && L.ring[L.fill]->count > 0) {
                    HostBatch& hb = *L.ring[L.fill];
                    if (L.in_flight > 0) {
                        // this GPU is busy (forwards of its slots are chained, a second one could not start anyway): let
                        // the partial batch grow until it is full or the running batch has finished — with few search
                        // threads this is what keeps all of them in ONE batch instead of two half-size ones
                        L.cv_work.wait(lk);
                        continue;
                    }
                    // the timer runs from the batch's first position or, if later, from the moment this GPU published its
                    // previous batch: the callers it just released need a few tens of microseconds to come back, and
                    // closing before that splits a handful of search threads into two alternating half-size batches
                    const auto deadline = std::max(hb.first, L.last_finish) + std::chrono::microseconds(L.cur_wait_us);
                    const auto now = std::chrono::steady_clock::now();
                    if (now >= deadline) {
                        // closed by the timer.  If nothing joined during the second half of the wait, every caller that was
                        // going to come is already in (few search threads, or a CPU-bound front-end): waiting that long was
                        // futile, halve it (the reference drops its wait to zero in this case, batch_forward_pipe.cc:139-143)
                        if (now - hb.last > std::chrono::microseconds(L.cur_wait_us / 2)) L.cur_wait_us /= 2;
                        idx = CloseFill(L);
                        B.n_timer++;
                        break;
                    }
                    L.cv_work.wait_until(lk, deadline);
                } else {
                    // idle: restore the configured wait step by step (batch_forward_pipe.cc:132-138)
                    const int w = B.wait_us.load(std::memory_order_relaxed);
                    L.cur_wait_us = std::min(w, L.cur_wait_us + std::max(1, w / 8));
                    L.cv_work.wait(lk);
                }
            }
            L.in_flight++;
        }
        HostBatch& hb = *L.ring[idx];
        while (hb.ready.load(std::memory_order_acquire) < hb.count) std::this_thread::yield();   // packers still writing
        RunHostBatch(e, r, s, hb);
        B.n_batches++;
        B.n_positions += hb.count;
        PublishBatch(hb);
        {
            std::lock_guard<std::mutex> lk(L.m);
            L.in_flight--;
            L.last_finish = std::chrono::steady_clock::now();
        }
        L.cv_work.notify_all();   // a partial batch that was growing behind this one may be closed now
    }
}

// Measurement aid: with SAYURI_B200_STATS_FILE set in the environment a thread appends "seconds batches positions" every
// 250 ms, so that a front-end run that is cut short (tools/selfplay_bench.py --window) still yields its evaluation rate.
static void StatsWriter(Batcher* Bp, std::string path) {
    const auto t0 = std::chrono::steady_clock::now();
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return;
    while (!Bp->quit.load(std::memory_order_relaxed)) {
        std::this_thread::sleep_for(std::chrono::milliseconds(250));
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(f, "%.3f %lld %lld\n", t, (long long)Bp->n_batches.load(), (long long)Bp->n_positions.load());
        std::fflush(f);
    }
    std::fclose(f);
}

// Stops the workers and FAILS every position that was claimed but not yet evaluated (its caller returns SB_ERR_STATE):
// nobody stays blocked on a batch that will never run.  The Batcher object itself is retired, not deleted: a caller that
// loaded the pointer just before the stop finds `quit` set and leaves.  Its pinned buffers are released once no call
// is in flight (immediately in the normal case).
static void StopBatcher(sb_engine* e) {
    std::lock_guard<std::mutex> g(e->batcher_start_mutex);
    Batcher* Bp = e->batcher.load(std::memory_order_acquire);
    if (!Bp) return;
    Batcher& B = *Bp;
    B.quit.store(true, std::memory_order_release);
    for (auto& ln : B.lanes) {
        { std::lock_guard<std::mutex> lk(ln->m); }   // callers / workers inside the lock have seen `quit` after this
        ln->cv_work.notify_all();
        ln->cv_space.notify_all();
    }
    for (auto& t : B.workers) t.join();
    B.workers.clear();
    for (auto& ln : B.lanes) {
        std::vector<int> pending;
        {
            std::lock_guard<std::mutex> lk(ln->m);
            for (int idx : ln->closed) pending.push_back(idx);
            ln->closed.clear();
            if (ln->fill >= 0 && ln->ring[ln->fill]->count > 0) {
                ln->ring[ln->fill]->readers.store(ln->ring[ln->fill]->count, std::memory_order_relaxed);
                pending.push_back(ln->fill);
            }
            ln->fill = -1;
        }
        for (int idx : pending) {
            HostBatch& hb = *ln->ring[idx];
            hb.rc = SB_ERR_STATE;
            hb.err = "the batcher was stopped (sb_reconfigure / sb_reload_weights / sb_destroy) before this position was evaluated";
            PublishBatch(hb);
        }
    }
    for (int spin = 0; spin < 2000 && B.outstanding.load(std::memory_order_acquire) > 0; ++spin)
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    if (B.outstanding.load(std::memory_order_acquire) == 0) B.FreeBuffers();   // else: kept until sb_destroy (open tickets)
    for (Replica& r : e->replicas) {
        if (r.device >= 0) cudaSetDevice(r.device);
        for (Slot& s : r.bslots) FreeSlot(s);
        r.bslots.clear();
        r.chain_tail = nullptr;
    }
    e->batcher.store(nullptr, std::memory_order_release);
}

// Starts the worker threads on first use.  Throws CudaError.
static Batcher* StartBatcher(sb_engine* e) {
    std::lock_guard<std::mutex> g(e->batcher_start_mutex);
    if (Batcher* live = e->batcher.load(std::memory_order_acquire)) return live;
    std::unique_ptr<Batcher> B(new Batcher);
    B->max_batch = e->max_batch;
    B->batch_size = e->batcher_batch > 0 ? std::min(e->batcher_batch, e->max_batch) : e->max_batch;
    B->wait_us = e->batcher_wait_us;
    // per lane: one batch per worker in flight, one filling, one spare — plus enough further entries that a few thousand
    // blocked callers all find a slot: callers that find no FILLING batch sleep on one condition variable and are all
    // woken when a batch is recycled (measured: 4096 callers on a 4 x 256 ring ran 10x slower than 256 callers)
    const int n_lanes = (int)e->replicas.size();
    const int n_ring = kBatcherSlots + 2 + std::min(14, std::max(0, 4096 / std::max(1, e->max_batch * n_lanes)));
    for (int g2 = 0; g2 < n_lanes; ++g2) {
        std::unique_ptr<Lane> L(new Lane);
        L->gpu = g2;
        L->cur_wait_us = e->batcher_wait_us;
        L->last_finish = std::chrono::steady_clock::now();
        for (int i = 0; i < n_ring; ++i) {
            std::unique_ptr<HostBatch> hb(new HostBatch);
            SB_CUDA(cudaHostAlloc(&hb->rec, (size_t)e->max_batch * sizeof(sb_packed_position), cudaHostAllocPortable));
            SB_CUDA(cudaHostAlloc(&hb->out, (size_t)e->max_batch * kOutFloats * sizeof(float), cudaHostAllocPortable));
            L->ring.push_back(std::move(hb));
            if (i > 0) L->free_list.push_back(i);
        }
        L->fill = 0;
        B->lanes.push_back(std::move(L));
    }
    for (Replica& r : e->replicas) AllocSlotVec(e, r, r.bslots, kBatcherSlots);
    Batcher* raw = B.get();
    e->batchers.push_back(std::move(B));
    for (int g2 = 0; g2 < n_lanes; ++g2)
        for (int k = 0; k < kBatcherSlots; ++k) raw->workers.emplace_back(BatchWorker, e, raw, g2, k);
    if (const char* stats_path = std::getenv("SAYURI_B200_STATS_FILE")) raw->workers.emplace_back(StatsWriter, raw, std::string(stats_path));
    e->batcher.store(raw, std::memory_order_release);
    return raw;
}

// thread -> lane affinity: every calling thread draws one ticket for its lifetime
static std::atomic<unsigned> g_lane_ticket{0};
static unsigned ThreadTicket() {
    thread_local unsigned t = g_lane_ticket.fetch_add(1, std::memory_order_relaxed);
    return t;
}

constexpr int kTicketFailed = 1;   // sb_eval_ticket::flags: the position could not be staged, the call returns an error

// First half of an evaluation: claim an entry of the lane's FILLING batch, pack the position into it.
static int EvalBegin(sb_engine* e, const float* planes, int board_size, int offset, sb_eval_ticket* t) {
    if (!e || !planes || !t) return SB_ERR_INVALID;
    if (board_size < 2 || board_size > e->geom.N) return Fail(e, SB_ERR_INVALID, "board size of a sample exceeds the NN canvas");
    if (offset < 0 || offset > 4) return Fail(e, SB_ERR_INVALID, "policy offset must be in [0, 4]");
    if (!e->weights_ready) return Fail(e, SB_ERR_STATE, "weights have not been loaded");
    Batcher* Bp = e->batcher.load(std::memory_order_acquire);
    if (!Bp) {
        try {
            Bp = StartBatcher(e);
        } catch (const CudaError& ce) {
            return Fail(e, SB_ERR_CUDA, ce.msg);
        }
    }
    Batcher& B = *Bp;
    B.outstanding.fetch_add(1, std::memory_order_acq_rel);
    const int lane = (int)(ThreadTicket() % (unsigned)B.lanes.size());
    Lane& L = *B.lanes[lane];
    HostBatch* hb = nullptr;
    int idx = -1, i = -1;
    uint32_t seq0 = 0;
    {
        std::unique_lock<std::mutex> lk(L.m);
        while (L.fill < 0 && !B.quit.load(std::memory_order_relaxed)) L.cv_space.wait(lk);
        if (B.quit.load(std::memory_order_relaxed)) {
            lk.unlock();
            B.outstanding.fetch_sub(1, std::memory_order_acq_rel);
            std::lock_guard<std::mutex> el(e->error_mutex);
            return Fail(e, SB_ERR_STATE, "the batcher is shutting down");
        }
        idx = L.fill;
        hb = L.ring[idx].get();
        i = hb->count++;
        seq0 = hb->done_seq.load(std::memory_order_relaxed);
        hb->last = std::chrono::steady_clock::now();
        if (i == 0) hb->first = hb->last;
        if (hb->count >= B.batch_size.load(std::memory_order_relaxed)) {
            // filled before its timer: traffic is high, a longer wait costs nothing and keeps batches full
            L.cur_wait_us = std::min(B.wait_us.load(std::memory_order_relaxed), L.cur_wait_us * 2 + 10);
            L.closed.push_back(CloseFill(L));
            B.n_full++;
            L.cv_work.notify_one();
        } else if (i == 0) {
            L.cv_work.notify_one();   // somebody has to watch this batch's timer
        }
    }
    // pack outside the lock, straight into the pinned batch record
    int flags = 0;
    if (!sb_pack_position(planes, board_size, offset, &hb->rec[i])) {
        float* raw = hb->raw.load(std::memory_order_acquire);
        if (!raw) {
            std::lock_guard<std::mutex> lk(L.m);
            raw = hb->raw.load(std::memory_order_relaxed);
            if (!raw) {
                if (cudaHostAlloc(&raw, (size_t)B.max_batch * SB_PLANE_FLOATS * sizeof(float), cudaHostAllocPortable) == cudaSuccess) {
                    hb->raw.store(raw, std::memory_order_release);
                } else {
                    cudaGetLastError();
                    raw = nullptr;
                }
            }
        }
        if (raw) {
            hb->rec[i].board_size = board_size;
            hb->rec[i].offset = offset;
            hb->rec[i].flags = SB_PACKED_RAW;
            std::memcpy(raw + (size_t)i * SB_PLANE_FLOATS, planes, (size_t)SB_INPUT_CHANNELS * board_size * board_size * sizeof(float));
            hb->any_raw.store(1, std::memory_order_relaxed);
        } else {
            // no staging memory for the fp32 planes: the entry runs as an empty position and this call FAILS
            std::memset(&hb->rec[i], 0, sizeof(sb_packed_position));
            hb->rec[i].board_size = board_size;
            hb->rec[i].offset = offset;
            flags |= kTicketFailed;
        }
        B.n_raw++;
    }
    hb->ready.fetch_add(1, std::memory_order_release);
    t->owner = Bp;
    t->batch = hb;
    t->seq = seq0;
    t->index = i;
    t->lane = lane | (idx << 8);
    t->flags = flags;
    t->board_size = board_size;
    t->offset = offset;
    return SB_OK;
}

// Second half: wait for (or poll) the batch, copy the caller's result out, recycle the batch behind the last reader.
// Returns 1 while the result is pending (block == false only).
static int EvalFinish(sb_engine* e, sb_eval_ticket* t, sb_output* out, bool block) {
    if (!e || !t || !t->owner || !t->batch) return SB_ERR_INVALID;
    Batcher& B = *static_cast<Batcher*>(t->owner);
    HostBatch* hb = static_cast<HostBatch*>(t->batch);
    const uint32_t seq0 = t->seq;
    if (!block && hb->done_seq.load(std::memory_order_acquire) == seq0) return 1;
    while (hb->done_seq.load(std::memory_order_acquire) == seq0) FutexWait(&hb->done_seq, seq0);
    int rc = hb->rc;
    const int i = t->index;
    if (rc == SB_OK && (t->flags & kTicketFailed)) {
        rc = SB_ERR_CUDA;
        std::lock_guard<std::mutex> lk(e->error_mutex);
        e->last_error = "no pinned staging memory for a position that cannot be packed (cudaHostAlloc failed)";
    } else if (rc == SB_OK) {
        if (out) {
            const float* src = hb->out + (size_t)i * kOutFloats;
            std::memcpy(out->probabilities, src, sizeof(float) * SB_MAX_INTERSECTIONS);
            std::memcpy(out->ownership, src + SB_MAX_INTERSECTIONS, sizeof(float) * SB_MAX_INTERSECTIONS);
            const float* m = src + 2 * SB_MAX_INTERSECTIONS;
            out->pass_probability = m[0];
            out->wdl[0] = m[1];
            out->wdl[1] = m[2];
            out->wdl[2] = m[3];
            out->stm_winrate = m[4];
            out->final_score = m[5];
            out->q_error = m[6];
            out->score_error = m[7];
            out->board_size = t->board_size;
            out->offset = t->offset;
            out->fp16 = e->precision == SB_PRECISION_FP16 ? 1 : 0;
        }
    } else {
        std::lock_guard<std::mutex> lk(e->error_mutex);
        e->last_error = hb->err;
    }
    if (hb->readers.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        // last reader recycles the batch
        Lane& L = *B.lanes[t->lane & 0xff];
        const int idx = t->lane >> 8;
        hb->count = 0;
        hb->ready.store(0, std::memory_order_relaxed);
        hb->any_raw.store(0, std::memory_order_relaxed);
        std::lock_guard<std::mutex> lk(L.m);
        if (!B.quit.load(std::memory_order_relaxed)) {
            if (L.fill < 0) {
                L.fill = idx;
                L.cv_space.notify_all();
            } else {
                L.free_list.push_back(idx);
            }
        }
    }
    t->owner = nullptr;
    t->batch = nullptr;
    B.outstanding.fetch_sub(1, std::memory_order_acq_rel);
    return rc;
}

static int EvalImpl(sb_engine* e, const float* planes, int board_size, int offset, sb_output* out) {
    if (!out) return SB_ERR_INVALID;
    sb_eval_ticket t;
    const int rc = EvalBegin(e, planes, board_size, offset, &t);
    if (rc) return rc;
    return EvalFinish(e, &t, out, true);
}

}  // namespace sb

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int sb_create(sb_engine** out, const sb_net_desc* desc, const sb_weights* w, const int* gpu_ids, int n_gpus,
              int board_size, int max_batch, int precision) {
    if (!desc) return Fail(nullptr, SB_ERR_INVALID, "null net description");
    if (!w) {
        if (desc->blocks < 0 || (desc->blocks > 0 && !desc->se_sizes)) return Fail(nullptr, SB_ERR_INVALID, "bad block count / se_sizes");
        return CreateImpl(out, nullptr, desc, gpu_ids, n_gpus, board_size, max_batch, precision);
    }
    HostNet net;
    std::string err;
    if (!NetFromAbi(desc, w, net, err)) return Fail(nullptr, SB_ERR_INVALID, err);
    return CreateImpl(out, &net, nullptr, gpu_ids, n_gpus, board_size, max_batch, precision);
}

int sb_create_from_file(sb_engine** out, const char* weights_path, const int* gpu_ids, int n_gpus, int board_size,
                        int max_batch, int precision) {
    if (!weights_path) return Fail(nullptr, SB_ERR_INVALID, "null weights path");
    HostNet net;
    std::string err;
    if (!LoadWeightsFile(weights_path, net, err)) return Fail(nullptr, SB_ERR_IO, err);
    return CreateImpl(out, &net, nullptr, gpu_ids, n_gpus, board_size, max_batch, precision);
}

int sb_reconfigure(sb_engine* e, int board_size, int max_batch) {
    if (!e) return SB_ERR_INVALID;
    // CudaForwardPipe::Construct semantics (cuda_forward_pipe.cc:56-77): non-positive values keep the
    // current setting; no rebuild if the board is unchanged and the batch fits.
    const int board = board_size > 0 ? board_size : e->geom.N;
    const int batch = max_batch > 0 ? max_batch : e->max_batch;
    if (board < 2 || board > SB_MAX_BOARD_SIZE) return Fail(e, SB_ERR_INVALID, "NN board size should be in [2, 19]");
    if (board == e->geom.N && batch <= e->max_batch) return SB_OK;
    for (Replica& r : e->replicas)
        for (Slot& s : r.slots)
            if (s.busy) return Fail(e, SB_ERR_STATE, "cannot reconfigure while a batch is in flight");
    StopBatcher(e);   // restarted by the next sb_eval with the new geometry (no Forward may be in flight, as in the reference)
    try {
        Configure(e, board, std::max(batch, board == e->geom.N ? e->max_batch : batch));
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

static int ReloadImpl(sb_engine* e, HostNet& net) {
    StopBatcher(e);
    if (net.blocks != e->net_shape.blocks || net.channels != e->net_shape.channels || net.P != e->net_shape.P ||
        net.V != e->net_shape.V || net.se_sizes() != e->se_sizes || net.block_types() != e->block_types ||
        net.inner_channels() != e->inner_channels || net.dw_kernels() != e->dw_kernels ||
        (net.replk ? net.p_dw_conv.k : 0) != e->replk_kernel)
        return Fail(e, SB_ERR_INVALID, "reload requires the same architecture; destroy and create for a new one");
    try {
        e->net_shape.act = net.act;
        e->net_shape.version = net.version;
        std::vector<uint8_t> blob = PackBlob(net, e->layout);
        for (Replica& r : e->replicas) {
            SB_CUDA(cudaSetDevice(r.device));
            SB_CUDA(cudaDeviceSynchronize());
            BuildReplica(e, r, &blob);
        }
        DistributeBlob(e, &blob);
        e->weights_ready = true;
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

int sb_reload_weights(sb_engine* e, const sb_net_desc* desc, const sb_weights* w) {
    if (!e) return SB_ERR_INVALID;
    HostNet net;
    std::string err;
    if (!NetFromAbi(desc, w, net, err)) return Fail(e, SB_ERR_INVALID, err);
    return ReloadImpl(e, net);
}

int sb_reload_weights_from_file(sb_engine* e, const char* weights_path) {
    if (!e || !weights_path) return SB_ERR_INVALID;
    HostNet net;
    std::string err;
    if (!LoadWeightsFile(weights_path, net, err)) return Fail(e, SB_ERR_IO, err);
    return ReloadImpl(e, net);
}

void sb_destroy(sb_engine* e) {
    if (!e) return;
    StopBatcher(e);
    for (Replica& r : e->replicas) {
        if (r.device >= 0) {
            cudaSetDevice(r.device);
            cudaDeviceSynchronize();
        }
        DestroyReplica(r);
    }
    DestroyComms(e);
    for (auto& b : e->batchers) b->FreeBuffers();
    delete e;
}

const char* sb_last_error(const sb_engine* e) { return e ? e->last_error.c_str() : g_create_error.c_str(); }
int sb_num_gpus(const sb_engine* e) { return e ? (int)e->replicas.size() : 0; }
int sb_num_slots(const sb_engine* e) { return e ? e->n_slots : 0; }
int sb_max_batch(const sb_engine* e) { return e ? e->max_batch : 0; }
int sb_board_size(const sb_engine* e) { return e ? e->geom.N : 0; }

int sb_get_net_desc(const sb_engine* e, sb_net_desc* d, int* se_sizes, int se_capacity) {
    if (!e || !d) return SB_ERR_INVALID;
    d->version = e->net_shape.version;
    d->input_channels = e->net_shape.input_channels;
    d->blocks = e->net_shape.blocks;
    d->channels = e->net_shape.channels;
    d->policy_channels = e->net_shape.P;
    d->value_channels = e->net_shape.V;
    d->activation = e->net_shape.act;
    d->se_sizes = se_sizes;
    d->block_types = nullptr;
    d->inner_channels = nullptr;
    d->dw_kernels = nullptr;
    d->policy_head_type = e->replk_kernel > 0 ? SB_POLICY_HEAD_REPLK : SB_POLICY_HEAD_NORMAL;
    d->policy_dw_kernel = e->replk_kernel;
    if (se_sizes) {
        if (se_capacity < e->net_shape.blocks) return SB_ERR_INVALID;
        for (int b = 0; b < e->net_shape.blocks; ++b) se_sizes[b] = e->se_sizes[b];
    }
    return SB_OK;
}

int sb_get_dw_desc(const sb_engine* e, int* dw_kernels, int capacity, int* policy_head_type, int* policy_dw_kernel) {
    if (!e || (dw_kernels && capacity < e->net_shape.blocks)) return SB_ERR_INVALID;
    if (dw_kernels)
        for (int b = 0; b < e->net_shape.blocks; ++b) dw_kernels[b] = e->dw_kernels[b];
    if (policy_head_type) *policy_head_type = e->replk_kernel > 0 ? SB_POLICY_HEAD_REPLK : SB_POLICY_HEAD_NORMAL;
    if (policy_dw_kernel) *policy_dw_kernel = e->replk_kernel;
    return SB_OK;
}

int sb_get_block_desc(const sb_engine* e, int* block_types, int* inner_channels, int capacity) {
    if (!e || capacity < e->net_shape.blocks) return SB_ERR_INVALID;
    for (int b = 0; b < e->net_shape.blocks; ++b) {
        if (block_types) block_types[b] = e->block_types[b];
        if (inner_channels) inner_channels[b] = e->inner_channels[b];
    }
    return SB_OK;
}

int sb_forward_batch(sb_engine* e, int gpu, int n, const float* const* planes, const int* board_sizes,
                     const int* policy_offsets, sb_output* out) {
    int rc = SubmitImpl(e, gpu, 0, n, nullptr, 0, planes, board_sizes, policy_offsets);
    if (rc) return rc;
    return WaitImpl(e, gpu, 0, out);
}

int sb_submit(sb_engine* e, int gpu, int slot, int n, const float* planes, long long plane_stride,
              const int* board_sizes, const int* policy_offsets) {
    if (plane_stride <= 0) return Fail(e, SB_ERR_INVALID, "plane_stride must be positive");
    return SubmitImpl(e, gpu, slot, n, planes, plane_stride, nullptr, board_sizes, policy_offsets);
}

int sb_wait(sb_engine* e, int gpu, int slot, sb_output* out) { return WaitImpl(e, gpu, slot, out); }

int sb_eval(sb_engine* e, const float* planes, int board_size, int policy_offset, sb_output* out) {
    return EvalImpl(e, planes, board_size, policy_offset, out);
}

int sb_eval_submit(sb_engine* e, const float* planes, int board_size, int policy_offset, sb_eval_ticket* ticket) {
    return EvalBegin(e, planes, board_size, policy_offset, ticket);
}
int sb_eval_poll(sb_engine* e, sb_eval_ticket* ticket, sb_output* out) { return EvalFinish(e, ticket, out, false); }
int sb_eval_wait(sb_engine* e, sb_eval_ticket* ticket, sb_output* out) { return EvalFinish(e, ticket, out, true); }

// Network::GetOutput(state, kAverage), /root/reference/src/neural/network.cc:258-282: the reference evaluates the 8 symmetric
// views of a position with 8 SERIAL Forward calls, post-processes each (TransformResult :361-411, ActivatePolicy :413-428) and
// averages.  Here the 8 views are built from the identity view (Encoder::SymmetryPlanes, encoder.cc:80-98: view[i] =
// planes[T(i)]), submitted as 8 tickets from the calling thread — they land in ONE batch — and post-processed with the
// same formulas in the same order.
static inline int SymmIndex(int n, int symm, int idx) {   // Symmetry::GetSymmetry, /root/reference/src/game/symmetry.cc:97-123
    int x = idx % n, y = idx / n;
    if (symm & 4) std::swap(x, y);
    if (symm & 2) x = n - 1 - x;
    if (symm & 1) y = n - 1 - y;
    return y * n + x;
}

int sb_eval_symm8(sb_engine* e, const float* planes, int board_size, int policy_offset, float temperature, sb_symm8_result* out) {
    if (!e || !planes || !out) return SB_ERR_INVALID;
    if (board_size < 2 || board_size > SB_MAX_BOARD_SIZE) return Fail(e, SB_ERR_INVALID, "board size out of range");
    if (!(temperature > 0.f)) return Fail(e, SB_ERR_INVALID, "temperature must be positive");
    const int ns = board_size * board_size;
    std::vector<float> views((size_t)8 * SB_INPUT_CHANNELS * ns);
    std::vector<int> table((size_t)8 * ns);
    for (int s = 0; s < 8; ++s) {
        int* t = table.data() + (size_t)s * ns;
        for (int i = 0; i < ns; ++i) t[i] = SymmIndex(board_size, s, i);
        float* v = views.data() + (size_t)s * SB_INPUT_CHANNELS * ns;
        for (int c = 0; c < SB_INPUT_CHANNELS; ++c)
            for (int i = 0; i < ns; ++i) v[(size_t)c * ns + i] = planes[(size_t)c * ns + t[i]];
    }
    sb_eval_ticket tk[8];
    int n_sub = 0, rc = SB_OK;
    for (; n_sub < 8; ++n_sub) {
        rc = EvalBegin(e, views.data() + (size_t)n_sub * SB_INPUT_CHANNELS * ns, board_size, policy_offset, &tk[n_sub]);
        if (rc) break;
    }
    std::unique_ptr<sb_output[]> raw(new sb_output[8]);
    for (int s = 0; s < n_sub; ++s) {
        const int r2 = EvalFinish(e, &tk[s], &raw[s], true);
        if (r2 && !rc) rc = r2;
    }
    if (rc) return rc;

    std::memset(out, 0, sizeof(*out));
    out->board_size = board_size;
    auto softplus_sq = [](float x) {
        if (x <= 20.f) x = std::log(1.f + std::exp(x));
        return (x * x) / 4.f;
    };
    std::vector<float> logits(ns + 1), prob(ns + 1);
    for (int s = 0; s < 8; ++s) {
        const sb_output& r = raw[s];
        const int* t = table.data() + (size_t)s * ns;
        // TransformResult: inverse symmetry (result[T(i)] = view[i]), tanh on the ownership
        for (int i = 0; i < ns; ++i) logits[t[i]] = r.probabilities[i];
        logits[ns] = r.pass_probability;
        // ActivatePolicy: softmax over the board and the pass move with temperature (utils/logits.h:22-39, double accumulator)
        const float alpha = *std::max_element(logits.begin(), logits.end());
        double denom = 0.0;
        for (int i = 0; i <= ns; ++i) {
            const double val = std::exp((double)(logits[i] - alpha) / (double)temperature);
            denom += val;
            prob[i] = (float)val;
        }
        for (int i = 0; i <= ns; ++i) prob[i] = (float)((double)prob[i] / denom);
        for (int i = 0; i < ns; ++i) out->probabilities[i] += prob[i] / 8;
        out->pass_probability += prob[ns] / 8;
        for (int i = 0; i < ns; ++i) out->ownership[t[i]] += std::tanh(r.ownership[i]) / 8;
        float w[3];
        {
            const float a3 = std::max(r.wdl[0], std::max(r.wdl[1], r.wdl[2]));
            double d3 = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double val = std::exp((double)(r.wdl[k] - a3));
                d3 += val;
                w[k] = (float)val;
            }
            for (int k = 0; k < 3; ++k) w[k] = (float)((double)w[k] / d3);
        }
        for (int k = 0; k < 3; ++k) out->wdl[k] += w[k] / 8;
        out->wdl_winrate += ((w[0] - w[2] + 1.f) / 2) / 8;
        out->stm_winrate += ((std::tanh(r.stm_winrate) + 1.f) / 2) / 8;
        out->final_score += (20 * r.final_score) / 8;
        out->q_error += (float)(0.25 * softplus_sq(r.q_error)) / 8;
        out->score_error += (150 * softplus_sq(r.score_error)) / 8;
    }
    return SB_OK;
}

int sb_batcher_config(sb_engine* e, int batch_size, int wait_us) {
    if (!e) return SB_ERR_INVALID;
    if (batch_size > e->max_batch) return Fail(e, SB_ERR_INVALID, "batch size exceeds max_batch");
    if (batch_size > 0) e->batcher_batch = batch_size;
    if (wait_us >= 0) e->batcher_wait_us = wait_us;
    std::lock_guard<std::mutex> g(e->batcher_start_mutex);
    if (Batcher* B = e->batcher.load(std::memory_order_acquire)) {
        if (batch_size > 0) B->batch_size.store(batch_size, std::memory_order_relaxed);
        if (wait_us >= 0) {
            B->wait_us.store(wait_us, std::memory_order_relaxed);
            for (auto& ln : B->lanes) {
                std::lock_guard<std::mutex> lk(ln->m);
                ln->cur_wait_us = wait_us;
            }
        }
        // a FILLING batch that is already over the new size is closed by the next arrival or its timer
    }
    return SB_OK;
}

int sb_batcher_stats(sb_engine* e, long long* out6) {
    if (!e || !out6) return SB_ERR_INVALID;
    std::lock_guard<std::mutex> g(e->batcher_start_mutex);
    for (int i = 0; i < 6; ++i) out6[i] = 0;
    if (Batcher* B = e->batcher.load(std::memory_order_acquire)) {
        out6[0] = B->n_batches;
        out6[1] = B->n_positions;
        out6[2] = B->n_full;
        out6[3] = B->n_timer;
        out6[4] = B->n_raw;
        out6[5] = (long long)B->workers.size();
    }
    return SB_OK;
}

// depth <= 0: `threads` blocking callers (one position per thread in flight, the front-end's model);
// depth  > 0: `threads` feeder threads, each keeping `depth` tickets in flight through sb_eval_submit / sb_eval_wait
static double EvalThroughput(sb_engine* e, const float* planes, int n_pos, int board_size, int threads, int depth, double seconds) {
    if (!e || !planes || n_pos < 1 || threads < 1 || seconds <= 0) return (double)SB_ERR_INVALID;
    sb_output warm;
    int rc = EvalImpl(e, planes, board_size, 0, &warm);   // starts the workers, surfaces configuration errors
    if (rc) return (double)rc;
    std::atomic<long long> total{0};
    std::atomic<int> failed{0};
    std::atomic<bool> stop{false};
    const size_t rec = (size_t)SB_PLANE_FLOATS;
    std::vector<std::thread> pool;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            std::unique_ptr<sb_output> o(new sb_output);
            long long n = 0;
            int i = t % n_pos;
            if (depth <= 0) {
                while (!stop.load(std::memory_order_relaxed)) {
                    if (EvalImpl(e, planes + (size_t)i * rec, board_size, i % 5, o.get())) {
                        failed.store(1);
                        break;
                    }
                    i = (i + 1) % n_pos;
                    ++n;
                }
            } else {
                std::vector<sb_eval_ticket> tk(depth);
                int head = 0, live = 0;
                while (!stop.load(std::memory_order_relaxed) && !failed.load(std::memory_order_relaxed)) {
                    while (live < depth) {
                        if (EvalBegin(e, planes + (size_t)i * rec, board_size, i % 5, &tk[(head + live) % depth])) {
                            failed.store(1);
                            break;
                        }
                        i = (i + 1) % n_pos;
                        ++live;
                    }
                    if (failed.load()) break;
                    if (EvalFinish(e, &tk[head], o.get(), true)) failed.store(1);
                    head = (head + 1) % depth;
                    --live;
                    ++n;
                }
                while (live > 0) {   // drain
                    EvalFinish(e, &tk[head], o.get(), true);
                    head = (head + 1) % depth;
                    --live;
                }
            }
            total += n;
        });
    }
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    stop.store(true);
    for (auto& th : pool) th.join();
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (failed.load()) return (double)SB_ERR_CUDA;
    return (double)total.load() / el;
}

double sb_eval_throughput(sb_engine* e, const float* planes, int n_pos, int board_size, int threads, double seconds) {
    return EvalThroughput(e, planes, n_pos, board_size, threads, 0, seconds);
}
double sb_eval_throughput_async(sb_engine* e, const float* planes, int n_pos, int board_size, int threads, int depth,
                                double seconds) {
    if (depth < 1) return (double)SB_ERR_INVALID;
    return EvalThroughput(e, planes, n_pos, board_size, threads, depth, seconds);
}

struct sb_host_net {
    HostNet net;
    std::vector<const std::vector<float>*> tensors;   // sb_weights order
};

int sb_host_net_load(sb_host_net** out, const char* weights_path) {
    if (!out || !weights_path) return SB_ERR_INVALID;
    *out = nullptr;
    std::unique_ptr<sb_host_net> h(new sb_host_net);
    std::string err;
    if (!LoadWeightsFile(weights_path, h->net, err)) return Fail(nullptr, SB_ERR_IO, err);
    auto conv = [&](const HostConv& c) {
        h->tensors.push_back(&c.w);
        h->tensors.push_back(&c.b);
    };
    auto fc = [&](const HostFC& f) {
        h->tensors.push_back(&f.w);
        h->tensors.push_back(&f.b);
    };
    const HostNet& n = h->net;
    conv(n.input_conv);
    for (const HostBlock& b : n.tower) {
        for (const HostConv& c : b.convs) conv(c);
        if (b.se_size > 0) {
            fc(b.squeeze);
            fc(b.excite);
        }
    }
    conv(n.p_hd_conv);
    if (n.replk) {
        conv(n.p_dw_conv);
        conv(n.p_pt_conv);
    }
    fc(n.p_inter_fc);
    conv(n.prob_conv);
    fc(n.pass_fc);
    conv(n.v_hd_conv);
    fc(n.v_inter_fc);
    conv(n.v_ownership);
    fc(n.v_misc);
    *out = h.release();
    return SB_OK;
}

void sb_host_net_free(sb_host_net* h) { delete h; }

int sb_host_net_desc(const sb_host_net* h, sb_net_desc* d, int* se_sizes, int* block_types, int* inner_channels, int* dw_kernels,
                     int capacity) {
    if (!h || !d) return SB_ERR_INVALID;
    const HostNet& n = h->net;
    if ((se_sizes || block_types || inner_channels || dw_kernels) && capacity < n.blocks) return SB_ERR_INVALID;
    d->version = n.version;
    d->input_channels = n.input_channels;
    d->blocks = n.blocks;
    d->channels = n.channels;
    d->policy_channels = n.P;
    d->value_channels = n.V;
    d->activation = n.act;
    d->se_sizes = se_sizes;
    d->block_types = block_types;
    d->inner_channels = inner_channels;
    d->dw_kernels = dw_kernels;
    d->policy_head_type = n.replk ? SB_POLICY_HEAD_REPLK : SB_POLICY_HEAD_NORMAL;
    d->policy_dw_kernel = n.replk ? n.p_dw_conv.k : 0;
    for (int b = 0; b < n.blocks; ++b) {
        if (se_sizes) se_sizes[b] = n.tower[b].se_size;
        if (block_types) block_types[b] = n.tower[b].type;
        if (inner_channels) inner_channels[b] = n.tower[b].inner;
        if (dw_kernels) dw_kernels[b] = n.tower[b].dw_kernel();
    }
    return SB_OK;
}

long long sb_host_net_tensor(const sb_host_net* h, int idx, const float** data) {
    if (!h || !data || idx < 0 || idx >= (int)h->tensors.size()) return -1;
    *data = h->tensors[idx]->data();
    return (long long)h->tensors[idx]->size();
}

void* sb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void sb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int sb_weights_blob(sb_engine* e, int gpu, void** device_ptr, size_t* bytes) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size() || !device_ptr || !bytes) return SB_ERR_INVALID;
    *device_ptr = e->replicas[gpu].blob;
    *bytes = e->layout.bytes;
    return SB_OK;
}

static int BlobCopy(sb_engine* e, int gpu, void* dst, const void* src, size_t bytes, bool import) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size() || !dst || !src) return SB_ERR_INVALID;
    if (bytes != e->layout.bytes) return Fail(e, SB_ERR_INVALID, "blob size mismatch: expected " + std::to_string(e->layout.bytes));
    try {
        SB_CUDA(cudaSetDevice(e->replicas[gpu].device));
        SB_CUDA(cudaDeviceSynchronize());
        SB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
        SB_CUDA(cudaDeviceSynchronize());
        if (import) e->weights_ready = true;
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

int sb_weights_export(sb_engine* e, int gpu, void* device_dst, size_t bytes) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size()) return SB_ERR_INVALID;
    return BlobCopy(e, gpu, device_dst, e->replicas[gpu].blob, bytes, false);
}

int sb_weights_import(sb_engine* e, int gpu, const void* device_src, size_t bytes) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size()) return SB_ERR_INVALID;
    return BlobCopy(e, gpu, e->replicas[gpu].blob, device_src, bytes, true);
}

int sb_weights_broadcast(sb_engine* e) {
    if (!e) return SB_ERR_INVALID;
    try {
        DistributeBlob(e, nullptr);
        e->weights_ready = true;
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

int sb_weights_stats(sb_engine* e, long long* out6) {
    if (!e || !out6) return SB_ERR_INVALID;
    out6[0] = e->stat_h2d_uploads;
    out6[1] = e->stat_d2d_fills;
    out6[2] = e->bcast_method;
    out6[3] = e->nccl_version;
    out6[4] = e->verify_ok;
    out6[5] = (long long)(e->bcast_ms * 1000.0);
    return SB_OK;
}

uint64_t sb_weights_checksum(sb_engine* e, int gpu) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size()) return 0;
    std::vector<uint8_t> host(e->layout.bytes);
    cudaSetDevice(e->replicas[gpu].device);
    if (cudaMemcpy(host.data(), e->replicas[gpu].blob, host.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : host) {
        h ^= b;
        h *= 1099511628211ull;
    }
    return h;
}

int sb_time_forward(sb_engine* e, int gpu, int slot, int iters, int flush_l2, float* ms_each, float* conv_ms,
                    int* conv_launches) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size() || slot < 0 || slot >= e->n_slots || iters < 0) return SB_ERR_INVALID;
    Replica& r = e->replicas[gpu];
    Slot& s = r.slots[slot];
    if (s.busy) return Fail(e, SB_ERR_STATE, "slot is busy");
    if (s.n < 1) return Fail(e, SB_ERR_STATE, "no batch has been uploaded to this slot yet");
    try {
        SB_CUDA(cudaSetDevice(r.device));
        if (flush_l2 && !r.flush_buf) {
            r.flush_bytes = (size_t)256 << 20;   // > 126 MB L2
            SB_CUDA(cudaMalloc(&r.flush_buf, r.flush_bytes));
        }
        for (int i = 0; i < iters; ++i) {
            if (flush_l2) SB_CUDA(cudaMemsetAsync(r.flush_buf, i & 0xff, r.flush_bytes, s.stream));
            SB_CUDA(cudaEventRecord(s.ev_a, s.stream));
            EnqueueForward(e, r, s, s.n);
            SB_CUDA(cudaEventRecord(s.ev_b, s.stream));
            CheckSlotError(s, cudaEventSynchronize(s.ev_b), "timed forward");
            float ms = 0.f;
            SB_CUDA(cudaEventElapsedTime(&ms, s.ev_a, s.ev_b));
            if (ms_each) ms_each[i] = ms;
        }
        if (conv_ms || conv_launches) {
            // conv kernel time = forward time - time of the bracketed non-convolution kernel groups, median of 5 passes
            std::vector<float> samples, totals;
            int launches = 0;
            for (int pass = 0; pass < 5; ++pass) {
                ConvTimer tm;
                tm.on = true;
                if (flush_l2) SB_CUDA(cudaMemsetAsync(r.flush_buf, pass & 0xff, r.flush_bytes, s.stream));
                SB_CUDA(cudaEventRecord(s.ev_a, s.stream));
                EnqueueForward(e, r, s, s.n, &tm);
                SB_CUDA(cudaEventRecord(s.ev_b, s.stream));
                CheckSlotError(s, cudaStreamSynchronize(s.stream), "profiled forward");
                float total = 0.f, others = 0.f;
                SB_CUDA(cudaEventElapsedTime(&total, s.ev_a, s.ev_b));
                const bool verbose = pass == 4 && std::getenv("SB_PROFILE_GROUPS") != nullptr;
                for (size_t i = 0; i + 1 < tm.ev.size(); i += 2) {
                    float ms = 0.f;
                    SB_CUDA(cudaEventElapsedTime(&ms, tm.ev[i], tm.ev[i + 1]));
                    others += ms;
                    if (verbose) std::fprintf(stderr, "sb_time_forward: non-conv group %zu: %.1f us\n", i / 2, ms * 1e3f);
                }
                if (verbose) std::fprintf(stderr, "sb_time_forward: forward %.1f us, other kernels %.1f us\n", total * 1e3f, others * 1e3f);
                samples.push_back(total - others);
                totals.push_back(total);
                launches = tm.conv_launches;
                for (cudaEvent_t ev : tm.ev) cudaEventDestroy(ev);
            }
            // report the pass with the median conv time, and (iters == 0) that pass's whole-forward time in ms_each[0] so
            // that the caller can apply the conv SHARE of this pass to its own timed steps
            std::vector<size_t> order(samples.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return samples[a] < samples[b]; });
            const size_t mid = order[order.size() / 2];
            if (iters == 0 && ms_each) ms_each[0] = totals[mid];
            std::sort(samples.begin(), samples.end());
            if (conv_ms) *conv_ms = samples[samples.size() / 2];
            if (conv_launches) *conv_launches = launches;
        }
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

long long sb_launch_count(const sb_engine* e) { return e ? e->launches.load() : 0; }

int sb_debug_read_trunk(sb_engine* e, int gpu, int slot, int sample, float* out) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size() || slot < 0 || slot >= e->n_slots || !out) return SB_ERR_INVALID;
    Replica& r = e->replicas[gpu];
    Slot& s = r.slots[slot];
    if (!s.trunk || sample < 0 || sample >= s.n) return Fail(e, SB_ERR_STATE, "no forward has run on this slot / bad sample");
    try {
        SB_CUDA(cudaSetDevice(r.device));
        SB_CUDA(cudaStreamSynchronize(s.stream));
        const Geom g = e->geom;
        const int C = e->net_shape.channels, R = s.trunk->rows, bs = s.sizes[sample];
        const size_t total = (size_t)R * s.trunk->channels;
        std::vector<__half> hi(total), lo(total);
        SB_CUDA(cudaMemcpy(hi.data(), s.trunk->hi, total * sizeof(__half), cudaMemcpyDeviceToHost));
        SB_CUDA(cudaMemcpy(lo.data(), s.trunk->lo, total * sizeof(__half), cudaMemcpyDeviceToHost));
        const bool split = Split(e);
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < bs; ++y)
                for (int x = 0; x < bs; ++x) {
                    const size_t i = act_index(g.row(sample, y, x), c, R);
                    out[(size_t)c * bs * bs + y * bs + x] = __half2float(hi[i]) + (split ? __half2float(lo[i]) : 0.f);
                }
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return SB_OK;
}

int sb_conv_stats(sb_engine* e, int gpu, int slot, long long* out, int capacity) {
    if (!e || gpu < 0 || gpu >= (int)e->replicas.size() || slot < 0 || slot >= e->n_slots || !out) return SB_ERR_INVALID;
    Replica& r = e->replicas[gpu];
    Slot& s = r.slots[slot];
    const int n = std::min(capacity, r.sm_count * 8);
    try {
        SB_CUDA(cudaSetDevice(r.device));
        SB_CUDA(cudaStreamSynchronize(s.stream));
        SB_CUDA(cudaMemcpy(out, s.d_stats, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
    } catch (const CudaError& ce) {
        return Fail(e, SB_ERR_CUDA, ce.msg);
    }
    return n;
}

int sb_set_option(sb_engine* e, const char* key, int value) {
    if (!e || !key) return SB_ERR_INVALID;
    if (!std::strcmp(key, "chunk_accumulate")) {   // split rung: 1 = fp32 RN re-accumulation of the main product per k-half (default)
        e->chunk_accumulate = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "acc_comp_ppb")) {
        e->acc_comp_ppb = value;
        return SB_OK;
    }
    if (!std::strcmp(key, "fuse_se_pool")) {
        e->fuse_se_pool = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "conv_dbg")) {
        e->conv_dbg = value;
        return SB_OK;
    }
    if (!std::strcmp(key, "pack_inputs")) {
        e->pack_inputs = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "pack_threads")) {
        e->pack_threads = std::max(1, value);
        return SB_OK;
    }
    if (!std::strcmp(key, "wide_n")) {   // takes effect at the next weight (re)load, when the tensor maps are rebuilt
        e->wide_n = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "resident_weights")) {
        e->resident_weights = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "chain_forwards")) {
        e->chain_forwards = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "small_batch_split")) {
        e->small_batch_split = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "pdl")) {
        e->use_pdl = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "pdl_aux")) {
        e->pdl_aux = value < 0 ? 0 : value > 2 ? 2 : value;
        return SB_OK;
    }
    if (!std::strcmp(key, "conv_chain")) {
        e->conv_chain = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "layer_overlap")) {
        e->layer_overlap = std::min(std::max(value, 0), 2);
        return SB_OK;
    }
    if (!std::strcmp(key, "tail_split")) {
        e->tail_split = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "nccl")) {   // weight broadcast at the next (re)load: 1 = ncclBroadcast (default), 0 = peer copies
        e->use_nccl = value ? 1 : 0;
        return SB_OK;
    }
    if (!std::strcmp(key, "stats_launch")) {
        e->stats_launch = value;
        return SB_OK;
    }
    if (!std::strcmp(key, "stats")) {
        e->collect_stats = value ? 1 : 0;
        return SB_OK;
    }
    return Fail(e, SB_ERR_INVALID, std::string("unknown option ") + key);
}

}  // extern "C"
