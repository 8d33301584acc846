// Shared definitions: canvas geometry, activations, fp16 hi/lo split helpers.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// ---- Canvas layout (DESIGN.md "Data layout in HBM") -------------------------------------------
// Activations are [rows] x [C] fp16 matrices (a `hi` tensor and, in split precision, a `lo` tensor with
// x ~= hi + lo) stored CHANNEL-BLOCKED ("C8"): element (row, c) sits at ((c / 8) * R + row) * 8 + c % 8, i.e.
// [C/8 chunks][R rows][8 channels = 16 bytes].  Consequences:
//   * a thread that owns one row (the TMEM epilogue: lane <-> row) and its 31 neighbours touch 32 consecutive
//     16-byte pieces: every global load/store of the epilogues is fully coalesced without staging;
//   * one 4-D TMA box [8 chunks][304 rows][16 B] lands in shared memory exactly as the UMMA no-swizzle K-major
//     "core matrix" layout (8 rows x 16 B contiguous, SBO = 128 B between row groups, LBO = 304*16 B between the
//     two K chunks of an MMA), where a shift by s rows is a plain +16*s bytes on the descriptor start address.
// Sample b, board cell (y, x) of an N x N canvas lives at row
//     kGuardRows + b * SS + y * P + x,     P = N + 1,  SS = (N + 1) * (N + 1)
// i.e. every board row carries ONE trailing halo cell and every sample ONE trailing halo row; the
// halo cell right of row y doubles as the halo left of row y+1, the halo row below sample b doubles
// as the halo above sample b+1.  All halo cells are kept at zero by the epilogues, so a 3x3 tap
// (ky, kx) of the convolution is the constant row shift (ky-1)*P + (kx-1): the implicit GEMM needs no
// im2col, only shifted views of one resident tile (19x19: 400 rows per sample for 361 real cells).
constexpr int kGuardRows = 32;   // zero rows in front of sample 0 (top halo of sample 0, TMA coords stay >= 0)
constexpr int kSuperRows = 256;  // rows per CTA work item: two UMMA M=128 tiles
constexpr int kSlabMargin = 24;  // slab rows loaded before/after the super tile (>= P + 1, multiple of 8)
constexpr int kSlabRows = kSuperRows + 2 * kSlabMargin;  // 304
constexpr int kMaxBoard = 19;    // /root/reference/src/game/types.h:5-7 (MAX_BOARD_SIZE)
constexpr int kMaxIntersections = kMaxBoard * kMaxBoard;
constexpr int kInputChannels = 43;   // /root/reference/src/neural/network_basic.h:10
constexpr int kInputChannelsPadded = 64;
constexpr int kMaxConvWidth = 512;   // widest conv output (Mixer feed-forward width of a 256-wide net is 384); bias staging in smem

// element (row, c) of a C8 tensor with R rows
__host__ __device__ inline size_t act_index(int row, int c, int R) {
    return ((size_t)(c >> 3) * (size_t)R + (size_t)row) * 8 + (size_t)(c & 7);
}

struct Geom {
    int N, P, SS;
    __host__ __device__ Geom() : N(0), P(0), SS(0) {}
    __host__ __device__ explicit Geom(int n) : N(n), P(n + 1), SS((n + 1) * (n + 1)) {}
    __host__ __device__ int row(int b, int y, int x) const { return kGuardRows + b * SS + y * P + x; }
    __host__ __device__ int rows_used(int batch) const { return kGuardRows + batch * SS; }
    __host__ __device__ int n_super(int batch) const { return (batch * SS + kSuperRows - 1) / kSuperRows; }
    __host__ __device__ int rows_alloc(int max_batch) const { return kGuardRows + n_super(max_batch) * kSuperRows + 64; }
};

// ---- Programmatic dependent launch ------------------------------------------------------------------------------
// Every kernel of a forward is launched with cudaLaunchAttributeProgrammaticStreamSerialization: the CTAs of kernel
// k+1 become resident as soon as SMs free up, run whatever does not depend on kernel k (for the convolution: barrier
// init, TMEM alloc, bias, first WEIGHT stages — constants) and block in griddepcontrol.wait only before touching
// activations written (or still read) by kernel k; the launch latency and the prologue disappear behind kernel k's tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- Activations: same ints and formulas as /root/reference/src/neural/activation.h:8-17,43-59 --
enum Act : int { kIdentity = 0, kReLU = 1, kELU = 2, kSELU = 3, kGELU = 4, kMISH = 5, kSwish = 6, kHardSwish = 7 };

// Compile-time activation: ONE formula is inlined per kernel instantiation (an 8-way runtime switch per
// element made the conv epilogue 25K instructions long and instruction-fetch bound).
// exp() goes through ex2.approx (rel. error ~2^-22) and the mish/swish quotients through rcp.approx: unbiased
// ~1e-7 relative perturbations, far below the 1e-4 parity bar.
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int ACT>
__device__ __forceinline__ float activate_t(float x) {
    if (ACT == kReLU) return fmaxf(x, 0.f);
    if (ACT == kELU) return x > 0.f ? x : (expf(x) - 1.f);
    if (ACT == kSELU) return x > 0.f ? (1.05070098f * x) : (1.05070098f * 1.67326324f * (expf(x) - 1.0f));
    if (ACT == kGELU) return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
    if (ACT == kMISH) {
        // x * tanh(log(1 + e^x)) == x * n / (n + 2), n = e^x (e^x + 2): no cancellation, one ex2 + one rcp.
        const float w = fast_exp(fminf(x, 20.f));
        const float n = w * (w + 2.f);
        return x * n * fast_rcp(n + 2.f);
    }
    if (ACT == kSwish) return x * fast_rcp(1.0f + fast_exp(-x));
    if (ACT == kHardSwish) return x >= 3.f ? x : x <= -3.f ? 0.f : (x * (x + 3.0f) * (1.0f / 6.0f));
    return x;
}

// Runtime-dispatched form for the tiny per-sample FC kernels (executed a handful of times per sample).
__device__ __noinline__ float activate(float x, int act) {
    switch (act) {
        case kReLU: return activate_t<kReLU>(x);
        case kELU: return activate_t<kELU>(x);
        case kSELU: return activate_t<kSELU>(x);
        case kGELU: return activate_t<kGELU>(x);
        case kMISH: return activate_t<kMISH>(x);
        case kSwish: return activate_t<kSwish>(x);
        case kHardSwish: return activate_t<kHardSwish>(x);
        default: return x;
    }
}

// Host-side dispatch of a kernel template over the activation enum.
#define SB_DISPATCH_ACT(act, ACT, ...)                                  \
    switch (act) {                                                      \
        case sb::kReLU: { constexpr int ACT = sb::kReLU; __VA_ARGS__; } break;           \
        case sb::kELU: { constexpr int ACT = sb::kELU; __VA_ARGS__; } break;             \
        case sb::kSELU: { constexpr int ACT = sb::kSELU; __VA_ARGS__; } break;           \
        case sb::kGELU: { constexpr int ACT = sb::kGELU; __VA_ARGS__; } break;           \
        case sb::kMISH: { constexpr int ACT = sb::kMISH; __VA_ARGS__; } break;           \
        case sb::kSwish: { constexpr int ACT = sb::kSwish; __VA_ARGS__; } break;         \
        case sb::kHardSwish: { constexpr int ACT = sb::kHardSwish; __VA_ARGS__; } break; \
        default: { constexpr int ACT = sb::kIdentity; __VA_ARGS__; } break;              \
    }

// Stage-wise (structure-of-arrays) activation of 16 values: every MUFU / FMA stage is issued for all 16
// elements before the next stage starts, so the ~100-cycle ex2 -> rcp dependency chain of one element is
// overlapped with its 15 neighbours (an element-by-element loop ran latency-bound: 110 cycles per element).
template <int ACT>
__device__ __forceinline__ void activate16(float (&v)[16]) {
    if (ACT == kMISH) {
        float w[16], n[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = fast_exp(fminf(v[i], 20.f));
#pragma unroll
        for (int i = 0; i < 16; ++i) n[i] = w[i] * (w[i] + 2.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = fast_rcp(n[i] + 2.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] * n[i] * w[i];
    } else if (ACT == kSwish) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = fast_exp(-v[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = fast_rcp(1.0f + w[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] * w[i];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = activate_t<ACT>(v[i]);
    }
}

// 16 floats -> 16 fp16 `hi` (+ 16 fp16 `lo` residues), packed two per register, stage-wise.
__device__ __forceinline__ void split16(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8], bool want_lo) {
    __half2 h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float a = fminf(fmaxf(v[2 * i], -60000.f), 60000.f);
        const float b = fminf(fmaxf(v[2 * i + 1], -60000.f), 60000.f);
        h[i] = __floats2half2_rn(a, b);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h[i]);
    }
    if (want_lo) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = __half22float2(h[i]);
            const __half2 l = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
    }
}

// ---- fp16 hi/lo split: v ~= hi + lo with ~22 significant bits ---------------------------------
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
    v = fminf(fmaxf(v, -60000.f), 60000.f);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

}  // namespace sb
