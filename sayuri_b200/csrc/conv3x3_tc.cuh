// conv3x3_tc — 3x3 same-pad convolution (+bias, +residual, *mask, activation) as an implicit GEMM on the
// 5th-gen tensor cores: TMA-staged NHWC tiles in 128B-swizzled shared memory -> tcgen05.mma (M=128,
// N=BN, K=16 per instruction, fp16 operands, fp32 accumulation in TMEM) -> tcgen05.ld epilogue.
//
// Replaces, for the tower and input convolutions, the reference's
//   Convolution<3>::Forward + Im2col          /root/reference/src/neural/blas/convolution.h:41-125
//   (or WinogradConvolution3::Forward         /root/reference/src/neural/blas/winograd_convolution3.cc:280-291)
//   followed by AddSpatialBiases::Forward     /root/reference/src/neural/blas/biases.cc:14-45
// and on the reference GPU path im2col/Winograd kernels + cuBLAS/cuDNN + add_spatial
//   (/root/reference/src/neural/cuda/cuda_kernels.cu:37-79,182-239,522-667).
//
// GEMM view: M = canvas rows (pixels of all samples, halo cells included), N = Cout, K = 9 * Cin.
// Because of the canvas layout (common.cuh) the A operand of tap (ky,kx) is the SAME row-major tile
// shifted by (ky-1)*P + (kx-1) rows, so one "slab" of 256 + 2*24 rows x 64 channels is loaded once per
// work item and k-half and re-used by all 9 taps through row-shifted UMMA descriptors (9x less
// activation traffic than per-tap loads).  Weight blocks [BN][64] stream through a 4-stage ring.
//
// Precision: SPLIT=true evaluates x*w as hi*hi + lo*hi + hi*lo with x = x_hi + x_lo, w = w_hi + w_lo
// (fp16 pairs, ~22 significant bits, fp32 accumulate): the "fp32-faithful" rung that meets the 1e-4
// parity bar.  SPLIT=false uses the hi parts only (the reference's own --fp16 trade).
//
// Warp roles (384 threads, 1 CTA/SM, persistent over work items):
//   warp 0 lane 0 : TMA producer for weight stages      warp 1 lane 0 : tcgen05.mma issuer
//   warp 2        : TMEM allocator                      warp 3 lane 0 : TMA producer for activation slabs
//   warps 4..11   : epilogue (warp%4 = TMEM lane quadrant, (warp-4)/4 = which M=128 tile of the item)
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

struct ConvParams {
    __half* out_hi;
    __half* out_lo;          // unused when !SPLIT
    const __half* res_hi;    // optional residual (same layout as out), nullptr if none
    const __half* res_lo;
    const float* bias;       // [cout]
    const uint8_t* mask;     // [rows]: 1 = real board cell of its sample, 0 = halo / off-board / padding
    int cout;                // real output channels (bias length)
    int out_pitch;           // row pitch (elements) of out/res: cout rounded up to 64
    int kh;                  // number of 64-channel K blocks per tap = padded Cin / 64
    int bn;                  // UMMA N = output channels per work item (multiple of 16, <= 128)
    int n_super;             // number of 256-row work items along M
    int n_ntiles;            // cout / bn
    int pitch;               // P = N + 1
    int act;                 // sb::Act applied after bias (+residual)
    int bo_mode;             // 0: descriptor base_offset = 0 ; 1: base_offset = (start_addr >> 7) & 7
    int* err;                // device int, receives a site code if a barrier wait times out
};

template <bool SPLIT>
struct ConvCfg {
    static constexpr int kParts = SPLIT ? 2 : 1;
    static constexpr int kSlabPartBytes = kSlabRows * 128;            // 304 rows x 64 fp16
    static constexpr int kSlabBytes = kParts * kSlabPartBytes;
    static constexpr int kNumSlabs = 2;
    static constexpr int kBStageBytes = 128 * 128;                    // up to 128 rows x 64 fp16
    static constexpr int kNumBStages = 4;
    static constexpr int kOffB = kNumSlabs * kSlabBytes;
    static constexpr int kOffBar = kOffB + kNumBStages * kBStageBytes;
    static constexpr int kOffBias = kOffBar + 256;
    static constexpr int kSmemBytes = kOffBias + 256 * 4 + 1024;      // + alignment slack
    static constexpr int kTmemCols = 512;                             // 2 stages x 2 tiles x bn (bn <= 128)
    static constexpr int kThreads = 384;
    static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
};

template <bool SPLIT>
__global__ void __launch_bounds__(384, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                  const ConvParams p) {
    using Cfg = ConvCfg<SPLIT>;
    const int KH = p.kh, BN = p.bn;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024 B alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t slab_addr = smem_base;
    const uint32_t bst_addr = smem_base + Cfg::kOffB;
    const uint32_t bar_addr = smem_base + Cfg::kOffBar;
    // barrier slots (8 B each)
    const uint32_t slab_full = bar_addr + 0, slab_empty = bar_addr + 16;          // [2] each
    const uint32_t b_full = bar_addr + 32, b_empty = bar_addr + 64;               // [4] each
    const uint32_t tmem_full = bar_addr + 96, tmem_empty = bar_addr + 112;        // [2] each
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + Cfg::kOffBar + 128);
    float* sbias = reinterpret_cast<float*>(smem_gen + Cfg::kOffBias);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.n_super * p.n_ntiles;

    for (int i = threadIdx.x; i < p.cout && i < 256; i += blockDim.x) sbias[i] = p.bias[i];

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(slab_full + 8 * i, 1);
            mbar_init(slab_empty + 8 * i, 1);
            mbar_init(tmem_full + 8 * i, 1);
            mbar_init(tmem_empty + 8 * i, 8);   // one arrive per epilogue warp
        }
        for (int i = 0; i < 4; ++i) {
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
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 3 && lane == 0) {
        // ===================== activation-slab producer =====================
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int st = item / p.n_ntiles;
            const int row_lo = kGuardRows + st * kSuperRows - kSlabMargin;
            for (int h = 0; h < KH; ++h, ++it) {
                const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
                mbar_wait(slab_empty + 8 * s, ph ^ 1u, p.err, 1);
                mbar_arrive_expect_tx(slab_full + 8 * s, Cfg::kSlabBytes);
#pragma unroll
                for (int part = 0; part < Cfg::kParts; ++part) {
                    const CUtensorMap* tm = part ? &tmA_lo : &tmA_hi;
                    const uint32_t dst = slab_addr + s * Cfg::kSlabBytes + part * Cfg::kSlabPartBytes;
                    tma_load_2d(dst, tm, h * 64, row_lo, slab_full + 8 * s);
                    tma_load_2d(dst + (kSlabRows / 2) * 128, tm, h * 64, row_lo + kSlabRows / 2, slab_full + 8 * s);
                }
            }
        }
    } else if (warp == 0 && lane == 0) {
        // ===================== weight-stage producer =====================
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int nt = item % p.n_ntiles;
            for (int h = 0; h < KH; ++h) {
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int part = 0; part < Cfg::kParts; ++part, ++it) {
                        const uint32_t s = it & 3u, ph = (it >> 2) & 1u;
                        mbar_wait(b_empty + 8 * s, ph ^ 1u, p.err, 2);
                        mbar_arrive_expect_tx(b_full + 8 * s, (uint32_t)BN * 128u);
                        tma_load_2d(bst_addr + s * Cfg::kBStageBytes, part ? &tmW_lo : &tmW_hi,
                                    tap * (KH * 64) + h * 64, nt * BN, b_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_f16(128, BN);
        uint32_t a_it = 0, b_it = 0, j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const uint32_t as = j & 1u, aph = (j >> 1) & 1u;
            mbar_wait(tmem_empty + 8 * as, aph ^ 1u, p.err, 3);
            tc_fence_after();
            for (int h = 0; h < KH; ++h, ++a_it) {
                const uint32_t s = a_it & 1u, sph = (a_it >> 1) & 1u;
                mbar_wait(slab_full + 8 * s, sph, p.err, 4);
                tc_fence_after();
                const uint32_t a_hi = slab_addr + s * Cfg::kSlabBytes;
                const uint32_t a_lo = a_hi + Cfg::kSlabPartBytes;
                for (int tap = 0; tap < 9; ++tap) {
                    const int shift = (tap / 3 - 1) * p.pitch + (tap % 3 - 1);
                    {   // weights hi  x  (activations hi [+ lo])
                        const uint32_t bs = b_it & 3u, bph = (b_it >> 2) & 1u;
                        mbar_wait(b_full + 8 * bs, bph, p.err, 5);
                        tc_fence_after();
                        const uint32_t b_addr = bst_addr + bs * Cfg::kBStageBytes;
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const uint32_t d = tmem_base + (as * 2 + t) * BN;
                            const uint32_t roff = (uint32_t)(kSlabMargin + t * 128 + shift) * 128u;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t aa = a_hi + roff + k * 32;
                                const uint64_t bd = umma_desc_sw128(b_addr + k * 32, 0);
                                umma_f16(d, umma_desc_sw128(aa, p.bo_mode ? (aa >> 7) : 0u), bd, idesc,
                                         (h | tap | k) != 0 ? 1u : 0u);
                                if (SPLIT) {
                                    const uint32_t al = a_lo + roff + k * 32;
                                    umma_f16(d, umma_desc_sw128(al, p.bo_mode ? (al >> 7) : 0u), bd, idesc, 1u);
                                }
                            }
                        }
                        umma_commit(b_empty + 8 * bs);
                        ++b_it;
                    }
                    if (SPLIT) {   // weights lo  x  activations hi
                        const uint32_t bs = b_it & 3u, bph = (b_it >> 2) & 1u;
                        mbar_wait(b_full + 8 * bs, bph, p.err, 6);
                        tc_fence_after();
                        const uint32_t b_addr = bst_addr + bs * Cfg::kBStageBytes;
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const uint32_t d = tmem_base + (as * 2 + t) * BN;
                            const uint32_t roff = (uint32_t)(kSlabMargin + t * 128 + shift) * 128u;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t aa = a_hi + roff + k * 32;
                                umma_f16(d, umma_desc_sw128(aa, p.bo_mode ? (aa >> 7) : 0u),
                                         umma_desc_sw128(b_addr + k * 32, 0), idesc, 1u);
                            }
                        }
                        umma_commit(b_empty + 8 * bs);
                        ++b_it;
                    }
                }
                umma_commit(slab_empty + 8 * s);   // slab reusable once every MMA reading it has retired
            }
            umma_commit(tmem_full + 8 * as);       // accumulators of this item are complete
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int t = (warp - 4) >> 2;
        const int q = warp & 3;
        uint32_t j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const int st = item / p.n_ntiles, nt = item % p.n_ntiles;
            const uint32_t as = j & 1u, aph = (j >> 1) & 1u;
            mbar_wait(tmem_full + 8 * as, aph, p.err, 7);
            tc_fence_after();
            const int row = kGuardRows + st * kSuperRows + t * 128 + q * 32 + lane;
            const bool live = p.mask[row] != 0;
            const size_t off = (size_t)row * p.out_pitch + (size_t)nt * BN;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (as * 2 + t) * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + c0, r);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + sbias[nt * BN + c0 + i];
                if (live && p.res_hi != nullptr) {
                    const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + off + c0);
                    uint4 a0 = rh[0], a1 = rh[1];
                    const __half* hh0 = reinterpret_cast<const __half*>(&a0);
                    const __half* hh1 = reinterpret_cast<const __half*>(&a1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] += __half2float(hh0[i]);
                        v[8 + i] += __half2float(hh1[i]);
                    }
                    if (SPLIT) {
                        const uint4* rl = reinterpret_cast<const uint4*>(p.res_lo + off + c0);
                        uint4 b0 = rl[0], b1 = rl[1];
                        const __half* ll0 = reinterpret_cast<const __half*>(&b0);
                        const __half* ll1 = reinterpret_cast<const __half*>(&b1);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v[i] += __half2float(ll0[i]);
                            v[8 + i] += __half2float(ll1[i]);
                        }
                    }
                }
                __align__(16) __half oh[16];
                __align__(16) __half ol[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = live ? activate(v[i], p.act) : 0.f;   // select, not multiply: garbage rows may hold NaN
                    split_f16(a, oh[i], ol[i]);
                }
                uint4* dh = reinterpret_cast<uint4*>(p.out_hi + off + c0);
                dh[0] = reinterpret_cast<const uint4*>(oh)[0];
                dh[1] = reinterpret_cast<const uint4*>(oh)[1];
                if (SPLIT) {
                    uint4* dl = reinterpret_cast<uint4*>(p.out_lo + off + c0);
                    dl[0] = reinterpret_cast<const uint4*>(ol)[0];
                    dl[1] = reinterpret_cast<const uint4*>(ol)[1];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * as);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace sb
