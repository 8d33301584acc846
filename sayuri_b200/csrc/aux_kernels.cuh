// HBM-/latency-bound kernels around the tensor-core convolution: input unpack + canvas placement,
// SE squeeze/excite, SE scale, head 1x1 convs, head pooling + FCs, output gather.  All plain SIMT with
// vectorised (16 B) row accesses; per-sample reductions run in a fixed order (batch-invariant results).
#pragma once
#include "common.cuh"
#include "../../include/sayuri_b200.h"

namespace sb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void load8(const __half* hi, const __half* lo, bool split, float (&v)[8]) {
    const uint4 a = *reinterpret_cast<const uint4*>(hi);
    const __half* ha = reinterpret_cast<const __half*>(&a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __half2float(ha[i]);
    if (split) {
        const uint4 b = *reinterpret_cast<const uint4*>(lo);
        const __half* hb = reinterpret_cast<const __half*>(&b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += __half2float(hb[i]);
    }
}

__device__ __forceinline__ void store8(__half* hi, __half* lo, bool split, const float (&v)[8]) {
    __align__(16) __half oh[8];
    __align__(16) __half ol[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_f16(v[i], oh[i], ol[i]);
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(oh);
    if (split) *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(ol);
}

// ------------------------------------------------------------------------------------------------
// unpack_planes: fp32 NCHW planes at each sample's NATIVE board size (InputData.planes,
// /root/reference/src/neural/network_basic.h:23-34, encoder.cc:31-50) -> NHWC canvas rows with Cin
// padded 43 -> 64, top-left placement on the N x N canvas with zeros elsewhere (the re-layout of
// BatchForwardPipe::SendQueryAndWait, batch_forward_pipe.cc:15-33), fp16 hi/lo split, and the per-row
// board mask (ApplyMask, cuda_forward_pipe.cc:636-682).  One thread per (canvas row, 8-channel group).
__global__ void unpack_planes_kernel(const float* __restrict__ planes, size_t sample_stride,
                                     const int* __restrict__ board_sizes, Geom g, int n, int n_rows, int R,
                                     __half* __restrict__ hi, __half* __restrict__ lo, bool split,
                                     uint8_t* __restrict__ mask) {
    pdl_launch_dependents();
    pdl_wait();   // the canvas may still be read by the previous forward's input convolution
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = idx / n_rows, r = idx - cg * n_rows;   // rows fastest: coalesced 16-byte pieces
    if (cg >= 8) return;
    const int b = r / g.SS, rem = r - b * g.SS;
    const int y = rem / g.P, x = rem - y * g.P;
    int bs = 0;
    if (b < n) bs = board_sizes[b];
    const bool live = (b < n) && (y < bs) && (x < bs);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cg * 8 + i;
        v[i] = (live && c < kInputChannels) ? planes[(size_t)b * sample_stride + (size_t)c * bs * bs + y * bs + x] : 0.f;
    }
    const size_t off = act_index(kGuardRows + r, cg * 8, R);
    store8(hi + off, lo + off, split, v);
    if (cg == 0) mask[kGuardRows + r] = live ? 1 : 0;
}

// unpack_packed: the same canvas construction from COMPACT records (sb_packed_position, include/sayuri_b200.h):
// plane c of sample b is scale[c] * bit-mask, so the host->device hop carries 2.2 KB per position instead of 62 KB.
// Samples flagged SB_PACKED_RAW read their fp32 planes from `raw` (same layout as unpack_planes) instead.
// Also scatters board sizes / policy offsets of the records into the d_meta arrays the later kernels read.
__global__ void unpack_packed_kernel(const sb_packed_position* __restrict__ rec, const float* __restrict__ raw,
                                     size_t sample_stride, Geom g, int n, int n_rows, int R, int max_batch,
                                     __half* __restrict__ hi, __half* __restrict__ lo, bool split,
                                     uint8_t* __restrict__ mask, int* __restrict__ meta) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        meta[idx] = rec[idx].board_size;
        meta[max_batch + idx] = rec[idx].offset;
    }
    const int cg = idx / n_rows, r = idx - cg * n_rows;   // rows fastest: coalesced 16-byte pieces
    if (cg >= 8) return;
    const int b = r / g.SS, rem = r - b * g.SS;
    const int y = rem / g.P, x = rem - y * g.P;
    int bs = 0;
    if (b < n) bs = rec[b].board_size;
    const bool live = (b < n) && (y < bs) && (x < bs);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (live) {
        const sb_packed_position& p = rec[b];
        const int cell = y * bs + x;
        if (p.flags & SB_PACKED_RAW) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cg * 8 + i;
                if (c < kInputChannels) v[i] = raw[(size_t)b * sample_stride + (size_t)c * bs * bs + cell];
            }
        } else {
            const int word = cell >> 5, bit = cell & 31;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cg * 8 + i;
                if (c < kInputChannels && ((p.bits[c][word] >> bit) & 1u)) v[i] = p.scale[c];
            }
        }
    }
    const size_t off = act_index(kGuardRows + r, cg * 8, R);
    store8(hi + off, lo + off, split, v);
    if (cg == 0) mask[kGuardRows + r] = live ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Pooling of a SLICE of the channels of one sample by one CTA (256 threads = 8 warps), for the multi-CTA-per-sample
// kernels below.  The CTA owns chunks [chunk0, chunk0 + ncl) (ncl <= 8); warp w takes chunk chunk0 + w % ncl and row
// part w / ncl of 8 / ncl equal row ranges; lanes stride the rows with every load of the thread in flight at once; a
// fixed shuffle tree and a fixed-order sum over the row parts finish the reduction.  Which CTA pools which channel
// and in which order depends only on (C, board geometry): results are batch-invariant.
// Writes s_sum / s_max [ncl * 8] (shared).  Needs s_part [2][8 warps][8] floats of scratch.
__device__ __forceinline__ void pool_slice_c8(const __half* __restrict__ hi, const __half* __restrict__ lo, bool split,
                                              const uint8_t* __restrict__ mask, int row0, int SS, int R, int chunk0, int ncl,
                                              float* s_part, float* s_sum, float* s_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int parts = 8 / ncl;
    const int cl = warp % ncl, part = warp / ncl;
    float s[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s[i] = 0.f;
        m[i] = -5000.f;   // "crazy negative value", se_unit.cc:22
    }
    if (part < parts) {
        const int per = (SS + parts - 1) / parts;
        const int r_end = min(SS, (part + 1) * per);
#pragma unroll 8
        for (int r = part * per + lane; r < r_end; r += 32) {
            const bool live = mask[row0 + r] != 0;
            float v[8];
            const size_t off = act_index(row0 + r, (chunk0 + cl) * 8, R);
            load8(hi + off, lo + off, split, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] += live ? v[i] : 0.f;
                m[i] = live ? fmaxf(m[i], v[i]) : m[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s_part[warp * 8 + i] = s[i];
            s_part[64 + warp * 8 + i] = m[i];
        }
    }
    __syncthreads();
    if (threadIdx.x < ncl * 8) {
        const int c = threadIdx.x >> 3, i = threadIdx.x & 7;
        float ss = 0.f, mm = -5000.f;
        for (int pt = 0; pt < parts; ++pt) {      // fixed order over the row parts
            ss += s_part[(pt * ncl + c) * 8 + i];
            mm = fmaxf(mm, s_part[64 + (pt * ncl + c) * 8 + i]);
        }
        s_sum[threadIdx.x] = ss;
        s_max[threadIdx.x] = mm;
    }
    __syncthreads();
}

// The last CTA of a sample to publish its slice runs the sample's fully-connected tail (threadfence + counter).
__device__ __forceinline__ bool last_cta_of_sample(int* counter, int n_ctas, int* s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(counter, 1);
        *s_flag = (prev == n_ctas - 1) ? 1 : 0;
        if (prev == n_ctas - 1) *counter = 0;      // ready for the next launch
    }
    __syncthreads();
    const bool last = *s_flag != 0;
    if (last) __threadfence();
    return last;
}

// Small fully-connected layer y = W x (+ b) for the per-sample tails (256 threads).  L = 1, 2, 4 or 8 adjacent lanes share one
// output (L chosen so that all outputs of the layer are computed in as few passes as possible: 256 outputs -> one thread
// each, 32 outputs -> 8 lanes each); a lane strides the input in 16-byte pieces and issues up to 8 weight loads before
// it starts to accumulate, a fixed shuffle tree over the L lanes finishes each output.  These tails are pure latency (one
// CTA per sample, weights from L2): with one warp per output and the outputs of a warp in sequence they took 11 us, with
// scalar strided loads and 8 sequential passes for the 2C excite outputs 18 us per SE unit.  x in shared memory, 16-byte
// aligned; calls `emit(o, sum)` from the first lane of each output.  The summation order depends only on (in, out).
template <typename Emit>
__device__ __forceinline__ void fc_tail_8lanes(const float* __restrict__ w, const float* x, int in, int out, Emit emit) {
    int L = 8;
    while (L > 1 && out * L > 256) L >>= 1;
    const int sub = threadIdx.x & (L - 1);
    const int per_pass = 256 / L;
    const bool vec = (in & 3) == 0;
    for (int o0 = 0; o0 < out; o0 += per_pass) {
        const int o = o0 + threadIdx.x / L;
        float acc = 0.f;
        if (o < out) {
            const float* wr = w + (size_t)o * in;
            if (vec) {
                const float4* w4 = reinterpret_cast<const float4*>(wr);
                const float4* x4 = reinterpret_cast<const float4*>(x);
                const int n4 = in >> 2;
                for (int c0 = sub; c0 < n4; c0 += 8 * L) {
                    float4 wv[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) wv[k] = (c0 + k * L < n4) ? w4[c0 + k * L] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (c0 + k * L < n4) {
                            const float4 xv = x4[c0 + k * L];
                            acc += wv[k].x * xv.x;
                            acc += wv[k].y * xv.y;
                            acc += wv[k].z * xv.z;
                            acc += wv[k].w * xv.w;
                        }
                    }
                }
            } else {
                for (int i = sub; i < in; i += L) acc += wr[i] * x[i];
            }
        }
        if (L > 1) acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (L > 2) acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (L > 4) acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (o < out && sub == 0) emit(o, acc);
    }
}

// se_pool_fc: GlobalPooling<false> + squeeze FC + excite FC of SEUnit::Forward
// (/root/reference/src/neural/blas/se_unit.cc:9-37,70-90; GPU twins cuda_kernels.cu:241-321 and the
// cuBLAS FCs cuda_layers.cc:975-1017).  grid (C/64, n): every CTA pools 64 channels of one sample into
// pooled[n][2C] (sums, maxima); the sample's last CTA (threadfence + counter) then runs the two FCs and writes
// sigmoid(gamma) and beta, gb[n][2C].  Mean divides by the sample's own n^2, (n-14)/10 uses the sample's own n.
// (Letting that last CTA also apply the SE scaling to its sample — one launch less — measured SLOWER: 95 us vs 54 us
// per SE unit at batch 256, one CTA per sample is too little parallelism for a 157 MB pass.)
template <int ACT>
__global__ void __launch_bounds__(256)
se_pool_fc_kernel(const __half* __restrict__ u_hi, const __half* __restrict__ u_lo, bool split,
                  const uint8_t* __restrict__ mask, const int* __restrict__ board_sizes, Geom g, int C, int R, int se,
                  const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ b2, float* pooled, int* counters, float* __restrict__ gb, int dbg) {
    extern __shared__ float sm[];
    __shared__ float s_part[128];
    __shared__ float s_sum[64], s_max[64];
    __shared__ int s_flag;
    pdl_launch_dependents();
    pdl_wait();   // u is the previous convolution's output
    const int b = blockIdx.y;
    const int chunk0 = blockIdx.x * 8;          // 8 chunks = 64 channels per CTA: 2 CTAs per sample at C = 128, so that
    const int ncl = min(8, (C >> 3) - chunk0);  // the whole grid is resident at once (48 registers: 5 CTAs per SM)
    const int row0 = kGuardRows + b * g.SS;
    if (!(dbg & 512)) pool_slice_c8(u_hi, u_lo, split, mask, row0, g.SS, R, chunk0, ncl, s_part, s_sum, s_max);
    if (dbg & 256) return;
    const int tid = threadIdx.x;
    if (tid < ncl * 8) {
        pooled[(size_t)b * 2 * C + chunk0 * 8 + tid] = s_sum[tid];
        pooled[(size_t)b * 2 * C + C + chunk0 * 8 + tid] = s_max[tid];
    }
    if (!last_cta_of_sample(counters + b, gridDim.x, &s_flag)) return;

    const int bs = board_sizes[b];
    float* pool = sm;               // [3C]
    float* hid = pool + 3 * C;      // [se]
    const float b_coeff = ((float)bs - 14.0f) / 10.f;   // se_unit.h:17-20
    const volatile float* pv = pooled + (size_t)b * 2 * C;   // written by other CTAs: bypass L1
    for (int c = tid; c < C; c += 256) {
        const float mean = pv[c] / (float)(bs * bs);
        pool[c] = mean;
        pool[C + c] = mean * b_coeff;
        pool[2 * C + c] = pv[C + c];
    }
    __syncthreads();
    fc_tail_8lanes(w1, pool, 3 * C, se, [&](int o, float v) { hid[o] = activate_t<ACT>(v + b1[o]); });   // squeeze, activation
    __syncthreads();
    fc_tail_8lanes(w2, hid, se, 2 * C, [&](int o, float v) {                                             // excite, identity
        v += b2[o];
        gb[(size_t)b * 2 * C + o] = o < C ? 1.0f / (1.0f + expf(-v)) : v;   // gamma = sigmoid, se_unit.cc:103
    });
}

// se_fc: the SE unit behind a convolution that already left pooling partials (conv3x3_tc2, POOL): part[group][2][C] holds
// the sum and the maximum of every channel over the board cells of each aligned group of canvas rows; a sample owns
// `gps` consecutive groups.  One CTA per sample adds its groups in a FIXED order (batch-invariant), then runs the squeeze
// and excite FCs exactly like se_pool_fc.  Replaces the pooling pass over the whole tensor (52 MB at batch 256, C = 128)
// by a 6.5 MB read.  (GlobalPooling + the two FCs of SEUnit::Forward, se_unit.cc:9-37,70-90.)
template <int ACT>
__global__ void __launch_bounds__(256)
se_fc_kernel(const float* __restrict__ part, int gps, const int* __restrict__ board_sizes, int C, int se,
             const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
             const float* __restrict__ b2, float* __restrict__ gb) {
    extern __shared__ float sm[];
    pdl_launch_dependents();
    pdl_wait();   // the partials are the previous convolution's output
    const int b = blockIdx.x, tid = threadIdx.x;
    const int bs = board_sizes[b];
    float* pool = sm;               // [3C]
    float* hid = pool + 3 * C;      // [se]
    const float b_coeff = ((float)bs - 14.0f) / 10.f;   // se_unit.h:17-20
    const float* base = part + (size_t)b * gps * 2 * C;
    // thread t owns entry t of the 2C-float group records (t < C: the sum of channel t, else the maximum of channel t - C):
    // every load of the CTA is one contiguous run; 8 groups are in flight at a time, added in group order
    for (int t = tid; t < 2 * C; t += 256) {
        const bool is_sum = t < C;
        float r = is_sum ? 0.f : -5000.f;
        for (int g0 = 0; g0 < gps; g0 += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (g0 + k < gps) ? base[(size_t)(g0 + k) * 2 * C + t] : (is_sum ? 0.f : -5000.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) r = is_sum ? r + v[k] : fmaxf(r, v[k]);
        }
        if (is_sum) {
            const float mean = r / (float)(bs * bs);
            pool[t] = mean;
            pool[C + t] = mean * b_coeff;
        } else {
            pool[C + t] = r;      // pool[2C + c]
        }
    }
    __syncthreads();
    fc_tail_8lanes(w1, pool, 3 * C, se, [&](int o, float v) { hid[o] = activate_t<ACT>(v + b1[o]); });   // squeeze, activation
    __syncthreads();
    fc_tail_8lanes(w2, hid, se, 2 * C, [&](int o, float v) {                                             // excite, identity
        v += b2[o];
        gb[(size_t)b * 2 * C + o] = o < C ? 1.0f / (1.0f + expf(-v)) : v;   // gamma = sigmoid, se_unit.cc:103
    });
}

// se_apply: x' = act(sigmoid(gamma) * u + beta + skip) on board cells, 0 elsewhere
// (SEUnit::SEProcess, se_unit.cc:92-128; GPU twin se_scale_kernel cuda_kernels.cu:391-440).  In place on u.
template <int ACT>
__global__ void se_apply_kernel(__half* __restrict__ u_hi, __half* __restrict__ u_lo,
                                const __half* __restrict__ x_hi, const __half* __restrict__ x_lo, bool split,
                                const uint8_t* __restrict__ mask, const float* __restrict__ gb, Geom g, int C,
                                int R, int n_rows) {
    pdl_launch_dependents();
    pdl_wait();   // gb comes from se_pool_fc
    const int groups = C >> 3;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = (int)(idx / n_rows), r = (int)(idx - (size_t)cg * n_rows);   // rows fastest
    if (cg >= groups) return;
    const int row = kGuardRows + r;
    const size_t off = act_index(row, cg * 8, R);
    float o[8];
    if (mask[row]) {
        const int b = r / g.SS;
        float u[8], x[8];
        load8(u_hi + off, u_lo + off, split, u);
        load8(x_hi + off, x_lo + off, split, x);
        const float* ga = gb + (size_t)b * 2 * C + cg * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = activate_t<ACT>(ga[i] * u[i] + ga[C + i] + x[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    store8(u_hi + off, u_lo + off, split, o);
}

// ------------------------------------------------------------------------------------------------
// dwconv: DepthwiseConvolution::Forward (convolution.cc:27-62) + bias + activation on the C8 canvas, optionally
// followed by "+ input" (AddSpatialBiasesPost, biases.cc:47-77: activation FIRST, then the residual) — the first
// stage of a Mixer block (blas_forward_pipe.cc:265-285) and the depthwise stage of the RepLK policy head
// (:443-457).  One thread per (canvas row, 8-channel chunk); the k x k taps of a chunk sit in shared memory; taps
// are bounds-checked against the sample's own board (a shift of more than one cell can land in a neighbouring
// board row or sample, which the one-cell zero halo does not cover).  w: fp32 [C][k*k], k odd <= 15.
template <int ACT>
__global__ void __launch_bounds__(128)
dwconv_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, __half* __restrict__ out_hi,
              __half* __restrict__ out_lo, bool split, const float* __restrict__ w, const float* __restrict__ bias,
              const int* __restrict__ board_sizes, Geom g, int n, int n_rows, int R_in, int R_out, int k, bool add_input) {
    // Shared memory: the k*k taps of this 8-channel chunk, and the input rows [r0 - halo, r0 + 128 + halo) of the chunk
    // converted to fp32 once (hi + lo), halo = (k/2) * (P + 1) rows: every tap of every thread is then one 32-byte
    // shared-memory read instead of two global loads and a conversion.
    extern __shared__ float dw_smem[];
    const int chunk = blockIdx.y;
    const int kk = k * k, pad = k >> 1;
    const int halo = pad * (g.P + 1);
    const int tile_rows = 128 + 2 * halo;
    float* sw = dw_smem;                     // [kk][8]
    float* sb8 = sw + kk * 8;                // [8]
    float* tile = sb8 + 8;                   // [tile_rows][8]
    for (int i = threadIdx.x; i < kk * 8; i += blockDim.x) {
        const int tap = i >> 3, c = i & 7;
        sw[i] = w[(size_t)(chunk * 8 + c) * kk + tap];
    }
    if (threadIdx.x < 8) sb8[threadIdx.x] = bias[chunk * 8 + threadIdx.x];
    const int r0 = blockIdx.x * 128;
    for (int t = threadIdx.x; t < tile_rows; t += blockDim.x) {
        const int grow = kGuardRows + r0 - halo + t;    // global canvas row of tile row t
        float v[8];
        if (grow >= 0 && grow < R_in) {
            const size_t off = act_index(grow, chunk * 8, R_in);
            load8(in_hi + off, in_lo + off, split, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float4* dst = reinterpret_cast<float4*>(tile + t * 8);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    const int r = r0 + threadIdx.x;
    if (r >= n_rows) return;
    const int b = r / g.SS, rem = r - b * g.SS;
    const int y = rem / g.P, x = rem - y * g.P;
    const int bs = b < n ? board_sizes[b] : 0;
    const bool live = (y < bs) && (x < bs);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (live) {
        const int t0 = threadIdx.x + halo;          // this thread's own row inside the tile
        for (int ky = 0; ky < k; ++ky) {
            const int yy = y + ky - pad;
            if ((unsigned)yy >= (unsigned)bs) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int xx = x + kx - pad;
                if ((unsigned)xx >= (unsigned)bs) continue;
                const float4* src = reinterpret_cast<const float4*>(tile + (t0 + (ky - pad) * g.P + (kx - pad)) * 8);
                const float4 a = src[0], c4 = src[1];
                const float* wt = sw + (ky * k + kx) * 8;   // same tap order as the reference loop
                acc[0] += a.x * wt[0];
                acc[1] += a.y * wt[1];
                acc[2] += a.z * wt[2];
                acc[3] += a.w * wt[3];
                acc[4] += c4.x * wt[4];
                acc[5] += c4.y * wt[5];
                acc[6] += c4.z * wt[6];
                acc[7] += c4.w * wt[7];
            }
        }
        const float* self = tile + t0 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] = activate_t<ACT>(acc[i] + sb8[i]);
            if (add_input) acc[i] += self[i];
        }
    }
    const size_t off = act_index(kGuardRows + r, chunk * 8, R_out);
    store8(out_hi + off, out_lo + off, split, acc);
}

// ------------------------------------------------------------------------------------------------
// (The two head-entry 1x1 convolutions, blas_forward_pipe.cc:430-442,513-522, run as ONE single-tap launch of
// the tensor-core conv kernel with the policy and value filters concatenated: pv[row][0:P] policy, [P:P+V] value.)
struct HeadWeights {
    const float* p_inter_w;  // [P][3P]
    const float* p_inter_b;  // [P]
    const float* pass_w;     // [5][P]
    const float* pass_b;     // [5]
    const float* v_inter_w;  // [3V][3V]
    const float* v_inter_b;  // [3V]
    const float* misc_w;     // [15][3V]
    const float* misc_b;     // [15]
    const float* prob_w;     // [5][P]
    const float* prob_b;     // [5]
    const float* own_w;      // [V]
    const float* own_b;      // [1]
};

// Per-sample output record handed back through the C ABI (sb_output in include/sayuri_b200.h).
constexpr int kOutFloats = 2 * kMaxIntersections + 8;

// head_fused: everything behind the head-entry convolution in ONE launch:
//   GlobalPooling<false> of the policy planes, GlobalPooling<true> of the value planes (se_unit.cc:9-68), the four small
//   FCs (blas_forward_pipe.cc:473-481,501-507,524-532,549-555), policy_conv += intermediate (:483-484), the P->5 and V->1
//   1x1 convs (:487-499,535-547) and FillOutputs (:597-618).
// grid (ceil((P+V)/32), n): every CTA pools up to 32 channels of pv (C8: P policy + V value channels) of one sample;
// the sample's LAST CTA runs the FCs and then writes the sample's output record: only the requested policy channel
// `offset` is evaluated, outputs are in the sample's NATIVE n x n order (the crop of batch_forward_pipe.cc:48-67) and
// zero-filled up to 361, misc values gathered to {pass[offset], wdl0..2, stm(3), final_score(8), q_error(13),
// score_error(14)}.
__global__ void __launch_bounds__(256)
head_fused_kernel(const __half* __restrict__ pv_hi, const __half* __restrict__ pv_lo, bool split, int R,
                  const uint8_t* __restrict__ mask, const int* __restrict__ board_sizes,
                  const int* __restrict__ offsets, Geom g, int P, int V, HeadWeights hw, int act, float* pooled,
                  int* counters, float* __restrict__ out) {
    extern __shared__ float sm[];
    __shared__ float s_part[128];
    __shared__ float s_sum[64], s_max[64];
    __shared__ float s_pass[5], s_misc[15];
    __shared__ int s_flag;
    pdl_launch_dependents();
    pdl_wait();   // pv is the head-entry convolution's output
    const int PV = P + V;
    const int b = blockIdx.y, bs = board_sizes[b];
    const int chunk0 = blockIdx.x * 4;
    const int ncl = min(4, (PV >> 3) - chunk0);
    pool_slice_c8(pv_hi, pv_lo, split, mask, kGuardRows + b * g.SS, g.SS, R, chunk0, ncl, s_part, s_sum, s_max);
    const int tid = threadIdx.x;
    if (tid < ncl * 8) {
        pooled[(size_t)b * 2 * PV + chunk0 * 8 + tid] = s_sum[tid];
        pooled[(size_t)b * 2 * PV + PV + chunk0 * 8 + tid] = s_max[tid];
    }
    if (!last_cta_of_sample(counters + b, gridDim.x, &s_flag)) return;

    float* ppool = sm;                       // [3P]
    float* vpool = ppool + 3 * P;            // [3V]
    float* spint = vpool + 3 * V;            // [P]
    float* svint = spint + P;                // [3V]
    const volatile float* pl = pooled + (size_t)b * 2 * PV;   // written by other CTAs: bypass L1
    const float b_diff = (float)bs - 14.0f;
    if (tid < PV) {
        const float mean = pl[tid] / (float)(bs * bs);
        if (tid < P) {
            ppool[tid] = mean;
            ppool[P + tid] = mean * (b_diff / 10.f);
            ppool[2 * P + tid] = pl[PV + tid];
        } else {
            const int v = tid - P;
            vpool[v] = mean;
            vpool[V + v] = mean * (b_diff / 10.f);
            vpool[2 * V + v] = mean * (b_diff * b_diff / 100.f - 0.1f);
        }
    }
    __syncthreads();
    fc_tail_8lanes(hw.p_inter_w, ppool, 3 * P, P, [&](int o, float v) { spint[o] = activate(v + hw.p_inter_b[o], act); });
    fc_tail_8lanes(hw.v_inter_w, vpool, 3 * V, 3 * V, [&](int o, float v) { svint[o] = activate(v + hw.v_inter_b[o], act); });
    __syncthreads();
    fc_tail_8lanes(hw.pass_w, spint, P, 5, [&](int o, float v) { s_pass[o] = v + hw.pass_b[o]; });
    fc_tail_8lanes(hw.misc_w, svint, 3 * V, 15, [&](int o, float v) { s_misc[o] = v + hw.misc_b[o]; });
    __syncthreads();

    const int off = offsets[b];
    float* o = out + (size_t)b * kOutFloats;
    for (int i = tid; i < kMaxIntersections + 8; i += 256) {
        if (i < kMaxIntersections) {
            float prob = 0.f, own = 0.f;
            if (i < bs * bs) {
                const int y = i / bs, x = i - y * bs;
                const int row = g.row(b, y, x);
                prob = hw.prob_b[off];
                own = hw.own_b[0];
                for (int c0 = 0; c0 < PV; c0 += 8) {     // P and V are multiples of 8: a chunk is all-policy or all-value
                    float v[8];
                    const size_t idx = act_index(row, c0, R);
                    load8(pv_hi + idx, pv_lo + idx, split, v);
                    if (c0 < P) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) prob += hw.prob_w[off * P + c0 + k] * (v[k] + spint[c0 + k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) own += hw.own_w[c0 - P + k] * v[k];
                    }
                }
            }
            o[i] = prob;
            o[kMaxIntersections + i] = own;
        } else {
            const int k = i - kMaxIntersections;
            float v;
            switch (k) {
                case 0: v = s_pass[off]; break;
                case 1: v = s_misc[0]; break;
                case 2: v = s_misc[1]; break;
                case 3: v = s_misc[2]; break;
                case 4: v = s_misc[3]; break;
                case 5: v = s_misc[8]; break;
                case 6: v = s_misc[13]; break;
                default: v = s_misc[14]; break;
            }
            o[2 * kMaxIntersections + k] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// conv3x3_simt: fp32 CUDA-core evaluation of the SAME convolution on the SAME canvas buffers.  It is
// the on-device cross-check for conv3x3_tc (tests compare the two layer by layer) and never the
// default; selected only with precision = SB_PRECISION_SIMT_DEBUG.
// wT: fp32 [9*CINp][cout] (k-major rows, cout contiguous).  One thread per (row, cout).
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, bool split, int cinp,
                    const float* __restrict__ wT, const float* __restrict__ bias,
                    const __half* __restrict__ res_hi, const __half* __restrict__ res_lo,
                    const uint8_t* __restrict__ mask, int cout, int n_rows, int pitch, int ntaps, int act,
                    __half* __restrict__ out_hi, __half* __restrict__ out_lo, int R) {
    const int co = blockIdx.y * 32 + threadIdx.x;
    const int r = blockIdx.x * 8 + threadIdx.y;
    if (r >= n_rows || co >= cout) return;
    const int row = kGuardRows + r;
    const size_t off = act_index(row, co, R);
    float v = 0.f;
    if (mask[row]) {
        float acc = 0.f;
        for (int tap = 0; tap < ntaps; ++tap) {
            const int src = ntaps == 1 ? row : row + (tap / 3 - 1) * pitch + (tap % 3 - 1);
            const float* w = wT + (size_t)tap * cinp * cout + co;
            for (int c = 0; c < cinp; ++c) {
                const size_t ai = act_index(src, c, R);
                float a = __half2float(in_hi[ai]);
                if (split) a += __half2float(in_lo[ai]);
                acc += a * w[(size_t)c * cout];
            }
        }
        acc += bias[co];
        if (res_hi) {
            acc += __half2float(res_hi[off]);
            if (split) acc += __half2float(res_lo[off]);
        }
        v = activate(acc, act);
    }
    __half h, l;
    split_f16(v, h, l);
    out_hi[off] = h;
    if (split) out_lo[off] = l;
}

}  // namespace sb
