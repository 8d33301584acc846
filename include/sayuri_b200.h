/*
 * sayuri_b200 — C ABI of the Blackwell-native (sm_100a) batched NN evaluation engine that drops in behind
 * Sayuri's NetworkForwardPipe / BatchForwardPipe plugin interface.
 *
 * Every entry point cites the reference interface it replaces (paths under /root/reference).  All
 * arguments are plain pointers and sizes; no C++ or torch types cross this boundary.  Functions return
 * 0 on success or a negative sb_status; sb_last_error() gives the message (the C++ shim turns non-zero
 * into std::runtime_error exactly like ReportCUDAErrors, src/neural/cuda/cuda_common.cc:55-62).
 *
 * There is NO CPU fallback: without a CUDA device (or with the library missing) every compute entry
 * point fails with SB_ERR_CUDA / the loader raises.
 */
#ifndef SAYURI_B200_H_
#define SAYURI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_MAX_BOARD_SIZE 19                 /* src/game/types.h:5-7 (MAX_BOARD_SIZE)            */
#define SB_MAX_INTERSECTIONS 361             /* kNumIntersections                                 */
#define SB_INPUT_CHANNELS 43                 /* src/neural/network_basic.h:10 (kInputChannels)    */
#define SB_PLANE_FLOATS (SB_INPUT_CHANNELS * SB_MAX_INTERSECTIONS) /* InputData::planes size      */

typedef enum sb_status {
    SB_OK = 0,
    SB_ERR_INVALID = -1,     /* bad argument / unsupported network (rejected loudly, never approximated) */
    SB_ERR_CUDA = -2,        /* CUDA runtime / driver error, no device, kernel fault                     */
    SB_ERR_IO = -3,          /* weight file unreadable or malformed                                      */
    SB_ERR_STATE = -4        /* call sequence error (e.g. wait on an idle slot)                          */
} sb_status;

/* Arithmetic of the convolution tower.  (SURVEY.md §7 "precision ladder".) */
typedef enum sb_precision {
    SB_PRECISION_FP32_SPLIT = 0, /* default: fp16 hi+lo operand split, 3 tcgen05 passes, fp32 accumulate:
                                    fp32-faithful, meets the 1e-4 parity bar vs the Eigen forward          */
    SB_PRECISION_FP16 = 1,       /* single fp16 pass, fp32 accumulate: the reference's own --fp16 trade
                                    (config.cc:33); OutputResult.fp16 = true                               */
    SB_PRECISION_SIMT_DEBUG = 2  /* fp32 CUDA-core convolution on the same buffers: on-device cross-check
                                    of the tensor-core kernel, test use only                               */
} sb_precision;

/* Tower block families (BlockBasic::Type, src/neural/description.h:88-132). */
typedef enum sb_block_type {
    SB_BLOCK_RESIDUAL = 0,           /* conv3x3, conv3x3 (+skip)                    blas_forward_pipe.cc:46-88    */
    SB_BLOCK_BOTTLENECK = 1,         /* 1x1 down, 2 x conv3x3, 1x1 up (+skip)       blas_forward_pipe.cc:90-162   */
    SB_BLOCK_NESTED_BOTTLENECK = 2,  /* 1x1 down, 2 inner residual blocks, 1x1 up   blas_forward_pipe.cc:164-263  */
    SB_BLOCK_MIXER = 3               /* depthwise k x k (+skip), 1x1 FFN up, 1x1 FFN down (+skip)       :265-312  */
} sb_block_type;

/* PolicyHeadType (src/neural/description.h:158, loader.cc:245-259) */
typedef enum sb_policy_head_type {
    SB_POLICY_HEAD_NORMAL = 0,
    SB_POLICY_HEAD_REPLK = 1         /* + depthwise k x k and 1x1 after the head-entry conv, blas_forward_pipe.cc:443-471 */
} sb_policy_head_type;

/* Network description == the scalar fields of DNNWeights (src/neural/description.h:164-215). */
typedef struct sb_net_desc {
    int version;            /* 3..5 (43-plane encoder, 5 policy planes, 15 misc outputs)           */
    int input_channels;     /* 43                                                                  */
    int blocks;             /* residual_blocks                                                     */
    int channels;           /* residual_channels                                                   */
    int policy_channels;    /* policy_head_channels                                                */
    int value_channels;     /* value_head_channels                                                 */
    int activation;         /* same ints as enum Activation, src/neural/activation.h:8-17          */
    const int* se_sizes;    /* [blocks]: 0 = no SE unit, >0 = squeeze width of ...-SE               */
    const int* block_types; /* [blocks] sb_block_type, or NULL = all SB_BLOCK_RESIDUAL             */
    const int* inner_channels; /* [blocks] bottleneck_channels of (Nested)Bottleneck blocks / feedforward_channels of
                                  Mixer blocks, or NULL                                                              */
    const int* dw_kernels;  /* [blocks] depthwise filter size of Mixer blocks, or NULL = 7                  */
    int policy_head_type;   /* sb_policy_head_type                                                 */
    int policy_dw_kernel;   /* depthwise filter size of the RepLK head (0 = 7)                     */
} sb_net_desc;

/* One tensor, fp32, caller-owned for the duration of the call only. */
typedef struct sb_tensor {
    const float* data;
    long long count;
} sb_tensor;

/*
 * Weights in the loader's tensor order (src/neural/loader.cc:658-747) AFTER ProcessWeights
 * (loader.cc:775-831): batch-norm already folded, so every layer contributes exactly two tensors
 * {weights, biases}: input conv; per block conv1, conv2 (Bottleneck: pre_btl_conv, conv1, conv2, post_btl_conv;
 * NestedBottleneck: pre_btl_conv, conv1..conv4, post_btl_conv; Mixer: dw_conv [C][k][k], conv1, conv2)
 * [, squeeze FC, excite FC]; policy head conv [, RepLK: p_dw_conv [P][k][k], p_pt_conv],
 * policy intermediate FC, prob conv, pass FC; value head conv, value intermediate FC, ownership conv,
 * misc FC.  Conv weights are UNTRANSFORMED OIHW (ConvLayer::GetWeights, never GetTransformF);
 * FC weights [out][in].
 */
typedef struct sb_weights {
    const sb_tensor* tensors;
    int n_tensors;
} sb_weights;

/* POD mirror of OutputResult (src/neural/network_basic.h:36-63): raw, pre-activation outputs. */
typedef struct sb_output {
    float probabilities[SB_MAX_INTERSECTIONS]; /* logits of the requested policy plane, native n*n order */
    float ownership[SB_MAX_INTERSECTIONS];     /* raw (tanh is applied by Network::TransformResult)      */
    float pass_probability;                    /* pass logit of the requested plane                      */
    float wdl[3];
    float stm_winrate;
    float final_score;
    float q_error;
    float score_error;
    int board_size;
    int offset;                                /* PolicyBufferOffset echoed                              */
    int fp16;                                  /* 1 if SB_PRECISION_FP16 was used                        */
} sb_output;

/* Compact, EXACT encoding of one InputData (src/neural/network_basic.h:23-34) for the host->device hop.  The
 * encoder (src/neural/encoder.cc:80-100,296-319) emits 37 planes of {0,1} and 6 board-constant planes, i.e. every
 * plane takes at most one non-zero value: plane c == scale[c] * bit-mask c.  2.2 KB instead of 62 KB per position. */
#define SB_PACKED_WORDS 12                     /* ceil(361 / 32) bit words per plane                       */
#define SB_PACKED_RAW 1                        /* flags: planes were not two-valued, raw fp32 planes follow */
typedef struct sb_packed_position {
    uint32_t bits[SB_INPUT_CHANNELS][SB_PACKED_WORDS]; /* bit i of plane c <=> planes[c][i] != 0, native n*n order */
    float scale[SB_INPUT_CHANNELS];                    /* the non-zero value of plane c (0 if the plane is empty)  */
    int32_t board_size;
    int32_t offset;                                    /* PolicyBufferOffset                                       */
    int32_t flags;
    int32_t reserved;
} sb_packed_position;                                  /* 2252 bytes */

typedef struct sb_engine sb_engine;

/* ---- lifetime: CudaForwardPipe::Initialize/Construct/Release/Destroy, ------------------------------
 *      src/neural/cuda/cuda_forward_pipe.cc:14-25,44-131; NNGraph::ConstructGraph :133-613            */

/* Build one replica of the net on each listed GPU for an N x N canvas and up to max_batch positions per
 * forward.  `w` may be NULL: the weight blob is then left to be filled through sb_weights_blob()
 * (NCCL broadcast from the rank that parsed the file).  Host tensors are copied; nothing is retained. */
int sb_create(sb_engine** out, const sb_net_desc* desc, const sb_weights* w, const int* gpu_ids, int n_gpus,
              int board_size, int max_batch, int precision);

/* Same, reading the reference weight-file format itself (text or float32bin):
 * DNNLoader::FromFile/Parse/FillWeights/ProcessWeights, src/neural/loader.cc:26-121,628-831. */
int sb_create_from_file(sb_engine** out, const char* weights_path, const int* gpu_ids, int n_gpus,
                        int board_size, int max_batch, int precision);

/* CudaForwardPipe::Construct(option, nullptr) on board/batch change (network.cc:494-498): keeps the
 * weights, re-allocates activations only if the board changed or max_batch grew. */
int sb_reconfigure(sb_engine* e, int board_size, int max_batch);

/* Weight hot-swap (same architecture): replaces the reference's process restart on new weights
 * (src/selfplay/engine.cc:63-90). */
int sb_reload_weights(sb_engine* e, const sb_net_desc* desc, const sb_weights* w);
int sb_reload_weights_from_file(sb_engine* e, const char* weights_path);

void sb_destroy(sb_engine* e);                 /* CudaForwardPipe::Destroy / NNGraph::DestroyGraph :1092-1136 */
const char* sb_last_error(const sb_engine* e); /* e may be NULL: message of the last failed sb_create*        */

int sb_num_gpus(const sb_engine* e);           /* CudaForwardPipe::GetNumWorkers, cuda_forward_pipe.cc:40-42  */
int sb_num_slots(const sb_engine* e);          /* pipeline depth per GPU (independent in-flight batches)      */
int sb_max_batch(const sb_engine* e);
int sb_board_size(const sb_engine* e);
int sb_get_net_desc(const sb_engine* e, sb_net_desc* desc, int* se_sizes, int se_capacity);
int sb_get_block_desc(const sb_engine* e, int* block_types, int* inner_channels, int capacity);
int sb_get_dw_desc(const sb_engine* e, int* dw_kernels, int capacity, int* policy_head_type, int* policy_dw_kernel);

/* ---- the hot path: CudaForwardPipe::BatchForward -> NNGraph::BatchForward, --------------------------
 *      src/neural/cuda/cuda_forward_pipe.cc:32-34,684-1018 (+ FillOutputs :1020-1090);
 *      the canvas re-layout of BatchForwardPipe::SendQueryAndWait (batch_forward_pipe.cc:15-33,48-67)
 *      happens on the device.                                                                          */

/* Blocking forward of n positions on replica `gpu` (index into gpu_ids).  planes[i] points at sample i's
 * InputData::planes: 43 * bs_i * bs_i floats, NCHW, packed at the sample's native board size.
 * One caller per (gpu, slot 0) at a time, like the reference's one worker thread per GPU. */
int sb_forward_batch(sb_engine* e, int gpu, int n, const float* const* planes, const int* board_sizes,
                     const int* policy_offsets, sb_output* out);

/* Asynchronous pair for overlapping H2D / compute / D2H of independent batches (slot < sb_num_slots).
 * planes: n samples `plane_stride` floats apart (SB_PLANE_FLOATS for an array of InputData::planes).
 * If the buffer came from sb_host_alloc() it is DMA'd in place and must stay valid until sb_wait. */
int sb_submit(sb_engine* e, int gpu, int slot, int n, const float* planes, long long plane_stride,
              const int* board_sizes, const int* policy_offsets);
int sb_wait(sb_engine* e, int gpu, int slot, sb_output* out);

/* ---- the batcher: NetworkForwardPipe::Forward for one position from ANY number of threads -----------------
 *      replaces BatchForwardPipe::SendQueryAndWait + Worker/GatherBatches (src/neural/batch_forward_pipe.cc:7-193)
 *      and the per-call host copies of NNGraph::BatchForward (cuda_forward_pipe.cc:694-701).  The calling thread
 *      packs its position straight into a pinned batch record (sb_pack_position), one worker thread per
 *      (GPU, stream) closes a batch when it is full or `wait_us` after its first position arrived (and no earlier
 *      than a stream is free: batches grow while the GPU is busy), runs it and wakes the callers.  Every GPU has
 *      its own lane (ring of batches, mutex, workers); a calling thread is bound to one lane for its lifetime
 *      (thread -> GPU affinity), so nothing on the per-evaluation path is shared between GPUs.                */

/* Exact packing of one position; returns 1 and fills *out, or returns 0 (out->flags = SB_PACKED_RAW) when some
 * plane holds two different non-zero values or a NaN.  Pure host function, thread-safe. */
int sb_pack_position(const float* planes, int board_size, int offset, sb_packed_position* out);
int sb_unpack_position(const sb_packed_position* rec, float* planes);   /* inverse, for tests */

/* Blocking, thread-safe evaluation of ONE position (planes: 43 * bs * bs floats at the native board size).
 * The first call starts the worker threads (2 per GPU, each with its own device slot and stream). */
int sb_eval(sb_engine* e, const float* planes, int board_size, int policy_offset, sb_output* out);

/* The same evaluation split in two for feeders that are NOT one OS thread per leaf: sb_eval_submit claims an entry of
 * the forming batch and packs the position (returns at once), sb_eval_poll returns 1 while the batch has not run and 0
 * once *out is filled (negative sb_status on failure), sb_eval_wait blocks.  A ticket must be collected exactly once (poll
 * until it stops returning 1, or wait); sb_eval == submit + wait.  Precondition shared with sb_eval: sb_destroy must not
 * be called while calls are in flight; sb_reconfigure / sb_reload_weights FAIL the positions that were claimed but not yet
 * evaluated (their callers return SB_ERR_STATE) instead of stranding them. */
typedef struct sb_eval_ticket {
    void* owner;            /* opaque */
    void* batch;            /* opaque */
    uint32_t seq;
    int32_t index;
    int32_t lane;
    int32_t flags;
    int32_t board_size;
    int32_t offset;
} sb_eval_ticket;
int sb_eval_submit(sb_engine* e, const float* planes, int board_size, int policy_offset, sb_eval_ticket* ticket);
int sb_eval_poll(sb_engine* e, sb_eval_ticket* ticket, sb_output* out);
int sb_eval_wait(sb_engine* e, sb_eval_ticket* ticket, sb_output* out);

/* Network::GetOutput(state, Network::kAverage) (src/neural/network.cc:258-282): the 8 symmetric views of one position,
 * each post-processed like Network::TransformResult + ActivatePolicy (network.cc:361-428: inverse symmetry, tanh ownership,
 * softmax wdl, winrates, score x 20, error transforms, policy softmax with `temperature` over board + pass) and averaged.
 * The reference makes 8 serial Forward calls; here the views are built from the identity view `planes` (43 * bs * bs,
 * Encoder::SymmetryPlanes, encoder.cc:80-98) and travel as 8 tickets from the calling thread, i.e. in ONE batch. */
typedef struct sb_symm8_result {
    float probabilities[SB_MAX_INTERSECTIONS]; /* averaged softmax policy, native order, identity orientation */
    float ownership[SB_MAX_INTERSECTIONS];     /* averaged tanh ownership                                    */
    float pass_probability;
    float wdl[3];
    float wdl_winrate;
    float stm_winrate;
    float final_score;
    float q_error;
    float score_error;
    int board_size;
} sb_symm8_result;
int sb_eval_symm8(sb_engine* e, const float* planes, int board_size, int policy_offset, float temperature,
                  sb_symm8_result* out);

/* BatchForwardPipe::SetForwardingSize (batch_size <= max_batch; <= 0 keeps) and the --gpu-waittime analogue in
 * microseconds (< 0 keeps; default 200). */
int sb_batcher_config(sb_engine* e, int batch_size, int wait_us);
/* out[0..5] = batches run, positions evaluated, batches closed full, batches closed by the timer,
 * positions that could not be packed (raw fp32 fallback), worker threads. */
int sb_batcher_stats(sb_engine* e, long long* out6);

/* Measurement helper (bench.py e2e leg, tools/): `threads` native host threads call sb_eval in a loop for
 * `seconds` over n_pos positions (planes: n_pos records SB_PLANE_FLOATS apart, native packing, pageable memory).
 * Returns evaluations per second (wall clock), or a negative sb_status. */
double sb_eval_throughput(sb_engine* e, const float* planes, int n_pos, int board_size, int threads, double seconds);
/* Same through sb_eval_submit / sb_eval_wait: `threads` feeder threads, each with `depth` positions in flight. */
double sb_eval_throughput_async(sb_engine* e, const float* planes, int n_pos, int board_size, int threads, int depth,
                                double seconds);

/* Pinned host memory (the reference's host_input_planes_ / host_output_* buffers, cuda_forward_pipe.cc:560-577). */
void* sb_host_alloc(size_t bytes);
void sb_host_free(void* p);

/* ---- weights on the device: one contiguous blob per replica (pre-packed fp16 hi/lo K-major conv
 *      matrices + fp32 biases/FCs) so that a reload is ONE collective instead of the reference's
 *      per-tensor MallocAndCopy (cuda_common.cc:228-243).                                              */
int sb_weights_blob(sb_engine* e, int gpu, void** device_ptr, size_t* bytes);
/* Device-to-device copies of the blob out of / into replica `gpu` (e.g. a torch.distributed / NCCL
 * broadcast buffer on the same device).  `bytes` must equal the blob size. */
int sb_weights_export(sb_engine* e, int gpu, void* device_dst, size_t bytes);
int sb_weights_import(sb_engine* e, int gpu, const void* device_src, size_t bytes);
uint64_t sb_weights_checksum(sb_engine* e, int gpu);   /* FNV-1a of the blob, to verify replicas agree */
/* Several replicas in ONE process (gpu_ids with more than one entry): sb_create* / sb_reload_weights* upload the blob
 * from host memory ONCE (into replica 0) and broadcast it device-to-device over NVLink: ncclBroadcast on an in-process
 * communicator (libnccl.so.2, opened at run time), cudaMemcpyPeerAsync where NCCL cannot serve the device list; a
 * device-side checksum of every replica is compared afterwards (mismatch = SB_ERR_CUDA).  This replaces the reference's
 * per-GPU host uploads of every tensor (cuda_forward_pipe.cc:85-116,440-552; cuda_common.cc:228-243).
 * sb_weights_broadcast repeats the broadcast + verification from replica 0 (after an sb_weights_import into it).
 * out6 = {host->device blob uploads so far, replicas filled device-to-device so far, method of the last broadcast
 * (0 = single replica, 1 = ncclBroadcast, 2 = peer copy), NCCL version code, replicas verified equal (1/0),
 * duration of the last broadcast in microseconds}. */
int sb_weights_broadcast(sb_engine* e);
int sb_weights_stats(sb_engine* e, long long* out6);

/* ---- measurement (bench.py; device timing on the engine's own stream with CUDA events) ------------- */

/* Re-run the forward of the batch last submitted to (gpu, slot), inputs resident in HBM, `iters` times;
 * ms_each[i] = device time of iteration i.  flush_l2 != 0 writes a buffer larger than L2 between
 * iterations.  conv_ms / conv_launches (optional) receive the device time and the count of the convolution-kernel
 * launches of one forward, measured in extra passes (median of 5) as forward time minus the time of the other kernels,
 * which are bracketed with events (the conv launches keep their back-to-back overlap). */
int sb_time_forward(sb_engine* e, int gpu, int slot, int iters, int flush_l2, float* ms_each,
                    float* conv_ms, int* conv_launches);
long long sb_launch_count(const sb_engine* e);  /* kernels launched by this engine so far */

/* ---- the weight-file reader on its own (host only, no CUDA): DNNLoader::FromFile + ProcessWeights -----------------
 *      (src/neural/loader.cc:26-121,628-831).  Lets a host feed sb_create() from a file it parsed here, and lets the
 *      CPU tests compare this reader with the oracle's, tensor by tensor, bit-exactly.                               */
typedef struct sb_host_net sb_host_net;
int sb_host_net_load(sb_host_net** out, const char* weights_path);          /* message: sb_last_error(NULL) */
void sb_host_net_free(sb_host_net* h);
/* Scalar description; the int arrays (NULL to skip) receive `blocks` entries each (capacity checked). */
int sb_host_net_desc(const sb_host_net* h, sb_net_desc* desc, int* se_sizes, int* block_types, int* inner_channels,
                     int* dw_kernels, int capacity);
/* Tensor `idx` in the order of struct sb_weights, BN folded; returns its element count and sets *data (owned by `h`), or -1
 * past the end. */
long long sb_host_net_tensor(const sb_host_net* h, int idx, const float** data);

/* ---- debugging / tests -------------------------------------------------------------------------- */
/* Tower output of sample `sample` of the last batch on (gpu, slot) as fp32 NCHW [channels][n*n]. */
int sb_debug_read_trunk(sb_engine* e, int gpu, int slot, int sample, float* out);
/* Cycle counters of the LAST conv3x3 launch on (gpu, slot), 8 int64 per CTA: {mma loop total, wait for free TMEM,
 * wait for activation slab, wait for weight stage, epilogue wait for accumulators, epilogue total, items, -}.
 * Filled only while option "stats" = 1.  Returns the number of int64 written (<= capacity) or a negative status. */
int sb_conv_stats(sb_engine* e, int gpu, int slot, long long* out, int capacity);
/* Named integer knobs: "stats" (collect conv kernel cycle counters, default 0). */
int sb_set_option(sb_engine* e, const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* SAYURI_B200_H_ */
