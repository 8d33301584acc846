// B200ForwardPipe: the C++ host side of the drop-in boundary.  Implements the reference's
// CudaForwardPipe interface (src/neural/cuda/cuda_forward_pipe.h:20-36, semantics of
// cuda_forward_pipe.cc:14-131) on top of the sayuri_b200 C ABI (include/sayuri_b200.h).
// Compiled only together with the reference front-end; it is not part of
// libsayuri_b200.so.
#ifdef USE_CUDA

#include "neural/cuda/cuda_forward_pipe.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "sayuri_b200.h"
#include "utils/format.h"
#include "utils/log.h"
#include "utils/option.h"

namespace {

void Check(int rc, const sb_engine* e) {
    // Same convention as ReportCUDAErrors (src/neural/cuda/cuda_common.cc:55-62): throw, the front-end
    // catches at mode level (src/main.cc:17-39).
    if (rc != SB_OK) throw std::runtime_error(std::string("sayuri_b200: ") + sb_last_error(e));
}

void PushConv(std::vector<sb_tensor>& t, ConvLayer& c) {
    // BN is already folded by DNNLoader::ProcessWeights (loader.cc:775-831); never GetTransformF().
    t.push_back({c.GetWeights().data(), (long long)c.GetWeights().size()});
    t.push_back({c.GetBiases().data(), (long long)c.GetBiases().size()});
}
void PushFc(std::vector<sb_tensor>& t, LinearLayer& l) {
    t.push_back({l.GetWeights().data(), (long long)l.GetWeights().size()});
    t.push_back({l.GetBiases().data(), (long long)l.GetBiases().size()});
}

}  // namespace

CudaForwardPipe::~CudaForwardPipe() {
    if (engine_) {
        sb_destroy(engine_);
        engine_ = nullptr;
    }
}

void CudaForwardPipe::Initialize(std::shared_ptr<DNNWeights> weights) {
    LOGGING << "Backend: sayuri_b200 (sm_100a tcgen05/TMA implicit-GEMM engine)\n";
    dump_gpu_info_ = true;
    auto option = ForwardPipeOption::Get()
                      .SetBoardSize(GetOption<int>("defualt_boardsize"))
                      .SetBatchSize(GetOption<int>("batch_size"));
    // SAYURI_B200_REF_BATCHER=1 keeps the reference's own queue + one worker per GPU (SendQueryAndWait / Worker,
    // batch_forward_pipe.cc:7-193) in front of sb_forward_batch, for A/B runs; the default is the engine's batcher.
    const char* env = std::getenv("SAYURI_B200_REF_BATCHER");
    ref_batcher_ = env && env[0] == '1';
    Construct(option, weights);
    // With our batcher no reference worker threads exist; AssignWorkers(0) only creates the (empty) thread group
    // that QuitWorkers joins at shutdown.
    BatchForwardPipe::AssignWorkers(ref_batcher_ ? num_gpus_ : 0);
}

OutputResult CudaForwardPipe::Forward(const InputData& input) {
    if (ref_batcher_) {
        return BatchForwardPipe::SendQueryAndWait(input);
    }
    // NetworkForwardPipe::Forward from any number of search threads (network.cc:179): the calling thread packs its
    // planes (InputData::planes is packed at the native board size, encoder.cc:31-50) straight into a pinned batch
    // record and blocks until its batch has run.
    sb_output o;
    const int offset = input.offset == PolicyBufferOffset::kDefault ? 0 : static_cast<int>(input.offset);
    Check(sb_eval(engine_, input.planes.data(), input.board_size, offset, &o), engine_);
    OutputResult r;
    const int ns = input.board_size * input.board_size;
    std::memcpy(r.probabilities.data(), o.probabilities, sizeof(float) * ns);
    std::memcpy(r.ownership.data(), o.ownership, sizeof(float) * ns);
    r.pass_probability = o.pass_probability;
    r.wdl[0] = o.wdl[0];
    r.wdl[1] = o.wdl[1];
    r.wdl[2] = o.wdl[2];
    r.stm_winrate = o.stm_winrate;
    r.final_score = o.final_score;
    r.q_error = o.q_error;
    r.score_error = o.score_error;
    r.offset = input.offset;
    r.board_size = input.board_size;
    r.komi = input.komi;
    r.fp16 = o.fp16 != 0;
    return r;
}

bool CudaForwardPipe::Valid() const {
    return weights_ != nullptr;
}

int CudaForwardPipe::GetNumWorkers() const {
    return num_gpus_;
}

void CudaForwardPipe::Construct(ForwardPipeOption option, std::shared_ptr<DNNWeights> weights) {
    // cuda_forward_pipe.cc:44-119
    if (weights) {
        weights_ = weights;
    }
    if (weights_ == nullptr) {
        return;   // dummy backend
    }
    int board_size = option.IsValidBoardSize() ? option.board_size : board_size_;
    int batch_size = option.IsValidBatchSize() ? option.batch_size : max_batch_per_nn_;
    board_size = std::max(board_size, GetOption<int>("fixed_nn_boardsize"));
    if (board_size == 0 || batch_size == 0) {
        LOGGING << "NN board size/batch size should be larger than zero.\n";
        return;
    }
    BatchForwardPipe::SetForwardingSize(batch_size);
    if (engine_ && !weights && board_size_ == board_size && batch_size <= max_batch_per_nn_) {
        ConfigureBatcher(batch_size);
        return;   // current engine already supports this configuration
    }
    if (engine_ && !weights) {
        // Reconstruct (network.cc:494-498): keep the weights, re-allocate activations.
        Check(sb_reconfigure(engine_, board_size, batch_size), engine_);
        board_size_ = board_size;
        max_batch_per_nn_ = sb_max_batch(engine_);   // what the engine really allocated (a board change allocates exactly `batch`)
        BatchForwardPipe::SetBoardSize(board_size);
        ConfigureBatcher(batch_size);
        return;
    }
    Release();
    board_size_ = board_size;
    max_batch_per_nn_ = batch_size;
    BatchForwardPipe::SetBoardSize(board_size);

    // user-specified GPUs first, else all devices (cuda_forward_pipe.cc:82-107; the engine validates ids)
    std::vector<int> gpus;
    const int specific = GetOptionCount("gpus");
    for (int i = 0; i < specific; ++i) {
        const int id = GetOption<int>("gpus", i);
        if (id >= 0) gpus.push_back(id);
    }

    DNNWeights& w = *weights_;
    const bool replk = w.policy_head_type == PolicyHeadType::kRepLK;
    std::vector<int> se(w.residual_blocks, 0), types(w.residual_blocks, SB_BLOCK_RESIDUAL), inner(w.residual_blocks, 0),
        dwk(w.residual_blocks, 0);
    std::vector<sb_tensor> t;
    PushConv(t, w.input_conv);
    for (int b = 0; b < w.residual_blocks; ++b) {
        BlockBasic* blk = w.tower[b].get();
        if (blk->IsResidualBlock()) {
            PushConv(t, blk->conv1);
            PushConv(t, blk->conv2);
        } else if (blk->IsBottleneckBlock() || blk->IsNestedBottleneckBlock()) {
            // loader order (loader.cc:416-555): pre 1x1, conv1, conv2 [, conv3, conv4], post 1x1
            types[b] = blk->IsBottleneckBlock() ? SB_BLOCK_BOTTLENECK : SB_BLOCK_NESTED_BOTTLENECK;
            inner[b] = blk->bottleneck_channels;
            PushConv(t, blk->pre_btl_conv);
            PushConv(t, blk->conv1);
            PushConv(t, blk->conv2);
            if (blk->IsNestedBottleneckBlock()) {
                PushConv(t, blk->conv3);
                PushConv(t, blk->conv4);
            }
            PushConv(t, blk->post_btl_conv);
        } else if (blk->IsMixerBlock()) {
            // loader order (loader.cc:556-607): depthwise k x k, ffn1 1x1, ffn2 1x1
            types[b] = SB_BLOCK_MIXER;
            inner[b] = blk->feedforward_channels;
            dwk[b] = blk->dw_conv.GetFilter();
            PushConv(t, blk->dw_conv);
            PushConv(t, blk->conv1);
            PushConv(t, blk->conv2);
        } else {
            throw std::runtime_error("sayuri_b200: unknown tower block type");
        }
        if (blk->apply_se) {
            se[b] = blk->se_size;
            PushFc(t, blk->squeeze);
            PushFc(t, blk->excite);
        }
    }
    PushConv(t, w.p_hd_conv);
    if (replk) {   // loader.cc:691-702
        PushConv(t, w.p_dw_conv);
        PushConv(t, w.p_pt_conv);
    }
    PushFc(t, w.p_inter_fc);
    PushConv(t, w.prob_conv);
    PushFc(t, w.pass_fc);
    PushConv(t, w.v_hd_conv);
    PushFc(t, w.v_inter_fc);
    PushConv(t, w.v_ownership);
    PushFc(t, w.v_misc);

    sb_net_desc d;
    d.version = w.version;
    d.input_channels = w.input_channels;
    d.blocks = w.residual_blocks;
    d.channels = w.residual_channels;
    d.policy_channels = w.policy_head_channels;
    d.value_channels = w.value_head_channels;
    d.activation = static_cast<int>(w.default_act);
    d.se_sizes = se.data();
    d.block_types = types.data();
    d.inner_channels = inner.data();
    d.dw_kernels = dwk.data();
    d.policy_head_type = replk ? SB_POLICY_HEAD_REPLK : SB_POLICY_HEAD_NORMAL;
    d.policy_dw_kernel = replk ? w.p_dw_conv.GetFilter() : 0;
    sb_weights sw{t.data(), (int)t.size()};
    const int precision = GetOption<bool>("fp16") ? SB_PRECISION_FP16 : SB_PRECISION_FP32_SPLIT;
    int rc = sb_create(&engine_, &d, &sw, gpus.empty() ? nullptr : gpus.data(), (int)gpus.size(), board_size_,
                       max_batch_per_nn_, precision);
    if (rc != SB_OK) {
        engine_ = nullptr;
        throw std::runtime_error(std::string("sayuri_b200: ") + sb_last_error(nullptr));
    }
    num_gpus_ = sb_num_gpus(engine_);
    ConfigureBatcher(max_batch_per_nn_);
    if (dump_gpu_info_) {
        LOGGING << Format("sayuri_b200: %d GPU replica(s), NN board %d, max batch %d, precision %s\n", num_gpus_,
                          board_size_, max_batch_per_nn_, precision == SB_PRECISION_FP16 ? "fp16" : "fp32-split");
    }
    dump_gpu_info_ = false;
}

void CudaForwardPipe::ConfigureBatcher(int batch_size) {
    // SetForwardingSize + --gpu-waittime (ms, default 2: config.cc:59).  Our timer runs from the arrival of a batch's
    // first position and a batch never waits for a busy GPU, so a tenth of the reference's wait is used.
    if (engine_) Check(sb_batcher_config(engine_, batch_size, GetOption<int>("gpu_waittime") * 100), engine_);
}

void CudaForwardPipe::Release() {
    if (engine_) {
        sb_destroy(engine_);
        engine_ = nullptr;
    }
}

void CudaForwardPipe::Destroy() {
    BatchForwardPipe::QuitWorkers();
    Release();
}

std::vector<OutputResult> CudaForwardPipe::BatchForward(int gpu, const std::vector<InputData>& inputs) {
    // Called by exactly one worker thread per GPU (batch_forward_pipe.cc:93-96).
    const int n = static_cast<int>(inputs.size());
    const int N = board_size_, NS = N * N;
    std::vector<const float*> planes(n);
    std::vector<int> sizes(n), offsets(n);
    // SendQueryAndWait (batch_forward_pipe.cc:15-33) has already re-laid smaller boards onto the N x N
    // canvas on the host; the C ABI takes native packed planes and does the placement on the device, so
    // undo that host re-layout for the (rare) mixed-size samples.
    std::vector<std::vector<float>> repacked;
    repacked.reserve(n);
    for (int i = 0; i < n; ++i) {
        const InputData& in = inputs[i];
        sizes[i] = in.board_size;
        offsets[i] = in.offset == PolicyBufferOffset::kDefault ? 0 : static_cast<int>(in.offset);
        if (in.board_size == N) {
            planes[i] = in.planes.data();
        } else {
            const int bs = in.board_size;
            repacked.emplace_back((size_t)kInputChannels * bs * bs);
            std::vector<float>& dst = repacked.back();
            for (int c = 0; c < kInputChannels; ++c)
                for (int y = 0; y < bs; ++y)
                    std::memcpy(&dst[((size_t)c * bs + y) * bs], &in.planes[(size_t)c * NS + (size_t)y * N], sizeof(float) * bs);
            planes[i] = dst.data();
        }
    }
    std::vector<sb_output> raw(n);
    Check(sb_forward_batch(engine_, gpu, n, planes.data(), sizes.data(), offsets.data(), raw.data()), engine_);
    std::vector<OutputResult> results(n);
    for (int i = 0; i < n; ++i) {
        OutputResult& r = results[i];
        const sb_output& o = raw[i];
        const int bs = sizes[i];
        if (bs == N) {
            std::memcpy(r.probabilities.data(), o.probabilities, sizeof(float) * NS);
            std::memcpy(r.ownership.data(), o.ownership, sizeof(float) * NS);
        } else {
            // back to canvas order; SendQueryAndWait crops it again (batch_forward_pipe.cc:48-67)
            for (int y = 0; y < bs; ++y)
                for (int x = 0; x < bs; ++x) {
                    r.probabilities[(size_t)y * N + x] = o.probabilities[y * bs + x];
                    r.ownership[(size_t)y * N + x] = o.ownership[y * bs + x];
                }
        }
        r.pass_probability = o.pass_probability;
        r.wdl[0] = o.wdl[0];
        r.wdl[1] = o.wdl[1];
        r.wdl[2] = o.wdl[2];
        r.stm_winrate = o.stm_winrate;
        r.final_score = o.final_score;
        r.q_error = o.q_error;
        r.score_error = o.score_error;
        r.offset = inputs[i].offset;
        r.board_size = bs;
        r.komi = inputs[i].komi;
        r.fp16 = o.fp16 != 0;
    }
    return results;
}

#endif
