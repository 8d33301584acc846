// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Thin extern "C" harness over the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libsayuri_ref_*.so).  It lets tests/ and
// bench.py's cpu_baseline / --impl reference leg call the reference's own loader and Eigen forward
// at full fp32 precision (no 6-decimal GTP dump):
//   DNNLoader::FromFile            /root/reference/src/neural/loader.cc:26-65
//   BlasForwardPipe::Forward       /root/reference/src/neural/blas/blas_forward_pipe.cc:314-563
//   Encoder::GetInputs             /root/reference/src/neural/encoder.cc:14-50
// Nothing in the product path (sayuri_b200/) may link or call this.
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "config.h"
#include "game/symmetry.h"
#include "neural/blas/blas_forward_pipe.h"
#include "neural/loader.h"
#include "neural/network_basic.h"
#include "utils/option.h"

namespace {
std::shared_ptr<DNNWeights> g_weights;
std::unique_ptr<BlasForwardPipe> g_pipe;
bool g_args_ready = false;

void EnsureArgs() {
    if (g_args_ready) return;
    // ArgsParser initialises the global option map, Zobrist, symmetry tables (config.cc:336-381).
    static char a0[] = "ref_harness";
    static char a1[] = "--quiet";
    static char a2[] = "-t";
    static char a3[] = "1";
    char* argv[] = {a0, a1, a2, a3};
    ArgsParser(4, argv);
    g_args_ready = true;
}
}  // namespace

extern "C" {

// Load a weights file with the reference loader.  winograd: 0 => Convolution<3> im2col path
// (primary oracle), 1 => WinogradConvolution3 (reference default, config.cc:32).  Returns 0 on success.
int ref_init(const char* weights_path, int winograd) {
    try {
        EnsureArgs();
        SetOption("winograd", (bool)winograd);
        g_weights = std::make_shared<DNNWeights>();
        DNNLoader::Get().FromFile(g_weights, std::string(weights_path));
        if (!g_weights->loaded) {
            g_weights.reset();
            return -1;
        }
        g_pipe = std::make_unique<BlasForwardPipe>();
        g_pipe->Initialize(g_weights);
        return 0;
    } catch (...) {
        return -2;
    }
}

// Net description: {version, input_channels, blocks, channels, P, V, activation, n_se_blocks}
int ref_net_info(int* out8) {
    if (!g_weights) return -1;
    out8[0] = g_weights->version;
    out8[1] = g_weights->input_channels;
    out8[2] = g_weights->residual_blocks;
    out8[3] = g_weights->residual_channels;
    out8[4] = g_weights->policy_head_channels;
    out8[5] = g_weights->value_head_channels;
    out8[6] = (int)g_weights->default_act;
    int nse = 0;
    for (auto& b : g_weights->tower) nse += b->apply_se ? 1 : 0;
    out8[7] = nse;
    return 0;
}

// One evaluation through BlasForwardPipe::Forward.  planes: 43*bs*bs floats, NCHW at native size.
// out layout (floats): prob[bs*bs] | own[bs*bs] | pass, wdl0, wdl1, wdl2, stm, score, q_err, score_err
int ref_forward(const float* planes, int board_size, float komi, int offset, float* out) {
    if (!g_pipe) return -1;
    InputData in;
    in.board_size = board_size;
    in.komi = komi;
    in.side_to_move = kBlack;
    in.offset = (PolicyBufferOffset)offset;
    const int s = board_size * board_size;
    std::memcpy(in.planes.data(), planes, sizeof(float) * kInputChannels * s);
    OutputResult r = g_pipe->Forward(in);
    std::memcpy(out, r.probabilities.data(), sizeof(float) * s);
    std::memcpy(out + s, r.ownership.data(), sizeof(float) * s);
    float* m = out + 2 * s;
    m[0] = r.pass_probability;
    m[1] = r.wdl[0];
    m[2] = r.wdl[1];
    m[3] = r.wdl[2];
    m[4] = r.stm_winrate;
    m[5] = r.final_score;
    m[6] = r.q_error;
    m[7] = r.score_error;
    return 0;
}

// Throughput of the reference Eigen forward: `threads` host threads each looping
// BlasForwardPipe::Forward (read-only weights, thread-local buffers: blas_forward_pipe.cc:340-371)
// over `n_pos` positions round-robin for at least `seconds`.  Returns total evals; *elapsed_s set.
long ref_time_forward(const float* planes, int n_pos, int board_size, int threads, double seconds,
                      double* elapsed_s) {
    if (!g_pipe) return -1;
    const int s = board_size * board_size;
    std::vector<InputData> inputs(n_pos);
    for (int i = 0; i < n_pos; ++i) {
        inputs[i].board_size = board_size;
        inputs[i].komi = 7.5f;
        inputs[i].side_to_move = kBlack;
        inputs[i].offset = PolicyBufferOffset::kNormal;
        std::memcpy(inputs[i].planes.data(), planes + (size_t)i * kInputChannels * s,
                    sizeof(float) * kInputChannels * s);
    }
    std::atomic<long> total{0};
    std::atomic<bool> stop{false};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            long n = 0;
            int i = t % n_pos;
            volatile float sink = 0.f;
            while (!stop.load(std::memory_order_relaxed)) {
                OutputResult r = g_pipe->Forward(inputs[i]);
                sink = sink + r.pass_probability;
                i = (i + 1) % n_pos;
                ++n;
            }
            total += n;
        });
    }
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    stop.store(true);
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    return total.load();
}

// Symmetry::TransformIndex (game/symmetry.h:35-37, tables built by symmetry.cc:97-123) for every cell of an N x N board:
// the gather index of Encoder::SymmetryPlanes (encoder.cc:80-100) and the scatter index of Network::TransformResult
// (network.cc:376-383).  Pins the oracle's index tables to the reference, bit-exactly.
int ref_symmetry_table(int board_size, int symmetry, int* out) {
    EnsureArgs();
    if (board_size < 2 || board_size > 19 || symmetry < 0 || symmetry >= Symmetry::kNumSymmetris) return -1;
    for (int i = 0; i < board_size * board_size; ++i) out[i] = Symmetry::Get().TransformIndex(board_size, symmetry, i);
    return 0;
}

}  // extern "C"
