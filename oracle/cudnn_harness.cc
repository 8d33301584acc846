// TEST / MEASUREMENT INFRASTRUCTURE — NOT PRODUCT CODE.
// Drives the UNMODIFIED reference GPU backend (CudaForwardPipe: cuDNN + cuBLAS + its own SIMT kernels,
// /root/reference/src/neural/cuda/cuda_forward_pipe.cc:14-119,684-1090) directly through its public
// BatchForward(gpu, inputs), bypassing GTP/MCTS/encoder, so that its forward throughput on this GPU can be put
// beside ours at the same batch sizes (BASELINE.json config 5: "vs reference cuDNN backend").  Host InputData in,
// host OutputResult out: the reference's own H2D/D2H and host re-layout are inside the timed region, like our e2e leg.
//
//   sayuri_cudnn_bench <weights> <planes.bin> <n_pos> <fp16 0|1> <seconds> <out.bin|-> <batch> [<batch> ...]
//
// planes.bin: n_pos x 43 x 361 float32 (19x19).  out.bin (optional): raw outputs of the first min(n_pos, batch)
// positions of the LAST batch size: prob[361] | own[361] | pass, wdl0..2, stm, score, q_err, score_err.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "config.h"
#include "neural/cuda/cuda_forward_pipe.h"
#include "neural/loader.h"
#include "neural/network_basic.h"
#include "utils/option.h"

int main(int argc, char** argv) {
    if (argc < 8) {
        std::fprintf(stderr, "usage: %s weights planes.bin n_pos fp16 seconds out.bin|- batch...\n", argv[0]);
        return 2;
    }
    const std::string wpath = argv[1];
    const int n_pos = std::atoi(argv[3]);
    const bool fp16 = std::atoi(argv[4]) != 0;
    const double seconds = std::atof(argv[5]);
    const std::string out_path = argv[6];
    std::vector<int> batches;
    for (int i = 7; i < argc; ++i) batches.push_back(std::atoi(argv[i]));
    int max_batch = 1;
    for (int b : batches) max_batch = b > max_batch ? b : max_batch;

    const int S = 361, PL = kInputChannels * S;
    std::vector<float> planes((size_t)n_pos * PL);
    {
        FILE* f = std::fopen(argv[2], "rb");
        if (!f || std::fread(planes.data(), sizeof(float), planes.size(), f) != planes.size()) {
            std::fprintf(stderr, "cannot read %s\n", argv[2]);
            return 2;
        }
        std::fclose(f);
    }

    // ArgsParser fills the global option map (config.cc:135-142,336-381).
    std::string bs = std::to_string(max_batch);
    std::vector<std::string> args = {"cudnn_bench", "--quiet", "-t", "1", "-g", "0", "-b", bs};
    if (!fp16) args.push_back("--no-fp16");
    std::vector<char*> av;
    for (auto& s : args) av.push_back(const_cast<char*>(s.c_str()));
    ArgsParser(static_cast<int>(av.size()), av.data());
    SetOption("batch_size", max_batch);

    auto weights = std::make_shared<DNNWeights>();
    DNNLoader::Get().FromFile(weights, wpath);
    if (!weights->loaded) {
        std::fprintf(stderr, "weights not loaded\n");
        return 3;
    }
    CudaForwardPipe pipe;
    pipe.Initialize(weights);

    std::vector<OutputResult> last;
    for (int B : batches) {
        std::vector<InputData> inputs(B);
        for (int i = 0; i < B; ++i) {
            inputs[i].board_size = 19;
            inputs[i].komi = 7.5f;
            inputs[i].side_to_move = kBlack;
            inputs[i].offset = PolicyBufferOffset::kNormal;
            std::memcpy(inputs[i].planes.data(), planes.data() + (size_t)(i % n_pos) * PL, sizeof(float) * PL);
        }
        for (int w = 0; w < 3; ++w) last = pipe.BatchForward(0, inputs);
        long iters = 0;
        auto t0 = std::chrono::steady_clock::now();
        double el = 0;
        do {
            last = pipe.BatchForward(0, inputs);
            ++iters;
            el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        } while (el < seconds);
        std::printf("ref-cudnn fp16=%d batch=%d iters=%ld ms_per_forward=%.4f evals_per_s=%.1f\n", (int)fp16, B, iters,
                    1e3 * el / iters, (double)iters * B / el);
        std::fflush(stdout);
    }
    if (out_path != "-") {
        FILE* f = std::fopen(out_path.c_str(), "wb");
        for (auto& r : last) {
            std::fwrite(r.probabilities.data(), sizeof(float), S, f);
            std::fwrite(r.ownership.data(), sizeof(float), S, f);
            float m[8] = {r.pass_probability, r.wdl[0], r.wdl[1], r.wdl[2],
                          r.stm_winrate,      r.final_score, r.q_error, r.score_error};
            std::fwrite(m, sizeof(float), 8, f);
        }
        std::fclose(f);
    }
    pipe.Destroy();
    return 0;
}
