// Include-path shadow of the reference header `neural/cuda/cuda_forward_pipe.h`
// (/root/reference/src/neural/cuda/cuda_forward_pipe.h:20-36).  Compiling the UNMODIFIED reference front-end
// with  -DUSE_CUDA -I<this shim dir> -I<reference>/src  makes Network (src/neural/network.cc:3-5,61-67) pick
// this class as its backend: same name, same virtual interface, implemented over the sayuri_b200 C ABI.
#pragma once

#ifdef USE_CUDA

#include <memory>
#include <vector>

#include "neural/batch_forward_pipe.h"
#include "neural/description.h"
#include "neural/network_basic.h"

struct sb_engine;

class CudaForwardPipe : public BatchForwardPipe {
public:
    virtual void Initialize(std::shared_ptr<DNNWeights> weights);

    virtual OutputResult Forward(const InputData& input);

    virtual bool Valid() const;

    virtual void Construct(ForwardPipeOption option, std::shared_ptr<DNNWeights> weights);

    virtual void Release();

    virtual void Destroy();

    virtual int GetNumWorkers() const;

    virtual std::vector<OutputResult> BatchForward(int gpu, const std::vector<InputData>& inputs);

    virtual ~CudaForwardPipe();

private:
    void ConfigureBatcher(int batch_size);

    sb_engine* engine_{nullptr};
    bool ref_batcher_{false};
    bool dump_gpu_info_{true};
    int max_batch_per_nn_{0};
    int board_size_{0};
    int num_gpus_{0};
};

using B200ForwardPipe = CudaForwardPipe;

#endif
