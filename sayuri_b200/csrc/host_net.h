// Host-side network model and weight-file reader of the engine (product code, no dependency on
// oracle/ or on the reference tree).  The file format and the batch-norm folding follow
//   DNNLoader::Parse / CheckMisc / FillWeights / ProcessWeights   /root/reference/src/neural/loader.cc:67-121,190-356,628-831
//   BatchNormLayer::LoadStddevs                                   /root/reference/src/neural/description.cc:70-85
// Tower blocks: ResidualBlock, BottleneckBlock, NestedBottleneckBlock, MixerBlock, each optionally with an SE unit
// (loader.cc:385-624); policy heads Normal and RepLK (loader.cc:684-729).  38-plane v1/v2 nets are REJECTED with a
// message, never approximated.
#pragma once
#include <string>
#include <vector>

#include "../../include/sayuri_b200.h"

namespace sb {

struct HostConv {
    int in = 0, out = 0, k = 0;
    bool depthwise = false;  // DepthwiseConvolution: in == 1, w = [out][k][k] (convolution.cc:27-62)
    std::vector<float> w;  // OIHW
    std::vector<float> b;
};
struct HostFC {
    int in = 0, out = 0;
    std::vector<float> w;  // [out][in]
    std::vector<float> b;
};
struct HostBlock {
    int type = SB_BLOCK_RESIDUAL;
    int inner = 0;                 // bottleneck_channels (0 for a plain residual block)
    // loader order: Residual {conv1, conv2}; Bottleneck {pre 1x1, conv1, conv2, post 1x1};
    // NestedBottleneck {pre 1x1, conv1, conv2, conv3, conv4, post 1x1}; Mixer {depthwise k x k, ffn1 1x1, ffn2 1x1}
    std::vector<HostConv> convs;
    int se_size = 0;  // 0 = no SE
    HostFC squeeze, excite;
    static int NumConvs(int type) {
        return type == SB_BLOCK_RESIDUAL ? 2 : type == SB_BLOCK_BOTTLENECK ? 4 : type == SB_BLOCK_NESTED_BOTTLENECK ? 6 : 3;
    }
    int dw_kernel() const { return type == SB_BLOCK_MIXER && !convs.empty() ? convs[0].k : 0; }
};
struct HostNet {
    int version = 0, input_channels = 0, blocks = 0, channels = 0, P = 0, V = 0, act = 0;
    HostConv input_conv;
    std::vector<HostBlock> tower;
    HostConv p_hd_conv;
    bool replk = false;            // PolicyHeadType RepLK: depthwise k x k + 1x1 after the head-entry conv
    HostConv p_dw_conv, p_pt_conv;
    HostFC p_inter_fc;
    HostConv prob_conv;
    HostFC pass_fc;
    HostConv v_hd_conv;
    HostFC v_inter_fc;
    HostConv v_ownership;
    HostFC v_misc;
    std::vector<int> se_sizes() const {
        std::vector<int> s;
        for (auto& b : tower) s.push_back(b.se_size);
        return s;
    }
    std::vector<int> block_types() const {
        std::vector<int> s;
        for (auto& b : tower) s.push_back(b.type);
        return s;
    }
    std::vector<int> inner_channels() const {
        std::vector<int> s;
        for (auto& b : tower) s.push_back(b.inner);
        return s;
    }
    std::vector<int> dw_kernels() const {
        std::vector<int> s;
        for (auto& b : tower) s.push_back(b.dw_kernel());
        return s;
    }
};

// Returns true on success; on failure `err` explains.
bool LoadWeightsFile(const std::string& path, HostNet& net, std::string& err);
bool NetFromAbi(const sb_net_desc* desc, const sb_weights* w, HostNet& net, std::string& err);
bool ValidateNet(const HostNet& net, std::string& err);

}  // namespace sb
