// Host-side network model and weight-file reader of the engine (product code, no dependency on
// oracle/ or on the reference tree).  The file format and the batch-norm folding follow
//   DNNLoader::Parse / CheckMisc / FillWeights / ProcessWeights   /root/reference/src/neural/loader.cc:67-121,190-356,628-831
//   BatchNormLayer::LoadStddevs                                   /root/reference/src/neural/description.cc:70-85
// Tower blocks: ResidualBlock, BottleneckBlock, NestedBottleneckBlock, each optionally with an SE unit
// (loader.cc:385-624).  Architectures outside this engine's scope (Mixer blocks, RepLK policy head, 38-plane
// v1/v2 nets) are REJECTED with a message, never approximated.
#pragma once
#include <string>
#include <vector>

#include "../../include/sayuri_b200.h"

namespace sb {

struct HostConv {
    int in = 0, out = 0, k = 0;
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
    // NestedBottleneck {pre 1x1, conv1, conv2, conv3, conv4, post 1x1}
    std::vector<HostConv> convs;
    int se_size = 0;  // 0 = no SE
    HostFC squeeze, excite;
    static int NumConvs(int type) { return type == SB_BLOCK_RESIDUAL ? 2 : type == SB_BLOCK_BOTTLENECK ? 4 : 6; }
};
struct HostNet {
    int version = 0, input_channels = 0, blocks = 0, channels = 0, P = 0, V = 0, act = 0;
    HostConv input_conv;
    std::vector<HostBlock> tower;
    HostConv p_hd_conv;
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
};

// Returns true on success; on failure `err` explains.
bool LoadWeightsFile(const std::string& path, HostNet& net, std::string& err);
bool NetFromAbi(const sb_net_desc* desc, const sb_weights* w, HostNet& net, std::string& err);
bool ValidateNet(const HostNet& net, std::string& err);

}  // namespace sb
