#include "host_net.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace sb {
namespace {

struct Shape {
    char kind;  // 'C' Convolution, 'D' DepthwiseConvolution, 'B' BatchNorm, 'F' FullyConnect
    int d[3];
};

// Cursor over the whole file held in memory: header lines are text, the parameter section is either
// one text line per tensor or a raw little-endian float32 stream with 0xFFFFFFFF terminators
// (loader.cc:833-898, utils/parse_float.cc:5-33).
class Reader {
public:
    explicit Reader(std::string data) : buf_(std::move(data)) {}
    bool binary = false;

    bool Line(std::string& out) {
        if (pos_ >= buf_.size()) return false;
        size_t e = buf_.find('\n', pos_);
        if (e == std::string::npos) e = buf_.size();
        out.assign(buf_, pos_, e - pos_);
        pos_ = e + 1;
        return true;
    }

    std::vector<float> Tensor(size_t expect) {
        std::vector<float> t;
        t.reserve(expect);
        if (binary) {
            for (;;) {
                if (pos_ + 4 > buf_.size()) throw std::runtime_error("truncated binary tensor");
                uint32_t u;
                std::memcpy(&u, buf_.data() + pos_, 4);
                pos_ += 4;
                if (u == 0xffffffffu) break;
                float f;
                std::memcpy(&f, &u, 4);
                t.push_back(f);
            }
        } else {
            std::string line;
            if (!Line(line)) throw std::runtime_error("missing tensor line");
            const char* q = line.c_str();
            for (;;) {
                char* end = nullptr;
                const double v = std::strtod(q, &end);  // the reference parses text via double
                if (end == q) break;
                t.push_back(static_cast<float>(v));
                q = end;
            }
        }
        if (t.size() != expect) {
            throw std::runtime_error("tensor size mismatch: expect " + std::to_string(expect) + " but got " +
                                     std::to_string(t.size()));
        }
        return t;
    }

private:
    std::string buf_;
    size_t pos_ = 0;
};

std::vector<std::string> Words(const std::string& line) {
    std::istringstream is(line);
    std::vector<std::string> w;
    std::string s;
    while (is >> s) w.push_back(s);
    return w;
}

std::string Lower(std::string s) {
    for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    return s;
}

int ActFromName(const std::string& name) {  // activation.h:19-41
    static const char* names[] = {"identity", "relu", "elu", "selu", "gelu", "mish", "swish", "hardswish"};
    const std::string l = Lower(name);
    for (int i = 0; i < 8; ++i)
        if (l == names[i]) return i;
    throw std::runtime_error("Unknown activation type.");
}

void ReadConv(Reader& r, const Shape& s, HostConv& c) {
    if (s.kind != 'C' && s.kind != 'D') throw std::runtime_error("expected a Convolution layer in the struct list");
    c.in = s.d[0];
    c.out = s.d[1];
    c.k = s.d[2];
    c.depthwise = s.kind == 'D';
    if (c.depthwise) {   // "DepthwiseConvolution 1 C k" (network.py writer)
        if (c.in != 1 && c.in != c.out) throw std::runtime_error("depthwise convolution shape is wrong");
        c.in = 1;
    }
    c.w = r.Tensor(static_cast<size_t>(c.in) * c.out * c.k * c.k);
    c.b = r.Tensor(static_cast<size_t>(c.out));
}

// Conv followed by BatchNorm, folded on the spot: scale = 1/std (v>=2) or 1/sqrt(var+1e-5) (v1);
// bias = (bias - mean) * scale; W[o,...] *= scale[o].   (loader.cc:776-789)
void ReadConvBn(Reader& r, const Shape* s, HostConv& c, bool v1) {
    ReadConv(r, s[0], c);
    if (s[1].kind != 'B' || s[1].d[0] != c.out) throw std::runtime_error("expected BatchNorm after Convolution");
    std::vector<float> mean = r.Tensor(static_cast<size_t>(c.out));
    std::vector<float> sd = r.Tensor(static_cast<size_t>(c.out));
    const size_t stride = static_cast<size_t>(c.in) * c.k * c.k;   // k*k for a depthwise layer (in == 1)
    for (int o = 0; o < c.out; ++o) {
        const float scale = v1 ? 1.0f / std::sqrt(sd[o] + 1e-5f) : 1.0f / sd[o];
        c.b[o] -= mean[o];
        for (size_t k = 0; k < stride; ++k) c.w[stride * o + k] *= scale;
        c.b[o] *= scale;
    }
}

void ReadFC(Reader& r, const Shape& s, HostFC& f) {
    if (s.kind != 'F') throw std::runtime_error("expected a FullyConnect layer in the struct list");
    f.in = s.d[0];
    f.out = s.d[1];
    f.w = r.Tensor(static_cast<size_t>(f.in) * f.out);
    f.b = r.Tensor(static_cast<size_t>(f.out));
}

void Parse(Reader& r, HostNet& net) {
    std::string line;
    if (!r.Line(line)) throw std::runtime_error("weights file is empty");
    {
        auto w = Words(line);
        if (w.size() < 2 || w[0] != "get" || w[1] != "main") throw std::runtime_error("weights file format is not acceptable");
    }
    std::map<std::string, std::string> info;
    std::vector<std::string> stack;
    std::vector<Shape> shapes;
    while (r.Line(line)) {
        auto w = Words(line);
        if (w.size() < 2 || w[0] != "get") continue;
        if (w[1] == "info") {
            while (r.Line(line)) {
                auto kv = Words(line);
                if (kv.empty() || kv[0][0] == '#') continue;
                if (kv[0] == "end") break;
                if (kv.size() >= 2) info[kv[0]] = kv[1];
            }
        } else if (w[1] == "stack") {
            while (r.Line(line)) {
                auto kv = Words(line);
                if (kv.empty() || kv[0][0] == '#') continue;
                if (kv[0] == "end") break;
                stack.push_back(kv[0]);
            }
        } else if (w[1] == "struct") {
            while (r.Line(line)) {
                auto kv = Words(line);
                if (kv.empty() || kv[0][0] == '#') continue;
                if (kv[0] == "end") break;
                Shape s{};
                const size_t nd = kv.size() - 1;
                for (size_t i = 0; i < nd && i < 3; ++i) s.d[i] = std::stoi(kv[i + 1]);
                if (kv[0] == "Convolution" && nd == 3) s.kind = 'C';
                else if (kv[0] == "DepthwiseConvolution" && nd == 3) s.kind = 'D';
                else if (kv[0] == "BatchNorm" && nd == 1) s.kind = 'B';
                else if (kv[0] == "FullyConnect" && nd == 2) s.kind = 'F';
                else throw std::runtime_error("layer shape is error");
                shapes.push_back(s);
            }
        } else if (w[1] == "parameters") {
            break;
        }
    }
    auto get = [&](const char* k) -> std::string {
        auto it = info.find(k);
        return it == info.end() ? std::string() : it->second;
    };
    net.version = get("Version").empty() ? 1 : std::stoi(get("Version"));
    r.binary = get("FloatType") == "float32bin";
    if (net.version >= 6) throw std::runtime_error("do not support this version");
    if (net.version < 3) throw std::runtime_error("v1/v2 networks (38 input planes) are not supported by sayuri_b200");
    net.input_channels = SB_INPUT_CHANNELS;
    if (get("InputChannels").empty() || std::stoi(get("InputChannels")) != net.input_channels)
        throw std::runtime_error("the number of input channels is wrong");
    if (!get("PolicyHeadType").empty()) {   // loader.cc:245-259
        const std::string t = Lower(get("PolicyHeadType"));
        if (t == "replk") net.replk = true;
        else if (t != "normal") throw std::runtime_error("policy head type '" + get("PolicyHeadType") + "' is not supported by sayuri_b200");
    }
    net.act = get("ActivationFunction").empty() ? 1 /* relu, loader.cc:261-265 */ : ActFromName(get("ActivationFunction"));
    if (get("ResidualBlocks").empty() || get("ResidualChannels").empty()) throw std::runtime_error("missing ResidualBlocks/ResidualChannels");
    net.blocks = std::stoi(get("ResidualBlocks"));
    net.channels = std::stoi(get("ResidualChannels"));
    const std::string p = net.version >= 5 ? get("PolicyHeadChannels") : get("PolicyExtract");
    const std::string v = net.version >= 5 ? get("ValueHeadChannels") : get("ValueExtract");
    if (p.empty() || v.empty()) throw std::runtime_error("missing policy/value head channels");
    net.P = std::stoi(p);
    net.V = std::stoi(v);

    if (stack.empty()) {  // loader.cc:270-292: ResidualBlock[-SE] inferred from the struct list
        size_t inner = 0;
        for (int b = 0; b < net.blocks; ++b) {
            std::string t = "ResidualBlock";
            inner += 4;
            if (inner + 2 < shapes.size() && shapes[inner + 2].kind == 'F') {
                t += "-SE";
                inner += 2;
            }
            stack.push_back(t);
        }
    }
    if (static_cast<int>(stack.size()) != net.blocks) throw std::runtime_error("stack size does not match ResidualBlocks");

    const bool v1 = net.version == 1;
    size_t off = 0;
    auto need = [&](size_t n) {
        if (off + n > shapes.size()) throw std::runtime_error("struct list is too short");
    };
    need(2);
    ReadConvBn(r, &shapes[off], net.input_conv, v1);
    off += 2;
    net.tower.resize(static_cast<size_t>(net.blocks));
    for (int b = 0; b < net.blocks; ++b) {
        std::string name = stack[static_cast<size_t>(b)];
        bool se = false;
        const size_t dash = name.find('-');
        if (dash != std::string::npos) {
            const std::string comp = name.substr(dash + 1);
            if (comp != "SE") throw std::runtime_error("block component '" + comp + "' is not supported by sayuri_b200");
            se = true;
            name = name.substr(0, dash);
        }
        HostBlock& blk = net.tower[static_cast<size_t>(b)];
        if (name == "ResidualBlock") blk.type = SB_BLOCK_RESIDUAL;
        else if (name == "BottleneckBlock") blk.type = SB_BLOCK_BOTTLENECK;
        else if (name == "NestedBottleneckBlock") blk.type = SB_BLOCK_NESTED_BOTTLENECK;
        else if (name == "MixerBlock") blk.type = SB_BLOCK_MIXER;
        else
            throw std::runtime_error("block type '" + name + "' is not supported by sayuri_b200");
        const int nc = HostBlock::NumConvs(blk.type);
        need(static_cast<size_t>(2 * nc));
        blk.convs.resize(static_cast<size_t>(nc));
        for (int q = 0; q < nc; ++q) {
            ReadConvBn(r, &shapes[off], blk.convs[static_cast<size_t>(q)], v1);
            off += 2;
        }
        blk.inner = blk.type == SB_BLOCK_RESIDUAL ? 0 : blk.type == SB_BLOCK_MIXER ? blk.convs[1].out : blk.convs[0].out;
        if (se) {
            need(2);
            ReadFC(r, shapes[off++], blk.squeeze);
            ReadFC(r, shapes[off++], blk.excite);
            blk.se_size = blk.squeeze.out;
        }
    }
    need(net.replk ? 14 : 10);
    ReadConvBn(r, &shapes[off], net.p_hd_conv, v1);
    off += 2;
    if (net.replk) {   // loader.cc:691-702
        ReadConvBn(r, &shapes[off], net.p_dw_conv, v1);
        off += 2;
        ReadConvBn(r, &shapes[off], net.p_pt_conv, v1);
        off += 2;
    }
    ReadFC(r, shapes[off++], net.p_inter_fc);
    ReadConv(r, shapes[off++], net.prob_conv);
    ReadFC(r, shapes[off++], net.pass_fc);
    ReadConvBn(r, &shapes[off], net.v_hd_conv, v1);
    off += 2;
    ReadFC(r, shapes[off++], net.v_inter_fc);
    ReadConv(r, shapes[off++], net.v_ownership);
    ReadFC(r, shapes[off++], net.v_misc);
    if (off != shapes.size()) throw std::runtime_error("struct list has unexpected extra layers");
    if (!r.Line(line) || Words(line).empty() || Words(line)[0] != "end")
        throw std::runtime_error("weights file format is not acceptable");
}

}  // namespace

bool ValidateNet(const HostNet& n, std::string& err) {
    auto conv_ok = [](const HostConv& c, int in, int out, int k) {
        return !c.depthwise && c.in == in && c.out == out && c.k == k && c.w.size() == static_cast<size_t>(in) * out * k * k &&
               c.b.size() == static_cast<size_t>(out);
    };
    auto dw_ok = [](const HostConv& c, int ch) {
        return c.depthwise && c.in == 1 && c.out == ch && c.k >= 3 && c.k <= 15 && (c.k & 1) &&
               c.w.size() == static_cast<size_t>(ch) * c.k * c.k && c.b.size() == static_cast<size_t>(ch);
    };
    // inner widths (bottleneck / feed-forward): any multiple of 8 up to 512 — weight rows are zero-padded to the UMMA
    // N granule (16) and the epilogue stores only the real channels; the tower width itself stays a multiple of 16 <= 256
    auto width_ok = [](int w) { return w >= 8 && w <= 512 && w % 8 == 0; };
    auto fc_ok = [](const HostFC& f, int in, int out) {
        return f.in == in && f.out == out && f.w.size() == static_cast<size_t>(in) * out && f.b.size() == static_cast<size_t>(out);
    };
    const int C = n.channels, P = n.P, V = n.V;
    if (n.version < 3 || n.version > 5) { err = "unsupported network version"; return false; }
    if (n.input_channels != SB_INPUT_CHANNELS) { err = "the number of input channels is wrong"; return false; }
    if (n.act < 0 || n.act > 7) { err = "Unknown activation type."; return false; }
    if (C < 16 || C > 256 || C % 16 != 0) { err = "residual channels must be a multiple of 16 in [16, 256]"; return false; }
    if (C > 256) { err = "residual channels above 256 are not supported"; return false; }
    if (P < 4 || V < 4 || (P + V) % 4 != 0 || P + V > 64) { err = "policy + value head channels must be a multiple of 4 and <= 64"; return false; }
    if (n.blocks < 0 || static_cast<int>(n.tower.size()) != n.blocks) { err = "tower size mismatch"; return false; }
    if (!conv_ok(n.input_conv, SB_INPUT_CHANNELS, C, 3)) { err = "the input layers are wrong"; return false; }
    for (int b = 0; b < n.blocks; ++b) {
        const HostBlock& k = n.tower[static_cast<size_t>(b)];
        if (k.type < SB_BLOCK_RESIDUAL || k.type > SB_BLOCK_MIXER || static_cast<int>(k.convs.size()) != HostBlock::NumConvs(k.type)) { err = "block " + std::to_string(b + 1) + " has an unsupported type"; return false; }
        if (k.type == SB_BLOCK_RESIDUAL) {
            if (!conv_ok(k.convs[0], C, C, 3) || !conv_ok(k.convs[1], C, C, 3)) { err = "residual block " + std::to_string(b + 1) + " is wrong"; return false; }
        } else if (k.type == SB_BLOCK_MIXER) {
            const int F = k.inner;
            if (!width_ok(F)) { err = "feed-forward channels of mixer block " + std::to_string(b + 1) + " (" + std::to_string(F) + ") must be a multiple of 8 in [8, 512]"; return false; }
            if (!dw_ok(k.convs[0], C) || !conv_ok(k.convs[1], C, F, 1) || !conv_ok(k.convs[2], F, C, 1)) { err = "the channels of mixer block " + std::to_string(b + 1) + " is wrong"; return false; }
        } else {
            const int I = k.inner;
            if (!width_ok(I)) { err = "bottleneck channels must be a multiple of 8 in [8, 512]"; return false; }
            const size_t last = k.convs.size() - 1;
            if (!conv_ok(k.convs[0], C, I, 1) || !conv_ok(k.convs[last], I, C, 1)) { err = "the outer channels of bottleneck block " + std::to_string(b + 1) + " is wrong"; return false; }
            for (size_t q = 1; q < last; ++q)
                if (!conv_ok(k.convs[q], I, I, 3)) { err = "the inner channels of bottleneck block " + std::to_string(b + 1) + " is wrong"; return false; }
        }
        if (k.se_size > 0 && (!fc_ok(k.squeeze, 3 * C, k.se_size) || !fc_ok(k.excite, k.se_size, 2 * C))) { err = "SE unit of block " + std::to_string(b + 1) + " is wrong"; return false; }
    }
    if (n.replk && (P % 8 != 0 || !dw_ok(n.p_dw_conv, P) || !conv_ok(n.p_pt_conv, P, P, 1))) { err = "the RepLK policy head is wrong"; return false; }
    if (!conv_ok(n.p_hd_conv, C, P, 1) || !fc_ok(n.p_inter_fc, 3 * P, P) || !conv_ok(n.prob_conv, P, 5, 1) || !fc_ok(n.pass_fc, P, 5)) { err = "the policy head is wrong"; return false; }
    if (!conv_ok(n.v_hd_conv, C, V, 1) || !fc_ok(n.v_inter_fc, 3 * V, 3 * V) || !conv_ok(n.v_ownership, V, 1, 1) || !fc_ok(n.v_misc, 3 * V, 15)) { err = "the value head is wrong"; return false; }
    return true;
}

bool LoadWeightsFile(const std::string& path, HostNet& net, std::string& err) {
    std::ifstream f(path, std::ifstream::binary | std::ifstream::in);
    if (!f.is_open()) {
        err = "Couldn't open weights file from " + path;
        return false;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        Reader r(ss.str());
        net = HostNet{};
        Parse(r, net);
    } catch (const std::exception& e) {
        err = std::string("Fail to load the network file! Cause: ") + e.what();
        return false;
    }
    return ValidateNet(net, err);
}

bool NetFromAbi(const sb_net_desc* d, const sb_weights* w, HostNet& net, std::string& err) {
    if (!d || !w || !w->tensors) { err = "null net description or weights"; return false; }
    net = HostNet{};
    net.version = d->version;
    net.input_channels = d->input_channels;
    net.blocks = d->blocks;
    net.channels = d->channels;
    net.P = d->policy_channels;
    net.V = d->value_channels;
    net.act = d->activation;
    if (net.blocks < 0 || net.blocks > 1024 || (net.blocks > 0 && !d->se_sizes)) { err = "bad block count / se_sizes"; return false; }
    int idx = 0;
    bool ok = true;
    auto take = [&](std::vector<float>& dst, size_t expect) {
        if (idx >= w->n_tensors || !w->tensors[idx].data || static_cast<size_t>(w->tensors[idx].count) != expect) {
            if (ok) err = "tensor " + std::to_string(idx) + ": expected " + std::to_string(expect) + " floats";
            ok = false;
            ++idx;
            return;
        }
        dst.assign(w->tensors[idx].data, w->tensors[idx].data + expect);
        ++idx;
    };
    auto conv = [&](HostConv& c, int in, int out, int k) {
        c.in = in; c.out = out; c.k = k;
        take(c.w, static_cast<size_t>(in) * out * k * k);
        take(c.b, static_cast<size_t>(out));
    };
    auto dwconv = [&](HostConv& c, int ch, int k) {
        c.in = 1; c.out = ch; c.k = k; c.depthwise = true;
        take(c.w, static_cast<size_t>(ch) * k * k);
        take(c.b, static_cast<size_t>(ch));
    };
    auto fc = [&](HostFC& f, int in, int out) {
        f.in = in; f.out = out;
        take(f.w, static_cast<size_t>(in) * out);
        take(f.b, static_cast<size_t>(out));
    };
    const int C = net.channels, P = net.P, V = net.V;
    conv(net.input_conv, net.input_channels, C, 3);
    net.tower.resize(static_cast<size_t>(net.blocks));
    for (int b = 0; b < net.blocks; ++b) {
        HostBlock& k = net.tower[static_cast<size_t>(b)];
        k.type = d->block_types ? d->block_types[b] : SB_BLOCK_RESIDUAL;
        if (k.type < SB_BLOCK_RESIDUAL || k.type > SB_BLOCK_MIXER) { err = "unsupported block type in the net description"; return false; }
        k.inner = k.type == SB_BLOCK_RESIDUAL ? 0 : (d->inner_channels ? d->inner_channels[b] : 0);
        if (k.type != SB_BLOCK_RESIDUAL && k.inner <= 0) { err = "bottleneck / mixer block without inner_channels"; return false; }
        k.convs.resize(static_cast<size_t>(HostBlock::NumConvs(k.type)));
        if (k.type == SB_BLOCK_RESIDUAL) {
            conv(k.convs[0], C, C, 3);
            conv(k.convs[1], C, C, 3);
        } else if (k.type == SB_BLOCK_MIXER) {
            const int kk = d->dw_kernels ? d->dw_kernels[b] : 7;
            if (kk < 3 || kk > 15 || !(kk & 1)) { err = "unsupported depthwise kernel size"; return false; }
            dwconv(k.convs[0], C, kk);
            conv(k.convs[1], C, k.inner, 1);
            conv(k.convs[2], k.inner, C, 1);
        } else {
            const size_t last = k.convs.size() - 1;
            conv(k.convs[0], C, k.inner, 1);
            for (size_t q = 1; q < last; ++q) conv(k.convs[q], k.inner, k.inner, 3);
            conv(k.convs[last], k.inner, C, 1);
        }
        k.se_size = d->se_sizes[b];
        if (k.se_size > 0) {
            fc(k.squeeze, 3 * C, k.se_size);
            fc(k.excite, k.se_size, 2 * C);
        }
    }
    conv(net.p_hd_conv, C, P, 1);
    net.replk = d->policy_head_type == SB_POLICY_HEAD_REPLK;
    if (d->policy_head_type != SB_POLICY_HEAD_NORMAL && !net.replk) { err = "unsupported policy head type"; return false; }
    if (net.replk) {
        const int kk = d->policy_dw_kernel > 0 ? d->policy_dw_kernel : 7;
        if (kk < 3 || kk > 15 || !(kk & 1)) { err = "unsupported depthwise kernel size"; return false; }
        dwconv(net.p_dw_conv, P, kk);
        conv(net.p_pt_conv, P, P, 1);
    }
    fc(net.p_inter_fc, 3 * P, P);
    conv(net.prob_conv, P, 5, 1);
    fc(net.pass_fc, P, 5);
    conv(net.v_hd_conv, C, V, 1);
    fc(net.v_inter_fc, 3 * V, 3 * V);
    conv(net.v_ownership, V, 1, 1);
    fc(net.v_misc, 3 * V, 15);
    if (!ok) return false;
    if (idx != w->n_tensors) { err = "unexpected number of tensors: consumed " + std::to_string(idx) + " of " + std::to_string(w->n_tensors); return false; }
    return ValidateNet(net, err);
}

}  // namespace sb
