"""Generate the committed golden fixtures under tests/golden/ (run in the build container only).

Pins the oracle: the reference ships no tests or golden vectors, so we take
outputs of the reference ITSELF:
  (a) the UNMODIFIED C++ loader + Eigen forward compiled into oracle/_ref (both the im2col path and the
      reference-default Winograd path), via oracle/ref_harness.cc;
  (b) the reference's independent PyTorch forward, /root/reference/train/torch/network.py:1121-1215,
      on the same exported weights — including the 9x9-on-19x19-canvas masked case that defines the
      mixed-board-size semantics (SURVEY.md Appendix B note 2).
The weight file is exported by the reference's own writer (network.py:1399-1481) from a seeded
Network(cfg) with randomised BatchNorm statistics so that BN folding is exercised.

Usage:  python tests/golden/make_golden.py        (needs /root/reference; rewrites the fixtures)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/train/torch")

from network import Network  # noqa: E402  (reference code, imported not copied)
from config import Config  # noqa: E402

from oracle.oracle_py import Reference  # noqa: E402
from sayuri_b200 import synth  # noqa: E402

CFG = {
    "NeuralNetwork": {
        "NNType": "Residual", "MaxBoardSize": 19, "ResidualChannels": 32, "PolicyHeadChannels": 8,
        "ValueHeadChannels": 8, "SeRatio": 4, "PolicyHeadType": "Normal", "Activation": "mish",
        "Stack": ["ResidualBlock", "ResidualBlock-SE", "ResidualBlock"],
    },
    "Train": {"TrainDirectory": "x", "StorePath": "x", "UseGPU": False},
}


CFG_BLOCKS = {
    "NeuralNetwork": {
        "NNType": "Residual", "MaxBoardSize": 19, "ResidualChannels": 32, "PolicyHeadChannels": 8,
        "ValueHeadChannels": 8, "SeRatio": 4, "PolicyHeadType": "Normal", "Activation": "mish",
        "Stack": ["BottleneckBlock-SE", "NestedBottleneckBlock", "ResidualBlock", "NestedBottleneckBlock-SE", "BottleneckBlock"],
    },
    "Train": {"TrainDirectory": "x", "StorePath": "x", "UseGPU": False},
}


CFG_MIXER = {
    "NeuralNetwork": {
        "NNType": "Residual", "MaxBoardSize": 19, "ResidualChannels": 32, "PolicyHeadChannels": 8,
        "ValueHeadChannels": 8, "SeRatio": 4, "PolicyHeadType": {"Type": "RepLK", "KernelSize": 7}, "Activation": "mish",
        "Stack": ["MixerBlock", "MixerBlock-SE", "ResidualBlock", {"Block": "MixerBlock", "Args": {"KernelSize": 5}}],
    },
    "Train": {"TrainDirectory": "x", "StorePath": "x", "UseGPU": False},
}


def randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, mod in net.named_modules():
            if hasattr(mod, "running_mean") and hasattr(mod, "running_var"):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 1.5 + 0.5)
                if getattr(mod, "beta", None) is not None:
                    mod.beta.copy_(torch.randn(mod.beta.shape, generator=g) * 0.1)
                if getattr(mod, "gamma", None) is not None:
                    mod.gamma.copy_(torch.rand(mod.gamma.shape, generator=g) * 0.8 + 0.6)


def main_symmetry():
    """Fourth fixture: Symmetry::TransformIndex of the compiled reference (game/symmetry.cc:97-123) for boards 2, 5, 9, 13,
    19 and all 8 symmetries: the bit-exact pin of the gather / scatter index work (encoder.cc:80-100, network.cc:376-383)."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsayuri_ref_v3.so"))
    lib.ref_symmetry_table.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    out = {}
    for n in (2, 5, 9, 13, 19):
        for sy in range(8):
            t = np.zeros(n * n, dtype=np.int32)
            assert lib.ref_symmetry_table(n, sy, t.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
            out["sym_%d_%d" % (n, sy)] = t
    np.savez_compressed(os.path.join(HERE, "golden_symmetry.npz"), **out)
    print("wrote golden_symmetry.npz")


def main_blocks():
    _fixture(CFG_BLOCKS, "btl_5bx32", 20260418, 9, 21)
    _fixture(CFG_MIXER, "mix_4bx32", 20260419, 10, 22)


def _fixture(cfg, tag, seed, bn_seed, pos_seed):
    """Second and third fixtures: the optional block families — BottleneckBlock and NestedBottleneckBlock,
    with and without SE (blas_forward_pipe.cc:90-263) — exported by the reference writer, evaluated by the
    UNMODIFIED compiled reference (im2col path) and cross-checked against the reference's PyTorch forward."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    net = Network(Config(json.dumps(cfg), is_file=False))
    net.eval()
    randomise_bn(net, bn_seed)
    wbin = os.path.join(HERE, "ref_%s.bin.txt" % tag)
    net.transfer_to_bin(wbin)
    out = {}
    ref = Reference(wbin, winograd=False)
    per = 2
    for bs in (9, 13, 19):
        x = synth.synth_positions(per, bs, seed=pos_seed)
        out["planes_%d" % bs] = x
        with torch.no_grad():
            pred, _ = net(torch.from_numpy(x).reshape(per, 43, bs, bs))
        prob5 = torch.stack([p[:, :-1] for p in pred[:5]], dim=1).numpy()
        for i in range(per):
            off = (i * 2 + bs) % 5
            r = ref.forward(x[i], bs, offset=off)
            v = np.concatenate([r["prob"], r["own"], r["misc"]])
            out["ref_%d_%d" % (bs, i)] = v
            out["offset_%d_%d" % (bs, i)] = np.int32(off)
            s = bs * bs
            d = np.abs(v[:s] - prob5[i, off]).max()
            d_own = np.abs(np.tanh(v[s:2 * s]) - pred[5].numpy()[i]).max()
            d_wdl = np.abs(v[2 * s + 1:2 * s + 4] - pred[6].numpy()[i]).max()
            print("%s net, bs %2d pos %d: C++ vs torch prob %.2e own %.2e wdl %.2e" % (tag, bs, i, d, d_own, d_wdl))
            assert max(d, d_own, d_wdl) < 5e-5
    np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % tag), **out)
    print("wrote", wbin, "golden_%s.npz" % tag)


def main():
    torch.manual_seed(20260417)
    np.random.seed(20260417)
    cfg = Config(json.dumps(CFG), is_file=False)
    net = Network(cfg)
    net.eval()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for name, mod in net.named_modules():
            if hasattr(mod, "running_mean") and hasattr(mod, "running_var"):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 1.5 + 0.5)
                if getattr(mod, "beta", None) is not None:
                    mod.beta.copy_(torch.randn(mod.beta.shape, generator=g) * 0.1)
                if getattr(mod, "gamma", None) is not None:
                    mod.gamma.copy_(torch.rand(mod.gamma.shape, generator=g) * 0.8 + 0.6)
    wbin = os.path.join(HERE, "ref_3bx32.bin.txt")
    wtxt = os.path.join(HERE, "ref_3bx32.txt")
    net.transfer_to_bin(wbin)
    net.transfer_to_text(wtxt)

    sizes = [9, 13, 19]
    per = 2
    out = {}
    ref = Reference(wbin, winograd=False)
    for bs in sizes:
        x = synth.synth_positions(per, bs, seed=11)
        out["planes_%d" % bs] = x
        for i in range(per):
            off = (i * 3 + bs) % 5
            r = ref.forward(x[i], bs, offset=off)
            out["ref_%d_%d" % (bs, i)] = np.concatenate([r["prob"], r["own"], r["misc"]])
            out["offset_%d_%d" % (bs, i)] = np.int32(off)
    # Reference-default Winograd path as a noise-floor record (one process = one option map, so re-init)
    refw = Reference(wbin, winograd=True)
    for bs in sizes:
        x = out["planes_%d" % bs]
        for i in range(per):
            off = int(out["offset_%d_%d" % (bs, i)])
            r = refw.forward(x[i], bs, offset=off)
            out["refwino_%d_%d" % (bs, i)] = np.concatenate([r["prob"], r["own"], r["misc"]])

    # (b) PyTorch forward: all 5 policy planes (raw), pass logits, tanh(ownership), raw wdl ...
    with torch.no_grad():
        for bs in sizes:
            x = torch.from_numpy(out["planes_%d" % bs]).reshape(per, 43, bs, bs)
            pred, _ = net(x)
            prob5 = torch.stack([p[:, :-1] for p in pred[:5]], dim=1).numpy()  # [B,5,s]
            pass5 = torch.stack([p[:, -1] for p in pred[:5]], dim=1).numpy()
            out["torch_prob5_%d" % bs] = prob5
            out["torch_pass5_%d" % bs] = pass5
            out["torch_own_tanh_%d" % bs] = pred[5].numpy()
            out["torch_wdl_%d" % bs] = pred[6].numpy()
            out["torch_q_tanh_%d" % bs] = pred[7].numpy()
            out["torch_scores20_%d" % bs] = pred[8].numpy()
        # canvas semantics: 9x9 and 13x13 samples on a 19x19 canvas (top-left, zero elsewhere; plane 42 = mask)
        for bs in (9, 13):
            xn = out["planes_%d" % bs].reshape(per, 43, bs, bs)
            canvas = np.zeros((per, 43, 19, 19), dtype=np.float32)
            canvas[:, :, :bs, :bs] = xn
            pred, _ = net(torch.from_numpy(canvas))
            prob5 = torch.stack([p[:, :-1] for p in pred[:5]], dim=1).reshape(per, 5, 19, 19)[:, :, :bs, :bs]
            native = out["torch_prob5_%d" % bs].reshape(per, 5, bs, bs)
            d = float(np.abs(prob5.numpy() - native).max())
            own = pred[5].reshape(per, 19, 19)[:, :bs, :bs].numpy()
            d2 = float(np.abs(own - out["torch_own_tanh_%d" % bs].reshape(per, bs, bs)).max())
            print("canvas-vs-native (PyTorch reference) board %d: prob %.3g own %.3g" % (bs, d, d2))
            assert d < 2e-5 and d2 < 2e-5
            out["torch_canvas_delta_%d" % bs] = np.float32(max(d, d2))

    # cross-check (a) vs (b) before writing
    for bs in sizes:
        for i in range(per):
            off = int(out["offset_%d_%d" % (bs, i)])
            s = bs * bs
            r = out["ref_%d_%d" % (bs, i)]
            d = np.abs(r[:s] - out["torch_prob5_%d" % bs][i, off]).max()
            d_own = np.abs(np.tanh(r[s:2 * s]) - out["torch_own_tanh_%d" % bs][i]).max()
            d_pass = abs(r[2 * s] - out["torch_pass5_%d" % bs][i, off])
            d_wdl = np.abs(r[2 * s + 1:2 * s + 4] - out["torch_wdl_%d" % bs][i]).max()
            dw = np.abs(r - out["refwino_%d_%d" % (bs, i)]).max()
            print("bs %2d pos %d: C++ vs torch prob %.2e own %.2e pass %.2e wdl %.2e | winograd-vs-im2col %.2e"
                  % (bs, i, d, d_own, d_pass, d_wdl, dw))
            assert max(d, d_own, d_pass, d_wdl) < 5e-5
    np.savez_compressed(os.path.join(HERE, "golden_3bx32.npz"), **out)
    print("wrote", wbin, wtxt, "golden_3bx32.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "blocks":
        main_blocks()      # only the block-family fixtures (the first one stays byte-identical)
    elif len(sys.argv) > 1 and sys.argv[1] == "symmetry":
        main_symmetry()
    else:
        main()
        main_blocks()
        main_symmetry()
