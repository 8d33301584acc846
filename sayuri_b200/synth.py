"""Synthetic weight files and positions for tests and benchmarks.

There is no network access for real checkpoints, so nets are random-initialised in the reference's
own weight-file format (writer: /root/reference/train/torch/network.py:1399-1481, reader:
/root/reference/src/neural/loader.cc:67-121; layout summarised in SURVEY.md Appendix A) and positions
follow the seeded generator of SURVEY.md §8(d).  Nothing here touches the GPU.
"""
import numpy as np

INPUT_CHANNELS = 43  # /root/reference/src/neural/network_basic.h:10
POLICY_OUTS = 5
VALUE_MISC = 15

# BASELINE.json configs -> (blocks, channels, policy head channels, value head channels); SURVEY.md §8.
NETS = {
    "6bx96": (6, 96, 24, 24),
    "10bx128": (10, 128, 24, 24),
    "15bx192": (15, 192, 32, 32),
    "20bx256": (20, 256, 32, 32),
}


def default_stack(blocks, se_every=3):
    """ResidualBlock with -SE on every third block (bash/configs/selfplay-setting.json:10-17)."""
    return ["ResidualBlock-SE" if (b + 1) % se_every == 0 else "ResidualBlock" for b in range(blocks)]


def synth_tensors(blocks, channels, P, V, seed=0, se_ratio=4, stack=None, activation="mish", policy_head="Normal",
                  dw_kernel=7):
    """Random tensors in loader order (loader.cc:658-747): list of (struct_line, [tensor, tensor])."""
    rng = np.random.default_rng(seed)
    stack = stack or default_stack(blocks)
    layers = []

    def conv(cin, cout, k, gain=1.6):
        std = gain / np.sqrt(cin * k * k)
        w = rng.normal(0.0, std, size=(cout, cin, k, k)).astype(np.float32)
        b = rng.normal(0.0, 0.05, size=(cout,)).astype(np.float32)
        layers.append(("Convolution %d %d %d" % (cin, cout, k), [w, b]))

    def dwconv(c, k, gain=1.2):   # "DepthwiseConvolution 1 C k", weights [C][1][k][k] (network.py writer)
        w = rng.normal(0.0, gain / k, size=(c, 1, k, k)).astype(np.float32)
        b = rng.normal(0.0, 0.05, size=(c,)).astype(np.float32)
        layers.append(("DepthwiseConvolution 1 %d %d" % (c, k), [w, b]))

    def bn(c):
        mean = rng.normal(0.0, 0.1, size=(c,)).astype(np.float32)
        std = rng.uniform(0.7, 1.4, size=(c,)).astype(np.float32)  # file stores sqrt(var+eps)/gamma
        layers.append(("BatchNorm %d" % c, [mean, std]))

    def fc(cin, cout, gain=1.0):
        w = rng.normal(0.0, gain / np.sqrt(cin), size=(cout, cin)).astype(np.float32)
        b = rng.normal(0.0, 0.1, size=(cout,)).astype(np.float32)
        layers.append(("FullyConnect %d %d" % (cin, cout), [w, b]))

    conv(INPUT_CHANNELS, channels, 3)
    bn(channels)
    for name in stack:
        base = name.split("-")[0]
        if base == "ResidualBlock":
            conv(channels, channels, 3)
            bn(channels)
            conv(channels, channels, 3, gain=0.7)
            bn(channels)
        elif base in ("BottleneckBlock", "NestedBottleneckBlock"):   # loader.cc:416-555; inner = channels // 2 like network.py:1054
            inner = channels // 2
            conv(channels, inner, 1)
            bn(inner)
            for _ in range(2 if base == "BottleneckBlock" else 4):
                conv(inner, inner, 3, gain=1.2)
                bn(inner)
            conv(inner, channels, 1, gain=0.7)
            bn(channels)
        elif base == "MixerBlock":   # loader.cc:556-607; ffn = 1.5 x channels like network.py:881
            ffn = int(1.5 * channels)
            dwconv(channels, dw_kernel)
            bn(channels)
            conv(channels, ffn, 1)
            bn(ffn)
            conv(ffn, channels, 1, gain=0.7)
            bn(channels)
        else:
            raise ValueError("unknown block type %s" % name)
        if name.endswith("-SE"):
            se = channels // se_ratio
            fc(3 * channels, se)
            fc(se, 2 * channels)
    conv(channels, P, 1)
    bn(P)
    if policy_head == "RepLK":   # loader.cc:691-702
        dwconv(P, dw_kernel)
        bn(P)
        conv(P, P, 1)
        bn(P)
    fc(3 * P, P)
    conv(P, POLICY_OUTS, 1)
    fc(P, POLICY_OUTS)
    conv(channels, V, 1)
    bn(V)
    fc(3 * V, 3 * V)
    conv(V, 1, 1)
    fc(3 * V, VALUE_MISC)
    info = dict(blocks=blocks, channels=channels, P=P, V=V, stack=stack, activation=activation, policy_head=policy_head)
    return info, layers


def write_weights(path, info, layers, binary=True, version=5):
    """Write the reference weight-file format (text `float32` or `float32bin`)."""
    head = ["get main", "get info", "NNType Residual", "Version %d" % version,
            "FloatType %s" % ("float32bin" if binary else "float32"),
            "InputChannels %d" % INPUT_CHANNELS, "ResidualChannels %d" % info["channels"],
            "ResidualBlocks %d" % info["blocks"], "PolicyHeadChannels %d" % info["P"],
            "ValueHeadChannels %d" % info["V"], "ValueMisc %d" % VALUE_MISC, "PolicyHeadType %s" % info.get("policy_head", "Normal"),
            "ActivationFunction %s" % info["activation"], "end info", "get stack"]
    head += list(info["stack"]) + ["end stack", "get struct"]
    head += [name for name, _ in layers] + ["end struct", "get parameters"]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        for _, tensors in layers:
            for t in tensors:
                flat = np.ascontiguousarray(t, dtype="<f4").ravel()
                if binary:
                    f.write(flat.tobytes() + b"\xff\xff\xff\xff")
                else:
                    f.write((" ".join(repr(float(v)) for v in flat) + "\n").encode())
        f.write(b"end parameters\n" if not binary else b"end parameters\n")
        f.write(b"end main")


def write_synth_net(path, name_or_shape, seed=0, binary=True, activation="mish", stack=None, policy_head="Normal",
                    dw_kernel=7):
    shape = NETS[name_or_shape] if isinstance(name_or_shape, str) else name_or_shape
    info, layers = synth_tensors(*shape, seed=seed, activation=activation, stack=stack, policy_head=policy_head,
                                 dw_kernel=dw_kernel)
    write_weights(path, info, layers, binary=binary)
    return info


def synth_positions(n, board_size, seed=20260417, komi_choices=(5.5, 6.5, 7.0, 7.5)):
    """Seeded synthetic encoder planes, NCHW at native board size: float32 [n, 43, bs*bs].

    Plane semantics follow /root/reference/src/neural/encoder.h:20-55: 24 history planes (own stones,
    opponent stones, last move for 8 past positions), 13 binary feature planes, 6 scalar planes
    (rule, wave, komi/20, -komi/20, intersections/361, ones).  SURVEY.md §8(d) recipe.
    """
    rng = np.random.default_rng(seed + board_size)
    s = board_size * board_size
    x = np.zeros((n, INPUT_CHANNELS, s), dtype=np.float32)
    for i in range(n):
        occ = rng.random(s)
        own = occ < 0.25
        opp = (occ >= 0.25) & (occ < 0.5)
        for h in range(8):
            flip = rng.random(s) < 0.03 * h
            x[i, 3 * h + 0] = own & ~flip
            x[i, 3 * h + 1] = opp & ~flip
            x[i, 3 * h + 2, rng.integers(0, s)] = 1.0
        for p in range(24, 37):
            x[i, p] = rng.random(s) < 0.1
        if rng.random() < 0.95:
            x[i, 24] = 0.0
        else:
            x[i, 24] = 0.0
            x[i, 24, rng.integers(0, s)] = 1.0
        komi = float(rng.choice(komi_choices))
        x[i, 37] = float(rng.integers(0, 2))
        x[i, 38] = float(rng.uniform(-1, 1))
        x[i, 39] = komi / 20.0
        x[i, 40] = -komi / 20.0
        x[i, 41] = s / 361.0
        x[i, 42] = 1.0
    return x
