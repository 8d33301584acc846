"""Shape coverage of the convolution kernel through the C ABI: every N-tile width class (16 ... 128, multi-tile, padded
to the UMMA granule), K-block counts 1..5, every block family, both policy heads, boards 2..19 — against the CPU oracle
at 1e-4 on the fp32-split rung.  (tools/fuzz_parity.py is the long form: 52 configurations, profiles/r01s2_fuzz_parity.log.)"""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # (C, P, V, stack, policy head, depthwise kernel, activation)
    (16, 8, 8, ["ResidualBlock-SE", "ResidualBlock"], "Normal", 7, "relu"),
    (48, 8, 24, ["BottleneckBlock", "NestedBottleneckBlock-SE"], "RepLK", 5, "mish"),          # inner 24: padded to N = 32
    (80, 16, 16, ["MixerBlock-SE", "ResidualBlock"], "Normal", 9, "swish"),                    # ffn 120: padded to 128
    (112, 40, 24, ["NestedBottleneckBlock", "MixerBlock"], "RepLK", 3, "mish"),               # ffn 168 -> 176 = 11 x 16
    (144, 24, 24, ["BottleneckBlock-SE", "MixerBlock"], "Normal", 7, "relu"),                  # 144 = 3 tiles of 48; ffn 216
    (160, 32, 32, ["ResidualBlock", "NestedBottleneckBlock"], "Normal", 7, "mish"),            # 2 tiles of 80
    (224, 16, 48, ["ResidualBlock-SE", "BottleneckBlock"], "RepLK", 7, "swish"),               # 2 tiles of 112
    (256, 24, 24, ["MixerBlock", "ResidualBlock-SE"], "Normal", 5, "mish"),                    # ffn 384 = 3 tiles of 128
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "C%d-%s" % (c[0], "+".join(s.split("Block")[0] for s in c[3])))
def test_conv_shapes_match_oracle(case):
    from oracle import oracle_py
    from sayuri_b200 import engine, synth
    oracle_py.build()
    C, P, V, stack, head, k, act = case
    path = os.path.join(tempfile.gettempdir(), "sb_shape_%d.bin" % C)
    synth.write_synth_net(path, (len(stack), C, P, V), seed=1000 + C, stack=stack, activation=act, policy_head=head, dw_kernel=k)
    sizes = [19, 2, 13, 9, 19, 7, 19]
    planes = [synth.synth_positions(1, bs, seed=7 * C + i)[0].ravel() for i, bs in enumerate(sizes)]
    offs = [i % 5 for i in range(len(sizes))]
    orc = oracle_py.Oracle(path)
    pipe = engine.B200ForwardPipe().initialize(path, 19, 8, gpus=[0])
    try:
        out = pipe.batch_forward(0, planes, sizes, offs)
        for i, bs in enumerate(sizes):
            ref = orc.forward(planes[i], bs, offs[i])
            s = bs * bs
            np.testing.assert_allclose(out[i]["probabilities"][:s], ref["prob"], rtol=0, atol=1e-4)
            np.testing.assert_allclose(out[i]["ownership"][:s], ref["own"], rtol=0, atol=1e-4)
            assert abs(float(out[i]["pass_probability"]) - float(ref["misc"][0])) < 1e-4
            np.testing.assert_allclose(np.asarray(out[i]["wdl"]), ref["misc"][1:4], rtol=0, atol=1e-4)
            assert not np.any(out[i]["probabilities"][s:])
    finally:
        pipe.destroy()
