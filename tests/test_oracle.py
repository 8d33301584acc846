"""Pins the CPU oracle (oracle/oracle_forward.c) against outputs of the reference itself.

Golden vectors: tests/golden/golden_3bx32.npz, produced by tests/golden/make_golden.py from the
UNMODIFIED reference C++ (loader.cc + blas_forward_pipe.cc, compiled into oracle/_ref) and the
reference's PyTorch forward (train/torch/network.py).  Tolerances are written next to each check.
"""
import numpy as np
import pytest

SIZES = (9, 13, 19)
PER = 2
# fp32 forward with a different summation order than Eigen's GEMM (+ the reference's -ffast-math):
ORACLE_VS_REF_ATOL = 2e-5


def _unpack(vec, bs):
    s = bs * bs
    return vec[:s], vec[s:2 * s], vec[2 * s:]


@pytest.mark.parametrize("fmt", ["bin", "txt"])
def test_oracle_matches_reference_golden(oracle_lib, golden, golden_weights_bin, golden_weights_txt, fmt):
    o = oracle_lib.Oracle(golden_weights_bin if fmt == "bin" else golden_weights_txt)
    assert (o.version, o.input_channels, o.blocks, o.channels, o.P, o.V, o.act, o.n_se) == (5, 43, 3, 32, 8, 8, 5, 1)
    for bs in SIZES:
        x = golden["planes_%d" % bs]
        for i in range(PER):
            off = int(golden["offset_%d_%d" % (bs, i)])
            got = o.forward(x[i], bs, offset=off)
            prob, own, misc = _unpack(golden["ref_%d_%d" % (bs, i)], bs)
            np.testing.assert_allclose(got["prob"], prob, rtol=0, atol=ORACLE_VS_REF_ATOL)
            np.testing.assert_allclose(got["own"], own, rtol=0, atol=ORACLE_VS_REF_ATOL)
            np.testing.assert_allclose(got["misc"], misc, rtol=0, atol=ORACLE_VS_REF_ATOL)


def test_oracle_matches_reference_golden_on_bottleneck_blocks(oracle_lib, golden_blocks, golden_blocks_weights):
    """BottleneckBlock[-SE] and NestedBottleneckBlock[-SE] (blas_forward_pipe.cc:90-263), mixed with a plain
    ResidualBlock, against the compiled reference's outputs."""
    o = oracle_lib.Oracle(golden_blocks_weights)
    assert (o.blocks, o.channels, o.P, o.V, o.act, o.n_se) == (5, 32, 8, 8, 5, 2)
    for bs in SIZES:
        x = golden_blocks["planes_%d" % bs]
        for i in range(PER):
            off = int(golden_blocks["offset_%d_%d" % (bs, i)])
            got = o.forward(x[i], bs, offset=off)
            prob, own, misc = _unpack(golden_blocks["ref_%d_%d" % (bs, i)], bs)
            np.testing.assert_allclose(got["prob"], prob, rtol=0, atol=ORACLE_VS_REF_ATOL)
            np.testing.assert_allclose(got["own"], own, rtol=0, atol=ORACLE_VS_REF_ATOL)
            np.testing.assert_allclose(got["misc"], misc, rtol=0, atol=ORACLE_VS_REF_ATOL)


def test_oracle_matches_reference_golden_on_mixer_blocks_and_replk_head(oracle_lib, golden_mixer, golden_mixer_weights):
    """MixerBlock[-SE] (blas_forward_pipe.cc:265-312: depthwise k x k, AddSpatialBiasesPost, 1x1 FFN) with kernel
    sizes 7 and 5, and the RepLK policy head (:443-471), against the compiled reference's outputs."""
    o = oracle_lib.Oracle(golden_mixer_weights)
    assert (o.blocks, o.channels, o.P, o.V, o.act, o.n_se) == (4, 32, 8, 8, 5, 1)
    for bs in SIZES:
        x = golden_mixer["planes_%d" % bs]
        for i in range(PER):
            off = int(golden_mixer["offset_%d_%d" % (bs, i)])
            got = o.forward(x[i], bs, offset=off)
            prob, own, misc = _unpack(golden_mixer["ref_%d_%d" % (bs, i)], bs)
            np.testing.assert_allclose(got["prob"], prob, rtol=0, atol=ORACLE_VS_REF_ATOL)
            np.testing.assert_allclose(got["own"], own, rtol=0, atol=ORACLE_VS_REF_ATOL)
            np.testing.assert_allclose(got["misc"], misc, rtol=0, atol=ORACLE_VS_REF_ATOL)


def test_oracle_matches_reference_pytorch_forward(oracle_lib, golden, golden_weights_bin):
    """Second, independent pin: all five policy planes, pass logits and the value outputs of
    train/torch/network.py:1121-1215 (which applies tanh / scaling inside forward)."""
    o = oracle_lib.Oracle(golden_weights_bin)
    for bs in SIZES:
        x = golden["planes_%d" % bs]
        for i in range(PER):
            t = o.forward_trace(x[i], bs, offset=0)
            np.testing.assert_allclose(t["all_prob"], golden["torch_prob5_%d" % bs][i], rtol=0, atol=5e-5)
            np.testing.assert_allclose(t["all_pass"], golden["torch_pass5_%d" % bs][i], rtol=0, atol=5e-5)
            np.testing.assert_allclose(np.tanh(t["own"]), golden["torch_own_tanh_%d" % bs][i], rtol=0, atol=2e-5)
            np.testing.assert_allclose(t["all_misc"][0:3], golden["torch_wdl_%d" % bs][i], rtol=0, atol=2e-5)
            np.testing.assert_allclose(np.tanh(t["all_misc"][3:8]), golden["torch_q_tanh_%d" % bs][i], rtol=0, atol=2e-5)
            np.testing.assert_allclose(20 * t["all_misc"][8:13], golden["torch_scores20_%d" % bs][i], rtol=0, atol=4e-4)


def test_winograd_noise_floor_recorded(golden):
    """The reference's own two conv paths (Winograd default vs im2col) differ by ~1e-6: the noise floor
    under the 1e-4 parity tolerance of BASELINE.json."""
    worst = 0.0
    for bs in SIZES:
        for i in range(PER):
            worst = max(worst, float(np.abs(golden["ref_%d_%d" % (bs, i)] - golden["refwino_%d_%d" % (bs, i)]).max()))
    assert worst < 2e-5


def test_reference_so_agrees_with_oracle_when_present(oracle_lib, golden, golden_weights_bin):
    """When oracle/_ref is built (container with /root/reference, or shipped to the GPU box),
    re-run the real reference and compare live."""
    if not oracle_lib.Reference.available():
        pytest.skip("oracle/_ref not built")
    r = oracle_lib.Reference(golden_weights_bin, winograd=False)
    o = oracle_lib.Oracle(golden_weights_bin)
    for bs in SIZES:
        x = golden["planes_%d" % bs]
        for i in range(PER):
            a = r.forward(x[i], bs, offset=i)
            b = o.forward(x[i], bs, offset=i)
            for k in ("prob", "own", "misc"):
                np.testing.assert_allclose(a[k], b[k], rtol=0, atol=ORACLE_VS_REF_ATOL)


def test_canvas_place_and_crop_bit_exact(oracle_lib):
    """batch_forward_pipe.cc:15-33 / :48-67 — pure index work, bit-exact."""
    rng = np.random.default_rng(3)
    for n in (9, 13, 19):
        planes = rng.standard_normal((43, n, n)).astype(np.float32)
        canvas = oracle_lib.Oracle.canvas_place(planes, 43, n, 19).reshape(43, 19, 19)
        want = np.zeros((43, 19, 19), dtype=np.float32)
        want[:, :n, :n] = planes
        assert np.array_equal(canvas, want)
        back = oracle_lib.Oracle.canvas_crop(canvas[5], n, 19).reshape(n, n)
        assert np.array_equal(back, planes[5])


def test_symmetry_tables_are_permutations_and_involutive_pairs(oracle_lib):
    """game/symmetry.cc:97-123.  Gather with T then scatter with T (network.cc:376-383) is identity."""
    rng = np.random.default_rng(4)
    for n in (9, 19):
        v = rng.standard_normal(n * n).astype(np.float32)
        for symm in range(8):
            t = oracle_lib.Oracle.symmetry_table(n, symm)
            assert sorted(t.tolist()) == list(range(n * n))
            gathered = v[t]                      # encoder.cc:92-95  buf[i] = plane[T(i)]
            scattered = np.empty_like(v)
            scattered[t] = gathered              # network.cc:379-381 out[T(i)] = in[i]
            assert np.array_equal(scattered, v)
        assert np.array_equal(oracle_lib.Oracle.symmetry_table(n, 0), np.arange(n * n))


def test_symmetry_tables_equal_the_reference_bit_exact(oracle_lib):
    """tests/golden/golden_symmetry.npz = Symmetry::TransformIndex of the compiled reference for boards 2..19 x 8 symmetries."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_symmetry.npz"))
    for n in (2, 5, 9, 13, 19):
        for symm in range(8):
            assert np.array_equal(oracle_lib.Oracle.symmetry_table(n, symm), g["sym_%d_%d" % (n, symm)]), (n, symm)


def test_loader_rejects_out_of_scope_and_malformed(oracle_lib, tmp_path):
    from sayuri_b200 import synth
    info, layers = synth.synth_tensors(1, 16, 4, 4, seed=0, stack=["ResidualBlock"])
    p = tmp_path / "bad.txt"
    info_bad = dict(info, stack=["BottleneckBlock"])
    synth.write_weights(str(p), info_bad, layers)
    with pytest.raises(RuntimeError):
        oracle_lib.Oracle(str(p))
    layers[0][1][0] = layers[0][1][0][:-1]  # truncated conv weight
    synth.write_weights(str(p), info, layers)
    with pytest.raises(RuntimeError):
        oracle_lib.Oracle(str(p))
