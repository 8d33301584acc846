"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the committed
golden vectors of the reference.  Needs a B200: run with `pytest -m gpu`.

Tolerances (written here, per BASELINE.json's north star):
  * default precision (fp16 hi/lo split, fp32 accumulate): 1e-4 absolute on every raw output
    (policy logits, ownership, pass, wdl, stm, score, errors) vs the reference Eigen forward;
  * index work (canvas placement / crop, policy-plane select, zero fill): bit-exact;
  * batch invariance: bit-exact (a position's result never depends on its batch neighbours);
  * SB_PRECISION_FP16 (the reference's own --fp16 trade, SELF_CHECK tolerance 0.2 L2,
    network.cc:333-359): 5e-2 absolute, reported not asserted tight.
"""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL = 1e-4
SIZES = (9, 13, 19)


def _misc(o):
    return np.array([o["pass_probability"], *o["wdl"], o["stm_winrate"], o["final_score"], o["q_error"], o["score_error"]],
                    dtype=np.float32)


def _check(out, ref, bs, atol=ATOL):
    s = bs * bs
    np.testing.assert_allclose(out["probabilities"][:s], ref["prob"], rtol=0, atol=atol)
    np.testing.assert_allclose(out["ownership"][:s], ref["own"], rtol=0, atol=atol)
    np.testing.assert_allclose(_misc(out), ref["misc"], rtol=0, atol=atol)
    assert not np.any(out["probabilities"][s:]) and not np.any(out["ownership"][s:])  # zero fill, exact


@pytest.fixture(scope="module")
def eng():
    from sayuri_b200 import engine
    engine.load_library()   # raises if the CUDA library is missing: no fallback
    return engine


@pytest.fixture(scope="module")
def golden_pipe(eng, golden_weights_bin):
    pipe = eng.B200ForwardPipe().initialize(golden_weights_bin, 19, 16, gpus=[0])
    yield pipe
    pipe.destroy()


def test_matches_reference_golden_vectors(golden_pipe, golden):
    """Outputs of the UNMODIFIED reference (tests/golden/make_golden.py) for 9/13/19 boards, every policy offset."""
    planes, sizes, offsets, refs = [], [], [], []
    for bs in SIZES:
        for i in range(2):
            planes.append(golden["planes_%d" % bs][i].ravel())
            sizes.append(bs)
            offsets.append(int(golden["offset_%d_%d" % (bs, i)]))
            v = golden["ref_%d_%d" % (bs, i)]
            s = bs * bs
            refs.append({"prob": v[:s], "own": v[s:2 * s], "misc": v[2 * s:]})
    out = golden_pipe.batch_forward(0, planes, sizes, offsets)
    for o, r, bs, off in zip(out, refs, sizes, offsets):
        _check(o, r, bs)
        assert o["board_size"] == bs and o["offset"] == off and o["fp16"] == 0


def test_all_policy_planes_match_pytorch_reference(golden_pipe, golden):
    """Policy-plane select (FillOutputs): plane k of the reference PyTorch forward for k = 0..4."""
    bs = 19
    x = golden["planes_19"][0].ravel()
    out = golden_pipe.batch_forward(0, [x] * 5, [bs] * 5, list(range(5)))
    for k in range(5):
        np.testing.assert_allclose(out[k]["probabilities"], golden["torch_prob5_19"][0, k], rtol=0, atol=ATOL)
        np.testing.assert_allclose(out[k]["pass_probability"], golden["torch_pass5_19"][0, k], rtol=0, atol=ATOL)
        # everything that does not depend on the offset is bit-identical across the five copies
        assert np.array_equal(out[k]["ownership"], out[0]["ownership"])
        assert np.array_equal(out[k]["wdl"], out[0]["wdl"])


def test_matches_reference_so_live_when_shipped(golden_pipe, golden_weights_bin, oracle_lib):
    if not oracle_lib.Reference.available():
        pytest.skip("oracle/_ref not shipped")
    from sayuri_b200 import synth
    ref = oracle_lib.Reference(golden_weights_bin, winograd=False)
    sizes = [19, 9, 13, 19]
    planes = [synth.synth_positions(1, bs, seed=900 + i)[0].ravel() for i, bs in enumerate(sizes)]
    out = golden_pipe.batch_forward(0, planes, sizes, [0, 1, 2, 3])
    for i, bs in enumerate(sizes):
        _check(out[i], ref.forward(planes[i], bs, offset=i), bs)


def test_batch_invariance_and_canvas_index_work_bit_exact(golden_pipe):
    """A position's outputs are bit-identical alone, duplicated, and buried in a mixed 9/13/19 batch at any slot:
    canvas placement/crop and per-sample reductions do not depend on the neighbours."""
    from sayuri_b200 import synth
    probe = {bs: synth.synth_positions(1, bs, seed=77)[0].ravel() for bs in SIZES}
    alone = {bs: golden_pipe.batch_forward(0, [probe[bs]], [bs], [2])[0] for bs in SIZES}
    filler = [synth.synth_positions(1, bs, seed=500 + i)[0].ravel() for i, bs in enumerate((19, 9, 13, 19, 9, 13, 19, 19, 13))]
    fsizes = [19, 9, 13, 19, 9, 13, 19, 19, 13]
    for pos in (0, 4, 9):
        for bs in SIZES:
            planes = filler[:pos] + [probe[bs]] + filler[pos:]
            sizes = fsizes[:pos] + [bs] + fsizes[pos:]
            out = golden_pipe.batch_forward(0, planes, sizes, [2] * len(sizes))
            for f in ("probabilities", "ownership", "pass_probability", "wdl", "stm_winrate", "final_score", "q_error", "score_error"):
                assert np.array_equal(out[pos][f], alone[bs][f]), (pos, bs, f)


def test_small_canvas_matches_large_canvas(eng, golden_weights_bin, oracle_lib):
    """CudaForwardPipe::Construct on a board change: a 9x9 net canvas gives the same results as 9x9 on the 19x19 canvas."""
    from sayuri_b200 import synth
    orc = oracle_lib.Oracle(golden_weights_bin)
    planes = [synth.synth_positions(1, 9, seed=300 + i)[0].ravel() for i in range(3)]
    pipe = eng.B200ForwardPipe().initialize(golden_weights_bin, 9, 4, gpus=[0])
    try:
        out9 = pipe.batch_forward(0, planes, [9] * 3, [0, 1, 4])
        for i in range(3):
            _check(out9[i], orc.forward(planes[i], 9, [0, 1, 4][i]), 9)
        with pytest.raises(RuntimeError, match="exceeds the NN canvas"):
            pipe.batch_forward(0, [synth.synth_positions(1, 13)[0].ravel()], [13], [0])
        pipe.construct(board_size=19, batch_size=-1)   # Reconstruct, network.cc:494-498
        assert pipe.board_size == 19
        out19 = pipe.batch_forward(0, planes, [9] * 3, [0, 1, 4])
        for i in range(3):
            _check(out19[i], orc.forward(planes[i], 9, [0, 1, 4][i]), 9)
        pipe.construct(batch_size=64)                   # grow the batch only
        assert pipe.max_batch == 64 and pipe.board_size == 19
        with pytest.raises(RuntimeError, match="batch size out of range"):
            pipe.batch_forward(0, planes * 30, [9] * 90, [0] * 90)
    finally:
        pipe.destroy()


@pytest.mark.parametrize("shape,stack", [
    ((2, 64, 8, 8), ["ResidualBlock-SE", "ResidualBlock"]),
    ((2, 96, 24, 24), ["ResidualBlock", "ResidualBlock-SE"]),
    ((3, 128, 24, 24), ["ResidualBlock", "ResidualBlock", "ResidualBlock-SE"]),
    ((2, 192, 32, 32), ["ResidualBlock", "ResidualBlock-SE"]),
    ((2, 256, 32, 32), ["ResidualBlock-SE", "ResidualBlock"]),
])
def test_channel_widths_of_all_baseline_nets(eng, oracle_lib, shape, stack):
    """Every tower width BASELINE.json names (96/128/192/256) through its own tile configuration, mixed board sizes."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_%dx%d.bin" % (shape[0], shape[1]))
    synth.write_synth_net(path, shape, seed=shape[1], stack=stack)
    orc = oracle_lib.Oracle(path)
    sizes = [19, 13, 9, 19, 19, 9, 13]
    planes = [synth.synth_positions(1, bs, seed=10 + i)[0].ravel() for i, bs in enumerate(sizes)]
    offsets = [i % 5 for i in range(len(sizes))]
    pipe = eng.B200ForwardPipe().initialize(path, 19, 8, gpus=[0])
    try:
        out = pipe.batch_forward(0, planes, sizes, offsets)
        for i, bs in enumerate(sizes):
            _check(out[i], orc.forward(planes[i], bs, offsets[i]), bs)
    finally:
        pipe.destroy()


@pytest.mark.parametrize("act", ["relu", "swish", "gelu", "hardswish", "elu", "selu", "identity"])
def test_every_activation(eng, oracle_lib, act):
    """activation.h:8-17,43-59."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_act_%s.bin" % act)
    synth.write_synth_net(path, (2, 32, 8, 8), seed=3, activation=act, stack=["ResidualBlock-SE", "ResidualBlock"])
    orc = oracle_lib.Oracle(path)
    sizes = [19, 9]
    planes = [synth.synth_positions(1, bs, seed=60 + i)[0].ravel() for i, bs in enumerate(sizes)]
    pipe = eng.B200ForwardPipe().initialize(path, 19, 4, gpus=[0])
    try:
        out = pipe.batch_forward(0, planes, sizes, [0, 3])
        for i, bs in enumerate(sizes):
            _check(out[i], orc.forward(planes[i], bs, [0, 3][i]), bs, atol=2e-4 if act in ("identity", "selu") else ATOL)
    finally:
        pipe.destroy()


def test_tensor_core_kernel_agrees_with_simt_cross_check(eng, golden_weights_bin):
    """conv3x3_tc vs conv3x3_simt (fp32 CUDA cores) on the same canvas buffers, trunk level."""
    from sayuri_b200 import synth
    sizes = [19, 13, 9, 19]
    planes = [synth.synth_positions(1, bs, seed=700 + i)[0].ravel() for i, bs in enumerate(sizes)]
    trunks = {}
    for prec in (eng.PRECISION_FP32_SPLIT, eng.PRECISION_SIMT_DEBUG):
        pipe = eng.B200ForwardPipe().initialize(golden_weights_bin, 19, 4, gpus=[0], precision=prec)
        try:
            pipe.batch_forward(0, planes, sizes, [0] * 4)
            trunks[prec] = [pipe.debug_read_trunk(0, 0, i, bs) for i, bs in enumerate(sizes)]
        finally:
            pipe.destroy()
    for a, b in zip(trunks[eng.PRECISION_FP32_SPLIT], trunks[eng.PRECISION_SIMT_DEBUG]):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)


def test_fp16_rung_is_close_and_flagged(eng, golden_weights_bin, oracle_lib):
    from sayuri_b200 import synth
    orc = oracle_lib.Oracle(golden_weights_bin)
    sizes = [19, 9, 13]
    planes = [synth.synth_positions(1, bs, seed=800 + i)[0].ravel() for i, bs in enumerate(sizes)]
    pipe = eng.B200ForwardPipe().initialize(golden_weights_bin, 19, 4, gpus=[0], precision=eng.PRECISION_FP16)
    try:
        out = pipe.batch_forward(0, planes, sizes, [0, 0, 0])
        for i, bs in enumerate(sizes):
            assert out[i]["fp16"] == 1
            _check(out[i], orc.forward(planes[i], bs, 0), bs, atol=5e-2)
    finally:
        pipe.destroy()


def test_async_pinned_submit_wait_equals_blocking_call(golden_pipe, eng):
    from sayuri_b200 import synth
    n = 6
    sizes = [19] * n
    x = synth.synth_positions(n, 19, seed=123).reshape(n, -1)
    blocking = golden_pipe.batch_forward(0, list(x), sizes, [1] * n)
    pinned = eng.PinnedArray((2, n, eng.PLANE_FLOATS))
    try:
        pinned.array[0] = x
        pinned.array[1] = x[::-1]
        outs = [np.zeros(n, dtype=eng.OUTPUT_DTYPE) for _ in range(2)]
        golden_pipe.submit(0, 0, pinned.array[0], sizes, [1] * n)
        golden_pipe.submit(0, 1, pinned.array[1], sizes, [1] * n)
        with pytest.raises(RuntimeError, match="slot is busy"):
            golden_pipe.submit(0, 1, pinned.array[1], sizes, [1] * n)
        golden_pipe.wait(0, 0, outs[0])
        golden_pipe.wait(0, 1, outs[1])
        with pytest.raises(RuntimeError, match="slot is idle"):
            golden_pipe.wait(0, 1, outs[1])
        for i in range(n):
            assert np.array_equal(outs[0][i]["probabilities"], blocking[i]["probabilities"])
            assert np.array_equal(outs[1][n - 1 - i]["probabilities"], blocking[i]["probabilities"])
            assert np.array_equal(outs[1][n - 1 - i]["ownership"], blocking[i]["ownership"])
    finally:
        pinned.free()


def test_reload_weights_hot_swap(eng, oracle_lib):
    from sayuri_b200 import synth
    d = tempfile.gettempdir()
    a, b = os.path.join(d, "sb_reload_a.bin"), os.path.join(d, "sb_reload_b.bin")
    synth.write_synth_net(a, (2, 32, 8, 8), seed=1, stack=["ResidualBlock", "ResidualBlock-SE"])
    synth.write_synth_net(b, (2, 32, 8, 8), seed=2, stack=["ResidualBlock", "ResidualBlock-SE"])
    x = synth.synth_positions(1, 19, seed=5)[0].ravel()
    pipe = eng.B200ForwardPipe().initialize(a, 19, 4, gpus=[0])
    try:
        ca = pipe.weights_checksum()
        _check(pipe.forward(x, 19), oracle_lib.Oracle(a).forward(x, 19), 19)
        pipe.reload(b)
        assert pipe.weights_checksum() != ca
        _check(pipe.forward(x, 19), oracle_lib.Oracle(b).forward(x, 19), 19)
        c = os.path.join(d, "sb_reload_c.bin")
        synth.write_synth_net(c, (3, 32, 8, 8), seed=2)
        with pytest.raises(RuntimeError, match="same architecture"):
            pipe.reload(c)
    finally:
        pipe.destroy()


def test_full_size_config2_properties(eng, oracle_lib):
    """BASELINE config 2 at full size (10bx128, 19x19, batch 256): size-independent properties plus a sampled
    oracle comparison.  (i) duplicates inside the batch are bit-identical, (ii) a permutation of the batch
    permutes the outputs bit-exactly, (iii) 6 sampled positions match the CPU oracle within 1e-4."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_10bx128.bin")
    synth.write_synth_net(path, "10bx128", seed=20260417)
    n = 256
    x = synth.synth_positions(n, 19, seed=20260419).reshape(n, -1)
    x[200] = x[3]
    x[255] = x[3]
    pipe = eng.B200ForwardPipe().initialize(path, 19, n, gpus=[0])
    try:
        out = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
        for f in ("probabilities", "ownership", "wdl"):
            assert np.array_equal(out[200][f], out[3][f]) and np.array_equal(out[255][f], out[3][f])
        perm = np.random.default_rng(0).permutation(n)
        outp = pipe.batch_forward(0, list(x[perm]), [19] * n, [0] * n)
        assert np.array_equal(outp["probabilities"], out["probabilities"][perm])
        assert np.array_equal(outp["ownership"], out["ownership"][perm])
        assert np.array_equal(outp["pass_probability"], out["pass_probability"][perm])
        orc = oracle_lib.Oracle(path)
        for i in (0, 1, 100, 128, 254, 255):
            _check(out[i], orc.forward(x[i], 19, 0), 19)
        assert np.isfinite(out["probabilities"]).all()
    finally:
        pipe.destroy()


FIELDS = ("probabilities", "ownership", "pass_probability", "wdl", "stm_winrate", "final_score", "q_error", "score_error")


def test_packed_records_equal_fp32_staging_bit_exact_and_raw_fallback(golden_pipe, oracle_lib, golden_weights_bin):
    """The compact host->device records (sb_pack_position) are an EXACT encoding: outputs are bit-identical to the
    fp32 staging path; a position with a many-valued plane travels raw inside the same batch and still matches the
    oracle."""
    from sayuri_b200 import synth
    sizes = [19, 9, 13, 19, 13]
    planes = [synth.synth_positions(1, bs, seed=900 + i)[0].ravel().copy() for i, bs in enumerate(sizes)]
    offs = [0, 1, 2, 3, 4]
    golden_pipe.set_option("pack_inputs", 0)
    plain = golden_pipe.batch_forward(0, planes, sizes, offs)
    golden_pipe.set_option("pack_inputs", 1)
    packed = golden_pipe.batch_forward(0, planes, sizes, offs)
    for f in FIELDS:
        assert np.array_equal(plain[f], packed[f]), f
    # make sample 3 unpackable: a plane with several different non-zero values
    rng = np.random.default_rng(5)
    planes[3].reshape(43, -1)[7] = rng.uniform(-1, 1, 361).astype(np.float32)
    from sayuri_b200 import engine
    assert not engine.pack_position(planes[3], 19)[1]
    mixed = golden_pipe.batch_forward(0, planes, sizes, offs)
    golden_pipe.set_option("pack_inputs", 0)
    mixed_plain = golden_pipe.batch_forward(0, planes, sizes, offs)
    golden_pipe.set_option("pack_inputs", 1)
    for f in FIELDS:
        assert np.array_equal(mixed[f], mixed_plain[f]), f
    for i in (0, 1, 2, 4):
        assert np.array_equal(mixed[i]["probabilities"], packed[i]["probabilities"])
    orc = oracle_lib.Oracle(golden_weights_bin)
    _check(mixed[3], orc.forward(planes[3], 19, 3), 19)


def test_batcher_eval_from_many_threads_is_bit_identical_to_batch_forward(eng, golden_weights_bin):
    """sb_eval (NetworkForwardPipe::Forward from many threads, the engine's own batcher) returns, for every position,
    exactly what sb_forward_batch returns for it: batch composition is decided by thread timing and must not
    matter.  Also covers a raw-fallback position and mixed board sizes."""
    import threading
    from sayuri_b200 import synth
    sizes = [(19, 13, 9)[i % 3] for i in range(48)]
    planes = [synth.synth_positions(1, bs, seed=1200 + i)[0].ravel().copy() for i, bs in enumerate(sizes)]
    planes[5].reshape(43, -1)[11, :20] = np.linspace(0.1, 0.9, 20, dtype=np.float32)   # unpackable
    offs = [i % 5 for i in range(48)]
    pipe = eng.B200ForwardPipe().initialize(golden_weights_bin, 19, 16, gpus=[0])
    try:
        ref = np.concatenate([pipe.batch_forward(0, planes[i:i + 16], sizes[i:i + 16], offs[i:i + 16]) for i in (0, 16, 32)])
        pipe.batcher_config(batch_size=8, wait_us=2000)
        got = [None] * 48

        def worker(t):
            for i in range(t, 48, 12):
                got[i] = pipe.eval(planes[i], sizes[i], offs[i])

        for rounds in range(2):
            ts = [threading.Thread(target=worker, args=(t,)) for t in range(12)]
            [t.start() for t in ts]
            [t.join() for t in ts]
            for i in range(48):
                for f in FIELDS:
                    assert np.array_equal(got[i][f], ref[i][f]), (rounds, i, f)
                assert got[i]["board_size"] == sizes[i] and got[i]["offset"] == offs[i]
        st = pipe.batcher_stats()
        assert st["positions"] == 96 and st["raw"] == 2 and st["workers"] == 2 and st["batches"] >= 12
        with pytest.raises(RuntimeError, match="exceeds the NN canvas"):
            pipe.eval(np.zeros(43 * 21 * 21, dtype=np.float32), 21, 0)
        # single caller: closed by the timer, still exact
        one = pipe.eval(planes[0], sizes[0], offs[0])
        assert np.array_equal(one["probabilities"], ref[0]["probabilities"])
        # reconfigure stops the workers; the next eval restarts them on the new geometry
        pipe.construct(batch_size=32)
        again = pipe.eval(planes[1], sizes[1], offs[1])
        assert np.array_equal(again["ownership"], ref[1]["ownership"])
    finally:
        pipe.destroy()


def test_tail_wave_split_is_bit_exact(eng):
    """conv3x3_tc2 splits the work items of a partial last wave into N-halves (conv_unit): same MMAs per output
    element in the same order, so the results are bit-identical with the split on and off.  Batch 100 of a 128-wide
    net gives 157 items over 74 CTA pairs: 2 full waves + 9 items, which do get split."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_4bx128.bin")
    synth.write_synth_net(path, (4, 128, 16, 16), seed=77)
    n = 100
    x = synth.synth_positions(n, 19, seed=31).reshape(n, -1)
    for prec in (eng.PRECISION_FP32_SPLIT, eng.PRECISION_FP16):
        pipe = eng.B200ForwardPipe().initialize(path, 19, n, gpus=[0], precision=prec)
        try:
            pipe.set_option("tail_split", 1)
            a = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
            pipe.set_option("tail_split", 0)
            b = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
            for f in FIELDS:
                assert np.array_equal(a[f], b[f]), (prec, f)
            assert np.isfinite(a["probabilities"]).all()
        finally:
            pipe.destroy()


def test_scheduling_knobs_do_not_change_a_single_bit(eng):
    """Resident weights (fp16 rung), programmatic dependent launch, chains of convolutions in one launch (conv_chain), tile-by-tile
    dependencies between consecutive convolution launches (layer_overlap), the small-batch N-tile split, the tail-wave split,
    chained forwards and the form of the MMA issue loop only change WHEN and WHERE the same MMAs run: outputs are bit-identical
    with every knob on and off."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_4bx128.bin")
    synth.write_synth_net(path, (4, 128, 16, 16), seed=77)
    for prec in (eng.PRECISION_FP32_SPLIT, eng.PRECISION_FP16):
        for n in (3, 160):
            x = synth.synth_positions(n, 19, seed=32).reshape(n, -1)
            pipe = eng.B200ForwardPipe().initialize(path, 19, n, gpus=[0], precision=prec)
            try:
                base = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
                assert np.isfinite(base["probabilities"]).all()
                for knob in ("resident_weights", "pdl", "pdl_aux", "small_batch_split", "chain_forwards", "layer_overlap", "tail_split",
                             "conv_chain"):
                    # pdl_aux, layer_overlap: 0 off, 1 small batches only (the default), 2 always
                    for value in ((0, 2) if knob in ("pdl_aux", "layer_overlap") else (0,)):
                        pipe.set_option(knob, value)
                        other = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
                        pipe.set_option(knob, 1)
                        for f in FIELDS:
                            assert np.array_equal(base[f], other[f]), (prec, n, knob, value, f)
                # the unrolled single-thread MMA issuer of the 3x3 convolutions against the general issue loop (conv_dbg 128)
                pipe.set_option("conv_dbg", 128)
                other = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
                pipe.set_option("conv_dbg", 0)
                for f in FIELDS:
                    assert np.array_equal(base[f], other[f]), (prec, n, "general MMA issue loop", f)
            finally:
                pipe.destroy()


def test_se_pooling_from_the_conv_epilogue_and_the_separate_pass_agree_with_the_oracle(eng, oracle_lib):
    """The SE unit's pooling comes from partials the last convolution of the block writes in its epilogue (16-row groups on
    the 19x19 canvas, 4-row groups on 9x9 / 13x13 canvases, fixed-order finalize) or, with fuse_se_pool = 0 and on canvases
    with an odd row count per sample, from a separate pass over the tensor.  Both meet the 1e-4 bar against the oracle on
    every block family, and the fused form is batch-invariant (bit-exact) like the rest of the path."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_sepool.bin")
    stack = ["ResidualBlock-SE", "BottleneckBlock-SE", "MixerBlock-SE", "NestedBottleneckBlock-SE"]
    synth.write_synth_net(path, (4, 64, 16, 16), seed=5, stack=stack)
    orc = oracle_lib.Oracle(path)
    for canvas, sizes in ((19, [19, 9, 13, 19, 7]), (13, [13, 9, 13]), (9, [9, 9, 5]), (12, [12, 9])):
        planes = [synth.synth_positions(1, bs, seed=90 + i)[0].ravel() for i, bs in enumerate(sizes)]
        offs = [i % 5 for i in range(len(sizes))]
        refs = [orc.forward(planes[i], bs, offs[i]) for i, bs in enumerate(sizes)]
        pipe = eng.B200ForwardPipe().initialize(path, canvas, 8, gpus=[0])
        try:
            for fuse in (1, 0):
                pipe.set_option("fuse_se_pool", fuse)
                out = pipe.batch_forward(0, planes, sizes, offs)
                for i, bs in enumerate(sizes):
                    _check(out[i], refs[i], bs)
                if fuse:
                    alone = pipe.batch_forward(0, planes[:1], sizes[:1], offs[:1])
                    shifted = pipe.batch_forward(0, planes[1:] + planes[:1], sizes[1:] + sizes[:1], offs[1:] + offs[:1])
                    for f in FIELDS:
                        assert np.array_equal(alone[0][f], out[0][f]) and np.array_equal(shifted[-1][f], out[0][f]), (canvas, f)
        finally:
            pipe.destroy()


def test_chunked_accumulation_settings_all_meet_the_bar(eng, oracle_lib):
    """Split rung: the main accumulator is re-accumulated in fp32 RN after every k-half (option chunk_accumulate; 0 = the
    whole K in one TMEM accumulator, the round-1 behaviour), with and without the truncation compensation.  Every setting
    meets 1e-4 on a shallow net; the deep nets are what tests/test_gpu_fullnets.py pins with the defaults."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_chunks.bin")
    synth.write_synth_net(path, (3, 192, 16, 16), seed=9, stack=["ResidualBlock", "BottleneckBlock-SE", "ResidualBlock-SE"])
    orc = oracle_lib.Oracle(path)
    sizes = [19, 13, 9, 19]
    planes = [synth.synth_positions(1, bs, seed=190 + i)[0].ravel() for i, bs in enumerate(sizes)]
    refs = [orc.forward(planes[i], bs, 0) for i, bs in enumerate(sizes)]
    pipe = eng.B200ForwardPipe().initialize(path, 19, 4, gpus=[0])
    try:
        for chunk, comp in ((1, 12), (1, 0), (0, 0)):
            pipe.set_option("chunk_accumulate", chunk)
            pipe.set_option("acc_comp_ppb", comp)
            out = pipe.batch_forward(0, planes, sizes, [0] * 4)
            for i, bs in enumerate(sizes):
                _check(out[i], refs[i], bs)
    finally:
        pipe.destroy()


def test_bottleneck_and_nested_bottleneck_blocks_match_reference_golden(eng, golden_blocks, golden_blocks_weights):
    """BottleneckBlock[-SE] / NestedBottleneckBlock[-SE] towers (blas_forward_pipe.cc:90-263) against
    outputs of the UNMODIFIED compiled reference, every board size in one mixed batch, 1e-4."""
    pipe = eng.B200ForwardPipe().initialize(golden_blocks_weights, 19, 8, gpus=[0])
    try:
        d = pipe.net_desc()
        assert d["block_types"] == [eng.BLOCK_BOTTLENECK, eng.BLOCK_NESTED_BOTTLENECK, eng.BLOCK_RESIDUAL,
                                    eng.BLOCK_NESTED_BOTTLENECK, eng.BLOCK_BOTTLENECK]
        assert d["inner_channels"] == [16, 16, 0, 16, 16] and d["se_sizes"] == [8, 0, 0, 8, 0]
        planes, sizes, offsets, refs = [], [], [], []
        for bs in SIZES:
            for i in range(2):
                planes.append(golden_blocks["planes_%d" % bs][i].ravel())
                sizes.append(bs)
                offsets.append(int(golden_blocks["offset_%d_%d" % (bs, i)]))
                v = golden_blocks["ref_%d_%d" % (bs, i)]
                s = bs * bs
                refs.append(dict(prob=v[:s], own=v[s:2 * s], misc=v[2 * s:]))
        out = pipe.batch_forward(0, planes, sizes, offsets)
        for o, r, bs in zip(out, refs, sizes):
            _check(o, r, bs)
    finally:
        pipe.destroy()


@pytest.mark.parametrize("prec_name", ["fp32_split", "fp16"])
def test_wide_bottleneck_tower_matches_oracle(eng, oracle_lib, prec_name):
    """A 128-wide tower mixing all three block families (inner width 64: one 64-channel K block, 1x1 convs as
    single-tap launches) against the CPU oracle; also through sb_create with explicit block descriptors."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_btl_6bx128.bin")
    stack = ["NestedBottleneckBlock", "BottleneckBlock-SE", "ResidualBlock", "NestedBottleneckBlock-SE", "BottleneckBlock", "ResidualBlock-SE"]
    synth.write_synth_net(path, (6, 128, 16, 16), seed=91, stack=stack)
    prec = eng.PRECISION_FP32_SPLIT if prec_name == "fp32_split" else eng.PRECISION_FP16
    atol = ATOL if prec_name == "fp32_split" else 5e-2
    sizes = [19, 9, 13, 19, 19]
    planes = [synth.synth_positions(1, bs, seed=70 + i)[0].ravel() for i, bs in enumerate(sizes)]
    orc = oracle_lib.Oracle(path)
    pipe = eng.B200ForwardPipe().initialize(path, 19, 8, gpus=[0], precision=prec)
    try:
        out = pipe.batch_forward(0, planes, sizes, [0, 1, 2, 3, 4])
        for i, bs in enumerate(sizes):
            _check(out[i], orc.forward(planes[i], bs, i), bs, atol=atol)
    finally:
        pipe.destroy()


def test_mixer_blocks_and_replk_head_match_reference_golden(eng, golden_mixer, golden_mixer_weights):
    """Rest of SURVEY.md §8 a22: MixerBlock[-SE] (depthwise 7x7 / 5x5 + FFN, blas_forward_pipe.cc:265-312) and the RepLK
    policy head (:443-471) against outputs of the UNMODIFIED compiled reference, mixed board sizes in one batch."""
    pipe = eng.B200ForwardPipe().initialize(golden_mixer_weights, 19, 8, gpus=[0])
    try:
        d = pipe.net_desc()
        assert d["block_types"] == [eng.BLOCK_MIXER, eng.BLOCK_MIXER, eng.BLOCK_RESIDUAL, eng.BLOCK_MIXER]
        assert d["inner_channels"] == [48, 48, 0, 48] and d["dw_kernels"] == [7, 7, 0, 5]
        assert d["policy_head_type"] == eng.POLICY_HEAD_REPLK and d["policy_dw_kernel"] == 7
        planes, sizes, offsets, refs = [], [], [], []
        for bs in SIZES:
            for i in range(2):
                planes.append(golden_mixer["planes_%d" % bs][i].ravel())
                sizes.append(bs)
                offsets.append(int(golden_mixer["offset_%d_%d" % (bs, i)]))
                v = golden_mixer["ref_%d_%d" % (bs, i)]
                s = bs * bs
                refs.append(dict(prob=v[:s], own=v[s:2 * s], misc=v[2 * s:]))
        out = pipe.batch_forward(0, planes, sizes, offsets)
        for o, r, bs in zip(out, refs, sizes):
            _check(o, r, bs)
    finally:
        pipe.destroy()


@pytest.mark.parametrize("prec_name", ["fp32_split", "fp16"])
def test_wide_mixer_tower_with_replk_head_matches_oracle(eng, oracle_lib, prec_name):
    """128-wide Mixer tower (feed-forward width 192 = two N tiles of 96, depthwise 7x7) with P = 24 behind a RepLK head
    (the 1x1 P -> P writes 24 of 32 padded output columns and must leave the value channels alone)."""
    from sayuri_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "sb_test_mix_4bx128.bin")
    stack = ["MixerBlock", "MixerBlock-SE", "BottleneckBlock", "MixerBlock"]
    synth.write_synth_net(path, (4, 128, 24, 24), seed=93, stack=stack, policy_head="RepLK")
    prec = eng.PRECISION_FP32_SPLIT if prec_name == "fp32_split" else eng.PRECISION_FP16
    atol = ATOL if prec_name == "fp32_split" else 5e-2
    sizes = [19, 9, 13, 19]
    planes = [synth.synth_positions(1, bs, seed=80 + i)[0].ravel() for i, bs in enumerate(sizes)]
    orc = oracle_lib.Oracle(path)
    pipe = eng.B200ForwardPipe().initialize(path, 19, 4, gpus=[0], precision=prec)
    try:
        out = pipe.batch_forward(0, planes, sizes, [0, 1, 2, 3])
        for i, bs in enumerate(sizes):
            _check(out[i], orc.forward(planes[i], bs, i), bs, atol=atol)
    finally:
        pipe.destroy()


def test_batcher_stress_every_result_is_the_callers_own(eng, golden_weights_bin):
    """64 threads hammer sb_eval with 24 different positions for ~4 000 calls in total, tiny batches and a short timer,
    so batches close by both paths thousands of times: every returned record must be bit-identical to the reference
    record of the position that was passed in (no cross-talk between ring entries, no stale batch re-use)."""
    import threading
    from sayuri_b200 import synth
    n_pos = 24
    sizes = [(19, 13, 9)[i % 3] for i in range(n_pos)]
    planes = [synth.synth_positions(1, bs, seed=3000 + i)[0].ravel().copy() for i, bs in enumerate(sizes)]
    offs = [i % 5 for i in range(n_pos)]
    pipe = eng.B200ForwardPipe().initialize(golden_weights_bin, 19, 16, gpus=[0])
    try:
        ref = np.concatenate([pipe.batch_forward(0, planes[i:i + 12], sizes[i:i + 12], offs[i:i + 12]) for i in (0, 12)])
        pipe.batcher_config(batch_size=5, wait_us=50)
        bad = []

        def worker(t):
            rng = np.random.default_rng(t)
            for _ in range(64):
                i = int(rng.integers(n_pos))
                o = pipe.eval(planes[i], sizes[i], offs[i])
                if not (np.array_equal(o["probabilities"], ref[i]["probabilities"]) and np.array_equal(o["ownership"], ref[i]["ownership"])
                        and o["pass_probability"] == ref[i]["pass_probability"] and o["board_size"] == sizes[i] and o["offset"] == offs[i]):
                    bad.append((t, i))

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(64)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not bad, bad[:5]
        st = pipe.batcher_stats()
        assert st["positions"] == 64 * 64 and st["full"] > 0 and st["timer"] > 0
    finally:
        pipe.destroy()


@pytest.mark.parametrize("fixture", ["ref_3bx32.bin.txt", "ref_btl_5bx32.bin.txt", "ref_mix_4bx32.bin.txt"])
def test_create_from_tensors_equals_create_from_file(eng, golden_weights_bin, fixture):
    """sb_create (net description + folded tensors in loader order: what the C++ shim passes from DNNWeights) builds the
    same engine as sb_create_from_file, for every block family and both policy heads: bit-identical outputs and blob."""
    from sayuri_b200 import synth
    path = os.path.join(os.path.dirname(golden_weights_bin), fixture)
    desc, tensors = eng.load_weights_file(path)
    sizes = [19, 13, 9, 19]
    planes = [synth.synth_positions(1, bs, seed=600 + i)[0].ravel() for i, bs in enumerate(sizes)]
    a = eng.B200ForwardPipe().initialize(path, 19, 4, gpus=[0])
    b = eng.B200ForwardPipe().initialize_from_tensors(desc, tensors, 19, 4, gpus=[0])
    try:
        assert a.weights_checksum() == b.weights_checksum()
        oa = a.batch_forward(0, planes, sizes, [0, 1, 2, 3])
        ob = b.batch_forward(0, planes, sizes, [0, 1, 2, 3])
        for f in FIELDS:
            assert np.array_equal(oa[f], ob[f]), f
        assert a.net_desc() == b.net_desc()
    finally:
        a.destroy()
        b.destroy()
