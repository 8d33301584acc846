"""CPU-side checks of the product library: it loads, exports every symbol include/sayuri_b200.h declares,
parses weight files exactly like the oracle, and fails LOUDLY without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "sayuri_b200", "libsayuri_b200.so")):
        g.build()
    from sayuri_b200 import engine
    return engine


def test_library_exports_every_declared_symbol(eng):
    header = open(os.path.join(ROOT, "include", "sayuri_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = eng.load_library()
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert sorted(eng.ABI_SYMBOLS) == declared


def test_output_struct_layout_matches_header(eng):
    assert ctypes.sizeof(eng.SbOutput) == (361 + 361 + 8) * 4 + 3 * 4
    assert eng.OUTPUT_DTYPE.itemsize == ctypes.sizeof(eng.SbOutput)
    assert eng.OUTPUT_DTYPE.fields["ownership"][1] == 361 * 4
    assert eng.OUTPUT_DTYPE.fields["board_size"][1] == (361 + 361 + 8) * 4


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(eng, golden_weights_bin):
    pipe = eng.B200ForwardPipe()
    with pytest.raises(RuntimeError, match="No executable GPU device"):
        pipe.initialize(golden_weights_bin, 19, 8)
    assert not pipe.valid()


def test_loader_errors_are_reported_before_touching_cuda(eng, tmp_path):
    from sayuri_b200 import synth
    pipe = eng.B200ForwardPipe()
    with pytest.raises(RuntimeError, match="Couldn't open weights file"):
        pipe.initialize(str(tmp_path / "missing.txt"), 19, 8)
    info, layers = synth.synth_tensors(1, 32, 8, 8, seed=0, stack=["ResidualBlock"])
    p = tmp_path / "bad.txt"
    synth.write_weights(str(p), dict(info, stack=["TransformerBlock"]), layers)
    with pytest.raises(RuntimeError, match="not supported by sayuri_b200"):
        pipe.initialize(str(p), 19, 8)
    synth.write_weights(str(p), dict(info, stack=["NestedBottleneckBlock"]), layers)   # stack and struct disagree
    with pytest.raises(RuntimeError, match="Fail to load the network file"):
        pipe.initialize(str(p), 19, 8)
    layers[2][1][0] = layers[2][1][0].ravel()[:-3]
    synth.write_weights(str(p), info, layers)
    with pytest.raises(RuntimeError, match="tensor size mismatch"):
        pipe.initialize(str(p), 19, 8)
    synth.write_weights(str(p), info, synth.synth_tensors(1, 32, 8, 8, seed=0, stack=["ResidualBlock"])[1], version=6)
    with pytest.raises(RuntimeError, match="do not support this version"):
        pipe.initialize(str(p), 19, 8)


def test_invalid_arguments_rejected(eng, golden_weights_bin):
    pipe = eng.B200ForwardPipe()
    with pytest.raises(RuntimeError, match="board size"):
        pipe.initialize(golden_weights_bin, 25, 8)
    with pytest.raises(RuntimeError, match="batch size"):
        pipe.initialize(golden_weights_bin, 19, 0)
    lib = eng.load_library()
    assert lib.sb_num_gpus(None) == 0
    assert lib.sb_submit(None, 0, 0, 1, None, 1, None, None) != 0
    assert lib.sb_wait(None, 0, 0, None) != 0


@pytest.mark.parametrize("which", ["residual", "bottleneck", "mixer", "synthetic"])
def test_engine_weight_reader_equals_oracle_reader_bit_exact(eng, oracle_lib, golden_weights_bin, golden_weights_txt, tmp_path, which):
    """The product's own file reader + BN folding (host_net.cc) against the oracle's (oracle_forward.c), tensor by tensor,
    bit for bit, for every block family / the RepLK head, text and binary formats (host only: runs without a GPU)."""
    from sayuri_b200 import synth
    golden_dir = os.path.dirname(golden_weights_bin)
    if which == "residual":
        paths = [golden_weights_bin, golden_weights_txt]
    elif which == "bottleneck":
        paths = [os.path.join(golden_dir, "ref_btl_5bx32.bin.txt")]
    elif which == "mixer":
        paths = [os.path.join(golden_dir, "ref_mix_4bx32.bin.txt")]
    else:
        p = str(tmp_path / "synth.txt")
        synth.write_synth_net(p, (3, 48, 8, 24), seed=5, binary=False, stack=["MixerBlock-SE", "NestedBottleneckBlock", "BottleneckBlock-SE"],
                              policy_head="RepLK", dw_kernel=5)
        paths = [p]
    for path in paths:
        desc, tensors = eng.load_weights_file(path)
        o = oracle_lib.Oracle(path)
        ref = [t for pair in o.tensors() for t in pair]
        assert len(tensors) == len(ref)
        for i, (a, b) in enumerate(zip(tensors, ref)):
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (path, i)
        assert (desc["blocks"], desc["channels"], desc["P"], desc["V"], desc["activation"]) == (o.blocks, o.channels, o.P, o.V, o.act)
        assert sum(1 for s in desc["se_sizes"] if s > 0) == o.n_se
