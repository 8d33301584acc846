"""Host logic of the compact position records (sb_pack_position / sb_unpack_position, sayuri_b200/csrc/host_pack.cc):
pure CPU code inside libsayuri_b200.so, so it runs without a GPU."""
import numpy as np
import pytest

from sayuri_b200 import engine, synth


@pytest.mark.parametrize("bs", [2, 7, 9, 13, 19])
def test_pack_roundtrip_is_exact(bs):
    x = synth.synth_positions(8, bs, seed=100 + bs)
    for i in range(8):
        rec, ok = engine.pack_position(x[i], bs, i % 5)
        assert ok
        assert rec.board_size == bs and rec.offset == i % 5 and rec.flags == 0
        back = engine.unpack_position(rec)
        assert np.array_equal(back, x[i].reshape(43, -1))
        # bits beyond bs*bs stay clear
        bits = np.ctypeslib.as_array(rec.bits).reshape(43, engine.PACKED_WORDS)
        n = bs * bs
        for w in range(engine.PACKED_WORDS):
            lo = w * 32
            valid = max(0, min(32, n - lo))
            mask = (1 << valid) - 1
            assert not np.any(bits[:, w] & ~np.uint32(mask))


def test_pack_scalar_planes_and_negative_zero():
    x = synth.synth_positions(1, 19, seed=1)[0].reshape(43, -1).copy()
    x[38] = -0.375          # a board-constant plane with a negative value
    x[2, 17] = -0.0         # negative zero is a zero
    rec, ok = engine.pack_position(x, 19)
    assert ok and rec.scale[38] == np.float32(-0.375)
    back = engine.unpack_position(rec)
    assert np.array_equal(back, np.where(x == 0, np.float32(0), x))


@pytest.mark.parametrize("pos", [0, 5, 31, 32, 200, 352, 360])
def test_pack_refuses_many_valued_planes_and_nan(pos):
    x = synth.synth_positions(1, 19, seed=2)[0].reshape(43, -1).copy()
    x[4] = 0
    x[4, 100] = 1.0
    y = x.copy()
    y[4, pos] = 0.5 if pos != 100 else 1.0
    y[4, (pos + 7) % 361] = 0.25
    rec, ok = engine.pack_position(y, 19)
    assert not ok and rec.flags == engine.PACKED_RAW
    z = x.copy()
    z[20, pos] = np.nan
    assert not engine.pack_position(z, 19)[1]
    with pytest.raises(ValueError):
        engine.unpack_position(rec)


def test_pack_rejects_bad_arguments():
    x = np.zeros(43 * 400, dtype=np.float32)
    assert not engine.pack_position(x, 1)[1]
    assert not engine.pack_position(x, 20)[1]
    with pytest.raises(ValueError):
        engine.pack_position(x[:100], 19)
