"""The shipped library is Blackwell-native at the instruction level: checked here, without a GPU, on the SASS of the built
sayuri_b200/libsayuri_b200.so (`cuobjdump -sass`; B200_PROFILING.md, "What proves a Blackwell-native kernel")."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sayuri_b200", "libsayuri_b200.so")


@pytest.fixture(scope="module")
def conv_kernels():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump is not installed")
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper(), "the library holds no sm_100a code"
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    conv = {f.split("\n")[0]: f for f in funcs if "conv3x3_tc2_kernel" in f.split("\n")[0]}
    assert conv, "conv3x3_tc2_kernel is not in the library"
    return conv, funcs


def test_every_conv_instantiation_uses_tcgen05_tmem_and_tma(conv_kernels):
    conv, _ = conv_kernels
    # 8 activations x {split rung, fp16 rung at 4 and 2 epilogue parts} + 3 pooled variants
    assert len(conv) >= 27, sorted(conv)
    for name, body in conv.items():
        assert len(re.findall(r"UTCHMMA\.2CTA", body)) >= 40, name          # tcgen05.mma.cta_group::2 (general + unrolled issue code)
        assert re.search(r"UTCBAR\.2CTA\.MULTICAST", body), name            # tcgen05.commit, multicast to both CTAs of the pair
        assert re.search(r"\bLDTM", body), name                             # tcgen05.ld
        assert re.search(r"UTMALDG\.2D\.2CTA", body) and re.search(r"UTMALDG\.3D\.2CTA", body), name   # TMA weight stages / slabs
        assert re.search(r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT", body), name  # mbarrier pipeline
        assert re.search(r"LDG\.E\.STRONG\.GPU", body) and re.search(r"REDG\.E\.ADD\.S32\.STRONG\.GPU", body), name   # tile counters


def test_no_legacy_tensor_core_path_anywhere(conv_kernels):
    _, funcs = conv_kernels
    for f in funcs:
        name = f.split("\n")[0]
        assert not re.search(r"\bHMMA\b|\bHGMMA\b|\bIMMA\b", f), "legacy mma.sync / wgmma instruction in " + name
