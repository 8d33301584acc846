"""profiles/<name>: instruction evidence of the shipped library, from `cuobjdump -sass` (runs without a GPU).
   python tools/sass_evidence.py profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sayuri_b200", "libsayuri_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_evidence.md")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
pat = collections.OrderedDict([
    ("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("UTCBAR.2CTA.MULTICAST", r"UTCBAR\.2CTA\.MULTICAST"), ("LDTM", r"\bLDTM"),
    ("UTMALDG.2D.2CTA", r"UTMALDG\.2D\.2CTA"), ("UTMALDG.3D.2CTA", r"UTMALDG\.3D\.2CTA"), ("UTMASTG", r"UTMASTG"),
    ("SYNCS.PHASECHK (mbarrier try_wait)", r"SYNCS\.PHASECHK"), ("UCGABAR (cluster barrier)", r"UCGABAR"),
    ("REDG...STRONG.GPU (tile counters)", r"REDG\.E\.ADD\.S32\.STRONG\.GPU"), ("LDG.E.STRONG.GPU (counter polls)", r"LDG\.E\.STRONG\.GPU"),
    ("FENCE.VIEW.ASYNC.G (proxy fence)", r"FENCE\.VIEW\.ASYNC\.G"), ("ACQBULK / griddepcontrol.wait", r"ACQBULK"),
    ("HMMA (legacy mma.sync)", r"\bHMMA"), ("HGMMA (wgmma)", r"HGMMA")])
total = collections.Counter()
conv = [f for f in funcs if "conv3x3_tc2_kernel" in f.split("\n")[0]]
for f in funcs:
    for k, rx in pat.items():
        total[k] += len(re.findall(rx, f))
lines = ["# SASS evidence of the shipped library (`cuobjdump -sass sayuri_b200/libsayuri_b200.so`, tools/sass_evidence.py)", "",
         "Instruction counts over the whole library (%d kernels, %d instantiations of conv3x3_tc2_kernel<SPLIT, ACT, POOL, PARTS>):" % (len(funcs), len(conv)),
         "```"]
for k in pat:
    lines.append("%-40s %d" % (k, total[k]))
lines += ["```",
          "tcgen05.mma -> UTCHMMA (cta_group::2), tcgen05.commit multicast -> UTCBAR.2CTA.MULTICAST, tcgen05.ld -> LDTM,",
          "cp.async.bulk.tensor (cta_group::2) -> UTMALDG.{2D,3D}.2CTA; no legacy HMMA, no UTMASTG (the epilogue stores 16-byte pieces",
          "straight from registers: with the C8 layout a warp's 32 rows are 32 consecutive pieces).", ""]


def body(f):
    return [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]


for tag, rx in (("split, mish, PARTS 4", r"ILb1ELi5ELb0ELi4E"), ("fp16, mish, PARTS 4", r"ILb0ELi5ELb0ELi4E")):
    f = next(f for f in conv if re.search(rx, f.split("\n")[0]))
    b = body(f)
    idx = [i for i, l in enumerate(b) if "UTCHMMA" in l]
    # instructions between consecutive MMA blocks of the unrolled 9-tap issuer (the fast path: every barrier wait succeeds)
    gaps = [idx[i + 1] - idx[i] for i in range(len(idx) - 1) if idx[i + 1] - idx[i] > 4]
    r2ur = sum("R2UR" in l for l in b[idx[0]:idx[-1]])
    lines += ["conv3x3_tc2_kernel<%s>: %d SASS instructions, %d UTCHMMA; between the first and the last UTCHMMA %d R2UR moves "
              "(the loop state of the issuing warp lives in uniform registers); instructions between consecutive MMA blocks: median %d." % (
                  tag, len(b), len(idx), r2ur, sorted(gaps)[len(gaps) // 2] if gaps else 0), ""]
    k = idx[len(idx) // 2]
    while k > 0 and "UTCHMMA" in b[k - 1] or (k > 1 and "UTCHMMA" in b[k - 2]):
        k -= 1
    lines += ["One tap of the unrolled issue code (%s): descriptor arithmetic in the uniform datapath, 4 K-steps of 16%s, the multicast commit that"
              " releases the weight stage:" % (tag, " into the main and the low-order accumulator" if "split" in tag else ""), "```"]
    lines += [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip()[:150] for l in b[max(k - 14, 0):k + 14]]
    lines += ["```", ""]
open(out_path, "w").write("\n".join(lines) + "\n")
print("wrote", out_path)
