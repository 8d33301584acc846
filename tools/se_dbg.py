import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
path = os.path.join(tempfile.gettempdir(), "sed.bin"); synth.write_synth_net(path, "10bx128", seed=20260417)
pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
b = 256
for prec in (0, 1):
    pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=prec)
    planes = [pos[i % 64] for i in range(b)]
    pipe.batch_forward(0, planes, [19]*b, [0]*b)
    for dbg in (0, 256, 512, 768):
        pipe.set_option("conv_dbg", dbg)
        sys.stderr.write("precision %d dbg %d\n" % (prec, dbg)); sys.stderr.flush()
        pipe.time_forward(0, 0, 0, flush_l2=True, profile_conv=True)
    pipe.destroy()
