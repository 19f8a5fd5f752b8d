"""Development: CUDA-event timing of the encoder and the decoder at a few batch sizes (ms per call, best of 5 after warm-up)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.vae import VaeWrapper


def best(fn, n=5, inner=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b) / inner)
    return min(t)


enc, dec = synth.make_encoder_net(seed=5), synth.make_decoder_net(seed=7)
for B in [int(a) for a in sys.argv[1:]] or [1, 64, 256]:
    vae = VaeWrapper(Config(), enc, max_batch=B, dec_net=dec)
    frames = np.stack([synth.depth_frame(seed=i % 4) for i in range(B)])
    vae.set_img(frames)
    te = best(lambda: vae.encode(to_numpy=False))
    td = best(lambda: vae.decode(to_numpy=False))
    print(f'B={B}: encode {te:.3f} ms ({B / te:.1f} k frames/s), decode {td:.3f} ms ({B / td:.1f} k frames/s)', flush=True)
    del vae
