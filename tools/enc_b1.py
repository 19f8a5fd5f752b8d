import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.vae import VaeWrapper
for B in (1, 2, 64):
    vae = VaeWrapper(Config(), synth.make_encoder_net(seed=5), max_batch=B)
    vae.set_img(np.stack([synth.depth_frame(seed=i % 4) for i in range(B)]))
    for _ in range(3): vae.encode(to_numpy=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): vae.encode(to_numpy=False)
    e1.record(); torch.cuda.synchronize()
    print('B', B, 'ms', e0.elapsed_time(e1) / 20)
