"""Development: run the encoder once at batch B (for ncu launch lists)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.vae import VaeWrapper
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vae = VaeWrapper(Config(), synth.make_encoder_net(seed=5), max_batch=B)
frames = np.stack([synth.depth_frame(seed=i % 4) for i in range(B)])
vae.set_img(frames)
for _ in range(2):
    lat = vae.encode(to_numpy=False)
torch.cuda.synchronize()
print('ok', float(lat.sum()))
