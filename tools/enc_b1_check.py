import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.vae import VaeWrapper
from oracle import encoder as OE
net = synth.make_encoder_net(seed=5)
frames = np.stack([synth.depth_frame(seed=i) for i in range(2)])
vae = VaeWrapper(Config(), net, max_batch=2)
vae.set_img(frames); lat = vae.encode()
vae.set_img(frames[1]); one = vae.encode()
ref = OE.encode(net, OE.preprocess(frames))
r = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
print('batch vs oracle', r(lat, ref), 'single vs oracle', r(one[0], ref[1]), 'single vs batch row', r(one[0], lat[1]))
