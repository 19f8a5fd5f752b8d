"""The whole hot path in the loop (BASELINE configs[3] / configs[4] data flow): depth frame -> encoder -> latent ->
reference generation -> SQP-RTI or full SQP -> plant -> collision label, for a handful of environments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', ['rti', 'sqp'])
def test_rollout_runs_and_is_shard_invariant(mode, cuda_device):
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.rollout import REC_WIDTH, rollout
    cfg = Config()
    if mode == 'sqp':
        cfg.sensor.is_depth, cfg.sensor.is_spherical = False, True
        cfg.sensor.vfov = cfg.sensor.hfov / cfg.sensor.aspect_ratio
    net, enc = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True), synth.make_encoder_net(seed=5)
    kw = dict(mode=mode, encode_every=1 if mode == 'sqp' else 0, device=cuda_device)
    rec = rollout(cfg, net, enc, 12, 6, first=0, **kw).cpu().numpy()
    assert rec.shape == (12, REC_WIDTH) and np.isfinite(rec).all()
    assert rec[:, 10].sum() == 0                        # no failed solves
    assert (rec[:, 15] < 4.0).all()                     # every environment moved towards its goal (4 m ahead)
    assert (rec[:, 14] >= 1.0).all() and (mode == 'rti') == bool((rec[:, 14] == 1.0).all())
    # shard invariance: environments [4, 12) run alone reproduce their rows (what the multi-GPU partition relies on)
    part = rollout(cfg, net, enc, 8, 6, first=4, **kw).cpu().numpy()
    assert np.array_equal(part, rec[4:])


def test_captured_rollout_step_is_bit_identical(cuda_device):
    """The RTI rollout step replayed as one CUDA graph gives exactly the records of the eager loop."""
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.rollout import rollout
    cfg = Config()
    net, enc = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True), synth.make_encoder_net(seed=5)
    a = rollout(cfg, net, enc, 16, 8, mode='rti', encode_every=0, device=cuda_device).cpu().numpy()
    b = rollout(cfg, net, enc, 16, 8, mode='rti', encode_every=0, device=cuda_device, graph=True).cpu().numpy()
    assert np.array_equal(a, b)
