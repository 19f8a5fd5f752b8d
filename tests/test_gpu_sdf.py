"""GPU parity of the fused SDF value+gradient kernel, through the C ABI.

Tolerances are north_star's: value 1e-5, gradient 1e-4, max-norm relative, against
(a) the committed outputs of the reference's own torch NeuralDF (fp32 and fp64) and (b) the oracle.
"""
import numpy as np
import pytest

from conftest import golden_sdf_inputs, relmax, sdf_cases

pytestmark = pytest.mark.gpu
VAL_TOL, GRAD_TOL = 1e-5, 1e-4


@pytest.fixture(scope='module')
def torch_mod(cuda_device):
    import torch
    return torch


@pytest.mark.parametrize('name', list(sdf_cases().keys()))
def test_sdf_matches_reference_fixture(name, torch_mod, cuda_device):
    torch = torch_mod
    from sdf_nmpc_b200.sdf import SdfEvaluator
    net, pts, lat, idx, G = golden_sdf_inputs(name)
    ev = SdfEvaluator(net, cuda_device)
    ev.set_latents(torch.from_numpy(lat).cuda())
    d, g = ev.eval(torch.from_numpy(pts).cuda(), pt2inst=torch.from_numpy(idx).cuda())
    d, g = d.cpu().numpy(), g.cpu().numpy()
    assert relmax(d, G['d_f32']) <= VAL_TOL, (name, relmax(d, G['d_f32']))
    assert relmax(g, G['g_f32']) <= GRAD_TOL, (name, relmax(g, G['g_f32']))
    assert relmax(d, G['d_f64']) <= VAL_TOL
    assert relmax(g, G['g_f64']) <= GRAD_TOL
    # the contiguous-block latent mapping (pts_per_inst) must agree with the explicit index
    n, n_lat = len(pts), lat.shape[0]
    d2, g2 = ev.eval(torch.from_numpy(pts).cuda(), pts_per_inst=n // n_lat)
    assert np.array_equal(d2.cpu().numpy(), d) and np.array_equal(g2.cpu().numpy(), g)


@pytest.mark.parametrize('n', [1, 7, 31, 32, 33, 1000, 4097])
def test_sdf_ragged_sizes_vs_oracle(n, torch_mod, cuda_device):
    torch = torch_mod
    from oracle import sdf as osdf
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.sdf import SdfEvaluator
    net = synth.make_sdf_net(seed=2, **synth.ARCH_B, bias_scale=0.05)
    pts = synth.sdf_query_points(n, seed=n)
    z = synth.latent(seed=n + 1)
    ev = SdfEvaluator(net, cuda_device)
    ev.set_latents(torch.from_numpy(z).cuda())
    d, g = ev.eval(torch.from_numpy(pts).cuda())
    dr, gr = osdf.value_and_grad(net, pts, z)
    scale_d, scale_g = np.abs(dr).max(), np.abs(gr).max()
    assert np.abs(d.cpu().numpy() - dr).max() <= VAL_TOL * max(scale_d, 1e-3)
    assert np.abs(g.cpu().numpy() - gr).max() <= GRAD_TOL * max(scale_g, 1e-3)
    # value-only call (served by the CUDA-core kernel, the value+gradient call by the tcgen05 kernel for n >= 64:
    # two implementations of the same operator, both inside the parity tolerance)
    d_only, none = ev.eval(torch.from_numpy(pts).cuda(), want_grad=False)
    assert none is None and np.abs(d_only.cpu().numpy() - dr).max() <= VAL_TOL * max(scale_d, 1e-3)


@pytest.mark.parametrize('impl', ['ffma', 'tc'])
def test_both_kernel_variants_meet_parity(impl, torch_mod, cuda_device, monkeypatch):
    """SDFNMPC_SDF_IMPL pins the CUDA-core or the tensor-core kernel; each must meet the tolerance on its own."""
    torch = torch_mod
    from sdf_nmpc_b200.sdf import SdfEvaluator
    monkeypatch.setenv('SDFNMPC_SDF_IMPL', impl)
    for name in ('archB', 'archA_bias'):
        net, pts, lat, idx, G = golden_sdf_inputs(name)
        ev = SdfEvaluator(net, cuda_device)
        ev.set_latents(torch.from_numpy(lat).cuda())
        d, g = ev.eval(torch.from_numpy(pts).cuda(), pt2inst=torch.from_numpy(idx).cuda())
        assert relmax(d.cpu().numpy(), G['d_f32']) <= VAL_TOL and relmax(g.cpu().numpy(), G['g_f32']) <= GRAD_TOL
        assert relmax(d.cpu().numpy(), G['d_f64']) <= VAL_TOL and relmax(g.cpu().numpy(), G['g_f64']) <= GRAD_TOL


def test_empty_input_and_errors(torch_mod, cuda_device):
    torch = torch_mod
    from sdf_nmpc_b200 import _lib, synth
    from sdf_nmpc_b200.sdf import SdfEvaluator
    net = synth.make_sdf_net(seed=2, **synth.ARCH_A)
    ev = SdfEvaluator(net, cuda_device)
    with pytest.raises(_lib.SdfNmpcError):          # eval before set_latents
        ev.eval(torch.zeros(4, 3, device='cuda'))
    ev.set_latents(torch.zeros(1, 128, device='cuda'))
    d, g = ev.eval(torch.zeros(0, 3, device='cuda'))
    assert d.shape == (0,) and g.shape == (0, 3)
    with pytest.raises(_lib.SdfNmpcError):          # more instances addressed than latents folded
        ev.eval(torch.zeros(8, 3, device='cuda'), pts_per_inst=2)


def test_host_entry_and_casadi_abi(torch_mod, cuda_device):
    """sdf_l4c / jac_sdf_l4c called exactly as CasADi-generated code calls them (double in/out, batch 1)."""
    from oracle import sdf as osdf
    from sdf_nmpc_b200.sdf import SdfEvaluator
    net, pts, lat, idx, G = golden_sdf_inputs('archB')
    ev = SdfEvaluator(net, cuda_device)
    d, g = ev.eval_host(pts[:100], lat[0])
    assert relmax(d, G['d_f32'][:100]) <= VAL_TOL and relmax(g, G['g_f32'][:100]) <= GRAD_TOL
    ev.bind_l4c()
    lib = ev.lib
    assert lib.sdf_l4c_n_in() == 1 and lib.sdf_l4c_n_out() == 1 and lib.jac_sdf_l4c_n_in() == 2
    sp = lib.sdf_l4c_sparsity_in(0)
    assert (sp[0], sp[1], sp[2]) == (131, 1, 1)
    sp = lib.jac_sdf_l4c_sparsity_out(0)
    assert (sp[0], sp[1], sp[2]) == (1, 131, 1)
    assert lib.jac_sdf_l4c_name_out(0) == b'jac_o0_i0'
    scale_g = np.abs(G['jacrev_f32'][:, :3]).max()
    for i in range(4):
        i0 = np.concatenate([pts[i], lat[idx[i]]]).astype(np.float64)
        val, jac = ev.l4c_call(i0)
        vo, go = osdf.l4casadi_call(net, i0)
        assert abs(val - G['d_f32'][i]) <= VAL_TOL * np.abs(G['d_f32']).max()
        assert np.abs(jac[:3] - G['jacrev_f32'][i, :3]).max() <= GRAD_TOL * scale_g
        assert abs(val - vo) <= VAL_TOL * np.abs(G['d_f32']).max() and np.abs(jac[:3] - go).max() <= GRAD_TOL * scale_g
        assert jac.shape == (131,) and np.all(jac[3:] == 0.0)


def test_full_size_properties(torch_mod, cuda_device):
    """Size-independent checks at microbench scale (1e6 points): determinism, tile-boundary independence
    (a point's result does not depend on its position in the batch) and a finite-difference check of the
    gradient on a random subset."""
    torch = torch_mod
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.sdf import SdfEvaluator
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    ev = SdfEvaluator(net, cuda_device)
    ev.set_latents(torch.from_numpy(synth.latent(seed=2)).cuda())
    n = 1_000_000
    pts = torch.from_numpy(synth.sdf_query_points(n, seed=3)).cuda()
    d1, g1 = ev.eval(pts)
    d2, g2 = ev.eval(pts)
    assert torch.equal(d1, d2) and torch.equal(g1, g2)
    perm = torch.randperm(n, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0))
    dp, gp = ev.eval(pts[perm])
    assert torch.equal(dp, d1[perm]) and torch.equal(gp, g1[perm])
    assert torch.isfinite(d1).all() and torch.isfinite(g1).all()
    sub = pts[:4096]
    eps = 1e-3
    for ax in range(3):
        e = torch.zeros(3, device='cuda'); e[ax] = eps
        dpl, _ = ev.eval(sub + e, want_grad=False)
        dmi, _ = ev.eval(sub - e, want_grad=False)
        fd = (dpl - dmi) / (2 * eps)
        assert (fd - g1[:4096, ax]).abs().max().item() <= 2e-2 * g1.abs().max().item()
