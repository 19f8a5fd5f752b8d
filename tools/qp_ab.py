"""Kernel A/B experiments (development aid): build variants of the library with extra -D flags and time the RTI step.

  here:        python tools/qp_ab.py build  tagA=-DFOO tagB="-DBAR -DBAZ" base=
  on the box:  python tools/qp_ab.py run    [B]        -> gpurun_out/qp_ab.json
Variants are written to sdf_nmpc_b200/lib/variants/ (git-ignored with the rest of lib/; they travel with gpurun)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, 'sdf_nmpc_b200', 'lib', 'variants')


def build(specs):
    from sdf_nmpc_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        tag, flags = spec.split('=', 1)
        objs = []
        for src in B.SOURCES:
            obj = os.path.join(VDIR, f'{tag}_{src[:-3]}.o')
            base_obj = os.path.join(B.LIBDIR, src.replace('.cu', '.o'))
            tu = 'sdf_api.cu' if tag.startswith('sdf') else ('encoder_api.cu' if tag.startswith('enc') else 'nmpc_api.cu')
            if src != tu and os.path.exists(base_obj):
                objs.append(base_obj)          # only one TU depends on the experiment flags (tags 'sdf*': the SDF TU)
                continue
            cmd = [B._nvcc()] + B.NVCC_FLAGS + flags.split() + ['-c', os.path.join(B.CSRC, src), '-o', obj]
            subprocess.run(cmd, check=True)
            objs.append(obj)
        lib = os.path.join(VDIR, f'lib_{tag}.so')
        subprocess.run([B._nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', lib] + objs +
                       ['-lcudart_static', '-ldl', '-lrt', '-lpthread'], check=True)
        print('built', lib)


def one(Bn):
    import torch
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    w = synth.nmpc_instances(Bn)
    mpc = BatchedNmpc(Config(), net, Bn)
    mpc.set_sdf_flag(1.0); mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
    mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
    x = torch.from_numpy(w['x0']).cuda()
    def step():
        nonlocal x
        mpc.set_x0(x); mpc.solve()
        x = mpc.plant_step(x.clone(), mpc.get_u(), 0.075)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    mpc.timing_begin(10)
    its = 0.0
    for _ in range(10):
        step()
        its += float(mpc.qp_iters.double().mean().item())
    ms, n = mpc.timing_end()
    from sdf_nmpc_b200.sdf import SdfEvaluator
    ev = SdfEvaluator(net)
    ev.set_latents(torch.from_numpy(synth.latent(seed=2)).cuda())
    n = 1_000_000
    pts = torch.from_numpy(synth.sdf_query_points(n, seed=3)).cuda()
    sdf = torch.empty(n, device='cuda'); grad = torch.empty(n, 3, device='cuda')
    for _ in range(3):
        ev.eval(pts, out=(sdf, grad))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ev.eval(pts, out=(sdf, grad))
    e1.record(); torch.cuda.synchronize()
    sweep = n / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e6
    from sdf_nmpc_b200.vae import VaeWrapper
    import numpy as np
    vae = VaeWrapper(Config(), synth.make_encoder_net(seed=5), max_batch=64)
    vae.set_img(np.stack([synth.depth_frame(seed=i % 4) for i in range(64)]))
    for _ in range(2):
        vae.encode(to_numpy=False)
    e0.record()
    for _ in range(5):
        vae.encode(to_numpy=False)
    e1.record(); torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1) / 5
    print(json.dumps(dict(enc64_ms=enc_ms, sdf_mpts=sweep, qp_ms=ms['qp'] / n, lin_ms=ms['linearize'] / n, sdf_ms=ms['sdf'] / n, iters=its / 10,
                          ok=int((mpc.status == 0).sum().item()))))


def run(Bn):
    out = {}
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith('.so'))
    for rep in range(2):
        for f in libs:
            env = dict(os.environ, SDFNMPC_LIB=os.path.join(VDIR, f))
            r = subprocess.run([sys.executable, __file__, 'one', str(Bn)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().split('\n')[-1] if r.returncode == 0 and r.stdout.strip() else None
            out.setdefault(f, []).append(json.loads(line) if line else dict(error=r.stderr[-400:]))
            print(f, out[f][-1], flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'qp_ab.json'), 'w'), indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build(sys.argv[2:])
    elif sys.argv[1] == 'one':
        one(int(sys.argv[2]))
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
