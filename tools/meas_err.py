import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from test_gpu_nmpc import _setup
from oracle import rti
for model in ('att','att_tau','acc'):
    torch, rti_, net, ocfg, prob, w, mpc = _setup(24, 'cuda:0', flag=1.0, model=model)
    xb, ub = rti.init_traj(prob['x0'], ocfg)
    for step in range(3):
        mpc.set_x0(prob['x0']); mpc.solve()
        xo, uo, out = rti.rti_step(net, xb, ub, prob, ocfg)
        x, u = mpc.get_matrices()
        print(model, step, 'u0', np.abs(mpc.get_u().cpu().numpy()-out['u0']).max(), 'u all', np.abs(u.cpu().numpy()-uo).max(), 'x all', np.abs(x.cpu().numpy()-xo).max(), flush=True)
        xb, ub = xo, uo
