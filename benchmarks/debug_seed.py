"""Debug aid: per-component errors of one configuration of tests/test_gpu_random_parity.py.
usage: python benchmarks/debug_seed.py SEED [fp64|fp32]"""
import sys

import numpy as np

sys.path[:0] = ['/root/repo', '/root/repo/oracle', '/root/repo/tests']
import zodi_oracle as oracle  # noqa: E402
import zodipy_b200 as zp  # noqa: E402
from test_gpu_random_parity import _case  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 13
precision = sys.argv[2] if len(sys.argv) > 2 else "fp64"
name, model_args, deg, u, obs, earth = _case(seed)
model = zp.Model(name=name, gauss_quad_degree=deg, precision=precision, **model_args)
ref = oracle.evaluate(model.spec, u, obs, earth)
noise = oracle.reference_rounding_noise(model.spec, u, obs, earth)
tot = np.abs(ref.sum(0))
got = model.evaluate_xyz(u, obs, earth, return_comps=True)
floor = 1e-6 if precision == "fp64" else 1.0
scale = np.maximum(np.abs(ref), floor * tot[None, :])
err = (np.abs(got - ref) - noise) / scale
print('x', model_args['x'], 'deg', deg, 'obs shape', obs.shape, 'kernel', model.device_model.kernel_name_for(u.shape[1], precision), 'n', u.shape[1])
et = np.abs(got.sum(0) - ref.sum(0)) / np.abs(ref).sum(0)
jt = np.nanargmax(et)
print('total: max err %.2e at pix %d (total %.4e) u=%s robs=%.4f' % (et[jt], jt, ref.sum(0)[jt], u[:, jt], np.linalg.norm(obs[:, jt if obs.shape[1] > 1 else 0])))
for ci, c in enumerate(model.spec['comps']):
    j = np.nanargmax(err[ci])
    print(c['label'], '%.2e' % err[ci, j], 'pix', j, 'ref %.6e got %.6e tot %.3e' % (ref[ci, j], got[ci, j], tot[j]), '| at worst-total pix: ref %.4e got %.4e' % (ref[ci, jt], got[ci, jt]))
