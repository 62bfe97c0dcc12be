import sys, numpy as np
sys.path[:0]=['/root/repo','/root/repo/oracle','/root/repo/tests']
import zodi_oracle as oracle, zodipy_b200 as zp
from test_gpu_random_parity import _case
seed=int(sys.argv[1]) if len(sys.argv)>1 else 13
name, model_args, deg, u, obs, earth = _case(seed)
model=zp.Model(name=name, gauss_quad_degree=deg, **model_args)
ref=oracle.evaluate(model.spec,u,obs,earth); noise=oracle.reference_rounding_noise(model.spec,u,obs,earth)
tot=np.abs(ref.sum(0))
got=model.evaluate_xyz(u,obs,earth,return_comps=True)
scale=np.maximum(np.abs(ref),1e-6*tot[None,:])
err=(np.abs(got-ref)-noise)/scale
print('obs shape',obs.shape,'kernel',model.device_model.kernel_name, 'n',u.shape[1])
for ci,c in enumerate(model.spec['comps']):
    j=np.nanargmax(err[ci]); print(c['label'],'%.2e'%err[ci,j],'pix',j,'ref %.6e got %.6e tot %.3e noise %.1e'%(ref[ci,j],got[ci,j],tot[j],noise[ci,j]),'robs %.4f'%np.linalg.norm(obs[:, j if obs.shape[1]>1 else 0]), 'u',u[:,j])
