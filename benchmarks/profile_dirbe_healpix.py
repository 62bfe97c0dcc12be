import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import zodipy_b200 as zp
from zodipy_b200 import engine
EARTH=np.array([[-0.3919640703],[0.9020953332],[0.0]])
m=zp.Model(zp.Quantity(25.0,'um'),precision='fp32')
for i in range(3):
    out=m.evaluate_healpix(1024, EARTH, device_out=True, out_dtype=np.float32)
torch.cuda.synchronize()
