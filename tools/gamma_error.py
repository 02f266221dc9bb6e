import sys; sys.path.insert(0,'.')
import numpy as np, torch
from pysp_b200 import engine
x=np.concatenate([np.linspace(0,1,2_000_001,dtype=np.float64), np.geomspace(1e-6,1,1_000_000)]).astype(np.float32)
y=engine.srgb_gamma(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
xd=x.astype(np.float64)
ref=np.where(xd<=np.float32(0.0031308), xd*np.float64(np.float32(12.92)), 1.055*np.power(xd,1/2.4)-0.055)
rel=np.abs(y-ref)/np.maximum(np.abs(ref),1e-3)
ref32=np.where(x<=np.float32(0.0031308), x*np.float32(12.92), (np.float32(1.055)*np.power(x,np.float32(1/2.4))-np.float32(0.055))).astype(np.float32)
rel32=np.abs(y-ref32.astype(np.float64))/np.maximum(np.abs(ref32),1e-3)
print("max rel err vs float64 formula: %.3e ; vs numpy float32 formula: %.3e"%(rel.max(), rel32.max()))
