import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from rbdreference_b200 import RBDReference, robots
from oracle.rbd_oracle import BatchOracle
def rel(x,r): return float(np.max(np.abs(x-r))/np.max(np.abs(r)))
for name in ["iiwa14","hyq","atlas"]:
    rb=robots.by_name(name); bo=BatchOracle(rb); n=rb.get_num_vel()
    rng=np.random.default_rng(5); B=256
    q,qd,qdd=rng.uniform(-np.pi,np.pi,(B,n)),rng.uniform(-1,1,(B,n)),rng.uniform(-1,1,(B,n))
    ref=bo.rnea_grad(q,qd,qdd); mref=bo.minv(q)
    for v in (0,1,2,3):
        RBDReference.set_kernel_variant(v)
        out=[]
        for dt in (torch.float64, torch.float32):
            eng=RBDReference(rb,dtype=dt)
            t=lambda x: torch.as_tensor(x,device="cuda").to(dt)
            out.append("%.1e/%.1e"%(rel(eng.rnea_grad(t(q),t(qd),t(qdd)).double().cpu().numpy(),ref), rel(eng.minv(t(q)).double().cpu().numpy(),mref)))
        print(name,"variant",v,"grad/minv f64",out[0],"f32",out[1])
RBDReference.set_kernel_variant(0)
