import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/lagrange-dual-learning_b200')
import numpy as np, torch
from ikl import ConstraintSpec, standard_config, fused, synthetic, _cabi
from oracle import ik_oracle as O, build_oracle as C
cfg=standard_config('cfg_joint_limits_and_4obs_dynamic')
spec=ConstraintSpec.from_config(cfg,3,0.005,1.0); ospec=O.spec_from_json(cfg,3,0.005,1.0)
B=(1<<26)+4
theta,q_des,obst=synthetic.make_batch(B,seed=5,device='cuda:0')
lam=list(synthetic.SHIPPED_LAMBDA_OBS4_DYN)
v_rows,l_rows,d_rows=fused.lagrangian_forward_backward(spec,theta,q_des,obst,lam); k1=_cabi.last_kernel_name()
thp,tgp,obp=synthetic.to_planes(theta,q_des,obst)
v_pl,l_pl,d_pl=fused.lagrangian_forward_backward(spec,thp,tgp,obp,lam); k2=_cabi.last_kernel_name()
torch.cuda.synchronize()
print(k1,k2)
t=time.time(); sums,abs_sums,_=C.lagrangian_f64(ospec,theta.cpu().numpy(),q_des.cpu().numpy(),obst.cpu().numpy(),np.array(lam),want_grad=False); print('C oracle %.1fs'%(time.time()-t))
for name,v in (('rows',v_rows),('planes',v_pl)):
    err=np.abs(v.cpu().numpy()-sums)/abs_sums
    print(name,'max err rel to sum|term| %.2e'%err.max())
full=(B//512)*512
dpl=d_pl.contiguous()
print('dtheta staged rows vs staged planes: max |diff| / max |dtheta| = %.2e  (staged rows adds the obstacle forces in the order k ^ r)' % (float((d_rows-dpl).abs().max())/float(dpl.abs().max())))
import os
os.environ['IKL_ROWS_TMA']='0'
v_d,l_d,d_d=fused.lagrangian_forward_backward(spec,theta,q_des,obst,lam); k3=_cabi.last_kernel_name()
print(k3, 'dtheta direct rows vs staged planes bitwise on full tiles:', torch.equal(d_d[:full], dpl[:full]), ' tail max diff %.2e'%float((d_d[full:]-dpl[full:]).abs().max()))
