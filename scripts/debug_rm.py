import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi
from mogp_b200.obsmodel import make_spec
eng=_capi.Engine(0)
rs=np.random.RandomState(150)
m=150
x=np.sort(rs.uniform(0,8,m)); x[:18]=0.0
y=(48-6*x+rs.randn(m)*3-30)/20
th=np.array([0.28,1.0,1.58,0.0133])
spec=make_spec('rbf',True,True,False)
def ref(xx):
    K=np.exp(-0.5*np.clip(-2*np.outer(xx,xx)+(xx*xx)[:,None]+(xx*xx)[None,:],0,None)/th[2]**2)
    Ky=K+np.eye(len(xx))*(th[3]+1e-8)
    return np.linalg.inv(np.linalg.cholesky(Ky))
slot=eng.cluster_create(spec,th); eng.cluster_set_data(slot,x,y)
M,_=eng.cluster_get_factor(slot)
print("infer M err",np.max(np.abs(M-ref(x))))
cur=x.copy(); cy=y.copy()
for (p,n) in [(5,7),(60,11),(m-30,4),(0,3),(64,16),(100,1)]:
    keep=np.r_[0:p,p+n:len(cur)]
    print("   leave-out:", eng.cluster_leave_out_score(slot,p,n,min(1,n-1)))
    eng.cluster_remove_rows(slot,p,n)
    M,al=eng.cluster_get_factor(slot)
    gx,gy=eng.cluster_get_data(slot)
    rr=cy[keep]+th[0]*cur[keep]
    zz=ref(cur[keep])@rr
    lml_ref=0.5*(-len(keep)*np.log(2*np.pi)+2*np.sum(np.log(np.diag(ref(cur[keep]))))-zz@zz)
    zM=M@rr
    lml_M=0.5*(-len(keep)*np.log(2*np.pi)+2*np.sum(np.log(np.diag(M)))-zM@zM)
    print("   lml engine %.12f  ref %.12f  from fetched M %.12f ; alpha err %.3e"%(eng.cluster_loglik(slot),lml_ref,lml_M,np.max(np.abs(al-M.T@zM))))
    R=ref(cur[keep])
    E=np.abs(M-R)
    i,j=np.unravel_index(np.argmax(E),E.shape)
    print("rm p=%d n=%d: max err %.3e at (%d,%d) ; rows with err>1e-9: %s"%(p,n,E.max(),i,j,np.where(E.max(1)>1e-9)[0][:20]), "upper nonzero", np.max(np.abs(np.triu(M,1))))
    xb,yb=cur[p:p+n],cy[p:p+n]
    eng.cluster_append(slot,xb,yb)
    cur=np.r_[cur[keep],xb]; cy=np.r_[cy[keep],yb]
    M,al=eng.cluster_get_factor(slot)
    E=np.abs(M-ref(cur))
    print("   append: max err %.3e rows>1e-9 %s"%(E.max(),np.where(E.max(1)>1e-9)[0][:20]))
