"""CPU-only: the host build of csrc/epnp_core.cuh + the RANSAC emulation of tests/host_harness against cv2.solvePnPRansac on
random planted problems (6 .. 4000 points, 0-85 % outliers, 0.3-6 px noise, quantised / coplanar coordinates): counts the
trials whose consensus set or whose rvec / tvec BITS differ.    python scripts/sweep_host_ransac.py <seed> <trials>
Result of 8 seeds x 3000 trials: profiles/r02_host_ransac_sweep.log (0 differences in 17 521 trials that have a model)."""
import sys, ctypes, resource, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, cv2
from tests.planted import K_LM
from tests.test_epnp_host import _case, P, CAM
resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
L=ctypes.CDLL(os.path.join(ROOT,'tests','host_harness','libepnp_host.so'))
L.host_ransac.restype=ctypes.c_int
seed=int(sys.argv[1]); N=int(sys.argv[2])
rng=np.random.RandomState(seed)
bad_set=bad_pose=tot=0
t0=time.time()
for trial in range(N):
    n=int(rng.choice([6,7,9,12,20,33,50,200,1000,4000]))
    of=float(rng.choice([0,0.1,0.3,0.5,0.7,0.85]))
    pw,uv=_case(rng,n,float(rng.choice([0.3,1.0,3.0,6.0])))
    if rng.rand()<0.3: pw[:,2]=np.round(pw[:,2])
    if rng.rand()<0.2: pw=np.round(pw)      # quantised like uint8 XYZ maps
    no=int(n*of); idx=rng.choice(n,no,replace=False); uv[idx]+=rng.uniform(-60,60,(no,2))
    ret,rvc,tvc,inl=cv2.solvePnPRansac(pw,uv.reshape(-1,1,2),K_LM,None,flags=cv2.SOLVEPNP_EPNP,reprojectionError=5,iterationsCount=100)
    rvec,tvec,mask,it=np.zeros(3),np.zeros(3),np.zeros(n,np.uint8),ctypes.c_int()
    r=L.host_ransac(P(pw),P(uv),n,P(CAM),ctypes.c_float(5.0),100,ctypes.c_double(0.99),P(rvec),P(tvec),mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),ctypes.byref(it))
    if (inl is None)!=(r<0):
        bad_set+=1; print('outcome mismatch',seed,trial,n,of); continue
    if inl is None: continue
    tot+=1
    if not np.array_equal(np.nonzero(mask)[0],inl[:,0]):
        bad_set+=1; print('SET mismatch',seed,trial,n,of,len(inl),int(mask.sum()))
    elif not (np.array_equal(rvec,rvc[:,0]) and np.array_equal(tvec,tvc[:,0])):
        bad_pose+=1; print('pose bits differ',seed,trial,n,of,len(inl),np.abs(rvec-rvc[:,0]).max(),np.abs(tvec-tvc[:,0]).max())
print('seed',seed,'trials with model',tot,'set mismatches',bad_set,'pose-bit mismatches',bad_pose,'%.0fs'%(time.time()-t0))
