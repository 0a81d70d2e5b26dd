import numpy as np, sys
a=np.loadtxt(sys.argv[1],dtype=np.uint64).astype(np.int64)
t0=a[a>0].min()
a=np.where(a>0,a-t0,-1)
names=["iss_wait_empty_start","iss_acc_empty_done","iss_first_full_done","iss_issue_done","ep_wait_full_start","ep_full_done","ep_ld_done_arrive","ep_compute_done","ep_bar_done","prod_wait_start","prod_wait_done"]
print("tile "+" ".join(n[:14].rjust(14) for n in names))
for i in list(range(0,12))+list(range(100,108)):
    print(f"{i:4d} "+" ".join(str(a[i,j]).rjust(14) for j in range(len(names))))
d=a[40:200]
print("mean per-tile period (ns), epilogue full_done:", np.diff(d[:,5]).mean())
print("ep: wait_full", (d[:,5]-d[:,4]).mean(), "ld+arrive", (d[:,6]-d[:,5]).mean(), "compute", (d[:,7]-d[:,6]).mean(), "bar", (d[:,8]-d[:,7]).mean())
print("issuer: wait acc_empty", (d[:,1]-d[:,0]).mean(), "wait first full", (d[:,2]-d[:,1]).mean(), "issue rest", (d[:,3]-d[:,2]).mean())
print("producer: wait empty at tile start", (d[:,10]-d[:,9]).mean())
raw=np.loadtxt(sys.argv[1],dtype=np.uint64).astype(np.int64)
if raw.shape[1] > 11 and raw[40:200,11].max() > 0:
    print("issuer: time inside 'issue rest' spent waiting for the later stages' data (ns):", raw[40:200,11].mean())
