import os, time, threading, sys
base=sys.argv[1]; n=64; sz=77<<20
d=os.path.join(base,'io_probe'); os.makedirs(d,exist_ok=True)
buf=os.urandom(1<<20)*77
for i in range(n):
    with open(f'{d}/f{i}','wb') as f: f.write(buf)
def rd(lo,hi):
    b=bytearray(sz)
    for i in range(lo,hi):
        with open(f'{d}/f{i}','rb',buffering=0) as f: f.readinto(b)
for nt in (1,4,16,64):
    t0=time.perf_counter(); ts=[threading.Thread(target=rd,args=(k*n//nt,(k+1)*n//nt)) for k in range(nt)]
    [t.start() for t in ts]; [t.join() for t in ts]
    dt=time.perf_counter()-t0; print(base,nt,'threads', round(n*sz/dt/1e9,1),'GB/s')
import shutil; shutil.rmtree(d)
