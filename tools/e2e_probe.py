import time, torch, os, sys
sys.path.insert(0, os.getcwd())
from pioneer_b200 import BatchedPioneerEnv
n=65536
env=BatchedPioneerEnv(n)
act=torch.zeros((n,6),dtype=torch.float32,pin_memory=True)
for _ in range(5): env.step_host(act)
t=time.perf_counter()
for _ in range(200): env.step_host(act)
dt=(time.perf_counter()-t)/200
print(os.environ.get("PNR_HOST_ONE_STREAM","two-streams"), "ms/step", dt*1e3, "env-steps/s %.3e"%(n/dt), "D2H GB/s", n*553/dt/1e9)
