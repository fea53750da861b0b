"""Developer probe (under gpurun): does cudaHostRegister work for a malloc'd 45 MB buffer in this container?"""
import ctypes, resource, time
import numpy as np, torch
print("RLIMIT_MEMLOCK", resource.getrlimit(resource.RLIMIT_MEMLOCK))
a = np.zeros(45_000_000 // 8 * 8 // 8, dtype=np.float64)
rt = torch.cuda.cudart()
torch.cuda.init()
t0 = time.perf_counter()
rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
print("cudaHostRegister rc", rc, "%.2f ms" % ((time.perf_counter() - t0) * 1e3))
d = torch.empty(a.size, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    torch.from_numpy(a).copy_(d); torch.cuda.synchronize()
    print("D2H 45 MB %.2f ms" % ((time.perf_counter() - t0) * 1e3))
