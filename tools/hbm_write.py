"""Developer check: pure-write, pure-read and copy bandwidth of this B200 (torch kernels), for the store-bound epilogues."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.quick_bench import timeit
for mb in (620, 2048):
    n = mb * (1 << 20) // 4
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    t, _ = timeit(lambda: x.fill_(1.0), iters=10, graph=True)
    print(f"{mb} MB fill  : {t*1e3:8.1f} us  {mb*1.048576/t/1e3:6.2f} TB/s written")
    t, _ = timeit(lambda: y.copy_(x), iters=10, graph=True)
    print(f"{mb} MB copy  : {t*1e3:8.1f} us  {2*mb*1.048576/t/1e3:6.2f} TB/s read+written")
    t, _ = timeit(lambda: x.sum(), iters=10, graph=True)
    print(f"{mb} MB sum   : {t*1e3:8.1f} us  {mb*1.048576/t/1e3:6.2f} TB/s read")
