"""dev helper: pinned H2D / D2H bandwidth as seen by the library's own pinned allocator."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fulgor_b200 as fg
n = 1 << 30
pin = fg.PinnedBuffer(n)
h = torch.from_numpy(pin.view(np.uint8, n))
d = torch.empty(n, dtype=torch.uint8, device="cuda")
tp = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, src in (("lib-pinned", h), ("torch-pinned", tp)):
    for it in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"H2D {name}: {n/dt/1e9:.1f} GB/s")
    for it in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); src.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"D2H {name}: {n/dt/1e9:.1f} GB/s")
os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv")
os.system("nvidia-smi topo -m | head -20")
