import torch, time
n = 92405760
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(15256576, dtype=torch.uint8).pin_memory(); d2 = torch.empty(15256576, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print(f"H2D 92MB pinned: {ms:.3f} ms  {n/ms/1e6:.1f} GB/s")
e0.record()
for _ in range(10): h2.copy_(d2, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print(f"D2H 15MB pinned: {ms:.3f} ms  {15256576/ms/1e6:.1f} GB/s")
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv", shell=True, capture_output=True, text=True).stdout)
