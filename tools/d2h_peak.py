"""Pure pinned D2H copy of one bench step worth of output (1.8 GB): the PCIe floor of the e2e number."""
import torch, time
n = 2999*600000
d = torch.empty(n, dtype=torch.int8, device="cuda"); d.zero_()
h = torch.empty(n, dtype=torch.int8, pin_memory=True)
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
ts=[]
for _ in range(5):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("pure D2H 1.8 GB pinned: ms", [round(t,2) for t in ts], "GB/s", round(n/min(ts)/1e6,1))
