"""PCIe ceilings of the box: pinned H2D alone, D2H alone, both directions at once (the e2e path moves 128 MiB in
and 64 MiB out per scan)."""
import torch
n = 128 << 20
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(n // 2, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_out = [torch.empty(n // 2, dtype=torch.uint8, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for r in range(reps):
        for i in range(4):
            if h2d:
                with torch.cuda.stream(s1): d_in[i].copy_(h_in[i], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[i].copy_(d_out[i], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return (reps * 4 * n / ms / 1e6 if h2d else 0.0), (reps * 4 * (n // 2) / ms / 1e6 if d2h else 0.0)
for _ in range(2): run(True, True, 2)
print("H2D alone GB/s: %.1f" % run(True, False)[0])
print("D2H alone GB/s: %.1f" % run(False, True)[1])
a, b = run(True, True)
print("both: H2D %.1f + D2H %.1f GB/s (128 MiB in per 64 MiB out)" % (a, b))
