"""Developer tool (GPU box): what the PCIe link gives for the e2e state round trip (109 MB each way at he30/ze63 Float32):
H2D alone, D2H alone, both directions at once — the floor of bench.py's e2e number when the step itself is hidden."""
import torch

n = 109_209_600 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


mb = n * 4 / 1e6
for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("H2D + D2H concurrently", both)):
    ms = timed(fn)
    print(f"{name:26s} {ms:7.3f} ms per 109.2 MB  = {mb / ms:6.1f} GB/s per direction")
