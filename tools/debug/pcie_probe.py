"""Debug: pinned-host copy rates on the box (alone and both directions at once), to tell whether bench.py's e2e figure is PCIe-bound.
usage: python tools/debug/pcie_probe.py [MB]"""
import sys, torch
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 182
n = mb * 1024 * 1024 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
for name, f in (("H2D alone", (1, 0)), ("D2H alone", (0, 1)), ("both", (1, 1))):
    run(*f, reps=2)
    ms = run(*f)
    print(f"{name:10s}: {ms:.3f} ms per {mb} MB step  ({mb / 1024 / ms * 1e3:.1f} GiB/s per direction)")
