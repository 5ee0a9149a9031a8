#!/usr/bin/env python
"""PCIe ceiling for the host-buffer (e2e) path: pinned H2D alone, D2H alone, both at once, in frame-sized pieces.
Prints GB/s; the e2e pipeline moves 45.6 MB up + 33.2 MB down per 4K frame."""
import json
import torch

dev = torch.device("cuda:0")
UP, DOWN, N = 45_619_200, 33_177_600, 24
hu = [torch.empty(UP, dtype=torch.uint8).pin_memory() for _ in range(4)]
hd = [torch.empty(DOWN, dtype=torch.uint8).pin_memory() for _ in range(4)]
du = [torch.empty(UP, dtype=torch.uint8, device=dev) for _ in range(4)]
dd = [torch.empty(DOWN, dtype=torch.uint8, device=dev) for _ in range(4)]
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_up.wait_stream(torch.cuda.current_stream()); s_dn.wait_stream(torch.cuda.current_stream())
    for i in range(N):
        if up:
            with torch.cuda.stream(s_up):
                du[i % 4].copy_(hu[i % 4], non_blocking=True)
        if down:
            with torch.cuda.stream(s_dn):
                hd[i % 4].copy_(dd[i % 4], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3


out = {}
for name, (u, d) in {"h2d_only": (1, 0), "d2h_only": (0, 1), "both": (1, 1)}.items():
    run(u, d)
    t = min(run(u, d) for _ in range(3))
    out[name] = {"s_per_frame": t / N, "h2d_gbs": UP * N * u / t / 1e9, "d2h_gbs": DOWN * N * d / t / 1e9, "frames_per_s": N / t}
print(json.dumps(out))
