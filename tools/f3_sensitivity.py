#!/usr/bin/env python
"""Where does k_fused3 wait?  The headline launch (32 x 4K) with (a) 32 distinct fg / bg frames (the bench), (b) ONE fg frame for all
32 (12.4 MB: L2 resident, luma / chroma loads become L2 hits), (c) ONE bg frame (33 MB, L2 resident), (d) both, (e) one out frame as
well.  If (b) is much faster than (a), the prefetch distance of the luma / chroma words is what the warps wait for."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lives_b200 as lb  # noqa: E402

FW, FH, IW, IH = 3840, 2160, 3840, 1608
dev = torch.device("cuda", 0)
eng = lb.Engine(device=0)
g = torch.Generator(device=dev); g.manual_seed(20)
B = 32
Y = [torch.randint(16, 236, (FH, FW), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
U = [torch.randint(16, 241, (FH // 2, FW // 2), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
V = [torch.randint(16, 241, (FH // 2, FW // 2), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
BG = [torch.randint(0, 256, (FH, FW * 4), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
OUT = [torch.empty((FH, FW * 4), dtype=torch.uint8, device=dev) for _ in range(B)]


def run(same_fg, same_bg, same_out, steps=50):
    fgs = [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_YUV420P, FW, FH, [Y[0 if same_fg else i].data_ptr(), U[0 if same_fg else i].data_ptr(), V[0 if same_fg else i].data_ptr()],
                                [FW, FW // 2, FW // 2], yuv_clamping=0, yuv_subspace=1) for i in range(B)]
    bgs = [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_RGBA32, FW, FH, [BG[0 if same_bg else i].data_ptr()], [FW * 4], gamma_type=1) for i in range(B)]
    outs = [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_RGBA32, FW, FH, [OUT[0 if same_out else i].data_ptr()], [FW * 4]) for i in range(B)]
    for _ in range(5):
        lb.fused_convert_letterbox_over_gamma_batch(fgs, bgs, outs, IW, IH, 0.5, 1, 2)
    eng.sync(); eng.timer_start()
    for _ in range(steps):
        lb.fused_convert_letterbox_over_gamma_batch(fgs, bgs, outs, IW, IH, 0.5, 1, 2)
    ms = eng.timer_stop_ms() / steps
    print("same fg %d bg %d out %d: %.4f ms per 32 frames = %.0f fps" % (same_fg, same_bg, same_out, ms, B / ms * 1e3), flush=True)


for cfg in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1)):
    run(*cfg)
