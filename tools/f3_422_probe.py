#!/usr/bin/env python
"""The headline chain with a planar 4:2:2 fg (3840x2160 YUV422P -> RGBA32, letterbox 3840x1608, alpha-over, gamma): 32 frames per
launch through k_fused3's IS422 instantiation (round 1 / early round 2: k_fused2).  PE_F3_NO_422=1 sends it back to k_fused2."""
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
for is422 in (1, 0):
    ch = FH if is422 else FH // 2
    Y = [torch.randint(16, 236, (FH, FW), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
    U = [torch.randint(16, 241, (ch, FW // 2), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
    V = [torch.randint(16, 241, (ch, FW // 2), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
    BG = [torch.randint(0, 256, (FH, FW * 4), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
    OUT = [torch.empty((FH, FW * 4), dtype=torch.uint8, device=dev) for _ in range(B)]
    pal = lb.WEED_PALETTE_YUV422P if is422 else lb.WEED_PALETTE_YUV420P
    fgs = [lb.Layer.wrap_device(eng, pal, FW, FH, [Y[i].data_ptr(), U[i].data_ptr(), V[i].data_ptr()], [FW, FW // 2, FW // 2],
                                yuv_clamping=0, yuv_subspace=1) for i in range(B)]
    bgs = [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_RGBA32, FW, FH, [BG[i].data_ptr()], [FW * 4], gamma_type=1) for i in range(B)]
    outs = [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_RGBA32, FW, FH, [OUT[i].data_ptr()], [FW * 4]) for i in range(B)]
    for _ in range(5):
        lb.fused_convert_letterbox_over_gamma_batch(fgs, bgs, outs, IW, IH, 0.5, 1, 2)
    eng.sync(); eng.timer_start()
    steps = 30
    for _ in range(steps):
        lb.fused_convert_letterbox_over_gamma_batch(fgs, bgs, outs, IW, IH, 0.5, 1, 2)
    ms = eng.timer_stop_ms() / steps
    alg = FW * FH * (1 + (1.0 if is422 else 0.5)) + 2 * FW * FH * 4
    print('{"fg": "%s", "ms_per_32_frames": %.4f, "fps": %.0f, "algorithmic_GBs": %.0f}' % ("YUV422P" if is422 else "YUV420P", ms, B / ms * 1e3, alg * B / ms / 1e6), flush=True)
    del Y, U, V, BG, OUT
