#!/usr/bin/env python
"""Event-timed GB/s of the two kernels added last in round 1 (k_slide_over, k_chroma_upsample_packed) on 4K frames, several
distinct frames per launch train so that the working set (> 126 MB) does not sit in L2.  Prints one JSON line per kernel."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import lives_b200 as lb  # noqa: E402
import pe_testlib as T  # noqa: E402

W, H, N, REPS = 3840, 2160, 8, 5


def main():
    eng = lb.Engine(device=0)
    rng = np.random.default_rng(0)
    out = []
    # slide over, RGBA32: 1 byte read + 1 byte written per output byte
    a = [lb.Layer.from_host(eng, 3, W, H, [T.make_packed(rng, W, H, 4)]) for _ in range(N)]
    b = [lb.Layer.from_host(eng, 3, W, H, [T.make_packed(rng, W, H, 4)]) for _ in range(N)]
    o = [lb.Layer.create(eng, 3, W, H) for _ in range(N)]
    for direction, mvl in ((1, 1), (3, 1)):
        for i in range(N):
            lb.slide_over(a[i], b[i], o[i], 100, direction, mvl, 0)
        eng.sync()
        eng.timer_start()
        for _ in range(REPS):
            for i in range(N):
                lb.slide_over(a[i], b[i], o[i], 100, direction, mvl, 0)
        ms = eng.timer_stop_ms() / (REPS * N)
        out.append({"kernel": "k_slide_over", "case": "RGBA32 4K direction %d" % direction, "us": ms * 1e3,
                    "GB/s": 2 * W * H * 4 / ms / 1e6})
    for l in a + b + o:
        l.free()
    # YUV420P -> YUVA8888: 1.5 bytes read + 4 written per pixel
    y, u, v = T.make_yuv_planar(rng, W, H, False, True)
    for opal, ps in ((589, 4), (588, 3), (544, 3), (3, 4), (1, 3)):
        tot = 0.0
        warm = [lb.Layer.from_host(eng, 512, W, H, [y, u, v]) for _ in range(N)]   # one untimed round: the pool owns blocks of the
        for l in warm:                                                              # destination size afterwards (no cudaMalloc below)
            assert lb.convert_layer_palette(l, opal, 0)
        eng.sync()
        for l in warm:
            l.free()
        for _ in range(REPS):
            ls = [lb.Layer.from_host(eng, 512, W, H, [y, u, v]) for _ in range(N)]
            eng.sync()
            eng.timer_start()
            for l in ls:
                assert lb.convert_layer_palette(l, opal, 0)
            tot += eng.timer_stop_ms()
            for l in ls:
                l.free()
        ms = tot / (REPS * N)
        out.append({"kernel": "k_yuv_planar_to_rgb_fast (single frame)" if opal < 512 else "k_quad_chroma (+ luma copy)" if opal == 544 else "k_chroma_upsample_packed", "case": "YUV420P -> %d 4K (pool warm; the call includes taking the destination block from the pool)" % opal,
                    "us": ms * 1e3, "GB/s": W * H * (1.5 + ps) / ms / 1e6})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
