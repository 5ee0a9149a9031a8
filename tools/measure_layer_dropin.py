#!/usr/bin/env python
"""Cost of the weed_layer_t drop-ins (libpe_weed_layer.so) on real layers built through the reference's libweed: the same
convert_layer_palette(4K YUV420P -> RGBA32) call with the caller's buffers pageable, page-locked around each transfer
(pe_weed_layer_set_pinning(1)), and page-locked once by the host (pe_host_register of the source planes, the bigblock scenario;
the new RGBA buffer still comes from malloc).  TEST INFRASTRUCTURE (needs oracle/_ref: the minihost + libweed)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import pe_testlib as T  # noqa: E402
from test_weed_layer import Host  # noqa: E402


def main():
    host = Host()
    from lives_b200 import _capi
    pe = _capi.lib()
    rng = np.random.default_rng(0)
    w, h = 3840, 2160
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    out = {}
    for mode in ("pageable", "transient_register", "host_registered"):
        host.lib.pe_weed_layer_set_pinning(1 if mode == "transient_register" else 0)
        ts = []
        for it in range(8):
            lay = host.layer(512, w, h, [y, u, v], subspace=1)
            if mode == "host_registered":
                for p in range(3):
                    rows = h if p == 0 else h // 2
                    pe.pe_host_register(host.mh.mh_layer_plane(lay, p), host.mh.mh_layer_rowstride(lay, p) * rows)
                ptrs = [host.mh.mh_layer_plane(lay, p) for p in range(3)]
            t0 = time.perf_counter()
            ok = host.lib.convert_layer_palette(lay, 3, 0)
            ts.append(time.perf_counter() - t0)
            assert ok == 1
            if mode == "host_registered":
                for p in ptrs:
                    pass  # (the drop-in released the planes through free(); registrations of freed ranges are dropped below)
            host.mh.mh_layer_free(lay)
        out[mode] = {"ms_per_call_median": 1e3 * float(np.median(ts[2:])), "fps": 1.0 / float(np.median(ts[2:]))}
    host.lib.pe_weed_layer_set_pinning(0)
    print(json.dumps({"what": "convert_layer_palette(weed_layer_t 4K YUV420P -> RGBA32) through libpe_weed_layer.so: 12.4 MB up, 33.2 MB down",
                      "modes": out}))


if __name__ == "__main__":
    main()
