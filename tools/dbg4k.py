import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import lives_b200 as lb, pe_testlib as T
eng = lb.Engine()
fw, fh, ow, oh, iw, ih = 3840, 2160, 3840, 2160, 3840, 1608
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(0)
fgs,bgs,outs=[],[],[]
for i in range(n):
    y,u,v = T.make_yuv_planar(rng, fw, fh, False, True)
    bg = T.make_packed(rng, ow, oh, 4)
    fgs.append(lb.Layer.from_host(eng, 512, fw, fh, [y,u,v], yuv_subspace=1))
    bgs.append(lb.Layer.from_host(eng, 3, ow, oh, [bg], gamma_type=-1))
    outs.append(lb.Layer.create(eng, 3, ow, oh))
for k in range(3):
    lb.fused_convert_letterbox_over_gamma_batch(fgs,bgs,outs,iw,ih,0.5,-1,1)
    eng.sync()
    print("step", k, "ok")
