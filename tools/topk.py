import sys; sys.path.insert(0,"/root/repo")
import numpy as np, torch, sara_b200 as sb
ctx=sb.SiftContext(3840,2160); ctx.set_profiling(True); ctx.set_octave_overlap(False)
f=torch.from_numpy(np.random.default_rng(0).random((2160,3840),dtype=np.float32)).cuda()
ts=[]
for i in range(8):
    ctx.pyramid_enqueue(0,f,sb.ImagePyramidParams(first_octave_index=0)); ctx.wait(0); ts.append(ctx.timings(0))
import os
print("SB_DBG", os.environ.get("SB_DBG"), "top kernel us %.1f" % (1e3*np.median([t["pyramid_top_kernel"] for t in ts[3:]])), "serial pyramid us %.1f" % (1e3*np.median([t["pyramid"] for t in ts[3:]])))
