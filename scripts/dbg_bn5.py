import sys; sys.path.insert(0,'.')
import torch, numpy as np
from tests import util, teacher
conf = util.make_conf(dtype="float32", base="xception", output_stride=8, image_size=97, refine=True, rate_mult=2)
for rep in range(2):
    res = teacher.run(conf, decision_forced=False)
    print(teacher.summarize(res))
    plan = res["plan"]
    name = "batch_normalization_5"
    g = res["teacher_grad"][name + "/out"]; y = res["teacher"][name + "/y"]
    act, yv, scale, shift, clog, shape = plan.act_sites[name][1:]
    z = (yv.float().view(-1, yv.shape[-1]) * scale + shift).view(shape).cpu().double()
    mask = (z > 0).double()
    s1 = (g * mask).sum((0,1,2))
    got = torch.from_numpy(plan.gradients()[name + "/beta"]).double()
    print("rep", rep, "dbeta: torch-from-teacher vs product max rel", float((s1-got).abs().max()/s1.abs().max()),
          "min|z|", float(z.abs().min()), "n(|z|<1e-6)", int((z.abs()<1e-6).sum()))
    bad = ((s1-got).abs() > 1e-4*s1.abs().max()).nonzero().flatten().tolist()
    print("bad channels", bad[:20])
    for c in bad[:3]:
        print(" ch", c, "s1", float(s1[c]), "got", float(got[c]), "sum|g|", float((g[...,c]*mask[...,c]).abs().sum()))
