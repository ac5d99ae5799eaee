"""Generates tests/golden/*.npz from the oracle (seed 1024, the reference's seed, ss.py:1798).  Small fixtures:
inputs, labels, low-resolution logits, loss and a few gradients of a tiny-image model; weights are regenerated from
`weight_seed` (a checksum guards against RNG drift)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deeplabv3plus_keras_b200.deeplab import ss_nw, ss_pw
from oracle import model as OM
from tests import util

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
CASES = {
    "xception_os16_65": dict(base="xception", output_stride=16, image_size=65, width=64),
    "mobilenetv2_os16_65": dict(base="mobilenetv2", output_stride=16, image_size=65, width=64, aspp=util.DEFAULT_ASPP),
    "xception_os8_br_49": dict(base="xception", output_stride=8, image_size=49, width=64, refine=True, rate_mult=2),
}
for name, case in CASES.items():
    conf = util.make_conf(**case)
    ss = util.build(conf)
    util.randomize_weights(ss.model, seed=1024)
    w = util.torch_weights(ss.model)
    out_hw = ss.model.outputs[0].shape[1:3]
    x, y = util.synthetic_batch(conf, 2, out_hw, seed=1024)
    data, l2, grads, out = OM.loss_and_grads(conf, w, torch.from_numpy(x).double(), torch.from_numpy(y), ss_pw, ss_nw)
    small = [k for k in sorted(grads) if grads[k].numel() <= 20000]
    keys = small[:: max(1, len(small) // 8)][:8]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), conf=np.array(conf, dtype=object), weight_seed=1024,
        weight_checksum=float(sum(float(np.abs(v.numpy()).sum()) for v in w.values())), x=x, y=y,
        logits=out["logits"].detach().numpy(), loss=float(data), l2=float(l2), pw=np.array(ss_pw), nw=np.array(ss_nw),
        grad_keys=np.array(keys), **{"grad/" + k: grads[k].numpy() for k in keys})
    print(name, "loss", float(data), "logits", tuple(out["logits"].shape), os.path.getsize(os.path.join(OUT, name + ".npz")))
