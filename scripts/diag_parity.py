"""Diagnostic (GPU): product vs oracle error statistics per config / dtype / image size; prints, never asserts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import model as OM
from tests import util
from tests.test_ops_gpu import NW, PW
from deeplabv3plus_keras_b200.engine import Plan

CASES = [
    ("x16", dict(base="xception", output_stride=16), (129, 257)),
    ("x8br", dict(base="xception", output_stride=8, refine=True, rate_mult=2), (97, 193)),
    ("m16", dict(base="mobilenetv2", output_stride=16, aspp=util.DEFAULT_ASPP), (129, 257)),
    ("m8br", dict(base="mobilenetv2", output_stride=8, refine=True, aspp=util.DEFAULT_ASPP), (96, 192)),
]
EMU = os.environ.get("EMU", "1") == "1"
only = sys.argv[1:] or None
for name, case, sizes in CASES:
    if only and name not in only:
        continue
    for size in sizes:
        for dtype in ("float32", "bfloat16"):
            conf = util.make_conf(dtype=dtype, image_size=size, **case)
            ss = util.build(conf); util.randomize_weights(ss.model)
            B = 2
            plan = Plan(ss.model, B, training=True)
            x, y = util.synthetic_batch(conf, B, plan.out_shape[1:3])
            plan.set_loss(PW, NW); plan.load_batch(x, y); plan.step_fwd_bwd(); plan.regularization(); torch.cuda.synchronize()
            w = util.torch_weights(ss.model)
            xin = torch.from_numpy(x)
            if dtype == "bfloat16": xin = xin.to(torch.bfloat16).float()
            t0 = time.time()
            data, l2, grads, out = OM.loss_and_grads(conf, w, xin.double(), torch.from_numpy(y), PW, NW,
                                                     emulate_bf16=(dtype == "bfloat16" and EMU))
            ref = out["logits"].detach().numpy(); got = plan.logits.buf.float().cpu().numpy()
            e = got - ref
            print(f"[{name} {size} {dtype}] logits max|e|/max|ref| {np.abs(e).max()/np.abs(ref).max():.3e}  rms(e)/rms(ref) {np.sqrt((e**2).mean())/np.sqrt((ref**2).mean()):.3e}"
                  f"  loss {plan.loss_value():.6f} vs {float(data+l2):.6f}  oracle {time.time()-t0:.1f}s")
            gg = plan.gradients(); lam = conf["hps"]["weight_decay"]; rows = []
            for k, g in grads.items():
                g = g.numpy().copy()
                if k.endswith("/kernel") and k.split("/")[0].startswith("conv2d"): g -= 2*lam*w[k].numpy()
                d = gg[k] - g
                rows.append((np.sqrt((d**2).mean())/max(np.sqrt((g**2).mean()),1e-12), np.abs(d).max()/max(np.abs(g).max(),1e-12), np.abs(g).max(), k))
            rows.sort(reverse=True)
            for r in rows[:4]: print(f"      grad rms-rel {r[0]:.3e} max-rel {r[1]:.3e} |g|max {r[2]:.2e}  {r[3]}")
            med = np.median([r[0] for r in rows]); print(f"      grad rms-rel median {med:.3e} over {len(rows)} tensors")
            # inference
            # well-conditioned inference: moving statistics := batch statistics of this batch
            st = OM.forward(conf, w, xin.double(), training=True, momentum_override=0.0)["new_stats"]
            nw_ = ss.model.named_weights()
            for k, v in st.items(): nw_[k][...] = v.numpy()
            w = util.torch_weights(ss.model)
            pl = ss.model.plan(B, training=False); pl.upload_weights(); probs = pl.predict(x)
            o2 = OM.forward(conf, w, xin.double(), training=False, emulate_bf16=(dtype == "bfloat16" and EMU)); rp = o2["probs"].numpy()
            agree = (pl.segment(x) == rp.argmax(-1)).mean()
            print(f"      inference probs max|e| {np.abs(probs-rp).max():.3e}  label agreement {agree:.5f}")
            del plan, pl; torch.cuda.empty_cache()
