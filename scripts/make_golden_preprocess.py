"""Golden vectors for the input-pipeline row from the REFERENCE'S OWN functions: the source of `resize` and
`resize_image_to_target_symmeric_size` is read from /root/reference at generation time (the module itself cannot be
imported: it imports tensorflow / cupy at the top, ss.py:38-47), compiled and executed here against the installed
scipy — nothing is copied into the repo but the resulting arrays (tests/golden/preprocess_*.npz)."""
import ast
import os
import sys

import numpy as np
from scipy import ndimage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/bodhi/deeplabv3plus_keras/semantic_segmentation.py"
tree = ast.parse(open(SRC).read())
wanted = {"resize", "resize_image_to_target_symmeric_size"}
mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted], type_ignores=[])
ns = {"np": np, "ndimage": ndimage, "DEVICE_CPU": -1}
exec(compile(mod, SRC, "exec"), ns)
ref_fit = ns["resize_image_to_target_symmeric_size"]

rng = np.random.default_rng(1024)
# small fixtures (VOC-like aspect ratios at a quarter of the size, an exact 2:1 down-scale whose interpolation hits
# exact .5 ties, odd extents)
CASES = {"landscape_125x94_to_129": (94, 125, 129), "portrait_83x125_to_56": (125, 83, 56),
         "square_64_to_97": (64, 64, 97), "down_128x256_to_64": (128, 256, 64), "odd_37x91_to_65": (37, 91, 65)}
out_dir = os.path.join(ROOT, "tests", "golden")
for name, (h, w, size) in CASES.items():
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    lab = rng.integers(0, 21, (h, w), dtype=np.uint8)
    blocks = rng.integers(0, 255, (max(h // 16, 1), max(w // 16, 1)), dtype=np.uint8)     # coherent regions + "void" 255
    lab = np.kron(blocks, np.ones((16, 16), dtype=np.uint8))[:h, :w] if h >= 16 and w >= 16 else lab
    lab = np.where(lab > 230, 255, lab % 25).astype(np.uint8)                           # some ids above 20
    image = 2.0 * (img / 255 - 0.5)
    image_p, *_ = ref_fit(image, size, device=-1)
    label = np.expand_dims(lab.copy(), axis=-1)
    label[label > 20] = 0
    label_p, *_ = ref_fit(label, size, device=-1)
    label_p[label_p > 20] = 0
    np.savez_compressed(os.path.join(out_dir, f"preprocess_{name}.npz"), image_u8=img, label_u8=lab, size=size,
                        image=image_p.astype(np.float32), label=label_p[..., 0].astype(np.uint8))
    print(name, image_p.shape, label_p.shape, int(label_p.max()))
