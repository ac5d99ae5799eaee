"""Input-pipeline row.  PINNED: the golden vectors were produced by executing the reference's own
`resize` / `resize_image_to_target_symmeric_size` (scripts/make_golden_preprocess.py); the oracle restatement
(oracle/preprocess.py) must reproduce them exactly, and the CUDA kernels must match both — labels bit-exact, images
within fp32 rounding of the fp64 result (tolerance 2e-6 absolute on values in (-1, 1))."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as OP

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "preprocess_*.npz")))
IDS = [os.path.basename(p)[11:-4] for p in GOLDEN]


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_reproduces_reference_vectors(path):
    g = np.load(path)
    image, label = OP.sample(g["image_u8"], g["label_u8"], int(g["size"]), 21)
    np.testing.assert_array_equal(image.astype(np.float32), g["image"])
    np.testing.assert_array_equal(label.astype(np.uint8), g["label"])


def test_geometry_matches_host_module():
    from deeplabv3plus_keras_b200.data import target_geometry
    for h, w, s in [(375, 500, 513), (500, 333, 224), (64, 64, 97), (1024, 2048, 512), (37, 91, 65), (500, 500, 224),
                    (1, 7, 16), (281, 500, 513)]:
        hp, wp, oy, ox, *_ = OP.geometry(h, w, s)
        assert target_geometry(h, w, s) == (hp, wp, oy, ox)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gpu_pipeline_matches_reference_vectors(dtype):
    from deeplabv3plus_keras_b200.data import preprocess_batch
    by_size = {}
    for p in GOLDEN:
        g = np.load(p)
        by_size.setdefault(int(g["size"]), []).append(g)
    for size, gs in by_size.items():
        imgs = [torch.from_numpy(g["image_u8"]).cuda() for g in gs]
        labs = [torch.from_numpy(g["label_u8"]).cuda() for g in gs]
        x, y = preprocess_batch(imgs, labs, size, 21, dtype=dtype)
        torch.cuda.synchronize()
        for i, g in enumerate(gs):
            np.testing.assert_array_equal(y[i].cpu().numpy().astype(np.uint8), g["label"])     # bit-exact
            got = x[i].float().cpu().numpy()
            tol = 2e-6 if dtype == torch.float32 else 8e-3                                     # bf16: 2^-8 relative
            assert np.abs(got - g["image"]).max() <= tol


@pytest.mark.gpu
def test_gpu_pipeline_ragged_batch_against_oracle():
    """A batch of differently sized samples (as a VOC batch is), 19 classes, against the scipy oracle."""
    from deeplabv3plus_keras_b200.data import preprocess_batch
    rng = np.random.default_rng(7)
    shapes = [(375, 500), (500, 375), (333, 500), (500, 500), (112, 97), (1, 40)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    labs = [rng.integers(0, 256, (h, w), dtype=np.uint8) for h, w in shapes]
    x, y = preprocess_batch([torch.from_numpy(a).cuda() for a in imgs], [torch.from_numpy(a).cuda() for a in labs],
                            224, 19)
    torch.cuda.synchronize()
    for i in range(len(shapes)):
        image, label = OP.sample(imgs[i], labs[i], 224, 19)
        np.testing.assert_array_equal(y[i].cpu().numpy(), label)
        assert np.abs(x[i].cpu().numpy() - image).max() <= 2e-6
    assert int(y.max()) <= 18
