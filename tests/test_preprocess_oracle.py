"""oracle/preprocess.py against the reference's own transform stack (torchvision transforms on PIL images,
datasets/dataset_h5.py:29-35), run live: Pillow and torchvision are installed in every container this suite runs in."""
import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms

from oracle import preprocess as O


def reference_transform(patch):
    t = transforms.Compose([transforms.Resize(224), transforms.ToTensor(),
                            transforms.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))])
    return t(Image.fromarray(patch)).numpy()


@pytest.mark.parametrize("size,seed", [(256, 0), (256, 1), (512, 2), (224, 3), (300, 4)])
def test_eval_transform_is_bit_exact(size, seed):
    rng = np.random.default_rng(seed)
    patch = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
    if seed == 1:      # smooth content with saturated regions (rounding / clipping paths)
        yy, xx = np.mgrid[0:size, 0:size]
        patch = np.stack([(yy * 255 // size), (xx * 255 // size), np.where((yy // 32 + xx // 32) % 2, 255, 0)], -1).astype(np.uint8)
    ref = reference_transform(patch)
    got = O.eval_transform(patch[None], 224)[0]
    resized = np.asarray(Image.fromarray(patch).resize((224, 224), Image.BILINEAR))
    np.testing.assert_array_equal(O.resize_bilinear_u8(patch, 224), resized)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_fp16_store_rounding_is_numpy_astype():
    x = torch.randn(1000, generator=torch.Generator().manual_seed(0)).numpy() * 3
    assert np.array_equal(x.astype(np.float16), torch.from_numpy(x).half().numpy())      # Step2_feature_extract.py:165
