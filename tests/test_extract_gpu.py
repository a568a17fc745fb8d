"""GPU parity of the feature-extraction loop: the fused resize/normalise kernel must be bit-exact with the Pillow-pinned
oracle (byte work), the fp16 cast identical to numpy's, and the loop equal to batched preprocessing + encoder."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,batch,seed", [(256, 5, 0), (512, 2, 1), (224, 3, 2), (300, 2, 3)])
def test_preprocess_bit_exact_vs_oracle(size, batch, seed):
    from acmil_b200.extract import preprocess
    from oracle import preprocess as O
    rng = np.random.default_rng(seed)
    patches = rng.integers(0, 256, (batch, size, size, 3), dtype=np.uint8)
    patches[0, : size // 2] = 255          # saturated region
    ref = O.eval_transform(patches, 224)
    got = preprocess(torch.from_numpy(patches).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_fp16_cast_and_extract_loop():
    from acmil_b200 import Struct
    from acmil_b200.extract import extract_feature, preprocess, to_fp16
    from acmil_b200.vit import CustomModel, VisionTransformer
    x = torch.randn(100000, generator=torch.Generator().manual_seed(0)) * 100
    assert np.array_equal(to_fp16(x.cuda()).cpu().numpy(), x.numpy().astype(np.float16))
    torch.manual_seed(0)
    enc = VisionTransformer(img_size=224, patch_size=16, embed_dim=96, depth=2, num_heads=3)
    model = CustomModel(Struct(n_class=2), enc).cuda().eval()
    rng = np.random.default_rng(5)
    patches = rng.integers(0, 256, (10, 256, 256, 3), dtype=np.uint8)
    feats = extract_feature(patches, model, batch_size=4)
    assert feats.shape == (10, 96) and feats.dtype == np.float32
    with torch.no_grad():
        _, ref = model(preprocess(torch.from_numpy(patches).cuda()), return_feature=True)
    np.testing.assert_allclose(feats, ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    with pytest.raises(RuntimeError):
        preprocess(torch.from_numpy(patches))
