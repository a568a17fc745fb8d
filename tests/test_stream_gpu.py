"""BASELINE.json configs[4] as a parity case: the end-to-end stream  uint8 patches -> Resize(224) + normalise -> patch
encoder -> fp16 feature store (Step2_feature_extract.py:35-71, 165) -> fp32 cast (Step3_WSI_classification_ACMIL.py:193)
-> ACMIL head  on the GPU against the chained oracles (Pillow-pinned preprocessing, encoder restatement, gated-pool
restatement) on the same seeded slide.  Small encoder / few patches so that the numpy side finishes in seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _np_state(m):
    return {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("encoder", ["vit", "resnet18"])
def test_extract_then_head_matches_oracles(encoder):
    from acmil_b200 import ACMIL_GA, Struct
    from acmil_b200.extract import extract_feature
    from acmil_b200.vit import CustomModel, VisionTransformer
    from oracle import gated_pool as OG, preprocess as OP, resnet as OR, vit as OV
    from resnet_common import randomize_bn

    n_patches = 150 if encoder == "vit" else 48
    rng = np.random.default_rng(11)
    patches = rng.integers(0, 256, (n_patches, 96, 96, 3), dtype=np.uint8)
    torch.manual_seed(3)
    if encoder == "vit":
        enc = VisionTransformer(img_size=224, patch_size=16, embed_dim=96, depth=2, num_heads=3)
        d_feat = 96
    else:
        from acmil_b200.resnet import resnet18
        enc = randomize_bn(resnet18(pretrained=False), 4)
        enc.class_classifier = torch.nn.Identity()
        enc.embed_dim = d_feat = 512
    extractor = CustomModel(Struct(n_class=2), enc).eval()
    head = ACMIL_GA(Struct(D_feat=d_feat, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).eval()
    p_ext, p_head = _np_state(extractor), _np_state(head)

    # ---- oracle chain on the host
    imgs = OP.eval_transform(patches, 224)
    if encoder == "vit":
        _, feats_ref = OV.custom_model_forward(p_ext, imgs, num_heads=3, patch=16)
    else:
        feats_ref = OR.resnet18_features({k[len("encoder."):]: v for k, v in p_ext.items() if k.startswith("encoder.")}, imgs)
    stored = feats_ref.astype(np.float16)                                   # the H5 'feat' dataset
    ref = OG.acmil_ga_forward(p_head, stored.astype(np.float32)[None])

    # ---- the GPU stream through the public API
    extractor, head = extractor.cuda(), head.cuda()
    feats = extract_feature(patches, extractor, batch_size=64)               # [N, D] fp32 on the host, like the reference
    np.testing.assert_allclose(feats, feats_ref, rtol=1e-3, atol=1e-3 * np.abs(feats_ref).max())
    bag = torch.from_numpy(feats.astype(np.float16)).cuda()                  # fp16 features, cast on the device
    with torch.no_grad():
        sub, slide, attn = head(bag[None])
    np.testing.assert_allclose(slide.cpu().numpy(), ref["slide"], rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(sub.cpu().numpy(), ref["sub"], rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(attn.cpu().numpy().reshape(ref["A_out"].shape), ref["A_out"], rtol=1e-3, atol=2e-3)
