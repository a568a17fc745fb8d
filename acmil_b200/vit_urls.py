"""Checkpoint registry of the reference's ViT-S encoders (models.py:113-123)."""


def get_pretrained_url(key):
    prefix = "https://github.com/lunit-io/benchmark-ssl-pathology/releases/download/pretrained-weights"
    names = {"BT": "bt_rn50_ep200.torch", "MoCoV2": "mocov2_rn50_ep200.torch", "SwAV": "swav_rn50_ep200.torch",
             "DINO_p16": "dino_vit_small_patch16_ep200.torch", "DINO_p8": "dino_vit_small_patch8_ep200.torch"}
    return f"{prefix}/{names.get(key)}"
