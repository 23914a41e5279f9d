def load_pretrained(*args, **kwargs):
    raise RuntimeError("pretrained timm weights are not available offline")
