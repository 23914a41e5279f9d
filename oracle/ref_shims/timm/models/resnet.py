def resnet26d(*a, **k):
    raise RuntimeError("not available in the oracle shims")


resnet50d = resnet26d
