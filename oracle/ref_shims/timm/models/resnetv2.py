class ResNetV2:
    def __init__(self, *a, **k):
        raise RuntimeError("not available in the oracle shims")
