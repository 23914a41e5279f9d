class ModelCheckpoint:
    def __init__(self, *a, **k):
        pass


class LearningRateMonitor:
    def __init__(self, *a, **k):
        pass
