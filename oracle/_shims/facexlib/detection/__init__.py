def init_detection_model(*a, **k):
    raise NotImplementedError
