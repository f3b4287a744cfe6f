def init_parsing_model(*a, **k):
    raise NotImplementedError
