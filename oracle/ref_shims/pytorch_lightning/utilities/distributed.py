def rank_zero_info(*args, **kwargs):
    pass


def rank_zero_only(fn):
    return fn
