def guess_format(path):
    return None
