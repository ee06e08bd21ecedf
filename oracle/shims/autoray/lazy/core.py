def lazy_cache(name, hasher=None):
    def deco(fn):
        return fn
    return deco
