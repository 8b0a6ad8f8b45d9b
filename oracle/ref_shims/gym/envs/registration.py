REGISTRY = {}


def register(id, entry_point=None, **kw):
    REGISTRY[id] = entry_point
