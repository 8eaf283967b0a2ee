import importlib


class EasyDict(dict):
    """dict with attribute access (`c.batch_size`), as the reference's option container."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


def get_obj_by_name(name):
    """'pkg.mod.Attr.sub' -> object: the longest importable module prefix, then attribute lookups."""
    parts = name.split(".")
    for cut in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:cut]))
        except ImportError:
            continue
        try:
            for attr in parts[cut:]:
                obj = getattr(obj, attr)
            return obj
        except AttributeError:
            continue
    raise ImportError("cannot resolve object name %r" % name)


def call_func_by_name(*args, func_name=None, **kwargs):
    assert func_name is not None
    fn = get_obj_by_name(func_name)
    assert callable(fn), func_name
    return fn(*args, **kwargs)


def construct_class_by_name(*args, class_name=None, **kwargs):
    """`construct_class_by_name(params=..., class_name='torch.optim.Adam', lr=...)` as sid_training_loop.py:291-292."""
    return call_func_by_name(*args, func_name=class_name, **kwargs)


def format_time(seconds):
    s = int(round(float(seconds)))
    if s < 60:
        return "%ds" % s
    if s < 3600:
        return "%dm %02ds" % (s // 60, s % 60)
    if s < 86400:
        return "%dh %02dm %02ds" % (s // 3600, (s // 60) % 60, s % 60)
    return "%dd %02dh %02dm" % (s // 86400, (s // 3600) % 24, (s // 60) % 60)
