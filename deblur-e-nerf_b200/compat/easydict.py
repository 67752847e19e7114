"""`easydict.EasyDict`: a dict whose items are attributes too, nested dicts wrapped on the way in.
Written for the behaviours the reference leans on (`EasyDict(yaml)`, `batch.diff = {}`, `pop`,
`**config.data`), not copied from the upstream package."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        dict.__init__(self)
        for source in (d or {}, kwargs):
            for key in source:
                self[key] = source[key]

    def __setitem__(self, key, value):
        value = _wrapped(value)
        dict.__setitem__(self, key, value)
        if isinstance(key, str):
            self.__dict__[key] = value

    def __setattr__(self, key, value):
        self[key] = value

    def __delitem__(self, key):
        dict.__delitem__(self, key)
        self.__dict__.pop(key, None)

    __delattr__ = __delitem__

    def pop(self, key, *default):
        self.__dict__.pop(key, None)
        return dict.pop(self, key, *default)

    def update(self, other=None, **kwargs):
        for source in (other or {}, kwargs):
            for key in dict(source):
                self[key] = source[key]

    def setdefault(self, key, default=None):
        if key not in self:
            self[key] = default
        return self[key]


def _wrapped(value):
    if type(value) is dict:
        return EasyDict(value)
    if isinstance(value, (list, tuple)) and not isinstance(value, str):
        return type(value)(_wrapped(v) for v in value)
    return value
