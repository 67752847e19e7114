"""Oracle restatement of the ``easydict`` package (TEST INFRASTRUCTURE).

``easydict`` (any 1.x; unpinned in the reference's ``environment.yml`` pip list)
is a dict whose keys are also attributes and which wraps nested plain dicts
recursively.  The reference's hot path relies on three behaviours
(``models/deblur_e_nerf.py:397,420,442,459,495``):

* ``EasyDict(batch)`` recurses into nested dicts,
* ``batch.diff = {}`` turns the assigned dict into an ``EasyDict``,
* ``pop`` removes both the item and the attribute.
"""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        merged = {} if d is None else dict(d)
        merged.update(kwargs)
        for key, value in merged.items():
            self[key] = value

    @classmethod
    def _wrap(cls, value):
        if isinstance(value, EasyDict):
            return value
        if isinstance(value, dict):
            return cls(value)
        if isinstance(value, (list, tuple)):
            return type(value)(cls._wrap(v) for v in value)
        return value

    def __setattr__(self, name, value):
        value = self._wrap(value)
        dict.__setitem__(self, name, value)
        object.__setattr__(self, name, value)

    __setitem__ = __setattr__

    def __delattr__(self, name):
        if name in self:
            dict.__delitem__(self, name)
        if name in self.__dict__:
            object.__delattr__(self, name)

    def update(self, e=None, **f):
        merged = dict(e or {})
        merged.update(f)
        for key, value in merged.items():
            self[key] = value

    def pop(self, key, *default):
        if key in self.__dict__:
            object.__delattr__(self, key)
        return dict.pop(self, key, *default)
