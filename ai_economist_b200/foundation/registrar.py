"""Name -> class registries (same surface as the reference's Registry, base/registrar.py:8-103:
case-insensitive `add` / `get` / `has` / `entries`, names may not contain '.')."""


class Registry:
    def __init__(self, base_class=None):
        self.base_class = base_class
        self._by_lower = {}
        self._names = []

    def add(self, cls):
        name = getattr(cls, "name", None)
        if not isinstance(name, str) or not name:
            raise ValueError("registered classes need a non-empty string `name`")
        if "." in name:
            raise NameError("Class name {} is illegally named (no '.' allowed).".format(name))
        if self.base_class is not None and not issubclass(cls, self.base_class):
            raise TypeError("{} must subclass {}".format(cls, self.base_class))
        self._by_lower[name.lower()] = cls
        if name not in self._names:
            self._names.append(name)
        return cls

    def get(self, name):
        key = name.lower()
        if key not in self._by_lower:
            raise KeyError('"{}" is not a name of a registered class'.format(name))
        return self._by_lower[key]

    def has(self, name):
        return name.lower() in self._by_lower

    @property
    def entries(self):
        return sorted(self._names)
