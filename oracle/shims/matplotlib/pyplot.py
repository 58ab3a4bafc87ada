"""pyplot stand-in: every call is swallowed; `h, = plt.plot(...)` unpacks to one inert handle."""
import sys as _sys


class _Inert:
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter((self,))

    def __getitem__(self, i):
        return self

    def __len__(self):
        return 1

    def __bool__(self):
        return True


_inert = _Inert()


def __getattr__(name):
    if name.startswith('__') and name.endswith('__'):
        raise AttributeError(name)
    return _inert
