"""Minimal stand-in for `gymnasium` (absent from this image) -- TEST INFRASTRUCTURE ONLY.

It exists so that the *unmodified* reference env classes under /root/reference can be
imported by oracle/ref_loader.py when golden vectors are (re)generated in the build
container.  Nothing in the product package imports this.
"""


class Env:
    metadata = {}

    def __init__(self, *args, **kwargs):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    @property
    def unwrapped(self):
        return self.env.unwrapped


from . import spaces  # noqa: E402,F401
