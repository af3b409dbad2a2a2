"""Minimal stand-ins for gymnasium.spaces used by the reference env constructors."""


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Discrete(Space):
    def __init__(self, n=0):
        self.n = n


class MultiDiscrete(Space):
    def __init__(self, nvec=()):
        self.nvec = nvec


class Dict(Space):
    def __init__(self, spaces=None):
        self.spaces = spaces or {}


class Tuple(Space):
    def __init__(self, spaces=()):
        self.spaces = spaces
