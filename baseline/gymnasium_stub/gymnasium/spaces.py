class Space:
    pass


class Box(Space):
    pass


class Discrete(Space):
    pass


class Dict(Space):
    pass


class MultiDiscrete(Space):
    pass


class Tuple(Space):
    pass
