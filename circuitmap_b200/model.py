"""placeholder -- replaced below"""
class Model:  # noqa
    pass
