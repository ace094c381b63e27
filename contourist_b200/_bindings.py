"""ctypes prototypes for the 2D / 4D / post-processing entry points."""


def bind(lib):
    pass
