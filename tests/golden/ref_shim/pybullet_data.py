"""Import stub (see pybullet.py)."""


def getDataPath():
    return ""
