"""Stub so that the reference's utils/utils.py (imports matplotlib at :11-12, not installed here) can be imported
by oracle/gen_golden.py.  Test infrastructure only."""


def rc(*a, **k):
    pass


def use(*a, **k):
    pass
