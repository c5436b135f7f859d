"""`colorama` stand-in for the reference decoder (decoder.py:28) — TEST INFRASTRUCTURE ONLY."""


class _Codes(object):
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Back = _Codes()
Style = _Codes()


def init(*a, **k):
    pass
