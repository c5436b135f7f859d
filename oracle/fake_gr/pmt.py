"""Identity `pmt` stand-in — TEST INFRASTRUCTURE ONLY.

The reference only round-trips Python values through PMT (`to_pmt` / `to_python` /
`cons` / `car` / `cdr`), so identity conversions preserve every value it produces.
"""


def to_pmt(x):
    return x


def to_python(x):
    return x


def intern(s):
    return s


def cons(a, b):
    return (a, b)


def car(p):
    return p[0]


def cdr(p):
    return p[1]
