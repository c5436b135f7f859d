"""GNU Radio / PMT binding for the blocks.

With GNU Radio installed the blocks subclass the real `gr.sync_block` and use the real `pmt`.
Without it (this repository's test and bench environments) a minimal in-process stand-in with the
same method names is used, which records tags and messages so a test harness can play scheduler.
"""
try:                                        # pragma: no cover - not available in the build image
    from gnuradio import gr                 # type: ignore
    import pmt                              # type: ignore
    HAVE_GNURADIO = hasattr(gr, "top_block")
except Exception:                           # ImportError or a partial install
    gr = None
    pmt = None
    HAVE_GNURADIO = False

if not HAVE_GNURADIO:
    class _Tag(object):
        __slots__ = ("offset", "key", "value", "srcid")

        def __init__(self, offset, key, value, srcid):
            self.offset, self.key, self.value, self.srcid = offset, key, value, srcid

    class _SyncBlock(object):
        """The slice of gr.sync_block the ADS-B blocks use."""

        def __init__(self, name=None, in_sig=None, out_sig=None):
            self._name, self._in_sig, self._out_sig = name, in_sig, out_sig
            self._history = 1
            self._tpp = None
            self._nitems_written = 0
            self._nitems_read = 0
            self._tags_out = []
            self._tags_in = []
            self._msgs = {}

        def name(self):
            return self._name

        def set_history(self, n):
            self._history = int(n)

        def history(self):
            return self._history

        def set_tag_propagation_policy(self, p):
            self._tpp = p

        def nitems_written(self, port):
            return self._nitems_written

        def nitems_read(self, port):
            return self._nitems_read

        def add_item_tag(self, port, offset, key, value, srcid=None):
            self._tags_out.append(_Tag(int(offset), key, value, srcid))

        def get_tags_in_range(self, port, start, end, key=None):
            sel = [t for t in self._tags_in if start <= t.offset < end and (key is None or t.key == key)]
            sel.sort(key=lambda t: t.offset)
            return sel

        def message_port_register_out(self, port):
            self._msgs.setdefault(port, [])

        def message_port_register_in(self, port):
            self._handlers = getattr(self, "_handlers", {})
            self._handlers.setdefault(port, None)

        def set_msg_handler(self, port, fn):
            self._handlers = getattr(self, "_handlers", {})
            self._handlers[port] = fn

        def message_port_pub(self, port, msg):
            self._msgs.setdefault(port, []).append(msg)

    class _Gr(object):
        sync_block = _SyncBlock
        basic_block = _SyncBlock
        TPP_DONT, TPP_ALL_TO_ALL, TPP_ONE_TO_ONE = 0, 1, 2

    class _Pmt(object):
        """Identity PMT: the blocks only round-trip Python values."""

        @staticmethod
        def to_pmt(x):
            return x

        @staticmethod
        def to_python(x):
            return x

        @staticmethod
        def intern(s):
            return s

        @staticmethod
        def cons(a, b):
            return (a, b)

        @staticmethod
        def car(p):
            return p[0]

        @staticmethod
        def cdr(p):
            return p[1]

    gr = _Gr()
    pmt = _Pmt()
