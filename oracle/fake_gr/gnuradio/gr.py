"""Minimal `gnuradio.gr` surface used by the reference blocks — TEST INFRASTRUCTURE ONLY.

Covers what framer.py / demod.py / decoder.py call:
`sync_block.__init__`, `set_history`, `set_tag_propagation_policy`, `nitems_written`,
`nitems_read`, `add_item_tag`, `get_tags_in_range`, `message_port_register_*`,
`set_msg_handler`, `message_port_pub`.  The driver (`oracle/run_reference.py`) owns
the item counters, the tag store and the message sinks.
"""

TPP_DONT = 0
TPP_ALL_TO_ALL = 1
TPP_ONE_TO_ONE = 2


class tag_t(object):
    __slots__ = ("offset", "key", "value", "srcid")

    def __init__(self, offset, key, value, srcid=None):
        self.offset = offset
        self.key = key
        self.value = value
        self.srcid = srcid


class sync_block(object):
    def __init__(self, name=None, in_sig=None, out_sig=None):
        self._name = name
        self._in_sig = in_sig
        self._out_sig = out_sig
        self._history = 1
        self._tpp = TPP_ALL_TO_ALL
        self._nitems_written = 0
        self._nitems_read = 0
        self._tags_out = []      # tags this block added (absolute offsets)
        self._tags_in = []       # tags visible on input port 0
        self._msgs = {}          # out port -> list of published messages
        self._handlers = {}

    # --- configuration -------------------------------------------------
    def set_history(self, n):
        self._history = int(n)

    def history(self):
        return self._history

    def set_tag_propagation_policy(self, p):
        self._tpp = p

    def name(self):
        return self._name

    # --- item counters / tags -------------------------------------------
    def nitems_written(self, port):
        return self._nitems_written

    def nitems_read(self, port):
        return self._nitems_read

    def add_item_tag(self, port, offset, key, value, srcid=None):
        self._tags_out.append(tag_t(int(offset), key, value, srcid))

    def get_tags_in_range(self, port, start, end, key=None):
        out = [t for t in self._tags_in
               if start <= t.offset < end and (key is None or t.key == key)]
        out.sort(key=lambda t: t.offset)
        return out

    # --- message ports ---------------------------------------------------
    def message_port_register_out(self, port):
        self._msgs.setdefault(port, [])

    def message_port_register_in(self, port):
        self._handlers.setdefault(port, None)

    def set_msg_handler(self, port, fn):
        self._handlers[port] = fn

    def message_port_pub(self, port, msg):
        self._msgs.setdefault(port, []).append(msg)


basic_block = sync_block
