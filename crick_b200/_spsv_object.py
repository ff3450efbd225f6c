"""Stream-Summary for ``SpaceSaving(dtype=object)`` -- host side by necessity.

Object items need ``PyObject_Hash`` / ``PyObject_RichCompareBool`` per element
(reference: crick/space_saving_stubs.c.in:20-29), which only the interpreter can evaluate, so
this one dtype cannot run on the GPU ("keep a host path for it").  It
is NOT a fallback for the numeric dtypes: int64 / float64 summaries always go through
libcrick_cuda.so.

The structure follows the reference's: an array-backed circular doubly linked list kept sorted
by (count desc, error asc), slots never move (space_saving_stubs.c.in:54-72), plus an
item -> slot map.  Every item is referenced exactly once (by its slot), like the reference's
single Py_INCREF per counter (:147-149, :203-204) -- the map is keyed by hash value and holds
slot numbers only.
"""

NIL = -1


def _same(a, b):
    # pyobject_cmp, space_saving_stubs.c.in:20-27: identity or ==, errors count as "different"
    if a is b:
        return True
    try:
        return bool(a == b)
    except Exception:
        return False


class ObjectSummary:
    __slots__ = ("capacity", "size", "head", "item", "count", "error", "next", "prev", "slots")

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.size = 0
        self.head = NIL
        self.item = []      # slot -> item (the only reference the summary holds)
        self.count = []
        self.error = []
        self.next = []
        self.prev = []
        self.slots = {}     # hash(item) -> [slot, ...]

    # ---- item -> slot map -------------------------------------------------------------
    def _find(self, item, h):
        for s in self.slots.get(h, ()):
            if _same(self.item[s], item):
                return s
        return NIL

    def _map_add(self, h, slot):
        self.slots.setdefault(h, []).append(slot)

    def _map_del(self, slot):
        h = hash(self.item[slot])
        lst = self.slots[h]
        lst.remove(slot)
        if not lst:
            del self.slots[h]

    # ---- list maintenance ---------------------------------------------------------------
    def _ge(self, a, cnt, err):
        # counter_ge, :112-118, against the pair (cnt, err)
        return self.count[a] > cnt or (self.count[a] == cnt and self.error[a] <= err)

    def _insert(self, c, prev):
        # counter_insert, :121-138: walk backwards from `prev` to the insertion point
        tail = self.prev[self.head]
        while True:
            if self._ge(prev, self.count[c], self.error[c]):
                break
            prev = self.prev[prev]
            if prev == tail:
                self.head = c
                break
        nxt = self.next[prev]
        self.next[c] = nxt
        self.prev[c] = prev
        self.prev[nxt] = c
        self.next[prev] = c

    def _new(self, item, count, error):
        # counter_new, :141-164
        c = self.size
        self.size += 1
        self.item.append(item)
        self.count.append(count)
        self.error.append(error)
        self.next.append(c)
        self.prev.append(c)
        if self.head == NIL:
            self.head = c
        else:
            self._insert(c, self.prev[self.head])
        return c

    def _rebalance(self, i):
        # rebalance, :167-183
        if self.head == i:
            return
        prev = self.prev[i]
        if self._ge(prev, self.count[i], self.error[i]):
            return
        nxt = self.next[i]
        self.prev[nxt] = prev
        self.next[prev] = nxt
        self._insert(i, prev)

    def _swap(self, i, item, h, count, error):
        # swap, :186-210: the slot keeps its place in the array, the old item is released
        self._map_del(i)
        self.item[i] = item
        self.count[i] = count
        self.error[i] = error
        self._map_add(h, i)
        self._rebalance(i)

    # ---- operations ---------------------------------------------------------------------
    def add(self, item, count):
        # spsv_add, :213-250 (a full summary ignores `count`: the new counter is min + 1)
        h = hash(item)            # unhashable -> TypeError before anything changes
        s = self._find(item, h)
        if s != NIL:
            self.count[s] += count
            self._rebalance(s)
        elif self.size == self.capacity:
            t = self.prev[self.head]
            m = self.count[t]
            self._swap(t, item, h, m + 1, m)
        else:
            self._map_add(h, self._new(item, count, 0))

    def set_state(self, items, counts, errors):
        # set_state, :253-286
        if len(items) > self.capacity:
            raise ValueError("deserialization failed, size > capacity")
        for it, c, e in zip(items, counts, errors):
            h = hash(it)
            if self._find(it, h) != NIL:
                raise ValueError("deserialization failed, duplicate items found")
            self._map_add(h, self._new(it, int(c), int(e)))

    def merge(self, other):
        # spsv_merge, :289-364 (Cafaro et al.).  Not replicated: the key the reference leaves
        # in the hash map when the second loop breaks (:334 / :345).
        if other.size == 0:
            return
        m1 = self.count[self.prev[self.head]] if self.size == self.capacity else 0
        m2 = other.count[other.prev[other.head]] if other.size == other.capacity else 0
        if other is self:
            # kh_get on the same table: every item is "in both"
            for i in range(self.size):
                self.count[i] += self.count[i]
                self.error[i] += self.error[i]
                self._rebalance(i)
            return
        for i in range(self.size):
            it = self.item[i]
            j = other._find(it, hash(it))
            if j != NIL:
                self.count[i] += other.count[j]
                self.error[i] += other.error[j]
            else:
                self.count[i] += m2
                self.error[i] += m2
            self._rebalance(i)
        j = other.head
        for _ in range(other.size):
            it, c, e = other.item[j], other.count[j] + m1, other.error[j] + m1
            h = hash(it)
            if self._find(it, h) == NIL:
                if self.size == self.capacity:
                    t = self.prev[self.head]
                    if self._ge(t, c, e):
                        break
                    self._swap(t, it, h, c, e)
                else:
                    self._map_add(h, self._new(it, c, e))
            j = other.next[j]

    def counters(self, k):
        # object_counters, space_saving.pyx:389-406: walk the list from the head
        out = []
        i = self.head
        for _ in range(min(self.size, k)):
            out.append((self.item[i], self.count[i], self.error[i]))
            i = self.next[i]
        return out
