"""Model of CPython's ``set`` table order for (int, int) tuples (oracle; test infra).

The reference's beam search (dummy_env_rectangular_pin.py:1356-1423) sorts
``points_to_visit - visited`` -- a set difference -- by distance with a stable
sort and keeps the first ``beam_width`` entries, so ties at the cut-off are
decided by the iteration order of that difference set.  The CUDA reward kernel
reproduces that order (csrc/pcb_beam.cuh); this module is the executable
specification the kernel was written from.  tests/test_pyset_model.py checks it
against real ``set`` objects of the running interpreter (CPython 3.8+ share the
tuple hash and the set probing scheme).
"""
M64 = (1 << 64) - 1
P1, P2, P5 = 11400714785074694791, 14029467366897019727, 2870177450012600261
LINEAR_PROBES, PERTURB_SHIFT = 9, 5


def tuple_hash(x, y):
    """Objects/tupleobject.c tuplehash for a 2-tuple of small non-negative ints."""
    acc = P5
    for lane in (x, y):
        acc = (acc + lane * P2) & M64
        acc = ((acc << 31) | (acc >> 33)) & M64
        acc = (acc * P1) & M64
    acc = (acc + (2 ^ (P5 ^ 3527539))) & M64
    return 1546275796 if acc == M64 else acc


class Table:
    """Open-addressing table of Objects/setobject.c (keys = indices into a point
    list; EMPTY = -1, DUMMY = -2)."""

    def __init__(self, hashes):
        self.h = hashes
        self.mask = 7
        self.slots = [-1] * 8
        self.fill = 0
        self.used = 0

    def _probe_free(self, slots, mask, h):
        # set_insert_clean / the miss path of set_add_entry
        perturb = h
        i = h & mask
        while True:
            if slots[i] == -1:
                return i
            if i + LINEAR_PROBES <= mask:
                for j in range(1, LINEAR_PROBES + 1):
                    if slots[i + j] == -1:
                        return i + j
            perturb >>= PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask

    def _resize(self, minused):
        size = 8
        while size <= minused:
            size <<= 1
        old = [k for k in self.slots if k >= 0]
        self.mask = size - 1
        self.slots = [-1] * size
        for k in old:
            self.slots[self._probe_free(self.slots, self.mask, self.h[k])] = k
        self.fill = self.used = len(old)

    def add(self, k):
        """set_add_entry for a key known to be absent (no dummies present)."""
        self.slots[self._probe_free(self.slots, self.mask, self.h[k])] = k
        self.fill += 1
        self.used += 1
        if self.fill * 5 >= self.mask * 3:
            self._resize(self.used * 4 if self.used <= 50000 else self.used * 2)

    def order(self):
        return [k for k in self.slots if k >= 0]

    def copy(self):
        """set_copy -> set_merge into a fresh table."""
        t = Table(self.h)
        if self.used == 0:
            return t
        if (t.fill + self.used) * 5 >= t.mask * 3:
            t._resize((t.used + self.used) * 2)
        if t.mask == self.mask and self.fill == self.used:
            t.slots = list(self.slots)
        else:
            for k in self.slots:
                if k >= 0:
                    t.slots[t._probe_free(t.slots, t.mask, self.h[k])] = k
        t.fill = t.used = self.used
        return t


def difference_order(targets, seen):
    """Iteration order of ``targets_set - seen_set`` (set_difference).
    targets: Table; seen: python set of keys."""
    if (targets.used >> 2) > len(seen):
        t = targets.copy()  # set_copy_and_difference: copy, then discard -> dummies
        return [k for k in t.order() if k not in seen]
    out = Table(targets.h)
    for k in targets.order():
        if k not in seen:
            out.add(k)
    return out.order()


def build(points):
    """set(points) for a list of distinct (x, y)."""
    t = Table([tuple_hash(x, y) for x, y in points])
    for k in range(len(points)):
        t.add(k)
    return t
