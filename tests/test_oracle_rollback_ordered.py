"""The reference's 5 RollbackOrdered unit tests (src/snapshot/rollback.rs:112-161) against the oracle's
restatement.  In the engine the same contract holds by construction: order == order_base + row, rows are
appended in insertion order and never re-used inside a run (bgr_spawn), and an unknown row is rejected."""
import ctypes as C

import pytest

from oracle_backend import load_oracle


class Ordered:
    def __init__(self, ids=(), handle=None):
        self.lib = load_oracle()
        self.h = C.c_void_p(handle if handle is not None else self.lib.orc_ordered_new())
        for n in ids:
            self.push(n)

    def push(self, n):
        self.lib.orc_ordered_push(self.h, n)

    def order(self, n):
        out = C.c_uint64()
        if self.lib.orc_ordered_order(self.h, n, C.byref(out)) != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return out.value

    def len(self):
        return self.lib.orc_ordered_len(self.h)

    def iter_sorted(self):
        buf = (C.c_uint64 * 64)()
        n = self.lib.orc_ordered_iter_sorted(self.h, buf, 64)
        return [buf[i] for i in range(n)]

    def clone(self):
        return Ordered(handle=self.lib.orc_ordered_clone(self.h))


def test_order_returns_insertion_index():  # rollback.rs:114-120
    ro = Ordered([10, 20, 30])
    assert ro.order(10) == 0 and ro.order(20) == 1 and ro.order(30) == 2


def test_iter_sorted_yields_insertion_order():  # rollback.rs:123-128
    assert Ordered([5, 3, 7, 1]).iter_sorted() == [5, 3, 7, 1]


def test_order_is_stable_after_more_pushes():  # rollback.rs:131-140
    ro = Ordered([0, 1])
    before = (ro.order(0), ro.order(1))
    ro.push(2)
    ro.push(3)
    assert (ro.order(0), ro.order(1)) == before


def test_order_unregistered_panics():  # rollback.rs:143-148
    ro = Ordered([0])
    with pytest.raises(RuntimeError, match="RollbackId was not registered in RollbackOrdered!"):
        ro.order(99)


def test_clone_is_independent():  # rollback.rs:151-160
    ro = Ordered([1, 2, 3])
    clone = ro.clone()
    clone.push(4)
    assert ro.len() == 3 and clone.len() == 4
    assert ro.order(1) == clone.order(1)
