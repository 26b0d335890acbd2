"""Two implementations of GgrsSnapshots<u32,u32> behind one test surface:
the oracle's (CPU restatement) and the engine's host-side ring bookkeeping (bgr_ring_* in the
product .so — pure host logic, runs without a GPU)."""
import ctypes as C

I32_MAX, I32_MIN = 2**31 - 1, -(2**31)


class RollbackPanic(Exception):
    pass


class OracleRing:
    def __init__(self, depth=None):
        from oracle_backend import load_oracle
        self.lib = load_oracle()
        self.h = C.c_void_p(self.lib.orc_ring_new(depth or 0, 1 if depth is not None else 0))

    def depth(self):
        return self.lib.orc_ring_depth(self.h)

    def set_depth(self, d):
        self.lib.orc_ring_set_depth(self.h, d)

    def push(self, frame, value):
        self.lib.orc_ring_push(self.h, frame, value)

    def confirm(self, frame):
        self.lib.orc_ring_confirm(self.h, frame)

    def rollback(self, frame):
        if self.lib.orc_ring_rollback(self.h, frame) != 0:
            raise RollbackPanic(self.lib.orc_last_error().decode())

    def get(self):
        v = C.c_uint32()
        if self.lib.orc_ring_get(self.h, C.byref(v)) != 0:
            raise RollbackPanic(self.lib.orc_last_error().decode())
        return v.value

    def peek(self, frame):
        v = C.c_uint32()
        return v.value if self.lib.orc_ring_peek(self.h, frame, C.byref(v)) else None

    def __del__(self):
        try:
            self.lib.orc_ring_free(self.h)
        except Exception:
            pass


class EngineRing:
    """bgr_ring_*: the host bookkeeping the engine uses to map frames to HBM slots.  The ring
    stores slot ids; the test keeps the u32 payload per slot like HBM keeps the planes."""

    def __init__(self, depth=None):
        from bevy_ggrs_b200 import capi
        self.lib = capi.load_library()
        self.cap = 64
        self.h = C.c_void_p(self.lib.bgr_ring_create(self.cap))
        self.payload = {}
        if depth is not None:
            self.set_depth(depth)

    def _err(self):
        return self.lib.bgr_last_error().decode()

    def depth(self):
        return self.lib.bgr_ring_depth(self.h)

    def set_depth(self, d):
        assert self.lib.bgr_ring_set_depth(self.h, d) == 0, self._err()

    def push(self, frame, value):
        slot = C.c_uint32()
        assert self.lib.bgr_ring_push(self.h, frame, C.byref(slot)) == 0, self._err()
        self.payload[slot.value] = value

    def confirm(self, frame):
        self.lib.bgr_ring_confirm(self.h, frame)

    def rollback(self, frame):
        slot = C.c_uint32()
        if self.lib.bgr_ring_rollback(self.h, frame, C.byref(slot)) != 0:
            raise RollbackPanic(self._err())

    def get(self):
        slot = C.c_uint32()
        if self.lib.bgr_ring_get(self.h, C.byref(slot)) != 0:
            raise RollbackPanic(self._err())
        return self.payload[slot.value]

    def peek(self, frame):
        slot = C.c_uint32()
        found = C.c_int32()
        self.lib.bgr_ring_peek(self.h, frame, C.byref(slot), C.byref(found))
        return self.payload[slot.value] if found.value else None

    def __del__(self):
        try:
            self.lib.bgr_ring_destroy(self.h)
        except Exception:
            pass
