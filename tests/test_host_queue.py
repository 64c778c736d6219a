"""CPU tests of the host-side ticket queue (sdr-server_b200/host/ticket_queue.c):
the same three behaviours the reference pins for its block queue in
test/test_queue.c:23-59 -- FIFO order, overwrite-the-newest on overflow, and
drain-before-poison-pill -- plus the detached-entry rule (src/queue.c:150-158)."""
import ctypes as C
import threading
import time


def make(pkg, size):
    H = pkg.host_lib()
    q = C.c_void_p()
    assert H.xl_tq_create(size, C.byref(q)) == 0
    return H, q


def test_fifo_order(pkg):
    """test/test_queue.c:23-33 (test_put_take)"""
    H, q = make(pkg, 4)
    for t in (10, 11, 12):
        H.xl_tq_put(q, t)
    for t in (10, 11, 12):
        assert H.xl_tq_take(q) == t
        H.xl_tq_complete(q)
    H.xl_tq_destroy(q)


def test_overflow_overwrites_newest(pkg, capfd):
    """test/test_queue.c:35-40 (test_overflow): capacity 2, three puts -> 1st and 3rd survive"""
    H, q = make(pkg, 2)
    H.xl_tq_put(q, 1)
    H.xl_tq_put(q, 2)
    H.xl_tq_put(q, 3)  # queue is full: overwrites the newest queued entry (2)
    assert H.xl_tq_overruns(q) == 1
    assert H.xl_tq_take(q) == 1
    H.xl_tq_complete(q)
    assert H.xl_tq_take(q) == 3
    H.xl_tq_complete(q)
    H.xl_tq_destroy(q)
    assert "<3>queue is full" in capfd.readouterr().err


def test_detached_entry_is_not_overwritten(pkg):
    """the entry being processed occupies a slot and is never overwritten (src/queue.c:150-158)"""
    H, q = make(pkg, 2)
    H.xl_tq_put(q, 1)
    assert H.xl_tq_take(q) == 1      # detached, still holds a slot
    H.xl_tq_put(q, 2)                # second slot
    H.xl_tq_put(q, 3)                # full -> overwrites 2, not the detached 1
    H.xl_tq_complete(q)
    assert H.xl_tq_take(q) == 3
    H.xl_tq_complete(q)
    H.xl_tq_destroy(q)


def test_drain_before_poison_pill(pkg):
    """test/test_queue.c:42-59 (test_terminated_only_after_fully_processed)"""
    H, q = make(pkg, 4)
    H.xl_tq_put(q, 7)
    H.xl_tq_put(q, 8)
    H.xl_tq_interrupt(q)
    assert H.xl_tq_take(q) == 7
    H.xl_tq_complete(q)
    assert H.xl_tq_take(q) == 8
    H.xl_tq_complete(q)
    assert H.xl_tq_take(q) == -1
    H.xl_tq_destroy(q)


def test_blocking_take_wakes_on_put_and_interrupt(pkg):
    H, q = make(pkg, 4)
    got = []

    def consumer():
        while True:
            t = H.xl_tq_take(q)
            got.append(t)
            if t == -1:
                return
            H.xl_tq_complete(q)

    th = threading.Thread(target=consumer)
    th.start()
    time.sleep(0.05)
    H.xl_tq_put(q, 42)
    time.sleep(0.05)
    H.xl_tq_interrupt(q)
    th.join(timeout=5)
    assert not th.is_alive()
    assert got == [42, -1]
    H.xl_tq_destroy(q)
