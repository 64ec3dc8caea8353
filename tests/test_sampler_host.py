"""Host logic of the device-resident sampler system (pyiid_b200/sim.py:
_DeviceSystem) against a fake backend: the look-ahead chains, the accounting of
the device state slots and the recovery from a dropped chain need no GPU."""
import collections
import gc

import numpy as np

from pyiid_b200 import _lib, sim


class FakeBackend(object):
    """Harmonic dynamics behind the chain interface of backend.Backend: slot ->
    (q, p); a chain is computed at begin and handed out by next; any other call
    drops it (as the native library does)."""

    def __init__(self, n):
        self.n = n
        self.slots = {}
        self.chain = None
        self.next_id = 0
        self.begun = []
        self.in_flight_writes = set()

    def _other_call(self):
        self.chain = None

    def state_upload(self, slot, q, p, f):
        self._other_call()
        self.slots[slot] = (np.array(q), np.array(p))

    def state_download(self, slot, want=('q', 'p', 'f')):
        self._other_call()
        q, p = self.slots[slot]
        return {'q': q, 'p': p, 'f': -2. * q}

    def leapfrog_chain_begin(self, src, dsts, step, center, target, potential, conv):
        self._other_call()
        assert len(set(dsts)) == len(dsts) and src not in dsts
        q, p = self.slots[src]
        out = []
        for d in dsts:
            p = p + 0.5 * step * (-2. * q)
            q = q + step * p
            p = p + 0.5 * step * (-2. * q)
            self.slots[d] = (q, p)
            out.append((float((q ** 2).sum()), 1., 0., 0.5 * float((p ** 2).sum()), q, p))
        self.next_id += 1
        self.chain = [self.next_id, collections.deque(out)]
        self.begun.append(len(dsts))
        return self.next_id

    def leapfrog_chain_next(self, chain_id):
        if self.chain is None or self.chain[0] != chain_id or not self.chain[1]:
            raise _lib.ChainDropped('dropped')
        return self.chain[1].popleft()


class Calc(object):
    target_data = None
    potential_name = 'rw'
    rw_to_eV = 1.


def make_system(n=5, slots=40):
    s = object.__new__(sim._DeviceSystem)
    s.calc = Calc()
    s.masses = np.ones((n, 1))
    s.be = FakeBackend(n)
    s.pool = sim._SlotPool(slots)
    s._ahead = collections.deque()
    s._ahead_prev = None
    s._ahead_step = 0.
    s._ahead_id = 0
    s._expected = 0
    s.evals = 0
    return s


def start_state(s, seed=0):
    rs = np.random.RandomState(seed)
    q, p = rs.normal(size=(s.be.n, 3)), rs.normal(size=(s.be.n, 3))
    slot = s.pool.take()
    s.be.state_upload(slot, q, p, -2. * q)
    return sim._DevState(q, p, float((q ** 2).sum()), -2. * q, 0.5 * float((p ** 2).sum()), slot,
                         s.pool)


def walk(s, st, n, step=0.01):
    out = []
    for _ in range(n):
        st = s.leapfrog(st, step)
        out.append(st)
    return out


def test_look_ahead_gives_the_step_by_step_states_and_returns_every_slot():
    s = make_system()
    st = start_state(s)
    ref = walk(s, st, 7)
    assert s.be.begun == [1] * 7
    s.be.begun = []
    s.expect(7)
    got = walk(s, st, 7)
    assert s.be.begun == [7]  # one chain
    for a, b in zip(ref, got):
        assert np.array_equal(a.q, b.q) and np.array_equal(a.p, b.p) and a.pe == b.pe
    # a subtree that stops early: the steps enqueued ahead go back to the pool
    s.expect(16)
    part = walk(s, st, 3)
    free_before = len(s.pool.free)
    s.close()
    assert len(s.pool.free) == free_before + 13
    del ref, got, part, st, a, b
    gc.collect()
    assert len(s.pool.free) == 40 and len(set(s.pool.free)) == 40


def test_another_trajectory_or_a_dropped_chain_starts_a_new_one():
    s = make_system()
    st = start_state(s)
    ref = walk(s, st, 6, 0.02)
    s.expect(6)
    a = s.leapfrog(st, 0.02)
    b = s.leapfrog(st, -0.02)  # not the continuation: the chain is void
    assert np.array_equal(a.q, ref[0].q) and not np.array_equal(b.q, ref[0].q)
    # the backend drops the chain behind the system's back (another call on it)
    s.expect(6)
    x = s.leapfrog(st, 0.02)
    s.be.state_download(x.slot)
    rest = walk(s, x, 5, 0.02)
    for r, g in zip(ref[1:], rest):
        assert np.array_equal(r.q, g.q) and np.array_equal(r.p, g.p)
    # a system that is simply dropped gives its look-ahead slots back
    s.expect(8)
    y = s.leapfrog(st, 0.02)
    pool = s.pool
    del s, a, b, x, y, rest, ref, st, r, g
    gc.collect()
    assert len(pool.free) == 40


def test_chain_length_is_capped_and_uncentred_steps_are_single():
    s = make_system(slots=200)
    st = start_state(s)
    s.expect(1000)
    walk(s, st, sim._DeviceSystem.CHAIN + 3)
    assert s.be.begun[0] == sim._DeviceSystem.CHAIN and s.be.begun[1] == sim._DeviceSystem.CHAIN
    s.close()
    s.be.begun = []
    s.expect(10)
    s.leapfrog(st, 0.01, center=False)
    assert s.be.begun == [1]
