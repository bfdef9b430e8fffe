"""Host logic of the observation tapes (Swarm heads, wrap-around move-to-top, step_many chunking,
independent X / A heads) against the reference's deque semantics (MRS.py:87-114) -- on CPU, with a
test double standing in for the C ABI (it records which slot every call writes; no compute)."""
import ctypes
import random
from collections import deque

import pytest
import torch

from mrsgym_b200 import _abi
from mrsgym_b200.core import Swarm


class FakeLib:
    """Implements the tape-touching ABI calls on the Swarm's (CPU) tape tensors: a step writes the
    step's serial number into the slot it was told to write."""

    def __init__(self, sw):
        self.sw = sw
        self.serial = 0

    def _slots(self, ptr_unused):
        return self.sw.X_tape, self.sw.A_tape

    def mrs_step(self, cfg, bufs, actions, slot_x, slot_a, stream):
        return self.mrs_step_many(cfg, bufs, actions, 1, slot_x, slot_a, stream)

    def mrs_step_many(self, cfg, bufs, actions, T, slot_x, slot_a, stream):
        X, A = self.sw.X_tape, self.sw.A_tape
        for t in range(T):
            self.serial += 1
            assert 0 <= slot_x - t < self.sw.L and 0 <= slot_a - t < self.sw.L
            X[slot_x - t] = float(self.serial)
            A[slot_a - t] = float(self.serial)
        return 0

    def mrs_observe(self, cfg, bufs, slot, write_X, write_A, stream):
        assert 0 <= slot < self.sw.L
        if write_X:
            self.sw.X_tape[slot] = float(self.serial)          # X of the current state = last step's serial
        if write_A:
            self.sw.A_tape[slot] = float(self.serial) + 0.5     # marks "A pushed outside step"
        return 0

    def mrs_tape_fill(self, cfg, bufs, which, src, dst_first, count, stream):
        tape = self.sw.X_tape if which == 1 else self.sw.A_tape
        assert 0 <= dst_first and dst_first + count <= self.sw.L
        assert not (0 <= src and dst_first <= src < dst_first + count)
        for i in range(count):
            tape[dst_first + i] = tape[src] if src >= 0 else 0.0
        return 0


def make_swarm(K, L, ring=False, fresh=False):
    sw = object.__new__(Swarm)              # bypass the CUDA-only constructor: host logic only
    sw.E, sw.N, sw.K, sw.L, sw.S, sw.D = 1, 1, K, L, 1, 1
    sw.cfg = _abi.MrsConfig()
    sw.cfg.E = sw.cfg.N = 1
    sw.cfg.K, sw.cfg.L = K, L
    sw.cfg.action_type = _abi.NO_ACTION
    sw.cfg.state_layout = _abi.X_POS_VEL
    sw.bufs = _abi.MrsBuffers()
    sw.X_tape = torch.full((L, 1, 1, 1), -1.0)
    sw.A_tape = torch.full((L, 1, 1, 1), -1.0)
    sw.device = torch.device('cpu')
    sw.launches = 0
    sw.hx = sw.ha = L - K - 1
    sw.ring = ring
    sw.fresh = fresh and not ring
    sw.generation = 0
    sw.a_empty = True
    sw.lib = FakeLib(sw)
    sw._stream = lambda: None
    sw._mrs_step = sw.lib.mrs_step                      # what Swarm.__init__ caches for the per-step call
    sw._cfg_ref, sw._bufs_ref = ctypes.byref(sw.cfg), ctypes.byref(sw.bufs)
    sw._launches_per_step = 1
    return sw


class RefRings:
    """The reference's two deques (MRS.calc_Xk / calc_Ak / reset, MRS.py:87-114,185-192)."""

    def __init__(self, K):
        self.K = K
        self.X, self.A = deque(), deque()

    def reset(self, x0):
        self.X, self.A = deque(), deque()
        self.push_X(x0)

    def push_X(self, x):
        self.X.appendleft(x)
        if len(self.X) > self.K + 1:
            self.X.pop()
        while len(self.X) < self.K + 1:
            self.X.append(x)

    def push_A(self, a):
        self.A.appendleft(a)
        if len(self.A) > self.K + 1:
            self.A.pop()
        while len(self.A) < self.K + 1:
            self.A.append(0.0)


@pytest.mark.parametrize('ring', [False, True, 'fresh'])
@pytest.mark.parametrize('K,L', [(0, 2), (0, 5), (1, 4), (2, 6), (3, 8), (3, 16), (5, 12)])
def test_tape_windows_follow_the_reference_deques(K, L, ring):
    rnd = random.Random(K * 100 + L)
    fresh = ring == 'fresh'           # fresh tapes: windows handed out earlier are never overwritten
    sw = make_swarm(K, L, ring is True, fresh)
    held = []
    ref = RefRings(K)
    sw.reset_windows()
    ref.reset(float(sw.lib.serial))
    for it in range(300):
        op = rnd.random()
        if op < 0.55:
            sw.step(None)
            ref.push_X(float(sw.lib.serial))
            ref.push_A(float(sw.lib.serial))
        elif op < 0.8:
            T = rnd.randint(1, 2 * L)
            sw.step_many(None, T)
            first = sw.lib.serial - T + 1
            for s in range(first, first + T):
                ref.push_X(float(s))
                ref.push_A(float(s))
        elif op < 0.92:
            sw.push_A()                                  # calc_Ak outside step: only the A ring shifts
            ref.push_A(float(sw.lib.serial) + 0.5)
        else:
            sw.reset_windows()
            ref.reset(float(sw.lib.serial))
        assert sw.X_window().flatten().tolist() == list(ref.X), (it, 'X')
        if fresh:
            if op > 0.97:
                sw.renew_tapes()
                assert sw.X_window().flatten().tolist() == list(ref.X)
            held.append((sw.X_window(), list(ref.X), sw.A_window(), None if sw.a_empty else list(ref.A)))
        got_A = sw.A_window().flatten().tolist()
        if len(ref.A) == 0:
            assert sw.a_empty                            # empty deque: no window yet (reference would raise)
        else:
            assert not sw.a_empty and got_A == list(ref.A), (it, 'A')
        assert 0 <= sw.hx < L and 0 <= sw.ha <= L
    for Xw, Xl, Aw, Al in held:
        assert Xw.flatten().tolist() == Xl
        assert Al is None or Aw.flatten().tolist() == Al
    assert not fresh or sw.generation > 300 // L


def test_capture_needs_room_and_action_checks():
    sw = make_swarm(2, 8)
    sw.cfg.action_type = _abi.ACTION_TYPES['set_speeds']
    with pytest.raises(ValueError):
        sw._check_actions(None)
    with pytest.raises(ValueError):
        sw._check_actions(torch.zeros(1, 1, 3))
    with pytest.raises(ValueError):
        sw._check_actions(torch.zeros(1, 1, 4, dtype=torch.float64))
    with pytest.raises(ValueError):
        sw._check_actions(torch.zeros(1, 1, 8)[..., ::2])
    sw._check_actions(torch.zeros(1, 1, 4))
    assert sw.max_chunk() == 6
    with pytest.raises(AttributeError):
        sw.set_action_type('set_nothing')
    assert sw.set_action_type('set_target_pos') == 3 and sw.set_action_type(None) == 0


@pytest.mark.parametrize('K,L', [(0, 1), (2, 3), (3, 4), (3, 10)])
def test_ring_mode_small_tapes(K, L):
    """ring mode needs only K_HOPS+1 slots: the window is the whole tape, rotated"""
    sw = make_swarm(K, L, ring=True)
    ref = RefRings(K)
    sw.reset_windows()
    ref.reset(float(sw.lib.serial))
    for it in range(5 * L + 3):
        if it % 4 == 3:
            sw.step_many(None, L)                        # a whole lap, like a graph replay
            for s in range(sw.lib.serial - L + 1, sw.lib.serial + 1):
                ref.push_X(float(s)); ref.push_A(float(s))
        else:
            sw.step(None)
            ref.push_X(float(sw.lib.serial)); ref.push_A(float(sw.lib.serial))
        assert sw.X_window().flatten().tolist() == list(ref.X)
        assert sw.A_window().flatten().tolist() == list(ref.A)
