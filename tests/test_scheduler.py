"""CPU: the continuous-batching scheduler's host logic against a scripted engine (no GPU, no library)."""
import pytest

from meshanything_b200.scheduler import SlotScheduler


class ScriptedEngine:
    """Each payload is the number of tokens its sequence will produce (<= max_new).  Mirrors the device contract:
    prefill picks token 0, every step adds one token to every unfinished slot, finished slots are frozen."""

    def __init__(self, slots, max_new):
        self.n, self.max_new = slots, max_new
        self.target = [0] * slots
        self.gen = [0] * slots
        self.fin = [1] * slots
        self.tag = [None] * slots
        self.log = []

    def prefill(self, slot, payload):
        assert self.fin[slot] == 1, "refill of a live slot"
        tag, length = payload
        assert 1 <= length <= self.max_new
        self.tag[slot], self.target[slot], self.gen[slot] = tag, length, 1
        self.fin[slot] = int(length == 1 or self.max_new == 1)
        self.log.append(("prefill", slot, tag))

    def step(self, n, max_ctx):
        for _ in range(n):
            live_ctx = [257 + self.gen[s] for s in range(self.n) if not self.fin[s]]
            if live_ctx:
                assert max_ctx >= max(live_ctx), "attention grid sized below a live slot's context"
            for s in range(self.n):
                if not self.fin[s]:
                    self.gen[s] += 1
                    if self.gen[s] >= self.target[s] or self.gen[s] >= self.max_new:
                        self.fin[s] = 1
            max_ctx += 1
        self.log.append(("step", n))

    def poll(self):
        return list(self.fin), list(self.gen)

    def fetch(self, slot, n):
        return (self.tag[slot], n)


@pytest.mark.parametrize("slots,poll_every", [(1, 4), (2, 3), (3, 32), (8, 1)])
def test_every_item_completes_with_its_own_length(slots, poll_every):
    max_new = 20
    lengths = [5, 20, 1, 7, 7, 13, 2, 20, 9, 3, 11]
    eng = ScriptedEngine(slots, max_new)
    sched = SlotScheduler(eng, slots, max_new, poll_every=poll_every)
    got = dict(sched.run([(f"item{i}", n) for i, n in enumerate(lengths)]))
    assert sorted(got) == list(range(len(lengths)))
    for i, n in enumerate(lengths):
        assert got[i] == (f"item{i}", n)
    assert sched.stats.prefills == len(lengths)
    # never more live sequences than slots, and every slot is refilled only after it finished (asserted in prefill)
    assert all(e[1] < slots for e in eng.log if e[0] == "prefill")


def test_short_sequences_do_not_wait_for_long_ones():
    """With 2 slots, one long sequence and many short ones: the short ones stream through the second slot while the
    long one runs -- far fewer batched steps than padded batches of 2 would need."""
    max_new = 64
    lengths = [64] + [4] * 10
    eng = ScriptedEngine(2, max_new)
    sched = SlotScheduler(eng, 2, max_new, poll_every=2)
    got = dict(sched.run([(i, n) for i, n in enumerate(lengths)]))
    assert len(got) == len(lengths)
    padded_steps = 63 + 5 * 3                # HF padding, batches of 2: (64,4) then five batches of short ones
    assert sched.stats.steps <= 70 < padded_steps
    assert sched.stats.finished_order[0] != 0   # a short item completes before the long one
    assert sched.stats.finished_order[-1] in (0, 10)


def test_empty_queue_and_single_token_cap():
    eng = ScriptedEngine(2, 1)
    sched = SlotScheduler(eng, 2, 1)
    assert list(sched.run([])) == []
    got = dict(sched.run([("a", 1), ("b", 1), ("c", 1)]))
    assert got == {0: ("a", 1), 1: ("b", 1), 2: ("c", 1)}
    assert sched.stats.steps == 0


def test_stuck_slot_is_reported():
    class Stuck(ScriptedEngine):
        def step(self, n, max_ctx):
            pass
    eng = Stuck(1, 4)
    with pytest.raises(RuntimeError):
        list(SlotScheduler(eng, 1, 4, poll_every=8).run([("x", 4)]))
