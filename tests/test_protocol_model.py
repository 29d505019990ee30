"""Exhaustive interleaving check of the cross-GPU combine protocol.

The fused kernels publish launch e's partial into slot set ``e & 1`` of every peer as self-validating words tagged with the
epoch and read their own slots once every source's tag equals e.  Nothing is ever reset.  This file model-checks that
schedule: every rank is a small state machine (write my slot on each peer; per source: poll the tag, then read the
payload), and a DFS explores ALL interleavings of the ranks' atomic steps for a few launches.  Properties:

* safety   -- a reader never observes a payload of a different launch than the tag it validated (no slot is overwritten
              while a reader of an older launch may still read it);
* liveness -- no reachable state deadlocks;
* control  -- the same model with ONE slot set instead of two violates safety (the checker can see the bug).

Word-level atomicity matches the kernel: one 8-byte store carries {value, tag}; the model's payload is the launch number.
"""
import sys
from functools import lru_cache

import pytest


def _explore(world: int, launches: int, slot_sets: int):
    """Returns (violation_found, deadlock_found, states_visited).  State = (per-rank program counter tuple, memory)."""
    # memory[dst][set][src] = tag (= payload launch number), 0 = never written
    # a rank's program for launch e (1-based): for dst in ranks: W(dst);  for src in ranks: P(src) until tag == e, then R(src)
    # program counter: (e, phase, idx): phase 0 = writing to dst idx, phase 1 = polling src idx, phase 2 = reading src idx
    sys.setrecursionlimit(100000)
    init_mem = tuple(tuple(tuple(0 for _ in range(world)) for _ in range(slot_sets)) for _ in range(world))
    init_pc = tuple((1, 0, 0) for _ in range(world))
    seen = set()
    violation = [False]
    deadlock = [False]
    stack = [(init_pc, init_mem)]
    while stack:
        pcs, mem = stack.pop()
        if (pcs, mem) in seen:
            continue
        seen.add((pcs, mem))
        progressed = False
        all_done = True
        for r in range(world):
            e, phase, idx = pcs[r]
            if e > launches:
                continue
            all_done = False
            s = e % slot_sets
            if phase == 0:      # store {payload = e, tag = e} into slot [s][r] of rank idx
                m = [list(map(list, x)) for x in mem]
                m[idx][s][r] = e
                nmem = tuple(tuple(tuple(y) for y in x) for x in m)
                npc = (e, 0, idx + 1) if idx + 1 < world else (e, 1, 0)
                stack.append((pcs[:r] + (npc,) + pcs[r + 1:], nmem))
                progressed = True
            elif phase == 1:    # poll: proceed only when the tag of source idx equals my launch
                if mem[r][s][idx] == e:
                    stack.append((pcs[:r] + ((e, 2, idx),) + pcs[r + 1:], mem))
                    progressed = True
                elif mem[r][s][idx] > e and slot_sets == 1:
                    # (single-buffer control) the tag ran past my launch: the payload I was waiting for is gone
                    violation[0] = True
            else:               # read the payload that was validated in phase 1: it must still belong to launch e
                if mem[r][s][idx] != e:
                    violation[0] = True
                npc = (e, 1, idx + 1) if idx + 1 < world else (e + 1, 0, 0)
                stack.append((pcs[:r] + (npc,) + pcs[r + 1:], mem))
                progressed = True
        if not all_done and not progressed:
            deadlock[0] = True
    return violation[0], deadlock[0], len(seen)


@pytest.mark.parametrize("world,launches", [(2, 6), (3, 3)])
def test_parity_double_buffering_is_safe_and_live(world, launches):
    violation, deadlock, n = _explore(world, launches, slot_sets=2)
    assert n > 200             # the search really explored the interleavings
    assert not violation, "a reader saw a payload of another launch"
    assert not deadlock


def test_single_slot_set_is_caught_by_the_checker():
    violation, _, _ = _explore(2, 3, slot_sets=1)
    assert violation, "negative control: one slot set must be unsafe"


# ----------------------------------------------------------------------------------------------------------------------
# launch tags from arrival counters (csrc/decode_comm.cuh::launch_tag): every CTA adds 1 to a 64-bit counter when it starts,
# CTA 0 adds 4096 - (grid - 1); tag = old / 4096 + 1.  Claim: for ANY arrival order inside a launch (launches themselves are
# ordered: a launch's arrivals happen after the previous launch has completed -- stream order or griddepcontrol.wait) every
# CTA of launch n computes the same tag n + 1, also when the grid size changes from launch to launch.
# ----------------------------------------------------------------------------------------------------------------------
def _tags_of_launch(counter: int, grid: int, order):
    tags = {}
    for cta in order:
        inc = (4096 - (grid - 1)) if cta == 0 else 1
        tags[cta] = counter // 4096 + 1
        counter += inc
    return counter, tags


def test_arrival_counter_tags_are_uniform_per_launch_for_any_arrival_order():
    import random

    rng = random.Random(0)
    counter = 0
    for launch in range(200):
        grid = rng.choice([1, 2, 7, 32, 147, 148, 4096])
        order = list(range(grid))
        rng.shuffle(order)
        if launch % 3 == 0:                       # CTA 0 first / last are the extreme cases
            order.remove(0)
            order = ([0] + order) if launch % 2 else (order + [0])
        counter, tags = _tags_of_launch(counter, grid, order)
        assert set(tags.values()) == {launch + 1}, (launch, grid)
        assert counter == 4096 * (launch + 1)      # every launch advances the counter by exactly 4096
    # the tag's parity (slot set of the cross-GPU words) alternates from launch to launch by construction


# ----------------------------------------------------------------------------------------------------------------------
# designated merger of the split merge: the CTA that owns a head's LAST tile merges the head after its own last tile, and only
# ever waits for CTAs with a LOWER block index.  Model: CTAs become resident in index order (a CTA with index i may start only
# when all j < i have started -- the hardware's dispatch order; here at most `resident` CTAs run at a time), every CTA first
# publishes all its partials (never blocks) and then waits for the parts of the heads it merges.  Liveness: the schedule
# always terminates, for any geometry, even with ONE resident CTA at a time.
# ----------------------------------------------------------------------------------------------------------------------
def _split(total, ncta):
    q, r = divmod(total, ncta)
    lo = [c * q + min(c, r) for c in range(ncta + 1)]
    return lo


@pytest.mark.parametrize("resident", [1, 2, 148])
def test_designated_merger_never_waits_for_a_later_cta(resident):
    import random

    rng = random.Random(1)
    for _ in range(300):
        heads = rng.randint(1, 40)
        tph = rng.randint(1, 70)
        total = heads * tph
        ncta = min(total, rng.choice([1, 3, 16, 148]))
        lo = _split(total, ncta)
        owner_of_tile = lambda t: max(c for c in range(ncta) if lo[c] <= t)
        published = set()                 # (cta, head) partials visible
        done = [False] * ncta
        started = 0
        running = []
        steps = 0
        while not all(done):
            steps += 1
            assert steps < 10 * ncta + 10, "split merge does not terminate"
            while started < ncta and len(running) < resident:      # dispatch in index order
                running.append(started)
                started += 1
            progressed = False
            for c in list(running):
                my_heads = range(lo[c] // tph, (lo[c + 1] - 1) // tph + 1) if lo[c + 1] > lo[c] else []
                for x in my_heads:                                   # streaming + tagged partial stores never block
                    published.add((c, x))
                merges = [x for x in my_heads if owner_of_tile((x + 1) * tph - 1) == c]
                need = {(owner_of_tile(t), x) for x in merges for t in range(x * tph, (x + 1) * tph)}
                assert all(src <= c for src, _ in need), "a merger would wait for a CTA dispatched after it"
                if need <= published:
                    done[c] = True
                    running.remove(c)
                    progressed = True
            assert progressed, "deadlock: a resident merger waits for a CTA that cannot become resident"
