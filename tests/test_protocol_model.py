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
