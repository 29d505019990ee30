"""The decode kernels' stream-K split (``decode_split`` in csrc/decode_simt.cu, shared by all three decode kernels; the
device side is ``dcomm::make_geom`` in csrc/decode_comm.cuh):
every 128-key tile is owned by exactly one CTA, CTA loads differ by at most one tile, and ``max_parts`` -- which sizes the
split-merge workspace -- really bounds the number of CTAs that touch one KV head (a violation would be a buffer overflow
in the kernel).  Pure host code: runs without a GPU."""
import pytest

from tree_attention_b200 import _build

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402


def _ranges(total, grid):
    q, rem = divmod(total, grid)
    lo = [c * q + min(c, rem) for c in range(grid + 1)]
    return lo


@settings(max_examples=300, deadline=None)
@given(b=st.integers(1, 4), hkv=st.integers(1, 64), s=st.integers(1, 300_000), ncta=st.sampled_from([1, 8, 132, 148, 296]))
def test_split_covers_every_tile_once_and_max_parts_bounds_sharing(b, hkv, s, ncta):
    C = _build.load()
    grid, max_parts = C.decode_split_for(b, hkv, s, ncta)
    tph = (s + 127) // 128
    total = b * hkv * tph
    assert 1 <= grid <= min(ncta, total)
    lo = _ranges(total, grid)
    assert lo[0] == 0 and lo[-1] == total
    sizes = [lo[c + 1] - lo[c] for c in range(grid)]
    assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
    # CTAs touching head x = those whose [lo, hi) intersects [x * tph, (x + 1) * tph)
    worst = 0
    for x in range(b * hkv):
        first = next(c for c in range(grid) if lo[c + 1] > x * tph)
        last = next(c for c in range(grid - 1, -1, -1) if lo[c] < (x + 1) * tph)
        worst = max(worst, last - first + 1)
    assert worst <= max_parts, (worst, max_parts, b, hkv, s, ncta)


@settings(max_examples=200, deadline=None)
@given(b=st.integers(1, 3), hkv=st.integers(1, 40), cap=st.integers(1, 40_000), frac=st.floats(0.0, 1.0),
       ncta=st.sampled_from([4, 132, 148]))
def test_max_parts_bounds_sharing_for_every_fill_level(b, hkv, cap, frac, ncta):
    """The kernels re-derive the split on the device from the run-time fill level ``kv_len <= capacity`` but keep the grid
    and the workspace sized for the capacity: ``max_parts`` must bound the CTAs per head for EVERY fill level (an empty
    shard still runs one masked tile per head)."""
    C = _build.load()
    grid, max_parts = C.decode_split_for(b, hkv, cap, ncta)
    s_eff = int(cap * frac)
    tph = max(1, (s_eff + 127) // 128)
    total = b * hkv * tph
    lo = _ranges(total, grid)                    # dcomm::cta_lo with tiles_q = total // grid (possibly 0)
    assert lo[-1] == total
    for x in range(b * hkv):
        owners = [c for c in range(grid) if lo[c + 1] > x * tph and lo[c] < (x + 1) * tph]
        assert owners and owners == list(range(owners[0], owners[-1] + 1))
        assert len(owners) <= max_parts, (len(owners), max_parts, b, hkv, cap, s_eff, ncta)


# ----------------------------------------------------------------------------------------------------------------------
# Executable model of the speed-weighted split that docs/NEXT.md section 1 proposes for the decode kernels (NOT in the kernels
# yet): each CTA claims a contiguous tile range whose length is proportional to the measured streaming rate of the SM it
# landed on, with ONE atomicAdd on a packed (claim count | tiles) counter.  The properties below are what the kernel-side
# merge logic would rely on; they hold for any block -> SM mapping and any claim order.
# ----------------------------------------------------------------------------------------------------------------------
def _claim_split(total, weights_by_sm, sm_of_claimer):
    wsum = sum(weights_by_sm)
    pieces, cursor = [], 0
    for k, sm in enumerate(sm_of_claimer):                  # k = claim order = value of the packed counter's count field
        ln = -(-total * weights_by_sm[sm] // wsum)           # ceil(T * w / W): the lengths sum to >= T
        pieces.append((min(cursor, total), min(cursor + ln, total)))
        cursor += ln
    return pieces


def test_weighted_claim_split_model():
    import random

    rng = random.Random(5)
    for _ in range(400):
        nsm = rng.choice([4, 16, 148])
        heads, tph = rng.randint(1, 64), rng.randint(1, 300)
        total = heads * tph
        weights = [rng.choice([100, 100, 100, 92, 87]) for _ in range(nsm)]   # 8 % and 13 % slower SMs, as measured
        sm_of_claimer = list(range(nsm))
        rng.shuffle(sm_of_claimer)                            # one CTA per SM, arbitrary claim order (PDL makes it vary)
        pieces = _claim_split(total, weights, sm_of_claimer)
        # 1. disjoint, ordered by claim index, covering every tile exactly once
        assert pieces[0][0] == 0 and pieces[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        # 2. balanced in TIME: every non-clipped piece costs the same to within one tile of rounding
        cost = [(hi - lo) / weights[sm] for (lo, hi), sm in zip(pieces, sm_of_claimer) if hi < total and hi > lo]
        if cost:
            assert max(cost) - min(cost) <= 1.0 / min(weights) + 1e-9
        # 3. per head: the pieces that overlap it are a contiguous run of claim indices; the merger (piece holding the head's
        #    last tile) has the HIGHEST claim index of the run, i.e. it only waits for earlier claimers
        min_len = min(hi - lo for lo, hi in pieces if hi > lo)
        for x in range(heads):
            lo_t, hi_t = x * tph, (x + 1) * tph
            run = [k for k, (lo, hi) in enumerate(pieces) if lo < hi_t and hi > lo_t]
            assert run == list(range(run[0], run[-1] + 1))
            merger = next(k for k, (lo, hi) in enumerate(pieces) if lo <= hi_t - 1 < hi)
            assert merger == run[-1]
            # 4. slot bound the workspace would be sized with (every piece but the last of the run is a whole claim)
            assert len(run) <= -(-tph // min_len) + 1 or pieces[run[-1]][1] == total
