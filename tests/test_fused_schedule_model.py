"""Host-side model of the phase sequence of the fused 2^20 kernel (csrc/fft_tma.cuh, tma_decode): the no-deadlock
argument in DESIGN.md rests on three properties of that sequence, checked here for many (batch, delay) pairs:
every tile of both passes appears exactly once, P2(g) comes after the whole of P1(g), and P1(h) comes after the whole
of P2(h - S) with S = delay + 2 scratch slots (so every dependency points at an earlier item)."""
import pytest

TPT = 256          # tiles per transform and pass (1024 lines / 4 lines per tile)


def decode(gi, B, D):
    """Python restatement of tma_decode(): global item index -> (pass, transform, tile)."""
    f, c = divmod(gi, TPT)
    if B <= D + 1:
        return (0, f, c) if f < B else (1, f - B, c)
    if f <= D:
        return (0, f, c)
    m, npairs = f - D - 1, B - D - 1
    if m < 2 * npairs:
        return (0, D + 1 + (m >> 1), c) if m & 1 else (1, m >> 1, c)
    return (1, npairs + (m - 2 * npairs), c)


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 7, 37, 128])
@pytest.mark.parametrize("D", [0, 1, 2, 3])
def test_phase_sequence_invariants(B, D):
    S = D + 2
    first, last = {}, {}
    seen = set()
    for f in range(2 * B):
        t, tf, c = decode(f * TPT, B, D)
        assert 0 <= tf < B
        for cc in (0, TPT - 1):                         # a phase is homogeneous
            assert decode(f * TPT + cc, B, D)[:2] == (t, tf)
        assert (t, tf) not in seen
        seen.add((t, tf))
        first[(t, tf)] = last[(t, tf)] = f
    assert len(seen) == 2 * B                           # every pass of every transform exactly once
    for g in range(B):
        assert first[(1, g)] > last[(0, g)]             # P2(g) after all of P1(g)
        if g >= S:
            assert first[(0, g)] > last[(1, g - S)]     # slot g mod S is free again


def test_slots_in_flight_never_exceed_S():
    B, D = 64, 1
    S = D + 2
    live = set()
    for f in range(2 * B):
        t, tf, _ = decode(f * TPT, B, D)
        if t == 0:
            live.add(tf)
            assert len({x % S for x in live}) == len(live) <= S      # no two live transforms share a slot
        else:
            live.discard(tf)
