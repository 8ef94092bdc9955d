"""Work partitioning of the sharded paths: batch rows and Pwelch segment ranges are
split contiguously over ranks; no data-path collective, except one lp-double all-gather for Pwelch."""


def batch_rows(rank, world, batch):
    """[r0, r1) of the global batch owned by `rank` (contiguous, remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def segment_count(lx, size, noverlap):
    """spectral.Segment's count (spectral/spectral.go:27-33)."""
    stride = size - noverlap
    if lx == size:
        return 1
    return (lx - size) // stride + 1 if lx > size else 0


def pwelch_segment_range(rank, world, nsamples, nfft, noverlap):
    """(s0, s1, x0, x1): segments [s0, s1) of the global signal and the sample range [x0, x1) they read
    (a halo of `noverlap` samples past the rank's last own sample)."""
    nsegs = segment_count(nsamples, nfft, noverlap)
    s0, s1 = batch_rows(rank, world, nsegs)
    stride = nfft - noverlap
    if s1 <= s0:
        return s0, s0, 0, 0
    return s0, s1, s0 * stride, (s1 - 1) * stride + nfft


def reduce_partials(partials):
    """Sum per-rank partial PSD sums in rank order (fixed order keeps runs reproducible)."""
    tot = partials[0].clone() if hasattr(partials[0], "clone") else partials[0].copy()
    for p in partials[1:]:
        tot += p
    return tot
