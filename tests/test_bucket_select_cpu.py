"""CPU model of the bucket selection (csrc/corner_select.cu: bucket_select_kernel) against the reference's
sort + greedy loop (cpp/src/templering_sfm.cpp:286-300).

The kernel never sorts the whole candidate list: candidates are grouped into score buckets (leading bits of an order code),
the buckets are walked from the best score down in gathers of whole buckets, a candidate that is already blocked by an
accepted corner is dropped BEFORE its gather is sorted, and only the survivors of a gather are sorted and run through the
greedy rounds (at most ALIVE unblocked candidates per round, the rest re-tested by the next round); a bucket whose
unblocked candidates exceed a gather is walked in runs of its sub-buckets (next bits of the order code).  The model below does
exactly that with the kernel's control flow in plain Python; random instances with distinct scores must give the
reference's corners, in the reference's order, for every bucket width, gather capacity and round size."""
import numpy as np
import pytest


def reference(cands, d, cap):
    """std::sort by score descending (scores are distinct here) + the greedy loop; the cap is tested after the push."""
    out = []
    for s, x, y in sorted(cands, key=lambda c: -c[0]):
        if all((x - ax) ** 2 + (y - ay) ** 2 >= d * d for ax, ay in out):
            out.append((x, y))
            if len(out) >= cap:
                break
    return out


def bucket_select(cands, d, cap, code_bits, bucket_bits, gather_cap, alive_cap, first_chunk):
    """bucket_select_kernel: order code = (max - score) scaled to code_bits (ascending code = descending score), buckets =
    leading bucket_bits of the code; `blocked` plays the per-frame pixel bitmap."""
    smax = max(c[0] for c in cands)
    smin = min(c[0] for c in cands)
    span = max(smax - smin, 1)
    shift = max(span.bit_length() - code_bits, 0)
    used = min(span.bit_length(), code_bits)
    bshift = max(used - bucket_bits, 0)
    words = [((smax - s) >> shift, x, y, s) for s, x, y in cands]
    nb = 1 << bucket_bits
    buckets = [[] for _ in range(nb)]
    rng = np.random.default_rng(len(cands))
    for w in words:  # unordered inside a bucket
        buckets[w[0] >> bshift].append(w)
    for b in buckets:
        rng.shuffle(b)
    ends = np.cumsum([len(b) for b in buckets])
    flat = [w for b in buckets for w in b]
    blocked = set()
    out = []

    def mark(x, y):
        for dy in range(-(d - 1), d):
            for dx in range(-(d - 1), d):
                if dx * dx + dy * dy < d * d:
                    blocked.add((x + dx, y + dy))

    def flush(surv):
        # sort by the word; equal order codes by the exact score (recomputed from the image on the device)
        surv.sort(key=lambda w: (w[0], -w[3]))
        t0, want = 0, first_chunk
        while t0 < len(surv) and len(out) < cap:
            chunk = surv[t0:t0 + want]
            alive, used_n = [], len(chunk)
            for k, w in enumerate(chunk):
                if d > 0 and t0 > 0 and (w[1], w[2]) in blocked:  # (the first round follows the filter directly)
                    continue
                if len(alive) == alive_cap:
                    used_n = k  # first candidate that does not fit: it starts the next round
                    break
                alive.append(w)
            # conflicts inside the round, in priority order (what the bit-row rounds compute in parallel)
            acc_round = []
            for w in alive:
                if all((w[1] - ax) ** 2 + (w[2] - ay) ** 2 >= d * d for ax, ay in acc_round):
                    acc_round.append((w[1], w[2]))
            for (x, y) in acc_round:
                if len(out) < cap:
                    out.append((x, y))
            if len(out) < cap and d > 0:
                for (x, y) in acc_round:
                    mark(x, y)
            t0 += used_n
            want = 2 * used_n if used_n < len(chunk) else 2 * want
            want = min(max(want, first_chunk), 8 * first_chunk)

    def unblocked(ws):
        return [w for w in ws if (w[1], w[2]) not in blocked] if d > 0 else list(ws)

    sbits = min(bshift, 3)  # the kernel takes 8 sub-bits; 3 make the runs visible on these small instances
    sshift = bshift - sbits
    pos, bi, n = 0, 0, len(flat)
    while pos < n and len(out) < cap:
        tgt = pos
        while bi < nb and ends[bi] - pos <= gather_cap:
            tgt = int(ends[bi])
            bi += 1
        if tgt == pos:  # a bucket that does not fit by itself is filtered anyway
            tgt = int(ends[bi])
            bi += 1
        surv = unblocked(flat[pos:tgt])
        if len(surv) > gather_cap:
            # only a single bucket can overflow: walk it in runs of sub-buckets whose (earlier counted) unblocked words fit
            if sbits == 0:
                return None  # the exact emulation takes the frame
            subh = [0] * (1 << sbits)
            for w in surv:
                subh[(w[0] >> sshift) & ((1 << sbits) - 1)] += 1
            lo = 0
            while lo < (1 << sbits) and len(out) < cap:
                if subh[lo] > gather_cap:
                    return None
                hi, tot = lo, subh[lo]
                while hi + 1 < (1 << sbits) and tot + subh[hi + 1] <= gather_cap:
                    hi += 1
                    tot += subh[hi]
                run = [w for w in unblocked(flat[pos:tgt]) if lo <= ((w[0] >> sshift) & ((1 << sbits) - 1)) <= hi]
                assert len(run) <= gather_cap
                flush(run)
                lo = hi + 1
        else:
            flush(surv)
        pos = tgt
    return out


@pytest.mark.parametrize("seed", range(60))
def test_bucket_selection_equals_sort_plus_greedy(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    side = int(rng.integers(8, 48))
    d = int(rng.choice([0, 1, 2, 3, 5, 8]))
    cap = int(rng.choice([1, 3, 10, 40, 1000]))
    # distinct integer scores with a heavy tail towards the threshold (as corner scores have), distinct pixels
    scores = rng.permutation(np.unique((rng.pareto(1.5, 4 * n) * 1000).astype(np.int64) + 1))[:n]
    pix = rng.permutation(side * side)[:len(scores)]
    cands = [(int(s), int(p % side), int(p // side)) for s, p in zip(scores, pix)]
    if not cands:
        return
    want = reference(cands, d, cap)
    for code_bits, bucket_bits, gather_cap, alive_cap, first_chunk in [(34, 11, 4096, 256, 512), (12, 4, 16, 4, 8), (6, 3, 5, 2, 4),
                                                                    (20, 1, 7, 3, 2), (3, 2, 2, 1, 1)]:
        got = bucket_select(cands, d, cap, code_bits, bucket_bits, gather_cap, alive_cap, first_chunk)
        if got is None:  # a sub-bucket alone exceeds the gather: the device hands the frame to the exact emulation
            continue
        assert got == want, (seed, code_bits, bucket_bits, gather_cap, alive_cap, first_chunk)
