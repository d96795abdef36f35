"""CPU: a line-by-line Python model of the two parallel look-back loops of the opt-in `lookback_parallel` kernels
(csrc/radix_sort.cu: 8 predecessor states per step and digit; csrc/rast_forward.cu: 32 per step by warp 0), run against
randomly delayed publications of AGGREGATE / INCLUSIVE states: the prefix must always be the sum back to and including the
nearest inclusive state, whatever is still unpublished when a batch is read.  (The kernels themselves are checked on the GPU
by tools/native/sort_check and rast_check; this pins the index arithmetic.)"""
import random

AGG, INCL = 1, 2


def batched_lookback(read, tile, batch=8):
    prefix, t, done, rounds = 0, tile - 1, False, 0
    while not done:
        v = [read(t - k) if t - k >= 0 else (INCL, 0) for k in range(batch)]
        consumed = 0
        for k in range(batch):                      # strictly in order; the first unpublished state ends the batch
            if not done and consumed == k and v[k][0] != 0:
                prefix += v[k][1]
                done = v[k][0] == INCL
                consumed = k + 1
        t -= consumed
        rounds += 1
        assert rounds < 100000
    return prefix


def warp_lookback(read, tile):
    prefix, t, done, rounds = 0, tile - 1, tile == 0, 0
    while not done:
        v = [read(t - lane) if t - lane >= 0 else (INCL, 0) for lane in range(32)]
        m_ready = sum(1 << l for l in range(32) if v[l][0] != 0)
        m_incl = sum(1 << l for l in range(32) if v[l][0] == INCL)
        inv = ~m_ready & 0xFFFFFFFF
        n_ready = 32 if m_ready == 0xFFFFFFFF else (inv & -inv).bit_length() - 1          # __ffs(~m_ready) - 1
        incl_in = m_incl & (0xFFFFFFFF if n_ready == 32 else (1 << n_ready) - 1)
        take = (incl_in & -incl_in).bit_length() if incl_in else n_ready                   # __ffs(incl_in)
        prefix += sum(v[l][1] for l in range(32) if l < take)
        if incl_in:
            done = True
        else:
            t -= take
        rounds += 1
        assert rounds < 100000
    return prefix


def test_parallel_lookbacks_sum_back_to_the_nearest_inclusive_state():
    rng = random.Random(1)
    for _ in range(1500):
        n = rng.randint(1, 300)
        counts = [rng.randint(0, 50) for _ in range(n)]
        inclusive = [sum(counts[:i + 1]) for i in range(n)]
        final = [(INCL, inclusive[i]) if i == 0 or rng.random() < 0.15 else (AGG, counts[i]) for i in range(n)]
        for fn in (batched_lookback, warp_lookback):
            delay = {i: (rng.randint(1, 3) if rng.random() < 0.3 else 0) for i in range(n)}

            def read(t):
                if delay[t] > 0:                    # not published yet: the kernel sees neither flag
                    delay[t] -= 1
                    return (0, 0)
                return final[t]
            assert fn(read, n) == inclusive[n - 1]
