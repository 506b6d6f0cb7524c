"""CPU check of the bookkeeping behind the fused tail (csrc/tail.cuh, partials_impl in csrc/gsa_api.cu), restated in Python (no GPU,
no oracle): for a shard that launches the tiles [t_lo, t_hi) of nrep replicates of tpr tiles each,
  * every group's expected arrival count equals the number of its launched tiles, every replicate's the number of its groups with
    launched tiles, and the launch completes exactly once;
  * the tiles a group sum reads are launched tiles or lie inside the ranges the host zeroes, [z_lo, t_lo) and [t_hi, z_hi);
  * the replicate sum reads exactly the groups that were summed;
  * the number of groups with launched tiles per replicate never exceeds the bound the host uses to decide whether to fuse."""
import random

TAIL_G = 32


def groups_per_replicate(tpr):
    return (tpr + TAIL_G - 1) // TAIL_G


def simulate(nrep, tpr, t_lo, t_hi, order):
    ngr = groups_per_replicate(tpr)
    gcount = [0] * (nrep * ngr)
    rcount = [0] * nrep
    top = 0
    summed_groups, summed_reps, finished = set(), set(), 0
    read_tiles = set()
    for tile in order:
        rl = tile // tpr
        tr0 = rl * tpr
        g = (tile - tr0) // TAIL_G
        g0, g1 = tr0 + g * TAIL_G, min(tr0 + g * TAIL_G + TAIL_G, tr0 + tpr)
        in_group = min(g1, t_hi) - max(g0, t_lo)
        gcount[rl * ngr + g] += 1
        assert gcount[rl * ngr + g] <= in_group
        if gcount[rl * ngr + g] != in_group:
            continue
        gcount[rl * ngr + g] = 0                                   # self-cleaning
        lo, hi = max(tr0, t_lo), min(tr0 + tpr, t_hi)
        gf, gl = (lo - tr0) // TAIL_G, (hi - 1 - tr0) // TAIL_G
        assert gf <= g <= gl
        read_tiles.update(range(g0, g1))
        summed_groups.add((rl, g))
        if gl != gf:
            rcount[rl] += 1
            if rcount[rl] != gl - gf + 1:
                continue
            rcount[rl] = 0
            assert all((rl, q) in summed_groups for q in range(gf, gl + 1))     # the replicate sum reads only finished groups
        assert rl not in summed_reps
        summed_reps.add(rl)
        top += 1
        if top == nrep:
            top = 0
            finished += 1
    assert finished == 1 and summed_reps == set(range(nrep))
    assert all(c == 0 for c in gcount) and all(c == 0 for c in rcount)       # counters are clean for the next launch
    return read_tiles, summed_groups


def test_group_and_replicate_counts_for_random_shards():
    rnd = random.Random(7)
    for _ in range(400):
        nrep = rnd.choice([1, 1, 2, 3, 4])
        tpr = rnd.choice([1, 5, 31, 32, 33, 100, 888, 1000, 7104])
        ntiles = nrep * tpr
        # a shard = a contiguous tile range touching EVERY one of its nrep replicates (rep_first .. rep_last are defined that way)
        t_lo = rnd.randrange(0, tpr)
        t_hi = rnd.randrange((nrep - 1) * tpr + 1, ntiles + 1)
        if t_hi <= t_lo:
            continue
        order = list(range(t_lo, t_hi))
        rnd.shuffle(order)                                             # CTAs finish in any order
        read, groups = simulate(nrep, tpr, t_lo, t_hi, order)
        # the host zeroes the members of the boundary groups that are not launched
        r0, r1 = t_lo // tpr * tpr, (t_hi - 1) // tpr * tpr
        z_lo = r0 + (t_lo - r0) // TAIL_G * TAIL_G
        z_hi = min(r1 + tpr, r1 + ((t_hi - 1 - r1) // TAIL_G + 1) * TAIL_G)
        assert z_lo <= t_lo and t_hi <= z_hi
        assert read <= set(range(z_lo, z_hi)), (nrep, tpr, t_lo, t_hi)
        assert set(range(t_lo, t_hi)) <= read
        # the host's bound on the groups with launched tiles in one replicate
        bound = min(groups_per_replicate(tpr), (t_hi - t_lo + TAIL_G - 1) // TAIL_G + 1)
        for rl in range(nrep):
            assert sum(1 for (r, _) in groups if r == rl) <= bound


def test_eight_rank_shards_of_config5_fuse():
    """config 5 on 8 GPUs: the replicate has ~7100 tiles, a rank launches ~888 of them: 28-29 groups <= TAIL_G, so the step stays one launch"""
    tpr = 7104
    for rank in range(8):
        t_lo, t_hi = rank * 888, (rank + 1) * 888
        bound = min(groups_per_replicate(tpr), (t_hi - t_lo + TAIL_G - 1) // TAIL_G + 1)
        assert bound <= TAIL_G
        simulate(1, tpr, t_lo, t_hi, list(range(t_lo, t_hi)))
