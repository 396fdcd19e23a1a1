"""The two parallel forms of the reference's pairwise_sum (src/utils.c:29-45) in simplemoc_b200/csrc/moc_kernels.cuh
-- the lane-spread sums of update_sources_coop_kernel (tree_slot, tree_combine, tree_depth) and the 256-thread
pairwise_sum_cta of the K2-K4 reductions -- perform the recursion's additions in the recursion's order: checked
SYMBOLICALLY for every group count the kernel accepts (1 <= G <= 512) and for region counts 1..1200 plus the
BASELINE configurations', not only for the sizes the GPU parity tests run
(tests/test_gpu_parity.py::test_reductions_bit_exact_on_identical_flux).

This file restates the index arithmetic of those device functions line by line in Python and compares
expression trees; it does not run the kernel (the GPU tests do) and imports nothing from oracle/."""
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS = os.path.join(os.path.dirname(HERE), "simplemoc_b200", "csrc", "moc_kernels.cuh")


def reference_tree(lo, n):
    """utils.c:29-45 as an expression tree: a leaf of <= 16 terms is summed left to right from 0."""
    if n <= 16:
        return ("seq", lo, n)
    half = n // 2
    return ("add", reference_tree(lo, half), reference_tree(lo + half, n - half))


def tree_depth(n):
    d = 0
    while n > 16:
        n -= n // 2
        d += 1
    return d


def tree_slot(n, depth, sub):
    lo, sz, split_mask, level = 0, n, 0, 0
    while level < depth:
        if sz <= 16:
            break
        split_mask |= 1 << level
        half = sz // 2
        if (sub >> (depth - 1 - level)) & 1:
            lo += half
            sz -= half
        else:
            sz = half
        level += 1
    owner = (sub & ((1 << (depth - level)) - 1)) == 0
    return lo, sz, owner, split_mask


def lane_spread_tree(n):
    """What lane 0 of a 2^depth-lane group holds after tree_combine, as an expression tree."""
    depth = tree_depth(n)
    lanes = 1 << depth
    slots = [tree_slot(n, depth, sub) for sub in range(lanes)]
    # a lane that owns no leaf carries the 0.f it was initialised with: it must never be added
    v = [("seq", lo, sz) if owner else None for lo, sz, owner, _ in slots]
    for level in range(depth - 1, -1, -1):
        span = 1 << (depth - level)
        right = [v[sub + span // 2] if (sub % span) + span // 2 < span and sub + span // 2 < lanes else None
                 for sub in range(lanes)]            # __shfl_down_sync(v, span / 2, width = 2^depth)
        for sub in range(lanes):
            if sub & (span - 1) == 0 and (slots[sub][3] >> level) & 1:
                assert v[sub] is not None and right[sub] is not None, (n, level, sub)
                v[sub] = ("add", v[sub], right[sub])
    return v[0]


@pytest.mark.parametrize("lo,hi", [(1, 129), (129, 257), (257, 385), (385, 513)])
def test_lane_spread_tree_is_the_reference_recursion(lo, hi):
    for n in range(lo, hi):
        assert tree_depth(n) <= 5                      # the kernel's precondition (G <= 512)
        assert lane_spread_tree(n) == reference_tree(0, n), n


def test_every_term_is_summed_exactly_once():
    for n in (1, 16, 17, 33, 104, 130, 200, 511, 512):
        depth = tree_depth(n)
        covered = []
        for sub in range(1 << depth):
            lo, sz, owner, _ = tree_slot(n, depth, sub)
            assert sz <= 16
            if owner:
                covered += list(range(lo, lo + sz))
        assert sorted(covered) == list(range(n)), n


def cta_tree(n):
    """pairwise_sum_cta (256 threads, the top 8 levels of the recursion spread over threads): what slots[0]
    holds at the end.  A thread's own subtree is summed by the device recursion `pairwise_sum`, which is the
    reference's (same split, same 16-term base case)."""
    D = 8
    slots = [None] * 256
    for t in range(256):
        lo, sz, depth = 0, n, 0
        while depth < D:
            if sz <= 16:
                break
            half = sz // 2
            if (t >> (D - 1 - depth)) & 1:
                lo += half
                sz -= half
            else:
                sz = half
            depth += 1
        if t & ((1 << (D - depth)) - 1) == 0:
            slots[t] = reference_tree(lo, sz)
    for level in range(D - 1, -1, -1):
        span = 1 << (D - level)
        new = list(slots)
        for t in range(0, 256, span):
            s2, split_all_the_way = n, True
            for d in range(level):
                if s2 <= 16:
                    split_all_the_way = False
                    break
                half = s2 // 2
                s2 = s2 - half if (t >> (D - 1 - d)) & 1 else half
            if split_all_the_way and s2 > 16:
                assert slots[t] is not None and slots[t + span // 2] is not None, (n, level, t)
                new[t] = ("add", slots[t], slots[t + span // 2])
        slots = new
    return slots[0]


def test_cta_tree_is_the_reference_recursion():
    """Region counts of every configuration in BASELINE.json (2 250, 6 750, 15 000, 67 500), every n up to 1 200,
    and sizes around the powers of two where subtrees become leaves."""
    sizes = set(range(1, 1201)) | {2250, 6750, 15000, 67500, 100003}
    for k in range(4, 17):
        sizes |= {(1 << k) - 1, 1 << k, (1 << k) + 1, 17 << (k - 4), (17 << (k - 4)) - 1}
    for n in sorted(sizes):
        assert cta_tree(n) == reference_tree(0, n), n


def test_the_model_follows_the_device_code():
    """Guard against the kernel and this restatement drifting apart: the lines the model mirrors are still there."""
    src = open(KERNELS).read()
    for needle in (r"if \(t\.sz <= 16\) break;",
                   r"t\.split_mask \|= 1u << level;",
                   r"if \(\(sub >> \(depth - 1 - level\)\) & 1\) \{ t\.lo \+= half; t\.sz -= half; \}",
                   r"t\.owner = \(sub & \(\(1 << \(depth - level\)\) - 1\)\) == 0;",
                   r"__shfl_down_sync\(0xffffffffu, v, span / 2, 1 << depth\)",
                   r"if \(\(sub & \(span - 1\)\) == 0 && \(\(t\.split_mask >> level\) & 1u\)\) v = __fadd_rn\(v, right\);",
                   r"while \(n > 16\) \{ n -= n / 2; d\+\+; \}",
                   r"if \(\(t >> \(D - 1 - depth\)\) & 1\) \{ lo \+= half; sz -= half; \}",
                   r"if \(\(t & \(\(1 << \(D - depth\)\) - 1\)\) == 0\) \{",
                   r"s2 = \(\(t >> \(D - 1 - d\)\) & 1\) \? s2 - half : half;",
                   r"if \(split_all_the_way && s2 > 16\) slots\[t\] = __fadd_rn\(slots\[t\], slots\[t \+ span / 2\]\);"):
        assert re.search(needle, src), needle


def test_fai_magic_remainder_is_exact():
    """emit_geometry (moc_walk_warp.cuh) takes `iq % fai` as iq - umulhi(iq, floor(2^32 / fai) + 1) * fai whenever
    iq < 2^20, with the reciprocal held in 32 bits (moc_sweep.inl), for 2 <= fai <= 63: exhaustive over that whole
    range.  fai = 1 has no 32-bit reciprocal (2^32 + 1 wraps to 1): the kernel answers 0 without it."""
    import numpy as np
    csrc = os.path.dirname(KERNELS)
    walk = open(os.path.join(csrc, "moc_walk_warp.cuh")).read()
    assert "const int fine = w.fai == 1 ? 0" in walk
    assert ": (unsigned)iq < (1u << 20) ? iq - (int)__umulhi((uint32_t)iq, w.fai_magic) * w.fai : iq % w.fai" in walk
    assert "w.fai_magic = (unsigned int)((1ull << 32) / (unsigned long long)std::max(h->F, 1)) + 1u;" in \
        open(os.path.join(csrc, "moc_sweep.inl")).read()
    create = open(os.path.join(csrc, "moc_device.cu")).read()
    assert "if (I->fai < 1) {" in create and "if (I->fai > 63 ||" in create
    iq = np.arange(1 << 20, dtype=np.uint64)

    def remainder(fai):
        magic = np.uint64(((1 << 32) // fai + 1) & 0xFFFFFFFF)      # (unsigned int)(...) + 1u
        return iq - ((iq * magic) >> np.uint64(32)) * np.uint64(fai)

    for fai in range(2, 64):
        assert np.array_equal(remainder(fai), iq % np.uint64(fai)), fai
    assert not np.array_equal(remainder(1), iq % np.uint64(1))      # why fai = 1 does not use it
