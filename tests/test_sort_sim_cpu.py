"""Index arithmetic of the per-tile merge sort (csrc/sfgs_binning.cu::merge_sort_smem) restated in Python: register
sort of VT keys per thread, then merge-path passes.  Guards the bounds of the binary search and of the sequential
merge (every shared-memory index the kernel forms is asserted in range here) for ragged list lengths."""
import random

import pytest

INF = (1 << 64) - 1


def merge_sort_sim(keys, VT):
    n = len(keys)
    nch = (n + VT - 1) // VT
    npad = nch * VT
    a = [INF] * npad
    for t in range(nch):
        k = [keys[t * VT + i] if t * VT + i < n else INF for i in range(VT)]
        for r in range(VT):                                   # odd-even transposition network
            for i in range(r & 1, VT - 1, 2):
                if k[i] > k[i + 1]:
                    k[i], k[i + 1] = k[i + 1], k[i]
        a[t * VT:t * VT + VT] = k
    src, dst = a, [None] * npad
    L = VT
    while L < npad:
        for t in range(nch):
            out0 = t * VT
            pair0 = out0 & ~(2 * L - 1)
            lenA = min(L, npad - pair0)
            lenB = min(L, max(0, npad - pair0 - L))
            A, B = src[pair0:pair0 + lenA], src[pair0 + L:pair0 + L + lenB]
            diag = out0 - pair0
            lo, hi = max(0, diag - lenB), min(diag, lenA)
            while lo < hi:
                mid = (lo + hi) >> 1
                assert 0 <= mid < lenA and 0 <= diag - 1 - mid < lenB
                if A[mid] <= B[diag - 1 - mid]:
                    lo = mid + 1
                else:
                    hi = mid
            ai, bi = lo, diag - lo
            assert 0 <= ai <= lenA and 0 <= bi <= lenB
            ka = A[ai] if ai < lenA else INF
            kb = B[bi] if bi < lenB else INF
            for i in range(VT):
                take_a = bi >= lenB or (ai < lenA and ka <= kb)
                dst[out0 + i] = ka if take_a else kb
                if take_a:
                    ai += 1
                    ka = A[ai] if ai < lenA else INF
                else:
                    bi += 1
                    kb = B[bi] if bi < lenB else INF
        src, dst = dst, src
        L *= 2
    return src[:n]


@pytest.mark.parametrize("VT,window", [(8, 1024), (16, 4096)])
def test_merge_path_sort_index_math(VT, window):
    rng = random.Random(VT)
    sizes = list(range(1, 70)) + [127, 128, 129, 255, 256, 257, 276, 511, 512, 513, 1000, window - 1, window]
    for n in sizes:
        if n > window:
            continue
        keys = rng.sample(range(1 << 50), n)                  # (depth, id) keys are unique
        assert merge_sort_sim(keys, VT) == sorted(keys), (VT, n)
