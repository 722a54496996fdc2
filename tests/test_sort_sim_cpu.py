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


def bitonic_network(k):
    """The register network of csrc/sfgs_binning.cu::warp_merge_sort (same loop nest, same compare direction)."""
    VT = len(k)
    kk = 2
    while kk <= VT:
        j = kk >> 1
        while j > 0:
            for i in range(VT):
                l = i ^ j
                if l > i:
                    up = (i & kk) == 0
                    lo, hi = min(k[i], k[l]), max(k[i], k[l])
                    k[i], k[l] = (lo, hi) if up else (hi, lo)
            j >>= 1
        kk <<= 1
    return k


@pytest.mark.parametrize("VT", [4, 8, 16])
def test_warp_sort_register_network_and_key_packing(VT):
    """One warp per tile (lists <= 32*VT): bitonic network per lane + the same merge-path passes; keys carry the reach
    mask below a 24-bit Gaussian id, which must not change the (depth, id) order and must unpack to the canonical key."""
    rng = random.Random(100 + VT)
    for _ in range(200):
        k = [rng.randrange(1 << 64) for _ in range(VT)]
        assert bitonic_network(list(k)) == sorted(k)
    k = [5, 5, 3, INF, INF, 1, 1, 0] * (VT // 4)                 # padded tails and ties
    assert bitonic_network(list(k[:VT])) == sorted(k[:VT])
    for n in [1, 2, VT - 1, VT, VT + 1, 31 * VT + 1, 32 * VT - 1, 32 * VT]:
        depth_id = rng.sample(range(1 << 56), n)
        packed = [((d >> 24) << 32) | ((d & 0xFFFFFF) << 8) | rng.randrange(256) for d in depth_id]
        canon = lambda p: (p & 0xFFFFFFFF00000000) | ((p >> 8) & 0xFFFFFF)      # noqa: E731  key_unpack
        out = merge_sort_sim(packed, VT)
        assert [canon(p) for p in out] == sorted(canon(p) for p in packed), (VT, n)


def warp_sort_sim(keys, VT, T=32):
    """csrc/sfgs_binning.cu::warp_merge_sort: ONE buffer; per pass every lane takes its merge-path split, the split of
    the next diagonal from its right neighbour (or the end of run A at the end of a pair), loads ca keys of A ascending
    and VT - ca keys of B descending (a bitonic sequence) and sorts them with the bitonic MERGE network."""
    n = len(keys)
    nch = (n + VT - 1) // VT
    assert nch <= T
    npad = nch * VT
    buf = [INF] * npad
    for t in range(nch):
        k = [keys[t * VT + i] if t * VT + i < n else INF for i in range(VT)]
        buf[t * VT:t * VT + VT] = bitonic_network(k)
    L = VT
    while L < npad:
        los, geo = [0] * (T + 1), []
        for lane in range(T):
            out0 = lane * VT
            pair0 = out0 & ~(2 * L - 1)
            lenA = min(L, max(0, npad - pair0))
            lenB = min(L, max(0, npad - pair0 - L))
            diag = out0 - pair0
            lo = 0
            if lane < nch:
                lo, hi = max(0, diag - lenB), min(diag, lenA)
                while lo < hi:
                    mid = (lo + hi) >> 1
                    assert 0 <= pair0 + mid < npad and 0 <= pair0 + L + diag - 1 - mid < npad and diag - 1 - mid < lenB
                    if buf[pair0 + mid] <= buf[pair0 + L + diag - 1 - mid]:
                        lo = mid + 1
                    else:
                        hi = mid
            los[lane] = lo
            geo.append((out0, pair0, lenA, lenB, diag))
        loaded = {}
        for lane in range(nch):
            out0, pair0, lenA, lenB, diag = geo[lane]
            pair_end = ((out0 + VT) & (2 * L - 1)) == 0 or lane + 1 >= nch
            a_end = lenA if pair_end else los[lane + 1]
            ai, bi = los[lane], diag - los[lane]
            ca = a_end - ai
            assert 0 <= ca <= VT and ai + ca <= lenA and bi + (VT - ca) <= lenB, (n, VT, L, lane)
            k = []
            for i in range(VT):
                if i < ca:
                    k.append(buf[pair0 + ai + i])
                else:
                    assert 0 <= bi + (VT - 1 - i) < lenB
                    k.append(buf[pair0 + L + bi + (VT - 1 - i)])
            loaded[lane] = k
        for lane in range(nch):
            k = loaded[lane]
            j = VT >> 1
            while j > 0:
                for i in range(VT):
                    l = i ^ j
                    if l > i and k[i] > k[l]:
                        k[i], k[l] = k[l], k[i]
                j >>= 1
            buf[lane * VT:lane * VT + VT] = k
        L *= 2
    return buf[:n]


@pytest.mark.parametrize("VT", [4, 8, 16])
def test_warp_sort_with_register_bitonic_merges(VT):
    rng = random.Random(7 + VT)
    sizes = sorted(set([1, 2, 3, VT - 1, VT, VT + 1, 2 * VT, 2 * VT + 1, 3 * VT, 5 * VT - 2, 17 * VT + 3, 31 * VT, 31 * VT + 1,
                        32 * VT - 1, 32 * VT] + [rng.randrange(1, 32 * VT + 1) for _ in range(60)]))
    for n in sizes:
        if n < 1:
            continue
        keys = rng.sample(range(1 << 60), n)
        assert warp_sort_sim(keys, VT) == sorted(keys), (VT, n)
    # duplicates (cannot occur in a tile, but the padded tail is made of equal keys) keep the multiset
    keys = [rng.randrange(8) for _ in range(19 * VT + 5)]
    assert warp_sort_sim(keys, VT) == sorted(keys)


@pytest.mark.parametrize("VT,T", [(8, 128), (16, 256)])
def test_cta_sort_with_register_bitonic_merges(VT, T):
    """merge_sort_smem (lists of 513..4096 keys): the same scheme with T threads and the neighbour's split read from
    shared memory; and the padded buffer index is injective inside the allocation."""
    rng = random.Random(11 + VT)
    for n in [513, 600, 1000, 1023, 1024, VT * T - 1, VT * T] + [rng.randrange(513, VT * T + 1) for _ in range(6)]:
        if n > VT * T:
            continue
        keys = rng.sample(range(1 << 60), n)
        assert warp_sort_sim(keys, VT, T) == sorted(keys), (VT, T, n)
    W = VT * T
    phys = [i + (i >> 4) for i in range(W)]
    assert len(set(phys)) == W and max(phys) < W + W // 16


@pytest.mark.parametrize("VT", [4, 8, 16])
def test_padded_buffer_is_bank_conflict_free_for_blocked_accesses(VT):
    """Shared-memory layout of the sort buffers (key i at i + i/16, 8-byte keys, 32 four-byte banks): when the 32 lanes of
    a warp access element j of their VT-key blocks — the register sort's stores and the merge passes' stores — every
    16-lane half of the request (what the hardware serves per wavefront for 8-byte accesses) touches 16 different bank
    pairs; without the pad word all lanes of a half hit the same few pairs (stride VT * 8 bytes)."""
    def bank_pair(i, padded):
        phys = i + (i >> 4) if padded else i
        return phys % 16                       # 16 bank pairs of 8 bytes = one 128-byte wavefront
    for j in range(VT):
        for half in (range(0, 16), range(16, 32)):
            padded = [bank_pair(lane * VT + j, True) for lane in half]
            plain = [bank_pair(lane * VT + j, False) for lane in half]
            assert len(set(padded)) == 16, (VT, j, padded)
            assert len(set(plain)) == 16 // VT, (VT, j, plain)      # unpadded: a VT-way conflict in every half
