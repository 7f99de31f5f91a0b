"""Host logic of the two-team fused coupling kernel (csrc/fused_coupling.cuh: fused_build_schedule), through the C-ABI test
hook `nf_fused_schedule` -- no GPU.  The kernel's three roles (producer, MMA issuer, epilogue teams) walk this list; the
invariants below are the ones their barrier protocol relies on."""
import ctypes as C

import pytest

import nfload

L3, L1, OTHER, FIRST, LAST = 0x80, 0x40, 0x20, 0x10, 0x20


def schedule(nch, chain, n_hoist, delay):
    lib = nfload.load()._capi.lib()
    buf = (C.c_ubyte * 64)()
    n = lib.nf_fused_schedule(nch, chain, n_hoist, delay, buf, 64)
    assert 0 < n <= 48
    return list(buf[:n])


@pytest.mark.parametrize("delay", [1, 4, 6, 9])
@pytest.mark.parametrize("chain", [1, 2])
@pytest.mark.parametrize("nch", [1, 2, 3, 4])
def test_schedule_invariants(nch, chain, delay):
    for n_hoist in sorted({min(nch, max(chain, 2)), nch}):
        items = schedule(nch, chain, n_hoist, delay)
        l2 = [e for e in items if not e & (L3 | L1)]
        l1 = [e for e in items if e & L1 and not e & L3]
        l3 = [e for e in items if e & L3]
        # every K chunk of every hidden chunk exactly once; first-Dense chunks and third-Dense slabs exactly once each
        assert sorted(e & 0xF for e in l2) == [(j << 2) | k for j in range(nch) for k in range(nch)]
        assert sorted(e & 3 for e in l1) == list(range(nch)) and sorted(e & 3 for e in l3) == list(range(nch))
        # hoisted chunks (next network) are 0 .. n_hoist - 1, the rest belongs to this network and precedes the chains reading it
        assert sorted(e & 3 for e in l1 if e & OTHER) == list(range(n_hoist))
        for e in l1:
            if not e & OTHER:
                k = e & 3
                readers = [i for i, x in enumerate(items) if not x & (L3 | L1) and (x & 3) == k]
                assert items.index(e) < min(readers)
        # hoisted first-Dense chunks come after every reader of the h1 planes they overwrite
        for e in l1:
            if e & OTHER:
                k = e & 3
                readers = [i for i, x in enumerate(items) if not x & (L3 | L1) and (x & 3) == k]
                assert items.index(e) > max(readers)
        # accumulation chains: K chunks of a hidden chunk in order, `chain` per chain, flags on the first / last of each chain,
        # nothing that touches the team's accumulator between the members of a chain
        for j in range(nch):
            ks = [(i, e) for i, e in enumerate(items) if not e & (L3 | L1) and ((e >> 2) & 3) == j]
            assert [e & 3 for _, e in ks] == list(range(nch))
            for i, e in ks:
                k = e & 3
                assert bool(e & FIRST) == (k % chain == 0)
                assert bool(e & LAST) == ((k + 1) % chain == 0 or k == nch - 1)
                if not e & FIRST:      # only third-Dense slabs (other accumulators) may sit between the members of a chain
                    between = items[ks[k - 1][0] + 1:i]
                    assert all(x & L3 for x in between)
        # third-Dense slabs in chunk order, wrapped (previous network) ones first; a slab of THIS network only after its chunk's last K chunk
        prev = [e & 3 for e in l3 if e & OTHER]
        cur = [e & 3 for e in l3 if not e & OTHER]
        assert prev == sorted(prev) and cur == sorted(cur)
        assert [e & 3 for e in l3] == prev + cur and (not prev or not cur or min(prev) > max(cur))
        for e in l3:
            if not e & OTHER:
                j = e & 3
                last = max(i for i, x in enumerate(items) if not x & (L3 | L1) and ((x >> 2) & 3) == j)
                assert items.index(e) > last


def test_schedule_epilogue_copy_has_the_same_order_of_everything_else():
    """The epilogue's copy differs only in where the third-Dense slabs sit (drained later)."""
    for nch in (1, 2, 3, 4):
        for chain in (1, 2):
            a = [e for e in schedule(nch, chain, nch, 4) if not e & L3]
            b = [e for e in schedule(nch, chain, nch, 6) if not e & L3]
            assert a == b


def test_schedule_rejects_bad_arguments():
    lib = nfload.load()._capi.lib()
    buf = (C.c_ubyte * 64)()
    assert lib.nf_fused_schedule(5, 2, 2, 4, buf, 64) < 0
    assert lib.nf_fused_schedule(4, 3, 3, 4, buf, 64) < 0
    assert lib.nf_fused_schedule(4, 2, 2, 0, buf, 64) < 0
