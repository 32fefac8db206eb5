"""The column sweep stages every level's input runs with 16-byte-granular bulk copies (k_eco.cu: fetch_level).  With an
odd numColumnsMax every other run starts 8 bytes off a boundary and the copy fetches the aligned superset of the run.
This is the arithmetic of that code, restated, checked over many shapes: every copy starts on a 16-byte boundary, has a
16-byte-multiple size, stays inside the array it reads (no byte before the first or behind the last element of a
(level, column) array), fits the stage row, and - together with the one element the last column's thread fetches itself
behind the last level of the last block - covers the block's columns."""
import pytest

BLOCK = 256
PITCH = BLOCK + 2


def fetch_plan(nL, nC, block, kk):
    col0 = block * BLOCK
    cols = min(BLOCK, nC - col0)
    last = col0 + BLOCK >= nC
    off = nC * kk
    mis = (off + col0) & 1
    need = cols + mis
    n = (need + 1) & ~1
    tail = n != need and last and kk == nL - 1
    if tail:
        n = need - 1
    return off + col0 - mis, n, mis, cols, tail


@pytest.mark.parametrize("nL", [1, 2, 4, 36, 60, 80])
def test_bulk_copies_of_the_sweep_stay_inside_their_arrays(nL):
    for nC in list(range(1, 70)) + [255, 256, 257, 258, 511, 512, 513, 29395, 29396]:
        if (nC & 1) and (nL & 1):
            continue   # slabs_are_bulk_copyable(): the tracer slabs alternate between the two alignments -> cp.async staging
        for block in range((nC + BLOCK - 1) // BLOCK):
            for kk in range(nL):
                start, n, mis, cols, tail = fetch_plan(nL, nC, block, kk)
                assert start % 2 == 0 and n % 2 == 0                      # 16-byte aligned address and size
                assert 0 <= start and start + n <= nL * nC                # inside the array
                assert n <= PITCH                                         # inside the stage row
                covered = set(range(start, start + n))
                if tail:
                    covered.add(nC * kk + block * BLOCK + cols - 1)       # the last column's own element
                assert set(range(nC * kk + block * BLOCK, nC * kk + block * BLOCK + cols)) <= covered
                # the thread of column c reads stage element c + mis
                assert mis in (0, 1) and (start + mis == nC * kk + block * BLOCK)
