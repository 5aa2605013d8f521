"""Specification check of the shared-memory panel layout of the fp64 Cholesky kernels (btkb_wpe.cu: pidx / k_wpe_chol; btkb_wide.cu:
blk_idx / k_mvdr_solve_wide_blk): two planes of doubles, row stride 20, the column XOR-swizzled with the low four bits
of the row,  idx(r, j) = 20 r + (j ^ (r & 15)).

The claim is that this ONE layout is free of bank conflicts for both access shapes of the kernels.  A shared-memory wavefront serves
32 banks of 4 bytes; a warp's 8-byte accesses are served half a warp at a time, so the 16 lanes of a half-warp must touch 16 different
bank PAIRS:

  (a) "every lane its own row, all lanes the same column" — scaling by the pivot, the TRSM rows, panel load / store;
  (b) the operand fragments of mma.sync.m8n8k4.f64: lane l reads row R0 + l / 4, column k0 + l % 4 (R0 a multiple of 8, k0 of 4).

The plain layouts fail one shape each (stride 20 without the swizzle: (a); stride 17: (b)), which is why the swizzle is there.
CPU only: pure index arithmetic."""
import itertools

import pytest

PS = 20


def pidx(r, j):
    return r * PS + (j ^ (r & 15))


def bank_pairs(doubles):
    return [(2 * d % 32) // 2 for d in doubles]       # an 8-byte element at double index d occupies banks 2d, 2d + 1 (mod 32)


def conflict_free(doubles):
    b = bank_pairs(doubles)
    return len(set(b)) == len(b)


@pytest.mark.parametrize("r0", [0, 1, 5, 16, 23, 100, 141])
def test_same_column_sixteen_consecutive_rows(r0):
    for j in range(16):
        assert conflict_free([pidx(r0 + l, j) for l in range(16)]), (r0, j)


@pytest.mark.parametrize("R0,k0", list(itertools.product([0, 8, 16, 24, 136], [0, 4, 8, 12])))
def test_mma_fragment_half_warps(R0, k0):
    for half in (0, 1):
        lanes = range(16 * half, 16 * half + 16)
        assert conflict_free([pidx(R0 + l // 4, k0 + l % 4) for l in lanes]), (R0, k0, half)


def test_a_row_of_the_panel_stays_inside_its_sixteen_columns():
    for r in range(64):
        assert sorted(pidx(r, j) - r * PS for j in range(16)) == list(range(16))


def test_the_plain_layouts_each_fail_one_shape():
    plain20 = lambda r, j: r * 20 + j
    plain17 = lambda r, j: r * 17 + j
    assert not conflict_free([plain20(l, 3) for l in range(16)])                              # rows 4 apart share a bank pair
    assert conflict_free([plain20(l // 4, 4 + l % 4) for l in range(16)])
    assert conflict_free([plain17(l, 3) for l in range(16)])
    assert not conflict_free([plain17(l // 4, 4 + l % 4) for l in range(16)])
