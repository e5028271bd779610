"""CPU: the host-side plan of the tile-fused stage kernel (cfd-2d_b200/csrc/fvm_tiling.h), checked
through the host-only hook cfd2d_tiling_plan (no GPU): the Hilbert renumbering is a permutation
that leaves halo cells in place, every invariant the kernel relies on holds (verified inside the
hook), and the ring-1 / perimeter-edge overhead of the tiles is what a compact tiling should give."""
import numpy as np
import pytest

from cfd2d_b200 import cases, decomp, fvm


@pytest.mark.parametrize("make,tc", [
    (lambda: cases.channel(48, 24, jitter=0.2, shuffle=True), 64),
    (lambda: cases.channel(48, 24, jitter=0.2, shuffle=True), 512),
    (lambda: cases.forward_step(30, 10, jitter=0.15), 100),
    (lambda: cases.strip(40, 10, jitter=0.2, shuffle=True), 32),
    (lambda: cases.strip(4, 2), 512),
])
def test_plan_invariants(make, tc):
    c = make()
    perm, st = fvm.tiling_plan(c.mesh, c.task, tile_cells=tc)
    nc = c.mesh.nc
    assert sorted(perm.tolist()) == list(range(nc))
    assert st["ntiles"] == -(-nc // tc)
    assert st["sum_ng"] >= nc and st["sum_ne"] >= c.mesh.ne
    assert st["interior"] == st["ntiles"] and st["boundary"] == 0      # serial mesh: no rank halo


def test_identity_order_when_hilbert_is_off():
    c = cases.channel(16, 8, shuffle=True)
    perm, _ = fvm.tiling_plan(c.mesh, c.task, tile_cells=64, hilbert=False)
    assert np.array_equal(perm, np.arange(c.mesh.nc))


def test_tiles_are_compact():
    """Regular 200x100x2 mesh, 512-cell tiles: a compact blob of 512 triangles has ~100 ring-1
    cells and ~6 % perimeter edges; a row-major strip would have ~512 and ~33 %."""
    c = cases.channel(200, 100)
    _, st = fvm.tiling_plan(c.mesh, c.task, tile_cells=512)
    nc, ne = c.mesh.nc, c.mesh.ne
    assert st["sum_ring"] / nc < 0.30, st
    assert st["sum_ne"] / ne < 1.12, st
    _, st0 = fvm.tiling_plan(c.mesh, c.task, tile_cells=512, hilbert=False)
    assert st0["sum_ring"] > 2 * st["sum_ring"]


def test_rank_local_mesh_has_boundary_tiles_and_fixed_halo():
    c = cases.channel(64, 32, jitter=0.1)
    part = decomp.slab_part(c.mesh, 2)
    rm = decomp.decompose(c.mesh, part, 2, only_rank=0)[0]
    perm, st = fvm.tiling_plan(rm.local, c.task, tile_cells=128, nc_owned=rm.nc)
    nc_ex = rm.local["cell_S"].shape[0]
    assert np.array_equal(perm[rm.nc:], np.arange(rm.nc, nc_ex))       # halo slice untouched
    assert sorted(perm[:rm.nc].tolist()) == list(range(rm.nc))
    assert st["boundary"] >= 1 and st["interior"] >= 1
    assert st["interior"] + st["boundary"] == st["ntiles"]


# ---- the pipelined tile kernel's plan (layout 2): per-tile blobs --------------------------------
@pytest.mark.parametrize("make,tc,bins", [
    (lambda: cases.channel(48, 24, jitter=0.2, shuffle=True, two_materials=True), 64, True),
    (lambda: cases.channel(48, 24, jitter=0.2, shuffle=True), 128, False),
    (lambda: cases.forward_step(30, 10, jitter=0.15), 100, True),
    (lambda: cases.strip(40, 10, jitter=0.2, shuffle=True), 32, False),
    (lambda: cases.strip(4, 2), 512, True),
])
def test_pipe_plan_invariants(make, tc, bins):
    """Every table k_stage_pipe reads is re-derived from the blob bytes inside cfd2d_pipe_plan."""
    c = make()
    st = fvm.pipe_plan(c.mesh, c.task, tile_cells=tc, dir_bins=bins)
    tc8 = (max(tc, 32) + 7) // 8 * 8
    assert st["ntiles"] == -(-c.mesh.nc // tc8)
    assert st["sum_ne"] >= c.mesh.ne
    assert st["interior"] == st["ntiles"] and st["boundary"] == 0
    assert st["blob_bytes"] % 16 == 0 and st["blob_max"] % 16 == 0


def test_pipe_plan_bytes_per_cell():
    """The static tables of a stage are ~150 B per cell at 128-cell tiles (64 B per edge, duplicated on
    tile perimeters, + centres, areas, slots, ring-1 tables) -- the three sweeps read ~290 B."""
    c = cases.channel(200, 100)
    st = fvm.pipe_plan(c.mesh, c.task, tile_cells=128, dir_bins=False)
    assert st["blob_bytes"] / c.mesh.nc < 190.0, st
    assert st["sum_ring1"] / c.mesh.nc < 0.45 and st["sum_ring2"] / c.mesh.nc < 0.55, st
    st256 = fvm.pipe_plan(c.mesh, c.task, tile_cells=256, dir_bins=False)
    assert st256["blob_bytes"] < st["blob_bytes"]


def test_pipe_plan_rank_local_mesh():
    c = cases.channel(64, 32, jitter=0.1)
    part = decomp.slab_part(c.mesh, 2)
    rm = decomp.decompose(c.mesh, part, 2, only_rank=0)[0]
    st = fvm.pipe_plan(rm.local, c.task, tile_cells=128, nc_owned=rm.nc)
    assert st["boundary"] >= 1 and st["interior"] >= 1
    assert st["interior"] + st["boundary"] == st["ntiles"]
