"""The N > 1 path on CPU: image-row tile ownership + the exchange step of
raytrace-miniapp_b200/dist.py, world_size 2 over gloo.  Each rank computes its tile with the CPU
oracle (standing in for the kernels), then the real exchange code runs; the result must equal
the single-process image bit for bit (each pixel is computed by exactly one rank, same ray
order) and I_ang to rounding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytrace_miniapp_b200 import abi, dist as rdist, problem_io
from conftest import GOLDEN, rel_l2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_problem(nx_keep):
    p, _ = problem_io.load_npz(os.path.join(GOLDEN, "ase_small.npz"))
    e = p.euv_beam
    e2 = abi.BeamGrid(e.x[:nx_keep], e.y, e.a[::3], e.b[::3], e.dx, e.dy, 3 * e.da, 3 * e.db,
                      dv=e.dv, dz=e.dz)
    return abi.Problem(e2, p.gain)


def _worker(rank, world, port, nx_keep, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle
        p = _small_problem(nx_keep)
        e = p.euv_beam
        n_pix = e.nx * e.ny
        lo, hi, per = rdist.tile_bounds(n_pix, world, rank)
        rays = p.rays().reshape(e.nx, e.ny, e.na * e.nb)
        pix = np.arange(n_pix)
        mine = pix[lo:hi]
        sel = np.concatenate([rays[q % e.nx, q // e.nx] for q in mine]) if mine.size else rays[:0, 0, 0]
        r = pyoracle.Oracle().trace_rays(p, sel, 1, 1.0)
        image = torch.from_numpy(r["image"].copy())
        I_ang = torch.from_numpy(r["I_ang"].copy())
        # rows outside the tile must be untouched by this rank
        rows = image.view(n_pix, e.nv)
        assert not rows[:lo].any() and not rows[hi:].any()
        rdist.exchange(image, I_ang, n_pix, e.nv, 1)
        if rank == 0:
            np.save(out + "_image.npy", image.numpy())
            np.save(out + "_iang.npy", I_ang.numpy())
        # row-cyclic decomposition: rows rank, rank + world, ...; exchange = sum of disjoint rows
        rows_sel = [j for j in range(e.ny) if j % world == rank]
        sel2 = np.concatenate([rays[i, j] for j in rows_sel for i in range(e.nx)])
        r2 = pyoracle.Oracle().trace_rays(p, sel2, 1, 1.0)
        img2, ang2 = torch.from_numpy(r2["image"].copy()), torch.from_numpy(r2["I_ang"].copy())
        # the gather form of the same exchange: every rank contributes its rows compactly
        # (1/world of the image), the gathered blocks are un-permuted into the image
        per = rdist.rows_per_rank(e.ny, world)
        row_elems = e.nx * e.nv
        part = torch.zeros(per * row_elems, dtype=torch.float64)
        mine_rows = img2.view(e.ny, row_elems)[rank::world]
        part.view(per, row_elems)[:mine_rows.shape[0]] = mine_rows
        gathered = torch.empty(world * per * row_elems, dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, part)
        img3 = rdist.unpermute_rows(gathered, torch.zeros_like(img2), e.ny, row_elems, world)
        rdist.exchange_rows(img2, ang2)
        assert torch.equal(img3, img2)
        if rank == 0:
            np.save(out + "_image_cyclic.npy", img2.numpy())
        # seeded-style exchange: plain sums of full-size partials
        a = torch.full((12,), float(rank + 1), dtype=torch.float64)
        b = torch.full((5,), 10.0 * (rank + 1), dtype=torch.float64)
        rdist.exchange(a, b, 4, 3, 2)
        assert torch.equal(a, torch.full((12,), 3.0, dtype=torch.float64))
        assert torch.equal(b, torch.full((5,), 30.0, dtype=torch.float64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx_keep", [60, 59], ids=["even_tiles", "ragged_tiles"])
def test_two_rank_tiles_equal_single_process(nx_keep, oracle, tmp_path):
    out = str(tmp_path / "r0")
    mp.spawn(_worker, args=(2, _free_port(), nx_keep, out), nprocs=2, join=True)
    p = _small_problem(nx_keep)
    full = oracle.create_image(p)
    img, ang = np.load(out + "_image.npy"), np.load(out + "_iang.npy")
    assert np.array_equal(img, full["image"])
    assert np.array_equal(np.load(out + "_image_cyclic.npy"), full["image"])
    assert rel_l2(ang, full["I_ang"]) < 1e-14
    assert np.linalg.norm(img) > 0


def test_tile_bounds_cover_everything():
    for n in (0, 1, 7, 1500, 4200, 4201):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = rdist.tile_bounds(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_bench_workloads_scale_rows_with_the_gpu_count():
    """bench.py --workload: every family keeps rays per GPU fixed under weak scaling (the image rows,
    the sharded axis, are multiplied by the GPU count) and is the fixed image under strong scaling."""
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    saved = bench.WORKLOAD
    try:
        for kind, rays1 in (("ase_medium", 2994600), ("s4", 1596000), ("s4x", 1596000), ("spectral128", 1596000)):
            bench.WORKLOAD = kind
            p1, name1 = bench.workload(1, "weak")
            p4, name4 = bench.workload(4, "weak")
            ps, _ = bench.workload(4, "strong")
            assert p1.n_rays == rays1 and p4.n_rays == 4 * rays1 and ps.n_rays == rays1, kind
            assert p4.euv_beam.ny == 4 * p1.euv_beam.ny and p4.euv_beam.nx == p1.euv_beam.nx
            assert "ny x4" in name4 and "ny x" not in name1
        bench.WORKLOAD = "spectral128"
        assert bench.workload(1, "weak")[0].euv_beam.nv == 128
    finally:
        bench.WORKLOAD = saved
