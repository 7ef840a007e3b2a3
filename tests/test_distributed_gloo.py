"""world_size=2 checks of the sharding rule and the single collective, on CPU with the gloo backend."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      SYNFERENCE_B200_QUIET="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from synference_b200 import distributed as D
    from synference_b200.utils import write_container
    assert D.rank_world() == (rank, world)
    a, b = D.shard_bounds(total, rank, world)
    full = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
    local = full[a:b].clone()                      # what this rank "synthesised"
    got = D.gather_rows(local, total)
    assert torch.equal(got, full), f"rank {rank}: gathered tensor differs"
    # per-rank library shards (columns = galaxies), merged on the host like combine_rank_files
    write_container(os.path.join(tmp, f"lib_{rank}.hdf5"),
                    {"Grid/Photometry": full[a:b].numpy().T.copy(), "Grid/Parameters": full[a:b, :2].numpy().T.copy()},
                    {"FilterCodes": ["a", "b", "c"], "ParameterNames": ["redshift", "log_mass"], "rank": rank,
                     "world_size": world})
    D.barrier()
    if rank == 0:
        out = D.merge_rank_shards([os.path.join(tmp, f"lib_{r}.hdf5") for r in range(world)], os.path.join(tmp, "lib.hdf5"))
        from synference_b200.utils import load_library_from_hdf5
        lib = load_library_from_hdf5(out)
        assert np.array_equal(lib["photometry"], full.numpy().T)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_shard_gather_merge_world2(tmp_path, total):
    mp.spawn(_worker, args=(2, _free_port(), total, str(tmp_path)), nprocs=2, join=True)


class _StubEngine:
    """Stands in for SynthEngine on CPU: 'photometry' is a known function of the parameters, so the rank plumbing of
    create_mock_library(multi_node=True) can run under gloo without a GPU (the kernels have their own GPU tests)."""

    def __init__(self, codes):
        self.filter_codes = list(codes)

    general = False

    def photometry(self, p, scaled=False, library_out=None):
        z = np.asarray(p.redshift, dtype=np.float64)
        base = np.stack([(j + 1) * (1.0 + z) for j in range(len(self.filter_codes))], 1).astype(np.float32)
        if library_out is not None:      # SynthEngine.photometry's contract: the mass-scaled block, transposed, in the same pass
            mat, col0 = library_out
            mat[:, col0:col0 + len(z)] = base.T.astype(np.float64) * (10.0 ** np.asarray(p.log_mass) / 1e9)
        return base


def _library_worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      SYNFERENCE_B200_QUIET="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synference_b200 as S
    from synference_b200 import distributed as D
    from synference_b200.synthetic import synthetic_grid
    raw = S.FilterCollection(filter_codes=["JWST/NIRCam.F150W", "JWST/NIRCam.F277W", "JWST/NIRCam.F444W"])
    lam = S.generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=8)
    inst = S.Instrument("JWST", filters=S.FilterCollection(filter_codes=raw.filter_codes, new_lam=lam))
    grid = synthetic_grid(lam)
    em = S.PacmanEmission(grid=grid, fesc=0.1, dust_curve=S.Calzetti2000())
    z = np.linspace(0.1, 6.0, n)
    sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.stack([np.full(n, 0.5), np.linspace(0.1, 0.8, n)], 1),
                                   redshifts=z, max_redshift=20)
    basis = S.GalaxyBasis("mn_basis", z, grid, em, sfhs, S.ZDistArray.delta(log10metallicity=np.full(n, -2.0)),
                          galaxy_params={"tau_v": np.linspace(0, 1, n)}, instrument=inst, redshift_dependent_sfh=True)
    basis._engine = lambda *a, **k: _StubEngine(inst.filters.filter_codes)
    logm = np.linspace(8.0, 10.0, n)
    cb = basis.create_mock_library("mn_lib", log_stellar_masses=logm, emission_model_key="emergent", out_dir=tmp,
                                   overwrite=True, batch_size=4, multi_node=True)
    a, b = D.shard_bounds(n, rank, world)
    assert cb.library_photometry.shape == (3, b - a), cb.library_photometry.shape
    want = np.stack([(j + 1) * (1.0 + z[a:b]) for j in range(3)], 0).astype(np.float32) * (10.0 ** logm[a:b] / 1e9)
    np.testing.assert_allclose(cb.library_photometry, want, rtol=1e-12)
    np.testing.assert_allclose(cb.library_parameters[0], z[a:b])
    assert os.path.exists(os.path.join(tmp, f"mn_lib_{rank}.hdf5"))
    D.barrier()
    if rank == 0:
        merged = S.load_library_from_hdf5(D.merge_rank_shards([os.path.join(tmp, f"mn_lib_{r}.hdf5") for r in range(world)],
                                                              os.path.join(tmp, "mn_lib.hdf5")))
        assert merged["photometry"].shape == (3, n)
        np.testing.assert_allclose(merged["parameters"][1], logm)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 13])
def test_create_mock_library_multi_node_world2(tmp_path, n):
    """the rank-slice path of create_mock_library (library.py:3127-3138) end to end under torch.distributed:
    every rank processes its contiguous slice in batches, writes and re-loads only its own pipeline files, saves its shard."""
    mp.spawn(_library_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
