"""File contract of the library builders on CPU: the real-HDF5 branch of the container (through a strict h5py stand-in),
the ``Model`` group, per-rank loading of pipeline files and the rank-file merge."""
import os
import sys

import numpy as np
import pytest

import synference_b200 as S
from synference_b200 import utils as U
from synference_b200.synthetic import synthetic_grid


@pytest.fixture
def h5py_stub(monkeypatch):
    from tests import fake_h5py
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    monkeypatch.delenv("SYNFERENCE_B200_FORCE_NPZ", raising=False)
    return fake_h5py


def _basis(n=6):
    raw = S.FilterCollection(filter_codes=["JWST/NIRCam.F150W", "JWST/NIRCam.F277W", "JWST/NIRCam.F444W"])
    lam = S.generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=8)
    inst = S.Instrument("JWST", filters=S.FilterCollection(filter_codes=raw.filter_codes, new_lam=lam))
    grid = synthetic_grid(lam)
    em = S.TotalEmission(grid=grid, dust_curve=S.Calzetti2000(slope=-0.2), dust_emission_model=S.Greybody(40.0, 1.5), fesc=0.1)
    z = np.linspace(0.5, 6.0, n)
    sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.stack([np.full(n, 0.5), np.linspace(0.1, 0.8, n)], 1),
                                   redshifts=z, max_redshift=20)
    zds = S.ZDistArray.delta(log10metallicity=np.linspace(-3, -1.5, n))
    b = S.GalaxyBasis("stub_basis", z, grid, em, sfhs, zds, galaxy_params={"tau_v": np.linspace(0, 1, n)}, instrument=inst,
                      redshift_dependent_sfh=True, build_library=False)
    b._create_matched_galaxies()
    return b, inst, grid


def test_h5py_branch_round_trips_the_model_block(tmp_path, h5py_stub):
    """with h5py present every attribute of the library must have an HDF5 type.  The Model block holds
    None, dicts, lists of strings and group attributes; it is written as the reference's ``Model`` group."""
    b, inst, grid = _basis()
    data, attrs = b._model_block({"emission_model_key": "total", "cat_type": "photometry", "none_value": None,
                                  "a_dict": {"x": 1.0}, "depths": [28.0, 29.0, 30.0]},
                                 {"tau_v": ("Av", lambda t: 1.086 * t["tau_v"])} if False else None)
    path = str(tmp_path / "lib.hdf5")
    datasets = {"Grid/Photometry": np.arange(18.0).reshape(3, 6), "Grid/Parameters": np.ones((2, 6))}
    datasets.update(data)
    attrs = dict(attrs, ParameterNames=["redshift", "log_mass"], FilterCodes=list(inst.filters.filter_codes), mixed=[1, "a", None])
    U.write_container(path, datasets, attrs, compress=6)
    with open(path, "rb") as fh:
        assert fh.read(4) == b"\x89HDF"               # the HDF5 branch ran, not the npz fallback
    d2, a2 = U.read_container(path)
    np.testing.assert_array_equal(d2["Grid/Photometry"], datasets["Grid/Photometry"])
    np.testing.assert_array_equal(d2["Model/depths"], [28.0, 29.0, 30.0])
    assert a2["Model@grid_name"] == grid.grid_name and a2["Model/EmissionModel@name"] == "TotalEmission"
    assert a2["Model/EmissionModel@dust_law"] == "Calzetti2000" and a2["Model/EmissionModel@dust_emission"] == "Greybody"
    assert a2["Model@sfh_class"] == "LogNormal" and a2["Model@metallicity_distribution_class"] == "DeltaConstant"
    assert list(a2["Model@filters"]) == list(inst.filters.filter_codes) and "Model@none_value" not in a2
    assert a2["mixed"] == ["1", "a", ""]
    np.testing.assert_allclose(d2["Model/Instrument/Filters/Header/Wavelengths"], np.asarray(inst.filters.lam))
    np.testing.assert_allclose(d2["Model/Instrument/Filters/JWST/NIRCam.F277W/Transmission"], inst.filters.filters[1].t)
    lib = S.load_library_from_hdf5(path)
    assert lib["photometry"].shape == (3, 6) and lib["parameter_names"] == ["redshift", "log_mass"]
    # and the strict stand-in really refuses what h5py refuses
    with pytest.raises(TypeError):
        h5py_stub.File(str(tmp_path / "x.hdf5"), "w").attrs.__setitem__("k", None)
    with pytest.raises(TypeError):
        h5py_stub.File(str(tmp_path / "x.hdf5"), "w").attrs.__setitem__("k", {"a": 1})


def test_save_library_with_model_block_both_containers(tmp_path, h5py_stub, monkeypatch):
    b, inst, grid = _basis()
    cb = S.CombinedBasis([b], np.full(6, 9.0), b.redshifts, ["total"], None, out_name="lib", out_dir=str(tmp_path))
    cb._extra_datasets, cb._extra_attrs = b._model_block({"emission_model_key": "total"})
    lib = {"photometry": np.ones((3, 6)), "parameters": np.ones((2, 6)), "parameter_names": ["redshift", "log_mass"],
           "filter_codes": list(inst.filters.filter_codes), "parameter_units": ["dimensionless", "log10_Msun"]}
    cb.save_library(lib, overwrite=True)
    _, a = U.read_container(cb.library_path)
    assert a["Model@emission_model_key"] == "total" and a["ParameterUnits"] == ["dimensionless", "log10_Msun"]
    monkeypatch.setenv("SYNFERENCE_B200_FORCE_NPZ", "1")
    cb.save_library(lib, overload_out_name="lib_npz", overwrite=True)
    _, a2 = U.read_container(cb.library_path)
    assert a2["Model@emission_model_key"] == "total" and a2["Model/EmissionModel@dust_attenuation_keys"] == a["Model/EmissionModel@dust_attenuation_keys"]


def _stub_pipeline_file(path, model, lo, hi, rank, world, codes, label="JWST", key="total"):
    n = hi - lo
    data = {f"Galaxies/Stars/Photometry/Fluxes/{key}/{label}/{c}": np.arange(lo, hi, dtype=float) + 100.0 * j for j, c in enumerate(codes)}
    data.update({"Galaxies/redshift": np.linspace(0, 1, 10)[lo:hi], "Galaxies/tau_v": np.arange(lo, hi) * 0.1,
                 "Galaxies/mass": np.full(n, 1e9), "Wavelengths": np.arange(5.0)})
    U.write_container(path, data, {"FilterCodes": codes, "InstrumentLabel": label, "rank": rank, "world_size": world,
                                   "galaxy_start": lo, "galaxy_stop": hi, "supp_names": [], "supp_units": []}, compress=False)


def test_load_bases_takes_only_this_ranks_files(tmp_path, monkeypatch):
    """with multi_node every rank compiles its own shard from ITS pipeline files; look-alike files
    that merely start with the model name are ignored."""
    from synference_b200 import library as L
    b, inst, grid = _basis(10)
    codes = list(inst.filters.filter_codes)
    out = str(tmp_path)
    _stub_pipeline_file(os.path.join(out, "stub_basis_rank0.hdf5"), "stub_basis", 0, 5, 0, 2, codes)
    _stub_pipeline_file(os.path.join(out, "stub_basis_rank1.hdf5"), "stub_basis", 5, 10, 1, 2, codes)
    _stub_pipeline_file(os.path.join(out, "stub_basis_v2.hdf5"), "stub_basis", 0, 10, 0, 1, codes)        # unrelated
    _stub_pipeline_file(os.path.join(out, "stub_basis_extra_lib.hdf5"), "stub_basis", 0, 10, 0, 1, codes)  # unrelated
    b.varying_param_names = ["redshift", "tau_v"]
    for rank, (lo, hi) in enumerate([(0, 5), (5, 10)]):
        monkeypatch.setattr(L._dist, "rank_world", lambda r=rank: (r, 2))
        cb = S.CombinedBasis([b], np.linspace(8, 10, 10), np.linspace(0, 1, 10), ["total"], None, out_name="lib", out_dir=out)
        mask = np.zeros(10, bool)
        mask[lo:hi] = True
        cb._mask, cb._multi_node = mask, True
        ent = cb.load_bases()["stub_basis"]
        np.testing.assert_array_equal(ent["observed_photometry"][codes[0]], np.arange(lo, hi, dtype=float))
        lib = cb.create_full_library(save=True, overwrite=True)
        assert lib["photometry"].shape == (3, hi - lo)
        np.testing.assert_allclose(lib["photometry"][1], (np.arange(lo, hi) + 100.0).astype(np.float32) * 10 ** np.linspace(8, 10, 10)[lo:hi] / 1e9)
        assert os.path.exists(os.path.join(out, f"lib_{rank}.hdf5"))
    # single process: the two rank files are one population
    monkeypatch.setattr(L._dist, "rank_world", lambda: (0, 1))
    os.remove(os.path.join(out, "stub_basis_v2.hdf5"))
    cb = S.CombinedBasis([b], np.linspace(8, 10, 10), np.linspace(0, 1, 10), ["total"], None, out_name="lib_all", out_dir=out)
    assert cb.create_full_library(save=False)["photometry"].shape == (3, 10)
    # merged library = concatenation of the rank shards, Model / attributes kept once
    from synference_b200.distributed import merge_rank_shards
    merged = S.load_library_from_hdf5(merge_rank_shards([os.path.join(out, f"lib_{r}.hdf5") for r in range(2)], os.path.join(out, "m.hdf5")))
    assert merged["photometry"].shape == (3, 10)


def test_combine_rank_files_slices_galaxies_along_the_first_axis(tmp_path):
    """utils.py:2214-2328: per-galaxy datasets go to [starts[r]:ends[r]], metadata groups are copied once, rank files go."""
    codes = ["a", "b"]
    out = str(tmp_path)
    for r, (lo, hi) in enumerate([(0, 4), (4, 7)]):
        _stub_pipeline_file(os.path.join(out, f"pipe_{r}.hdf5"), "pipe", lo, hi, r, 2, codes)
    path = S.combine_rank_files(2, os.path.join(out, "pipe_0.hdf5"), 7, [0, 4], [4, 7])
    assert path == os.path.join(out, "pipe.hdf5") and not os.path.exists(os.path.join(out, "pipe_1.hdf5"))
    d, a = U.read_container(path)
    np.testing.assert_array_equal(d["Galaxies/Stars/Photometry/Fluxes/total/JWST/b"], np.arange(7.0) + 100.0)
    assert d["Wavelengths"].shape == (5,) and a["world_size"] == 1 and a["galaxy_stop"] == 7


def test_containers_without_h5py_round_trip(tmp_path, monkeypatch):
    """write_container / read_container when h5py is absent: uncompressed files are the raw container (JSON header + arrays at
    aligned offsets, one write per array), compressed ones a deflated .npz; both give back dtypes, shapes, order-independent
    contents and the attribute block (None, dicts and group attributes included)."""
    monkeypatch.setenv("SYNFERENCE_B200_FORCE_NPZ", "1")
    rng = np.random.default_rng(0)
    data = {"Grid/Photometry": rng.random((3, 1001)), "Grid/Parameters": np.asfortranarray(rng.random((4, 7))),
            "Model/names": np.array(["ab", "cde"]), "empty": np.zeros((0, 5)), "scalar": np.float64(3.0),
            "f32": rng.random((5, 2)).astype(np.float32), "i32": np.arange(7, dtype=np.int32), "flags": np.array([True, False])}
    attrs = {"ParameterNames": ["a", "b"], "Model@key": "total", "none": None, "nested": {"q": 1.5, "l": [1, 2]}, "n": 3}
    for compress in (False, True, 3):
        path = str(tmp_path / f"c_{compress}.hdf5")
        assert U.write_container(path, data, attrs, compress=compress) == path
        with open(path, "rb") as fh:
            magic = fh.read(8)
        assert (magic == b"SB2CONT1") == (compress is False)
        got, a = U.read_container(path)
        assert set(got) == set(data) and a == attrs
        for k, v in data.items():
            v = np.asarray(v)
            assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
        assert got["Grid/Photometry"].flags.writeable
    with pytest.raises(TypeError):
        U.write_container(str(tmp_path / "bad.hdf5"), {"o": np.array([{"a": 1}], dtype=object)}, {}, compress=False)
