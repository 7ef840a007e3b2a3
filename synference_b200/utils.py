"""Hot-path helpers mirroring ``src/synference/utils.py`` of the reference.

Only the helpers on the mock-library path are present:
constant-R wavelength grid (``utils.py:257-289``), min/max grid limits
(``utils.py:115-126``), asinh magnitude conversions (``utils.py:647-805``), the
scaling checks (``utils.py:929-988``) and the library file reader
(``utils.py:37-112``).
"""

from __future__ import annotations

import json
import logging
import os

import numpy as np

from .units import Angstrom, Jy, Quantity, has_units, nJy, strip_units

logger = logging.getLogger("synference_b200")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s | %(name)s | %(levelname)s | %(message)s"))
    logger.addHandler(_h)
    # INFO on rank 0, WARNING elsewhere (utils.py:2331-2376 semantics, with RANK instead of MPI)
    logger.setLevel(logging.INFO if int(os.environ.get("RANK", "0")) == 0 else logging.WARNING)
    if os.environ.get("SYNFERENCE_B200_QUIET"):
        logger.setLevel(logging.WARNING)


def calculate_min_max_wav_grid(filterset, max_redshift, min_redshift=0):
    """Wavelength limits so every filter stays on the grid for min_z <= z <= max_z (utils.py:115-126)."""
    lo, hi = filterset.get_non_zero_lam_lims()
    return lo / (1 + max_redshift), hi / (1 + min_redshift)


def generate_constant_R(R=300, start=1 * Angstrom, end=9e5 * Angstrom, auto_start_stop=False,
                        filterset=None, **kwargs):
    """lambda_{i+1} = lambda_i (1 + 0.5/R), built by repeated multiplication (utils.py:257-289)."""
    if auto_start_stop and filterset is not None:
        start, end = calculate_min_max_wav_grid(filterset, **kwargs)
    start_v = float(strip_units(start, "Angstrom"))
    end_v = float(strip_units(end, "Angstrom"))
    assert start_v < end_v, "Start wavelength must be less than end wavelength."
    assert R > 0, "R must be greater than 0."
    x = [start_v]
    while x[-1] < end_v:
        x.append(x[-1] * (1.0 + 0.5 / R))
    return Quantity(x, Angstrom)


# ---- asinh magnitudes (utils.py:647-805) ------------------------------------

_POGSON = 2.5 * np.log10(np.e)


def _jy(x):
    return strip_units(x, "Jy") if has_units(x) else np.asarray(x, dtype=float)


def _broadcast_b(f_b, like):
    f_b = _jy(f_b)
    if f_b.ndim == 1 and like.ndim == 2:
        assert f_b.shape[0] == like.shape[0], "Flux softening must match the number of filters."
        return np.tile(f_b, (like.shape[1], 1)).T
    if f_b.ndim == 0:
        return np.full_like(like, float(f_b), dtype=float)
    assert f_b.shape == like.shape, "Flux and flux softening must have the same shape."
    return f_b


def f_jy_to_asinh(f_jy, f_b=5 * nJy):
    f = _jy(f_jy)
    b = _broadcast_b(f_b, f)
    return -_POGSON * (np.arcsinh(f / (2 * b)) + np.log(b / 3631.0))


def f_jy_err_to_asinh(f_jy, f_jy_err, f_b=5 * nJy):
    f, e = _jy(f_jy), _jy(f_jy_err)
    assert f.shape == e.shape, "Flux and flux error must have the same shape."
    b = _broadcast_b(f_b, f)
    return _POGSON * e / np.sqrt(f**2 + (2 * b) ** 2)


def asinh_to_f_jy(f_asinh, f_b=5 * nJy):
    m = np.asarray(f_asinh, dtype=float)
    b = _broadcast_b(f_b, m)
    return Quantity(2 * b * np.sinh(-m / _POGSON - np.log(b / 3631.0)), Jy)


def asinh_err_to_f_jy(f_asinh, f_asinh_err, f_b=5 * nJy):
    m, me = np.asarray(f_asinh, dtype=float), np.asarray(f_asinh_err, dtype=float)
    b = _broadcast_b(f_b, m)
    f = 2 * b * np.sinh(-m / _POGSON - np.log(b / 3631.0))
    return Quantity(me * np.sqrt(f**2 + (2 * b) ** 2) / _POGSON, Jy)


def asinh_to_snr(f_asinh, f_asinh_err, f_b=5 * nJy):
    f = np.asarray(asinh_to_f_jy(f_asinh, f_b))
    return f / np.asarray(asinh_err_to_f_jy(f_asinh, f_asinh_err, f_b))


# ---- scaling checks (utils.py:929-988) ---------------------------------------

def check_scaling(arr) -> bool:
    """True when a quantity scales linearly with stellar mass (its unit carries Msun)."""
    return has_units(arr) and "Msun" in str(arr.units) and "log" not in str(arr.units)


def check_log_scaling(arr) -> bool:
    return has_units(arr) and "log10" in str(arr.units) and "Msun" in str(arr.units)


# ---- library container --------------------------------------------------------
# The reference writes HDF5 (library.py:4074-4153).  h5py is not installable here, so
# the same logical layout (dataset paths + attrs) is kept in a flat container (raw arrays behind a JSON
# header when uncompressed, a deflated .npz otherwise); when h5py is importable a real HDF5 file is written instead.

def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


def _h5_attr(h5py, v):
    """A value h5py can store as an attribute (it has no encoding for ``None``, dicts or mixed lists), or ``None`` to skip."""
    if v is None:
        return None
    if isinstance(v, dict):
        return json.dumps(_jsonable(v))
    if isinstance(v, (list, tuple, np.ndarray)):
        items = list(v.tolist() if isinstance(v, np.ndarray) else v)
        if not items:
            return np.zeros(0, dtype=np.float64)
        if all(isinstance(x, (bool, int, float, np.integer, np.floating, np.bool_)) for x in items):
            return np.asarray(items)
        return np.array(["" if x is None else (json.dumps(_jsonable(x)) if isinstance(x, (dict, list, tuple)) else str(x))
                         for x in items], dtype=h5py.string_dtype())
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    return v


def write_container(path, datasets: dict, attrs: dict, compress=True):
    """Write ``{dataset path: array}`` and ``{attr: value}``; returns the path actually written.

    An attribute key ``"Group/Sub@name"`` belongs to that group (or dataset), a plain key to the file root -- the ``Model``
    block of a library is a group with attributes and sub-groups (``library.py:2017-2132``).
    ``compress``: False / 0 none, True deflate at zlib's default level, 1 ... 9 that deflate level."""
    level = 6 if compress is True else int(compress or 0)
    if _have_h5py() and not os.environ.get("SYNFERENCE_B200_FORCE_NPZ"):
        import h5py
        with h5py.File(path, "w") as f:
            for k, v in datasets.items():
                kw = dict(compression="gzip", compression_opts=level) if level and np.ndim(v) else {}
                f.create_dataset(k, data=v, **kw)
            for k, v in attrs.items():
                v = _h5_attr(h5py, v)
                if v is None:
                    continue
                if "@" in k:
                    where, name = k.rsplit("@", 1)
                    obj = f[where] if where in f else f.require_group(where)
                    obj.attrs[name] = v
                else:
                    f.attrs[k] = v
        return path
    if level == 0:
        return _write_raw_container(path, datasets, attrs)
    import zipfile
    payload = {k.replace("/", "::"): np.asarray(v) for k, v in datasets.items()}
    payload["__attrs__"] = np.frombuffer(json.dumps(_jsonable(attrs)).encode(), dtype=np.uint8)
    if level == 6:
        with open(path, "wb") as fh:  # keep the reference's file name, whatever its suffix
            np.savez_compressed(fh, **payload)
        return path
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_DEFLATED, compresslevel=level, allowZip64=True) as zf:
        for k, v in payload.items():
            with zf.open(k + ".npy", "w", force_zip64=True) as f:
                np.lib.format.write_array(f, np.asanyarray(v), allow_pickle=False)
    return path


# Uncompressed container without h5py: magic, header length, JSON header ({dataset: dtype, shape, offset}, attrs), then the
# arrays' bytes at 64-byte aligned offsets.  One write per array straight from its memory: no CRC pass and no staging copy
# (an .npz member costs both), so a library build's files go out at page-cache speed from the writer threads.
_RAW_MAGIC = b"SB2CONT1"


def _write_raw_container(path, datasets: dict, attrs: dict):
    arrays, table, off = [], [], 0
    for k, v in datasets.items():
        a = np.asanyarray(v)
        if a.dtype.hasobject:
            raise TypeError(f"dataset {k}: object arrays cannot be stored")
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        off = (off + 63) // 64 * 64
        table.append([k, a.dtype.str, list(a.shape), off, int(a.nbytes)])
        arrays.append((off, a))
        off += a.nbytes
    header = json.dumps({"datasets": table, "attrs": _jsonable(attrs)}).encode()
    start = (16 + len(header) + 63) // 64 * 64
    with open(path, "wb", buffering=0) as fh:
        fh.write(_RAW_MAGIC + len(header).to_bytes(8, "little") + header + b"\0" * (start - 16 - len(header)))
        pos = 0
        for o, a in arrays:
            if o > pos:
                fh.write(b"\0" * (o - pos))
            if a.nbytes:
                fh.write(a.reshape(-1).view(np.uint8))
            pos = o + a.nbytes
    return path


def _read_raw_container(path):
    with open(path, "rb") as fh:
        head = fh.read(16)
        n = int.from_bytes(head[8:16], "little")
        meta = json.loads(fh.read(n).decode())
        start = (16 + n + 63) // 64 * 64
        out = {}
        for k, dt, shape, off, nbytes in meta["datasets"]:
            dt = np.dtype(dt)
            fh.seek(start + off)
            a = np.fromfile(fh, dtype=dt, count=nbytes // dt.itemsize if dt.itemsize else 0)
            out[k] = a.reshape(shape)
    return out, meta["attrs"]


def _jsonable(x):
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer, np.bool_)):
        return x.item()
    if isinstance(x, bytes):
        return x.decode("utf-8", "replace")
    return x


def _plain(v):
    """h5py attribute value -> plain Python (bytes -> str, string arrays -> list of str)."""
    if isinstance(v, bytes):
        return v.decode("utf-8", "replace")
    if isinstance(v, np.ndarray):
        if v.dtype.kind in ("O", "S", "U"):
            return [_plain(x) for x in v.tolist()]
        return v.tolist()
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    return v


def read_container(path):
    """Inverse of :func:`write_container` -> ``(datasets, attrs)``; group / dataset attributes come back under
    ``"Group/Sub@name"`` keys."""
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic == _RAW_MAGIC:
        return _read_raw_container(path)
    if magic[:2] == b"PK":
        d = np.load(path, allow_pickle=False)
        attrs = json.loads(bytes(d["__attrs__"]).decode()) if "__attrs__" in d.files else {}
        return {k.replace("::", "/"): d[k] for k in d.files if k != "__attrs__"}, attrs
    import h5py
    out, attrs = {}, {}
    with h5py.File(path, "r") as f:
        def visit(name, obj):
            if isinstance(obj, h5py.Dataset):
                out[name] = obj[()]
            for k in obj.attrs:
                attrs[f"{name}@{k}"] = _plain(obj.attrs[k])
        f.visititems(visit)
        attrs.update({k: _plain(f.attrs[k]) for k in f.attrs})
    return out, attrs


def load_library_from_hdf5(hdf5_path, photometry_key="Grid/Photometry", parameters_key="Grid/Parameters",
                           filter_codes_attr="FilterCodes", parameters_attr="ParameterNames",
                           supp_key="Grid/SupplementaryParameters", spectra_key="Grid/Spectra"):
    """Read a library written by ``CombinedBasis.save_library`` (utils.py:37-112)."""
    data, attrs = read_container(hdf5_path)
    out = {"parameters": data[parameters_key],
           "filter_codes": list(attrs.get(filter_codes_attr, [])),
           "parameter_names": list(attrs.get(parameters_attr, [])),
           "parameter_units": list(attrs.get("ParameterUnits", [])),
           "photometry_units": attrs.get("PhotometryUnits", "nJy")}
    if photometry_key in data:
        out["photometry"] = data[photometry_key]
    if spectra_key in data:
        out["spectra"] = data[spectra_key]
    if supp_key in data:
        out["supplementary_parameters"] = data[supp_key]
        out["supplementary_parameter_names"] = list(attrs.get("SupplementaryParameterNames", []))
        out["supplementary_parameter_units"] = list(attrs.get("SupplementaryParameterUnits", []))
    return out


def combine_rank_files(size, filepath, num_galaxies, starts, ends):
    """Combine per-rank PIPELINE files into one (``utils.py:2214-2328``, same signature): ``filepath`` is any of the rank
    files ``<stem>_<rank>.hdf5``; per-galaxy datasets (first axis = galaxies) land in ``[starts[r]:ends[r]]`` of datasets
    sized ``num_galaxies``, the ``Instruments`` / ``EmissionModel`` / ``Model`` groups and the wavelength axis are copied
    once from rank 0, attributes come from the first file that has them; the rank files are removed.  Returns the path."""
    ext = filepath.split(".")[-1]
    path_no_ext = ".".join(filepath.split(".")[:-1])
    new_path = "_".join(path_no_ext.split("_")[:-1]) + f".{ext}"
    temp_path = "_".join(path_no_ext.split("_")[:-1]) + "_<rank>." + ext
    static = ("Instruments/", "EmissionModel/", "Model/", "Wavelengths")
    out, attrs = {}, {}
    for rank in range(size):
        data, a = read_container(temp_path.replace("<rank>", str(rank)))
        for k, v in a.items():
            attrs.setdefault(k, v)
        for k, v in data.items():
            v = np.asarray(v)
            if k.startswith(static) or v.ndim == 0:
                out.setdefault(k, v)
                continue
            if k not in out:
                out[k] = np.zeros((num_galaxies,) + v.shape[1:], dtype=v.dtype)
            out[k][starts[rank]:ends[rank], ...] = v
    attrs["rank"], attrs["world_size"] = 0, 1
    attrs["galaxy_start"], attrs["galaxy_stop"] = 0, int(num_galaxies)
    write_container(new_path, out, attrs, compress=False)
    for rank in range(size):
        os.remove(temp_path.replace("<rank>", str(rank)))
    return new_path

