"""A small in-memory stand-in for ``h5py`` (absent from the build container) that enforces the same restrictions on what
can be stored: attributes and datasets must have a native HDF5 type -- ``None``, dicts, numpy unicode arrays and mixed
lists raise ``TypeError`` exactly where h5py would.  It lets the real-HDF5 branch of ``utils.write_container`` /
``read_container`` execute in the CPU test suite.  Files are pickles behind the HDF5 magic number."""
import pickle

import numpy as np

MAGIC = b"\x89HDF\r\n\x1a\n"
_VLEN = {"vlen": str}


def string_dtype(encoding="utf-8", length=None):
    return np.dtype("O", metadata=_VLEN)


def _check(value, what):
    if value is None:
        raise TypeError(f"Object dtype dtype('O') has no native HDF5 equivalent ({what} is None)")
    if isinstance(value, dict):
        raise TypeError(f"Object dtype dtype('O') has no native HDF5 equivalent ({what} is a dict)")
    if isinstance(value, (str, bytes, bool, int, float, np.generic)):
        return value
    arr = np.asarray(value)
    if arr.dtype.kind == "O" and (arr.dtype.metadata or {}).get("vlen") is not str:
        raise TypeError(f"Object dtype dtype('O') has no native HDF5 equivalent ({what})")
    if arr.dtype.kind == "U":
        raise TypeError(f"No conversion path for dtype: {arr.dtype!r} ({what})")
    return arr


class _Attrs(dict):
    def __setitem__(self, key, value):
        super().__setitem__(key, _check(value, f"attribute {key!r}"))


class Dataset:
    def __init__(self, data):
        self._data = np.array(_check(data, "dataset"))
        self.attrs = _Attrs()

    def __getitem__(self, idx):
        return self._data[idx] if idx != () else (self._data if self._data.ndim else self._data[()])

    @property
    def shape(self):
        return self._data.shape


class Group:
    def __init__(self):
        self._items = {}
        self.attrs = _Attrs()

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node._items:
                if not create:
                    raise KeyError(path)
                node._items[part] = Group()
            node = node._items[part]
        return node

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        return self._walk(path)

    def __iter__(self):
        return iter(self._items)

    def keys(self):
        return self._items.keys()

    def require_group(self, path):
        node = self._walk(path, create=True)
        if not isinstance(node, Group):
            raise TypeError(f"{path} is a dataset")
        return node

    create_group = require_group

    def create_dataset(self, name, data=None, **kw):
        parts = [p for p in name.split("/") if p]
        parent = self._walk("/".join(parts[:-1]), create=True)
        if parts[-1] in parent._items:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        parent._items[parts[-1]] = Dataset(data)
        return parent._items[parts[-1]]

    def visititems(self, func, _prefix=""):
        for k, v in self._items.items():
            name = f"{_prefix}{k}"
            func(name, v)
            if isinstance(v, Group):
                v.visititems(func, _prefix=name + "/")


class File(Group):
    def __init__(self, path, mode="r"):
        super().__init__()
        self._path, self._mode = path, mode
        if mode in ("r", "a", "r+"):
            with open(path, "rb") as fh:
                assert fh.read(len(MAGIC)) == MAGIC, "not an HDF5 file"
                root = pickle.load(fh)
            self._items, self.attrs = root._items, root.attrs

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self._mode != "r":
            root = Group()
            root._items, root.attrs = self._items, self.attrs
            with open(self._path, "wb") as fh:
                fh.write(MAGIC)
                pickle.dump(root, fh)
