"""Checkpoint reader with the reference's interface (reference: tinyfusers/storage/unpicker.py:15-87):
`load_weights(path)` reads a torch zip checkpoint (`<name>/data.pkl` + `<name>/data/<key>` storages) without torch
and returns the unpickled object with every tensor as an fp32 numpy array.

Host-side glue next to the hot path. Two things the reference gets wrong are done right
here, because real checkpoints need them: half-precision storages are decoded as IEEE fp16 (the reference reads
them with struct format 'f', unpicker.py:36), and a tensor's storage offset and strides are honoured (the reference
reshapes the whole storage, unpicker.py:25-28), so views that share one storage load correctly."""
import collections
import io
import pickle
import zipfile

import numpy as np

_STORAGE_DTYPES = {
    "FloatStorage": np.dtype("<f4"), "HalfStorage": np.dtype("<f2"), "DoubleStorage": np.dtype("<f8"),
    "IntStorage": np.dtype("<i4"), "LongStorage": np.dtype("<i8"), "ShortStorage": np.dtype("<i2"),
    "CharStorage": np.dtype("i1"), "ByteStorage": np.dtype("u1"), "BoolStorage": np.dtype("?"),
    "BFloat16Storage": "bfloat16",
}


class _StorageType:
    def __init__(self, name):
        self.name, self.dtype = name, _STORAGE_DTYPES[name]


class TypedStorage:
    """One `<name>/data/<key>` entry, decoded lazily."""

    def __init__(self, archive, prefix, dtype, file_index, device, num_elements):
        self._archive, self._prefix = archive, prefix
        self.dtype, self.file_index, self.device, self.num_elements = dtype, file_index, device, num_elements
        self._data = None

    def __call__(self):
        if self._data is None:
            raw = self._archive.read(f"{self._prefix}/data/{self.file_index}")
            if self.dtype == "bfloat16":      # bf16 = upper half of an fp32 word
                u16 = np.frombuffer(raw, dtype="<u2")
                self._data = (u16.astype(np.uint32) << 16).view(np.float32)
            else:
                self._data = np.frombuffer(raw, dtype=self.dtype)
        return self._data


def _rebuild_tensor_v2(storage, storage_offset, size, stride, requires_grad=False, backward_hooks=None, metadata=None):
    flat = storage()
    size, stride = tuple(size), tuple(stride)
    if len(size) == 0:
        view = flat[storage_offset:storage_offset + 1].reshape(())
    else:
        itemsize = flat.dtype.itemsize
        view = np.lib.stride_tricks.as_strided(flat[storage_offset:], shape=size,
                                               strides=tuple(s * itemsize for s in stride), writeable=False)
    if view.dtype.kind == "f":
        return np.array(view, dtype=np.float32, order="C")     # always a fresh, writable fp32 array
    return np.array(view, order="C")


def _rebuild_parameter(data, requires_grad=False, backward_hooks=None, *rest):
    return data


class _Opaque:
    """Stand-in for classes the checkpoint mentions but the weights do not need (training callbacks, ...)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["state"] = state


class TorchUnpickler(pickle.Unpickler):
    def __init__(self, fh, archive, prefix):
        super().__init__(fh)
        self._archive, self._prefix = archive, prefix
        self._storages = {}

    def persistent_load(self, saved_id):
        assert saved_id[0] == 'storage', f"unknown persistent id {saved_id[0]!r}"
        _, type_class, file_index, device, num_elements = saved_id[:5]
        st = self._storages.get(file_index)
        if st is None:
            dtype = type_class.dtype if isinstance(type_class, _StorageType) else np.dtype(type_class)
            st = self._storages[file_index] = TypedStorage(self._archive, self._prefix, dtype, file_index, device, num_elements)
        return st

    def find_class(self, module, name):
        if module == 'collections' and name == 'OrderedDict':
            return collections.OrderedDict
        if module == 'torch._utils' and name in ('_rebuild_tensor_v2', '_rebuild_tensor'):
            return _rebuild_tensor_v2
        if module == 'torch._utils' and name == '_rebuild_parameter':
            return _rebuild_parameter
        if module == 'torch' and name in _STORAGE_DTYPES:
            return _StorageType(name)
        if module == 'torch' and name == 'Size':
            return tuple
        if module in ('numpy.core.multiarray', 'numpy._core.multiarray') and name == 'scalar':
            return np.core.multiarray.scalar if hasattr(np, "core") else np._core.multiarray.scalar
        if module == 'numpy' and name == 'dtype':
            return np.dtype
        if module == '_codecs' and name == 'encode':
            import _codecs
            return _codecs.encode
        if module.startswith('pytorch_lightning'):
            return _Opaque
        raise pickle.UnpicklingError(f"global {module}.{name} is not supported")


def load_weights(weight_path):
    if not zipfile.is_zipfile(weight_path):
        raise NameError(f"File format not supported: {weight_path}")
    with zipfile.ZipFile(weight_path, 'r') as archive:
        pkl = next(n for n in archive.namelist() if n.endswith('/data.pkl') and n.count('/') == 1)
        prefix = pkl.split('/', 1)[0]
        return TorchUnpickler(io.BytesIO(archive.read(pkl)), archive, prefix).load()
