"""Minimal DLPack consumer (ctypes): turns any object with ``__dlpack__`` that lives on a CUDA device into a raw
device pointer for the C ABI's ``mmd_set_state_dev`` / ``mmd_get_state_dev`` -- the zero-copy path for JAX, CuPy
and torch arrays (north-star: "ctypes plus DLPack so JAX/NumPy arrays pass zero-copy").  No torch import here."""

import ctypes as C

kDLCUDA, kDLCUDAManaged = 2, 13
kDLFloat = 2


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("device", DLDevice),
        ("ndim", C.c_int),
        ("dtype", DLDataType),
        ("shape", C.POINTER(C.c_int64)),
        ("strides", C.POINTER(C.c_int64)),
        ("byte_offset", C.c_uint64),
    ]


class DLManagedTensor(C.Structure):
    pass


DLManagedTensor._fields_ = [
    ("dl_tensor", DLTensor),
    ("manager_ctx", C.c_void_p),
    ("deleter", C.CFUNCTYPE(None, C.POINTER(DLManagedTensor))),
]

_api = C.pythonapi
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_SetName.restype = C.c_int
_api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]


class DeviceArray:
    """A borrowed view of a DLPack producer's CUDA float64 buffer: ``ptr`` (int), ``shape``; `release()` hands the
    tensor back to the producer (call it after the consuming work has been ordered on the stream)."""

    def __init__(self, obj, stream=None, device=None, shape=None, name="array"):
        if not hasattr(obj, "__dlpack__"):
            raise TypeError(f"{name}: object of type {type(obj).__name__} does not implement __dlpack__")
        if hasattr(obj, "__dlpack_device__"):
            dev_type, dev_id = obj.__dlpack_device__()
            if int(dev_type) not in (kDLCUDA, kDLCUDAManaged):
                raise ValueError(f"{name}: DLPack device type {int(dev_type)} is not CUDA (there is no CPU path)")
        # DLPack protocol: the consumer passes ITS stream; the producer makes the data safe to read on it
        # (1 = legacy default stream, 2 = per-thread default stream, other ints = cudaStream_t values)
        cap = obj.__dlpack__(stream=stream) if stream is not None else obj.__dlpack__()
        if not _api.PyCapsule_IsValid(cap, b"dltensor"):
            raise ValueError(f"{name}: __dlpack__ did not return a 'dltensor' capsule")
        self._cap = cap
        self._mt = C.cast(_api.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(DLManagedTensor))
        _api.PyCapsule_SetName(cap, b"used_dltensor")   # we own the tensor now: the capsule destructor must not free it
        t = self._mt.contents.dl_tensor
        try:
            if t.device.device_type not in (kDLCUDA, kDLCUDAManaged):
                raise ValueError(f"{name}: not a CUDA tensor")
            if device is not None and t.device.device_id != device:
                raise ValueError(f"{name}: lives on cuda:{t.device.device_id}, the chains on cuda:{device}")
            if (t.dtype.code, t.dtype.bits, t.dtype.lanes) != (kDLFloat, 64, 1):
                raise ValueError(f"{name}: dtype must be float64")
            self.shape = tuple(t.shape[i] for i in range(t.ndim))
            if t.strides:
                expect = 1
                for i in range(t.ndim - 1, -1, -1):
                    if self.shape[i] != 1 and t.strides[i] != expect:
                        raise ValueError(f"{name}: must be C-contiguous")
                    expect *= self.shape[i]
            if shape is not None:
                n_have, n_want = 1, 1
                for v in self.shape:
                    n_have *= v
                for v in shape:
                    n_want *= v
                if n_have != n_want or self.shape[0] != shape[0]:
                    raise ValueError(f"{name}: shape {self.shape} does not match {tuple(shape)}")
            self.ptr = (t.data or 0) + t.byte_offset
        except Exception:
            self.release()
            raise

    def release(self):
        if self._mt is not None:
            mt, self._mt = self._mt, None
            if mt.contents.deleter:
                mt.contents.deleter(mt)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
