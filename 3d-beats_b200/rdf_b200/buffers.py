"""Device buffers for the RDF path: torch tensors behind the small pycuda-GPUArray surface the reference's callers
use (`.fill .set .get .reshape .shape .dtype .size .itemsize __cuda_array_interface__`; SURVEY 8b "buffer convention").

`GPUArray` / `to_gpu` mirror pycuda.gpuarray (used as `cu_array` throughout src/decision_tree.py and the scripts);
`GpuBuffer` mirrors src/engine/buffer.py:10-39 minus the OpenGL interop (`.cu()` returns the GPUArray).
PyTorch is only the allocator and stream owner here; kernels receive raw pointers through the C ABI.
"""
import numpy as np
import torch

_NP2TORCH = {
    np.dtype(np.uint8): torch.uint8,
    np.dtype(np.int8): torch.int8,
    np.dtype(np.uint16): torch.uint16,
    np.dtype(np.int16): torch.int16,
    np.dtype(np.uint32): torch.uint32,
    np.dtype(np.int32): torch.int32,
    np.dtype(np.uint64): torch.uint64,
    np.dtype(np.int64): torch.int64,
    np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}
# unsigned 16/32/64-bit tensors have few torch ops; fills and copies go through a same-width signed view
_SIGNED_VIEW = {torch.uint16: torch.int16, torch.uint32: torch.int32, torch.uint64: torch.int64}


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError('rdf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def torch_dtype(dtype):
    if isinstance(dtype, torch.dtype):
        return dtype
    if dtype is int:                      # the reference uses the removed aliases np.int / np.float (SURVEY note N3)
        dtype = np.int64
    if dtype is float:
        dtype = np.float64
    return _NP2TORCH[np.dtype(dtype)]


class GPUArray:
    """C-contiguous device array with the pycuda.gpuarray.GPUArray methods the RDF callers rely on."""

    def __init__(self, shape, dtype=np.float32, tensor=None, device=None):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        shape = tuple(int(s) for s in shape)
        if tensor is None:
            tensor = torch.empty(shape, dtype=torch_dtype(dtype), device=device or default_device())
        else:
            assert tuple(tensor.shape) == shape and tensor.is_contiguous()
        self.tensor = tensor
        self.shape = shape
        self.dtype = _TORCH2NP[tensor.dtype]

    # -- pycuda surface ------------------------------------------------------------------------------------
    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def itemsize(self):
        return self.dtype.itemsize

    @property
    def nbytes(self):
        return self.size * self.itemsize

    @property
    def ptr(self):
        return self.tensor.data_ptr()

    @property
    def gpudata(self):
        return self.tensor.data_ptr()

    @property
    def __cuda_array_interface__(self):
        return {'shape': self.shape, 'typestr': self.dtype.str, 'data': (self.tensor.data_ptr(), False), 'version': 3,
                'strides': None}

    def _signed(self):
        sv = _SIGNED_VIEW.get(self.tensor.dtype)
        return self.tensor.view(sv) if sv is not None else self.tensor

    def fill(self, value):
        t = self._signed()
        v = np.asarray(value).astype(self.dtype)
        if t.dtype != self.tensor.dtype:
            v = v.view(_TORCH2NP[t.dtype])
        t.fill_(v.item())
        return self

    def set(self, ary):
        if isinstance(ary, GPUArray):
            assert ary.shape == self.shape and ary.dtype == self.dtype
            self._signed().copy_(ary._signed())
            return self
        ary = np.ascontiguousarray(ary)
        assert ary.shape == self.shape, f'shape mismatch {ary.shape} vs {self.shape}'
        assert ary.dtype == self.dtype, f'dtype mismatch {ary.dtype} vs {self.dtype}'
        src = torch.from_numpy(ary.view(_TORCH2NP[self._signed().dtype]) if self._signed().dtype != self.tensor.dtype else ary)
        self._signed().copy_(src)
        return self

    def get(self, ary=None):
        host = self._signed().cpu().numpy()
        if host.dtype != self.dtype:
            host = host.view(self.dtype)
        if ary is not None:
            ary[...] = host
            return ary
        return host

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return GPUArray(shape, tensor=self.tensor.view(*shape))

    def __getitem__(self, idx):
        t = self.tensor[idx]
        if not t.is_contiguous():
            raise ValueError('GPUArray slices must stay contiguous')
        return GPUArray(tuple(t.shape), tensor=t)

    def __len__(self):
        return self.shape[0]

    def __bool__(self):
        # pycuda GPUArrays are truthy by length; the reference's `if filter_images` relies on it
        # (src/decision_tree.py:313,324).  New code must test `is not None`.
        return self.size > 0


def to_gpu(ary):
    ary = np.ascontiguousarray(ary)
    out = GPUArray(ary.shape, dtype=ary.dtype)
    out.set(ary)
    return out


def zeros(shape, dtype=np.float32):
    return GPUArray(shape, dtype=dtype).fill(0)


class GpuBuffer:
    """src/engine/buffer.py:10-39 without OpenGL: `.cu()` hands out the device array, `.gl()` is unsupported."""

    def __init__(self, shape, dtype, data_ptr=None, gl_buffer_flag=None):
        self.shape = tuple(shape) if not isinstance(shape, (int, np.integer)) else (int(shape),)
        self.dtype = _TORCH2NP[torch_dtype(dtype)]
        self._cu = GPUArray(self.shape, dtype=self.dtype)
        if data_ptr is not None:
            self._cu.set(np.asarray(data_ptr, dtype=self.dtype).reshape(self.shape))

    def cu(self):
        return self._cu

    def gl(self):
        raise RuntimeError('GpuBuffer.gl(): OpenGL interop is outside the RDF hot path (SURVEY section 2.1)')


def as_gpuarray(obj):
    """Accept a GPUArray, a torch CUDA tensor, or anything with .cu() (GpuBuffer) and return a GPUArray view."""
    if isinstance(obj, GPUArray):
        return obj
    if hasattr(obj, 'cu') and callable(obj.cu):
        return as_gpuarray(obj.cu())
    if isinstance(obj, torch.Tensor):
        if not obj.is_cuda:
            raise ValueError('expected a CUDA tensor')
        if not obj.is_contiguous():
            raise ValueError('expected a contiguous tensor')
        return GPUArray(tuple(obj.shape), tensor=obj)
    raise TypeError(f'cannot interpret {type(obj).__name__} as a device array')
