"""Host-side mirror of the reference's Sobol interface (src/sobol_sensitivity.jl) over libgsa_cuda.so.

Reference API                                            here
---------------------------------------------------------------------------------------------------
Sobol(; order=[0,1], nboot=1, conf_level=0.95)  :77-83   Sobol(order=(0, 1), nboot=1, conf_level=0.95)
SobolResult (S1, S1_Conf_Int, S2, S2_Conf_Int,  :85-92   SobolResult (same field names and order)
             ST, ST_Conf_Int)
gsa(f, ::Sobol, A, B; batch, Ei_estimator)      :110     gsa(f, Sobol(...), A, B, batch=True, Ei_estimator=...)
gsa(f, ::Sobol, p_range; samples, ...)          :407     gsa(f, Sobol(...), p_range, samples=...)
gsa_sobol_all_y_analysis(method, all_y, d, n,   :169     gsa_sobol_all_y_analysis(method, all_y, d, n, Ei_estimator)
                         Ei_estimator, ...)
QuasiMonteCarlo.generate_design_matrices(n, lb, ub,      generate_design_matrices(n, lb, ub, num_mats=2)
                         SobolSample(), num_mats)

`f` is a DeviceModel (a model evaluated inside the fused CUDA kernel: "ishigami", "sobol_g", "linear", ...) or,
as in the reference, any Python callable: with batch=True it receives the d x cols matrix of all design
points (built on the GPU) and returns a vector / nout x cols matrix; the estimator kernels then run on it.

Shapes follow Julia: designs are (d, N) with one sample per column; results are S1/ST (d,) or (nout, d),
S2 (d, d) or (d, d, nout); the *_Conf_Int fields are None when nboot == 1.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import EI_ESTIMATORS, LOC_DEVICE, LOC_HOST, GsaError, SobolOpts, SobolResultC, check, lib


@dataclass
class Sobol:
    """`struct Sobol <: GSAMethod` (reference :77-83).  Only `2 in order` matters: S1 and ST are always computed."""
    order: Sequence[int] = (0, 1)
    nboot: int = 1
    conf_level: float = 0.95


@dataclass
class SobolResult:
    """`mutable struct SobolResult` (reference :85-92) -- same fields, same order."""
    S1: Any
    S1_Conf_Int: Any
    S2: Any
    S2_Conf_Int: Any
    ST: Any
    ST_Conf_Int: Any


@dataclass
class DeviceModel:
    """A model evaluated on the GPU inside the fused kernel; replaces the Julia closure `f`.

    name    "ishigami" {a, b} | "sobol_g" a[d] | "linear"/"linear_gaussian" W (nout x d) | "ishi_linear"
    params  array of parameters (a 2-D W is flattened column-major, as Julia would store it)
    """
    name: str
    params: Any = field(default_factory=lambda: np.zeros(0))

    def __post_init__(self):
        self.params = np.ascontiguousarray(np.asarray(self.params, dtype=np.float64).ravel(order="F"))
        self.handle = None            # user models live in the handle they were registered with
        if self.name.startswith("user:"):
            return                    # model_id is filled in by from_fatbin
        mid = C.c_int()
        check(lib().gsa_model_lookup(self.name.encode(), C.byref(mid)))
        self.model_id = mid.value

    USER_KINDS = {"generic": 0, "product": 1, "additive": 2}

    @classmethod
    def from_fatbin(cls, image, symbol: str, nout: int = 1, params=(), handle: "Optional[GsaHandle]" = None,
                    kind: str = "generic") -> "DeviceModel":
        """A user-supplied model: `image` (bytes or a path) is a fatbin / cubin / PTX built with
        `nvcc -rdc=true -gencode arch=compute_100a,code=sm_100a` that defines

            kind="generic"   extern "C" __device__ void   <symbol>(const double* x, int d, const double* p, int np, double* y, int nout);
            kind="product"   extern "C" __device__ double <symbol>(int k, double x_k, const double* p, int np);   f(x) = prod_k of it
            kind="additive"  extern "C" __device__ double <symbol>(int k, double x_k, const double* p, int np);   f(x) = sum_k of it

        It is linked into the fused kernels with nvJitLink (gsa_model_register_fatbin / gsa_model_register_separable).  Separable
        kinds run first/total order on the O(d)-per-sample product kernel instead of the O(d^2) generic one."""
        if isinstance(image, (str, bytes)) and not isinstance(image, bytes):
            with open(image, "rb") as f:
                image = f.read()
        h = handle or default_handle()
        m = cls("user:" + symbol, params)
        mid = C.c_int()
        buf = C.create_string_buffer(bytes(image), len(image))
        if kind == "generic":
            check(lib().gsa_model_register_fatbin(h.ptr, buf, len(image), symbol.encode(), int(nout), C.byref(mid)))
        else:
            check(lib().gsa_model_register_separable(h.ptr, buf, len(image), symbol.encode(), cls.USER_KINDS[kind], C.byref(mid)))
        m.model_id, m.handle = mid.value, h
        return m

    def nout(self, d: int) -> int:
        n = C.c_int()
        check(lib().gsa_model_nout(self.handle.ptr if self.handle else None, self.model_id, self.params.size, d, C.byref(n)))
        return n.value


class GsaHandle:
    """Owns a gsa_handle (one CUDA device, its stream and scratch buffers)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        check(lib().gsa_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().gsa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ptr(self):
        return self._h

    def set_stream(self, cuda_stream: Optional[int] = None):
        """Issue the handle's work on an external CUDA stream (an integer cudaStream_t, 0 = legacy default stream);
        None = back to the handle-owned stream."""
        if cuda_stream is None:
            check(lib().gsa_set_stream(self._h, None, 0))
        else:
            check(lib().gsa_set_stream(self._h, C.c_void_p(cuda_stream), 1))

    def synchronize(self):
        check(lib().gsa_synchronize(self._h))

    def last_timing(self):
        ms, kms, n = C.c_float(), C.c_float(), C.c_int()
        check(lib().gsa_last_timing(self._h, C.byref(ms), C.byref(kms), C.byref(n)))
        return ms.value, kms.value, n.value


class _BorrowedHandle:
    """A gsa_handle owned by someone else (a GsaMulti): only its pointer is used (DeviceModel.nout)."""

    def __init__(self, ptr):
        self.ptr = ptr


class PeerGroup:
    """K5 over NVLink peer memory (csrc/peer.cu), one process per GPU: the ranks of a torch.distributed group exchange the
    IPC handles of their mailboxes once; after that `allreduce` is ONE small kernel per rank on the handle's stream -- no
    collective library on the data path, bit-identical sums on every rank, result optionally written straight into pinned
    host memory.

        pg = PeerGroup(handle, max_doubles=nboot * nout * nstat)           # collective: every rank of the group
        p = sobol_partials(model, method, A_local, B_local, N, col0, handle=handle)
        pg.allreduce(p, nstat, host_out=pinned)                            # collective
    """

    HANDLE_BYTES = 64

    def __init__(self, handle: "GsaHandle", max_doubles: int, group=None):
        import os
        import torch
        import torch.distributed as dist
        self.h = handle
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        on_gpu = dist.get_backend(group) == "nccl"

        def all_ok(ok: bool, what: str, err):
            """Every local step is followed by an agreement: either all ranks go on, or all raise -- a rank that failed alone
            must not leave its peers waiting inside the next collective."""
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
            if on_gpu:
                flag = flag.cuda()
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                lib().gsa_peer_disconnect(handle.ptr)
                self.h = None
                raise GsaError(2, f"peer all-reduce unavailable: {what} failed on " + (f"this rank ({err})" if not ok else "another rank"))

        buf = (C.c_ubyte * self.HANDLE_BYTES)()
        err = None
        try:
            if os.environ.get("GSA_PEER_TEST_FAIL") == str(self.rank):      # test hook: this rank cannot export its mailbox
                raise GsaError(2, "GSA_PEER_TEST_FAIL")
            check(lib().gsa_peer_export(handle.ptr, int(max_doubles), C.addressof(buf)))
        except GsaError as e:
            err = e
        all_ok(err is None, "gsa_peer_export", err)
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
        if on_gpu:
            mine = mine.cuda()
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
        self._blob = C.create_string_buffer(blob, len(blob))
        try:
            check(lib().gsa_peer_connect(handle.ptr, self.rank, self.world, C.addressof(self._blob)))
        except GsaError as e:
            err = e
        all_ok(err is None, "gsa_peer_connect", err)      # (also the barrier: every mailbox is mapped before the first exchange)

    def allreduce(self, partials, nstat: int, host_out=None):
        """In-place sum of the CUDA tensor `partials` over the ranks; `host_out` (a pinned CPU tensor of the same size) also
        receives the sums, written by the kernel.  Asynchronous on torch's current stream."""
        hp = host_out.data_ptr() if host_out is not None else None
        with _TorchStream(self.h, True):
            check(lib().gsa_peer_allreduce(self.h.ptr, partials.data_ptr(), partials.numel(), int(nstat), hp))
        return partials

    def status(self) -> int:
        st = C.c_int()
        check(lib().gsa_peer_status(self.h.ptr, C.byref(st)))
        return st.value

    def close(self):
        import torch.distributed as dist
        if getattr(self, "h", None) is not None:
            try:
                if dist.is_initialized():
                    dist.barrier(group=self.group)      # nobody unmaps a mailbox a peer may still write to
                lib().gsa_peer_unmap(self.h.ptr)
                if dist.is_initialized():
                    dist.barrier(group=self.group)      # nobody frees a mailbox a peer still has mapped
            finally:
                lib().gsa_peer_disconnect(self.h.ptr)
                self.h = None


class GsaMulti:
    """Single-process, multi-device form (gsa_multi): one handle per device, the partial sums are gathered by one kernel over
    NVLink peer memory.

        m = GsaMulti([0, 1, 2, 3]);  res = m.gsa(DeviceModel("sobol_g", a), Sobol(), A, B)      # A, B on the host"""

    def __init__(self, devices=None, ndev: Optional[int] = None):
        if devices is None:
            if ndev is None:
                import torch
                ndev = torch.cuda.device_count()
            devices = list(range(ndev))
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        self._m = C.c_void_p()
        check(lib().gsa_multi_create(arr, len(self.devices), C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None) and self._m.value:
            lib().gsa_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def model_from_fatbin(self, image, symbol: str, nout: int = 1, params=(), kind: str = "generic") -> "DeviceModel":
        """A user `__device__` model linked once and loaded on every device of this object (gsa_multi_model_register_*); see
        DeviceModel.from_fatbin for the image format and the kinds."""
        if isinstance(image, str):
            with open(image, "rb") as f:
                image = f.read()
        m = DeviceModel("user:" + symbol, params)
        mid = C.c_int()
        buf = C.create_string_buffer(bytes(image), len(image))
        if kind == "generic":
            check(lib().gsa_multi_model_register_fatbin(self._m, buf, len(image), symbol.encode(), int(nout), C.byref(mid)))
        else:
            check(lib().gsa_multi_model_register_separable(self._m, buf, len(image), symbol.encode(), DeviceModel.USER_KINDS[kind], C.byref(mid)))
        h = C.c_void_p()
        check(lib().gsa_multi_handle(self._m, 0, C.byref(h)))
        m.model_id, m.handle = mid.value, _BorrowedHandle(h)
        return m

    def gsa(self, f: "DeviceModel", method: "Sobol", A, B, *, batch: bool = True, Ei_estimator="Jansen1999") -> "SobolResult":
        Am, Bm = _Mat(A), _Mat(B)
        if Am.loc != LOC_HOST or Bm.loc != LOC_HOST:
            raise ValueError("GsaMulti.gsa takes host designs (each device streams its own column shard)")
        if (Am.d, Am.N) != (Bm.d, Bm.N):
            raise ValueError("A and B must have the same shape")
        d, N = Am.d, Am.N
        nout = f.nout(d)
        o = _opts(method, Ei_estimator, LOC_HOST)
        rc, finish, _ = _alloc_result(nout, d, method.nboot, 2 in method.order, nout > 1)
        check(lib().gsa_multi_sobol_run(self._m, C.byref(o), f.model_id, f.params.ctypes.data, f.params.size, Am.ptr, Bm.ptr,
                                        d, N, C.byref(rc)))
        return finish()

    def gsa_p_range(self, f: "DeviceModel", method: "Sobol", p_range, samples: int, *, Ei_estimator="Jansen1999") -> "SobolResult":
        """gsa(f, Sobol(...), p_range; samples) over all devices with the design generated inside the fused kernel."""
        lb = np.ascontiguousarray([float(p[0]) for p in p_range]); ub = np.ascontiguousarray([float(p[1]) for p in p_range])
        d = lb.size
        nout = f.nout(d)
        o = _opts(method, Ei_estimator, LOC_DEVICE)
        rc, finish, _ = _alloc_result(nout, d, method.nboot, 2 in method.order, nout > 1)
        check(lib().gsa_multi_sobol_run_generated(self._m, C.byref(o), f.model_id, f.params.ctypes.data, f.params.size,
                                                  lb.ctypes.data, ub.ctypes.data, d, int(samples), C.byref(rc)))
        return finish()


def pin_host(a) -> None:
    """cudaHostRegister a NumPy array in place (gsa_host_register): host designs are then streamed at PCIe speed without staging."""
    check(lib().gsa_host_register(a.ctypes.data, a.nbytes))


def unpin_host(a) -> None:
    check(lib().gsa_host_unregister(a.ctypes.data))


_default = {}


def default_handle(device: int = -1) -> GsaHandle:
    if device < 0:
        try:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        except ImportError:
            device = 0
    if device not in _default:
        _default[device] = GsaHandle(device)
    return _default[device]


# ---------------------------------------------------------------------------------------------------
# array plumbing
# ---------------------------------------------------------------------------------------------------
class _TorchStream:
    """Run the handle's work on torch's current CUDA stream while device tensors are involved, so that it is ordered
    with whatever produced them (the handle-owned stream is non-blocking and would not wait for torch's kernels)."""

    def __init__(self, h: "GsaHandle", active: bool):
        self.h, self.active = h, active

    def __enter__(self):
        if self.active:
            import torch
            self.h.set_stream(torch.cuda.current_stream().cuda_stream)
        return self

    def __exit__(self, *exc):
        if self.active:
            self.h.set_stream(None)
        return False


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Mat:
    """A d x N design in Julia layout (the d values of a sample contiguous): pointer + location."""

    def __init__(self, x):
        if _is_torch(x):
            import torch
            if x.dim() != 2:
                raise ValueError("design matrices must be 2-D (d, N)")
            t = x.detach()
            if t.dtype != torch.float64:
                t = t.to(torch.float64)
            tt = t.t()                         # (N, d) view
            if not tt.is_contiguous():
                tt = tt.contiguous()
            self.keep = tt
            self.d, self.N = int(x.shape[0]), int(x.shape[1])
            self.ptr = tt.data_ptr()
            self.loc = LOC_DEVICE if tt.is_cuda else LOC_HOST
            self.device = tt.device.index if tt.is_cuda else None
        else:
            a = np.asarray(x)
            if a.ndim != 2:
                raise ValueError("design matrices must be 2-D (d, N)")
            a = np.asfortranarray(a, dtype=np.float64)
            self.keep = a
            self.d, self.N = a.shape
            self.ptr = a.ctypes.data
            self.loc = LOC_HOST
            self.device = None


def _opts(method: Sobol, Ei_estimator, loc: int) -> SobolOpts:
    est = Ei_estimator.lstrip(":") if isinstance(Ei_estimator, str) else Ei_estimator
    if est not in EI_ESTIMATORS:
        # the reference fails late with a BoundsError on an unknown symbol (:202-236); we fail early
        raise ValueError(f"unknown Ei_estimator {Ei_estimator!r}; expected one of {sorted(EI_ESTIMATORS)}")
    return SobolOpts(int(2 in method.order), int(method.nboot), float(method.conf_level), EI_ESTIMATORS[est], loc)


def _alloc_result(nout: int, d: int, nboot: int, second: bool, multi: bool):
    z = lambda *s: np.zeros(s, dtype=np.float64, order="F")
    s1, st = z(nout, d), z(nout, d)
    s1c, stc = (z(nout, d), z(nout, d)) if nboot > 1 else (None, None)
    s2 = z(d, d, nout) if second else None
    s2c = z(d, d, nout) if second and nboot > 1 else None
    p = lambda a: None if a is None else a.ctypes.data_as(_lib._dp)
    rc = SobolResultC(p(s1), p(s1c), p(s2), p(s2c), p(st), p(stc))

    def finish() -> SobolResult:
        v = (lambda a: None if a is None else (a if multi else a[0].copy()))
        m = (lambda a: None if a is None else (a if multi else a[:, :, 0].copy()))
        return SobolResult(v(s1), v(s1c), m(s2), m(s2c), v(st), v(stc))

    return rc, finish, (s1, s1c, s2, s2c, st, stc)


# ---------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------
def generate_design_matrices(n: int, lb, ub, num_mats: int = 2, *, device: bool = False, col0: int = 0,
                             ncols: Optional[int] = None, handle: Optional[GsaHandle] = None):
    """QuasiMonteCarlo.generate_design_matrices(n, lb, ub, SobolSample(), num_mats) on the GPU, bit-exact.

    Returns `num_mats` matrices of shape (d, ncols): NumPy (Fortran order) by default, or, with device=True,
    CUDA torch tensors that view the Julia layout (strides (1, d)).  `col0`/`ncols` select a column range
    (used to generate one rank's shard in place)."""
    h = handle or default_handle()
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    if lb.shape != ub.shape or lb.ndim != 1:
        raise ValueError("lb and ub must be vectors of equal length")
    d = lb.size
    ncols = n - col0 if ncols is None else ncols
    ptrs = (C.c_void_p * num_mats)()
    if device:
        import torch
        mats = [torch.empty((ncols, d), dtype=torch.float64, device="cuda") for _ in range(num_mats)]
        for m in range(num_mats):
            ptrs[m] = mats[m].data_ptr()
        with _TorchStream(h, True):
            check(lib().gsa_sobol_design(h.ptr, n, d, num_mats, lb.ctypes.data, ub.ctypes.data, col0, ncols, ptrs, LOC_DEVICE))
        return [m.t() for m in mats]
    mats = [np.empty((d, ncols), dtype=np.float64, order="F") for _ in range(num_mats)]
    for m in range(num_mats):
        ptrs[m] = mats[m].ctypes.data
    check(lib().gsa_sobol_design(h.ptr, n, d, num_mats, lb.ctypes.data, ub.ctypes.data, col0, ncols, ptrs, LOC_HOST))
    return mats


# ---------------------------------------------------------------------------------------------------
# K2..K4
# ---------------------------------------------------------------------------------------------------
def gsa_sobol_all_y_analysis(method: Sobol, all_y, d: int, n: int, Ei_estimator="Jansen1999", *,
                             handle: Optional[GsaHandle] = None) -> SobolResult:
    """gsa_sobol_all_y_analysis (reference :169-405) on the GPU.  `all_y`: vector (scalar model) or nout x cols matrix
    (NumPy or torch, host or device), blocks per replicate [A | B | AB_1..AB_d (| BA_1..BA_d)]."""
    h = handle or default_handle()
    second = 2 in method.order
    step = 2 * d + 2 if second else d + 2
    if _is_torch(all_y):
        import torch
        t = all_y.detach().to(torch.float64)
        multi = t.dim() == 2
        tt = (t.t().contiguous() if multi else t.contiguous())      # (cols, nout): Julia layout
        nout = int(t.shape[0]) if multi else 1
        cols = int(t.shape[-1])
        ptr, loc, keep = tt.data_ptr(), (LOC_DEVICE if tt.is_cuda else LOC_HOST), tt
    else:
        a = np.asarray(all_y, dtype=np.float64)
        multi = a.ndim == 2
        a = np.asfortranarray(a) if multi else np.ascontiguousarray(a)
        nout = a.shape[0] if multi else 1
        cols = a.shape[-1]
        ptr, loc, keep = a.ctypes.data, LOC_HOST, a
    if cols != method.nboot * n * step:
        raise ValueError(f"all_y has {cols} columns, expected nboot*n*step = {method.nboot * n * step}")
    o = _opts(method, Ei_estimator, loc)
    rc, finish, _ = _alloc_result(nout, d, method.nboot, second, multi)
    with _TorchStream(h, loc == LOC_DEVICE):
        check(lib().gsa_sobol_analyze_y(h.ptr, C.byref(o), ptr, nout, d, n, C.byref(rc)))
    del keep
    return finish()


def _all_points(h: GsaHandle, method: Sobol, A: _Mat, B: _Mat, want_torch: bool):
    second = 2 in method.order
    n = A.N // method.nboot
    cols = method.nboot * n * (2 * A.d + 2 if second else A.d + 2)
    o = _opts(method, "Jansen1999", A.loc)
    if want_torch:
        import torch
        P = torch.empty((cols, A.d), dtype=torch.float64, device="cuda")
        with _TorchStream(h, True):
            check(lib().gsa_sobol_all_points(h.ptr, C.byref(o), A.ptr, B.ptr, A.d, A.N, P.data_ptr(), LOC_DEVICE))
        return P.t()
    P = np.empty((A.d, cols), dtype=np.float64, order="F")
    check(lib().gsa_sobol_all_points(h.ptr, C.byref(o), A.ptr, B.ptr, A.d, A.N, P.ctypes.data, LOC_HOST))
    return P


def gsa(f, method: Sobol, A, B=None, *, batch: bool = False, Ei_estimator="Jansen1999", samples: Optional[int] = None,
        handle: Optional[GsaHandle] = None, fused_design: bool = True) -> SobolResult:
    """gsa(f, method::Sobol, A, B; batch, Ei_estimator) (reference :110-168), or, when `B is None` and `samples` is
    given, gsa(f, method::Sobol, p_range; samples) (:407-417) with the design generated on the GPU."""
    if not isinstance(method, Sobol):
        raise TypeError("only the Sobol method is implemented by this package")
    h = handle or getattr(f, "handle", None) or default_handle()
    if B is None:
        if samples is None:
            raise TypeError("gsa(f, Sobol(), p_range; samples) needs `samples`")
        p_range = A
        lb = [float(p[0]) for p in p_range]
        ub = [float(p[1]) for p in p_range]
        dev = isinstance(f, DeviceModel)
        if dev and fused_design:
            # fused design generation: the kernel generates the points it evaluates (no A/B in memory at all)
            d = len(lb)
            nout = f.nout(d)
            o = _opts(method, Ei_estimator, LOC_HOST)
            rc, finish, _ = _alloc_result(nout, d, method.nboot, 2 in method.order, nout > 1)
            lbv, ubv = np.ascontiguousarray(lb, dtype=np.float64), np.ascontiguousarray(ub, dtype=np.float64)
            code = lib().gsa_sobol_run_generated(h.ptr, C.byref(o), f.model_id, f.params.ctypes.data, f.params.size,
                                                 lbv.ctypes.data, ubv.ctypes.data, d, int(samples), C.byref(rc))
            if code == _lib.GSA_OK:
                return finish()
            if code != 3:            # GSA_ERR_UNSUPPORTED: no generating kernel for this model/options -> materialise the design
                check(code)
        AB = generate_design_matrices(samples, lb, ub, 2 * method.nboot, device=dev, handle=h)     # :408-413
        if dev:
            import torch
            A = torch.cat([m.t() for m in AB[:method.nboot]], dim=0).t()                           # :414-415
            B = torch.cat([m.t() for m in AB[method.nboot:]], dim=0).t()
        else:
            A = np.concatenate(AB[:method.nboot], axis=1)
            B = np.concatenate(AB[method.nboot:], axis=1)
        return gsa(f, method, A, B, batch=batch, Ei_estimator=Ei_estimator, handle=h)

    Am, Bm = _Mat(A), _Mat(B)
    if (Am.d, Am.N) != (Bm.d, Bm.N):
        raise ValueError("A and B must have the same shape")
    if Am.loc != Bm.loc:
        raise ValueError("A and B must both be on the host or both on the device")
    d, N = Am.d, Am.N
    second = 2 in method.order

    if isinstance(f, DeviceModel):
        nout = f.nout(d)
        o = _opts(method, Ei_estimator, Am.loc)
        rc, finish, _ = _alloc_result(nout, d, method.nboot, second, nout > 1)
        with _TorchStream(h, Am.loc == LOC_DEVICE):
            check(lib().gsa_sobol_run(h.ptr, C.byref(o), f.model_id, f.params.ctypes.data, f.params.size,
                                      Am.ptr, Bm.ptr, d, N, C.byref(rc)))
        return finish()

    if not callable(f):
        raise TypeError("f must be a DeviceModel or a callable")
    # closure path: fused design on the GPU (fuse_designs :94-108), user code evaluates it, estimator kernels reduce it
    n = N // method.nboot
    on_dev = Am.loc == LOC_DEVICE
    P = _all_points(h, method, Am, Bm, on_dev)
    if batch:
        all_y = f(P)                                                                               # :143
    else:
        cols = P.shape[1]
        ys = [np.asarray(f(P[:, i].cpu().numpy() if on_dev else P[:, i]), dtype=np.float64) for i in range(cols)]   # :151
        all_y = np.stack(ys, axis=-1) if ys and ys[0].ndim else np.asarray(ys, dtype=np.float64)
        if all_y.ndim > 2:
            all_y = all_y.reshape(-1, cols)
    return gsa_sobol_all_y_analysis(method, all_y, d, n, Ei_estimator, handle=h)


# ---------------------------------------------------------------------------------------------------
# sharded form (one process per GPU): local partial sums -> all-reduce -> finalize
# ---------------------------------------------------------------------------------------------------
def sample(n: int, lb, ub, *, device: bool = False, handle: Optional["GsaHandle"] = None):
    """QuasiMonteCarlo.sample(n, lb, ub, SobolSample()): the d x n unrandomised Sobol design the reference's other methods draw
    (call sites src/regression_sensitivity.jl:151, src/delta_sensitivity.jl:170, src/easi_sensitivity.jl:164,
    src/mutual_information_sensitivity.jl:140) -- K1 with a single matrix, bit-exact."""
    return generate_design_matrices(n, lb, ub, 1, device=device, handle=handle)[0]


def shard_columns(N: int, rank: int, world: int):
    """Contiguous column range [col0, col0+ncols) of rank `rank`."""
    base, rem = divmod(N, world)
    col0 = rank * base + min(rank, rem)
    return col0, base + (1 if rank < rem else 0)


def nstat(method: Sobol, d: int, Ei_estimator="Jansen1999") -> int:
    o = _opts(method, Ei_estimator, LOC_HOST)
    k = C.c_int()
    check(lib().gsa_sobol_nstat(C.byref(o), d, C.byref(k)))
    return k.value


def sobol_partials(f: DeviceModel, method: Sobol, A_local, B_local, N: int, col0: int, *, Ei_estimator="Jansen1999",
                   handle: Optional[GsaHandle] = None, out=None):
    """Per-replicate partial sums of this rank's column shard, as a CUDA torch tensor (nboot, nout, nstat).
    The sums are additive: `torch.distributed.all_reduce(partials)` then `sobol_finalize`.  Runs on torch's
    current CUDA stream; asynchronous."""
    import torch
    h = handle or getattr(f, "handle", None) or default_handle()
    Am, Bm = _Mat(A_local), _Mat(B_local)
    d = Am.d
    nout = f.nout(d)
    o = _opts(method, Ei_estimator, Am.loc)
    ns = nstat(method, d, Ei_estimator)
    if out is None:
        out = torch.empty((method.nboot, nout, ns), dtype=torch.float64, device="cuda")
    with _TorchStream(h, True):
        check(lib().gsa_sobol_partials(h.ptr, C.byref(o), f.model_id, f.params.ctypes.data, f.params.size, Am.ptr, Bm.ptr,
                                       d, N, col0, Am.N, out.data_ptr()))
    return out


def sobol_partials_generated(f: DeviceModel, method: Sobol, lb, ub, samples: int, col0: int, ncols: int, *,
                             Ei_estimator="Jansen1999", handle: Optional[GsaHandle] = None, out=None):
    """Like sobol_partials, but the shard's design is generated inside the fused kernel (gsa_sobol_partials_generated):
    global columns [col0, col0+ncols) of the nboot*samples-column design of gsa(f, ::Sobol, p_range; samples)."""
    import torch
    h = handle or getattr(f, "handle", None) or default_handle()
    lbv, ubv = np.ascontiguousarray(lb, dtype=np.float64), np.ascontiguousarray(ub, dtype=np.float64)
    d = lbv.size
    nout = f.nout(d)
    o = _opts(method, Ei_estimator, LOC_DEVICE)
    if out is None:
        out = torch.empty((method.nboot, nout, nstat(method, d, Ei_estimator)), dtype=torch.float64, device="cuda")
    with _TorchStream(h, True):
        check(lib().gsa_sobol_partials_generated(h.ptr, C.byref(o), f.model_id, f.params.ctypes.data, f.params.size,
                                                 lbv.ctypes.data, ubv.ctypes.data, d, int(samples), col0, ncols, out.data_ptr()))
    return out


def sobol_finalize(method: Sobol, partials, d: int, N: int, *, Ei_estimator="Jansen1999") -> SobolResult:
    """(:318-404) summed partials (nboot, nout, nstat) -> SobolResult.  Pure host code."""
    if _is_torch(partials):
        partials = partials.detach().cpu().numpy()
    p = np.ascontiguousarray(partials, dtype=np.float64)
    nboot, nout, _ = p.shape
    o = _opts(method, Ei_estimator, LOC_HOST)
    rc, finish, _ = _alloc_result(nout, d, nboot, 2 in method.order, nout > 1)
    check(lib().gsa_sobol_finalize(C.byref(o), p.ctypes.data, nout, d, N // nboot, C.byref(rc)))
    return finish()


def reduce_and_finalize(method: Sobol, partials, d: int, N: int, *, Ei_estimator="Jansen1999", group=None) -> SobolResult:
    """K5 through torch.distributed (NCCL on GPUs; gloo in the CPU tests): the per-rank partial sums (nboot, nout, nstat) -- a
    few KB -- are all-gathered and added IN RANK ORDER by gsa_sobol_combine, which adds the (hi, lo) pairs of the
    double-double entries error-free (a plain all_reduce(SUM) would add hi and lo words independently and lose the
    mean^2 >> var robustness the single-GPU path has); then the host epilogue.  Every rank returns the same SobolResult."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        p = partials.contiguous()
        parts = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(parts, p, group=group)
        stacked = np.ascontiguousarray(torch.stack(parts).cpu().numpy(), dtype=np.float64)
        out = np.empty(p.shape, dtype=np.float64)
        check(lib().gsa_sobol_combine(stacked.ctypes.data, world, int(p.numel()), int(p.shape[-1]), out.ctypes.data))
        partials = out
    return sobol_finalize(method, partials, d, N, Ei_estimator=Ei_estimator)


class ShardedStep:
    """A prepared `gsa_sobol_run_sharded` call (one process per GPU): everything that does not change from step to step --
    options, pointers, result buffers -- is built once, `run()` is ONE ctypes call that launches the fused kernel on this rank's
    shard, all-reduces the partial sums over the handle's peer group (for product-form models inside the same kernel: its last
    CTA pushes them over NVLink) and runs the epilogue.  Without a connected peer group the shard is the whole design.

        pg = PeerGroup(handle, nboot * nout * nstat)              # once, collective
        step = ShardedStep(model, method, A_local, B_local, N, col0, handle=handle)
        res = step.run()                                          # collective: every rank, same arguments

    `p_range=(lb, ub)` instead of A_local/B_local: the shard's points are generated inside the kernel (columns
    [col0, col0 + ncols) of the nboot*samples-column design of gsa(f, ::Sobol, p_range; samples), N = nboot*samples)."""

    def __init__(self, f: DeviceModel, method: Sobol, A_local, B_local, N: int, col0: int, *, Ei_estimator="Jansen1999",
                 handle: Optional[GsaHandle] = None, p_range=None, ncols: Optional[int] = None, torch_stream: bool = True):
        self.h = handle or getattr(f, "handle", None) or default_handle()
        self.f, self.method = f, method
        if p_range is None:
            self.Am, self.Bm = _Mat(A_local), _Mat(B_local)
            if (self.Am.d, self.Am.N) != (self.Bm.d, self.Bm.N) or self.Am.loc != self.Bm.loc:
                raise ValueError("A_local and B_local must have the same shape and location")
            d, ncols, loc = self.Am.d, self.Am.N, self.Am.loc
        else:
            self.lb = np.ascontiguousarray(p_range[0], dtype=np.float64)
            self.ub = np.ascontiguousarray(p_range[1], dtype=np.float64)
            d, loc = self.lb.size, LOC_HOST
            if ncols is None:
                raise ValueError("ncols is required with p_range")
        nout = f.nout(d)
        self.o = _opts(method, Ei_estimator, loc)
        self.rc, self.finish, self._keep = _alloc_result(nout, d, method.nboot, 2 in method.order, nout > 1)
        L = lib()
        if p_range is None:
            self._call = lambda: L.gsa_sobol_run_sharded(self.h.ptr, C.byref(self.o), f.model_id, f.params.ctypes.data, f.params.size,
                                                         self.Am.ptr, self.Bm.ptr, d, N, col0, ncols, C.byref(self.rc))
        else:
            samples = N // method.nboot
            self._call = lambda: L.gsa_sobol_run_sharded_generated(self.h.ptr, C.byref(self.o), f.model_id, f.params.ctypes.data,
                                                                   f.params.size, self.lb.ctypes.data, self.ub.ctypes.data, d,
                                                                   samples, col0, ncols, C.byref(self.rc))
        if torch_stream and (loc == LOC_DEVICE or p_range is not None):
            import torch
            self.h.set_stream(torch.cuda.current_stream().cuda_stream)     # ordered after whatever produced the tensors

    def run(self) -> SobolResult:
        check(self._call())
        return self.finish()


def gsa_sharded(f: DeviceModel, method: Sobol, A_local, B_local, N: int, col0: int, *, Ei_estimator="Jansen1999",
                group=None, handle: Optional[GsaHandle] = None, peers: Optional["PeerGroup"] = None) -> SobolResult:
    """gsa(f, Sobol(...), A, B; batch=true) with the d x N design sharded by contiguous column ranges over the ranks of a
    torch.distributed group (one process per GPU): local fused kernel -> one all-reduce -> epilogue.
    `peers` (a PeerGroup built once on the same handle) runs the whole step as one `gsa_sobol_run_sharded` call: the all-reduce
    goes over NVLink peer memory (csrc/peer.cu, fused into the kernel's tail where possible) instead of torch.distributed."""
    h = handle or (peers.h if peers is not None else None)
    if peers is None:
        p = sobol_partials(f, method, A_local, B_local, N, col0, Ei_estimator=Ei_estimator, handle=h)
        return reduce_and_finalize(method, p, _Mat(A_local).d, N, Ei_estimator=Ei_estimator, group=group)
    step = ShardedStep(f, method, A_local, B_local, N, col0, Ei_estimator=Ei_estimator, handle=h)
    try:
        return step.run()
    finally:
        h.set_stream(None)
