"""ctypes binding of libpnode_cuda.so -- the Python stand-in for the Julia `ccall` host (include/pnode.h).

The library is built in-tree by ``build.py``.  There is no fallback: if the shared object is missing it
is (re)built, and if that fails the import error propagates; solving without a CUDA device raises
``PnodeError`` (status PNODE_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

EK0, EK1, DIAGONAL_EK1 = 0, 1, 2
COV_AUTO, COV_DENSE, COV_KRONECKER, COV_BLOCKS = 0, 1, 2, 3
PRIOR_IWP, PRIOR_IOUP = 0, 1
DIFF_DYNAMIC, DIFF_FIXED, DIFF_DYNAMIC_MV, DIFF_FIXED_MV = 0, 1, 2, 3
RETCODES = {0: "Success", 1: "MaxIters", 2: "DtLessThanMin", 3: "Unstable"}

STATUS = {0: "OK", -1: "ArgumentError", -2: "DimensionMismatch", -3: "CompileError", -4: "CUDAError",
          -5: "Unsupported", -6: "NoDevice"}

# pnode_field
F_T_END, F_U_FINAL, F_U_FINAL_COV, F_X_FINAL, F_X_FINAL_COVFAC, F_LOGLIK, F_DIFFUSION, F_DT_LAST = range(8)
F_NACCEPT, F_NREJECT, F_NF, F_RETCODE, F_STEP_OFFSETS, F_T, F_U, F_U_COV, F_DIFFUSIONS, F_X, F_X_COVFAC = range(8, 19)
F_DIFFUSION_MV, F_DIFFUSIONS_MV, F_SAVEAT_U, F_SAVEAT_U_COV = 19, 20, 21, 22
F_FENRIR_LOGLIK = 23
F_DATA_LOGLIK = 24
_FIELD_DTYPE = {F_NACCEPT: np.int64, F_NREJECT: np.int64, F_NF: np.int64, F_RETCODE: np.int32,
                F_STEP_OFFSETS: np.int64}

EXPORTS = [
    "pnode_device_count", "pnode_last_error", "pnode_abi_version", "pnode_create", "pnode_destroy",
    "pnode_compile_check", "pnode_compile_seconds", "pnode_solve_ensemble", "pnode_solve_ensemble_device",
    "pnode_result_info_get", "pnode_result_field_bytes", "pnode_result_copy", "pnode_result_device_ptr",
    "pnode_result_free", "pnode_measure_fp64_peak", "pnode_solve_ensemble_sharded", "pnode_shard_range",
]


class PnodeError(RuntimeError):
    def __init__(self, status, msg):
        self.status = status
        self.kind = STATUS.get(status, str(status))
        super().__init__(f"{self.kind}: {msg}")


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("d", C.c_int32), ("q", C.c_int32), ("n_p", C.c_int32),
        ("alg", C.c_int32), ("cov", C.c_int32), ("prior", C.c_int32), ("diffusion", C.c_int32),
        ("calibrate", C.c_int32), ("smooth", C.c_int32), ("adaptive", C.c_int32), ("save_everystep", C.c_int32),
        ("save_state", C.c_int32), ("lanes_per_traj", C.c_int32), ("reserved0", C.c_int32),
        ("maxiters", C.c_int64), ("initial_diffusion", C.c_double),
        ("t0", C.c_double), ("t1", C.c_double), ("dt", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double),
        ("beta1", C.c_double), ("beta2", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double),
        ("gamma", C.c_double), ("qsteady_min", C.c_double), ("qsteady_max", C.c_double), ("qoldinit", C.c_double),
        ("rhs_src", C.c_char_p), ("rhs_name", C.c_char_p),
        ("tstops", C.POINTER(C.c_double)), ("n_tstops", C.c_int64),
        ("forced_diffusion", C.POINTER(C.c_double)), ("n_forced", C.c_int64),
        ("rate_parameter", C.c_double),
        ("saveat", C.POINTER(C.c_double)), ("n_saveat", C.c_int64),
        ("mass_matrix", C.POINTER(C.c_double)),
        ("second_order", C.c_int32), ("reserved1", C.c_int32),
        ("data_t", C.POINTER(C.c_double)), ("n_data", C.c_int64), ("data_u", C.POINTER(C.c_double)),
        ("n_obs", C.c_int32), ("data_forward", C.c_int32),
        ("observation_matrix", C.POINTER(C.c_double)), ("observation_noise_cov", C.POINTER(C.c_double)),
        ("observation_noise_var", C.c_double),
        ("pn_observation_noise", C.POINTER(C.c_double)), ("pn_observation_noise_var", C.c_double),
        ("rate_matrix", C.POINTER(C.c_double)), ("update_rate_parameter", C.c_int32), ("reserved2", C.c_int32),
    ]


class ResultInfo(C.Structure):
    _fields_ = [
        ("n_traj", C.c_int64), ("total_saved", C.c_int64), ("total_accepted", C.c_int64),
        ("total_rejected", C.c_int64), ("d", C.c_int32), ("q", C.c_int32), ("D", C.c_int32),
        ("lanes_per_traj", C.c_int32), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("n_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
    ]


_lib = None


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building if necessary) libpnode_cuda.so and declare the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    lib = C.CDLL(path)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    lib.pnode_device_count.restype = C.c_int
    lib.pnode_last_error.restype = C.c_char_p
    lib.pnode_abi_version.restype = C.c_int
    lib.pnode_create.restype = C.c_int
    lib.pnode_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.pnode_destroy.argtypes = [vp]
    lib.pnode_compile_check.restype = C.c_int
    lib.pnode_compile_check.argtypes = [C.POINTER(Config), C.POINTER(C.c_int64)]
    lib.pnode_compile_seconds.restype = C.c_double
    lib.pnode_compile_seconds.argtypes = [vp]
    lib.pnode_solve_ensemble.restype = C.c_int
    lib.pnode_solve_ensemble.argtypes = [vp, C.c_int64, dp, dp, C.POINTER(vp)]
    lib.pnode_solve_ensemble_device.restype = C.c_int
    lib.pnode_solve_ensemble_device.argtypes = [vp, C.c_int64, vp, vp, C.POINTER(vp)]
    lib.pnode_result_info_get.restype = C.c_int
    lib.pnode_result_info_get.argtypes = [vp, C.POINTER(ResultInfo)]
    lib.pnode_result_field_bytes.restype = C.c_int64
    lib.pnode_result_field_bytes.argtypes = [vp, C.c_int]
    lib.pnode_result_copy.restype = C.c_int
    lib.pnode_result_copy.argtypes = [vp, C.c_int, vp, C.c_size_t]
    lib.pnode_result_device_ptr.restype = vp
    lib.pnode_result_device_ptr.argtypes = [vp, C.c_int]
    lib.pnode_result_free.argtypes = [vp]
    lib.pnode_measure_fp64_peak.restype = C.c_int
    lib.pnode_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.pnode_solve_ensemble_sharded.restype = C.c_int
    lib.pnode_solve_ensemble_sharded.argtypes = [C.POINTER(vp), C.c_int32, C.c_int64, dp, dp, C.POINTER(vp)]
    lib.pnode_shard_range.restype = None
    lib.pnode_shard_range.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise PnodeError(rc, load().pnode_last_error().decode(errors="replace"))


def device_count() -> int:
    return load().pnode_device_count()


def make_config(*, d, q, n_p, rhs_name, rhs_src=None, alg=EK1, cov=COV_AUTO, diffusion=DIFF_DYNAMIC, calibrate=True,
                initial_diffusion=1.0, smooth=False, adaptive=True, save_everystep=False, save_state=False,
                lanes_per_traj=0, t0=0.0, t1=1.0, dt=0.0, abstol=1e-6, reltol=1e-3, dtmin=0.0, dtmax=0.0,
                maxiters=1_000_000, beta1=0.0, beta2=0.0, qmin=0.2, qmax=10.0, gamma=0.9, qsteady_min=1.0,
                qsteady_max=1.0, qoldinit=1e-4, tstops=None, forced_diffusion=None, device=0, prior=PRIOR_IWP,
                rate_parameter=0.0, saveat=None, mass_matrix=None, second_order=False, data_t=None, data_u=None,
                observation_matrix=None, observation_noise_cov=None, pn_observation_noise=None, data_forward=False,
                update_rate_parameter=False) -> Config:
    """`rate_parameter`: a number, a length-d vector or a d x d matrix (IOUP, src/priors/ioup.jl:81-101)."""
    c = Config()
    rate_arr = np.asarray(rate_parameter, dtype=np.float64) if rate_parameter is not None else np.asarray(0.0)
    c.update_rate_parameter = int(update_rate_parameter)
    c.second_order = int(second_order)
    c.data_forward = int(data_forward)
    c.struct_size = C.sizeof(Config)
    c.device, c.d, c.q, c.n_p = device, d, q, n_p
    c.alg, c.cov, c.prior, c.diffusion = alg, cov, prior, diffusion
    c.rate_parameter = float(rate_arr) if rate_arr.ndim == 0 else 0.0
    c.calibrate, c.smooth, c.adaptive = int(calibrate), int(smooth), int(adaptive)
    c.save_everystep, c.save_state, c.lanes_per_traj = int(save_everystep), int(save_state), lanes_per_traj
    c.maxiters, c.initial_diffusion = int(maxiters), float(initial_diffusion)
    c.t0, c.t1, c.dt, c.abstol, c.reltol = float(t0), float(t1), float(dt or 0.0), float(abstol), float(reltol)
    c.dtmin, c.dtmax = float(dtmin), float(dtmax)
    c.beta1, c.beta2, c.qmin, c.qmax, c.gamma = beta1, beta2, qmin, qmax, gamma
    c.qsteady_min, c.qsteady_max, c.qoldinit = qsteady_min, qsteady_max, qoldinit
    c.rhs_name = rhs_name.encode()
    c.rhs_src = rhs_src.encode() if rhs_src is not None else None
    keep = []
    if tstops is not None and len(tstops):
        a = np.ascontiguousarray(tstops, dtype=np.float64)
        keep.append(a)
        c.tstops, c.n_tstops = a.ctypes.data_as(C.POINTER(C.c_double)), len(a)
    if forced_diffusion is not None and len(forced_diffusion):
        a = np.ascontiguousarray(forced_diffusion, dtype=np.float64)
        keep.append(a)
        c.forced_diffusion, c.n_forced = a.ctypes.data_as(C.POINTER(C.c_double)), len(a)
    if saveat is not None and len(saveat):
        a = np.ascontiguousarray(saveat, dtype=np.float64)
        keep.append(a)
        c.saveat, c.n_saveat = a.ctypes.data_as(C.POINTER(C.c_double)), len(a)
    if data_t is not None and len(data_t):   # Fenrir data log-likelihood
        dp_ = C.POINTER(C.c_double)
        a = np.ascontiguousarray(data_t, dtype=np.float64)
        b = np.ascontiguousarray(np.asarray(data_u, dtype=np.float64).reshape(len(a), -1))
        keep += [a, b]
        c.data_t, c.n_data, c.data_u, c.n_obs = a.ctypes.data_as(dp_), len(a), b.ctypes.data_as(dp_), b.shape[1]
        if observation_matrix is not None:
            h = np.ascontiguousarray(np.asarray(observation_matrix, dtype=np.float64).reshape(b.shape[1], -1))
            keep.append(h)
            c.observation_matrix = h.ctypes.data_as(dp_)
        cov = np.asarray(observation_noise_cov, dtype=np.float64)
        if cov.ndim == 0:
            c.observation_noise_var = float(cov)
        else:
            cv = np.ascontiguousarray(cov.reshape(b.shape[1], b.shape[1]))
            keep.append(cv)
            c.observation_noise_cov = cv.ctypes.data_as(dp_)
    if pn_observation_noise is not None:
        # a Number / UniformScaling, a Diagonal (1-d array) or a d x d covariance matrix (cov2psdmatrix)
        cov = np.asarray(pn_observation_noise, dtype=np.float64)
        if cov.ndim == 0:
            c.pn_observation_noise_var = float(cov)
        else:
            cv = np.ascontiguousarray(np.diag(cov) if cov.ndim == 1 else cov.reshape(d, d))
            keep.append(cv)
            c.pn_observation_noise = cv.ctypes.data_as(C.POINTER(C.c_double))
    if mass_matrix is not None:
        a = np.ascontiguousarray(np.asarray(mass_matrix, dtype=np.float64).reshape(d, d))
        keep.append(a)
        c.mass_matrix = a.ctypes.data_as(C.POINTER(C.c_double))
    if rate_arr.ndim >= 1:   # vector rate -> Diagonal(r), matrix rate as is (to_sde, ioup.jl:85-93)
        if rate_arr.ndim == 1 and rate_arr.shape[0] != d or rate_arr.ndim == 2 and rate_arr.shape != (d, d):
            raise ValueError("rate_parameter must be a number, a vector of length d or a d x d matrix")
        a = np.ascontiguousarray(np.diag(rate_arr) if rate_arr.ndim == 1 else rate_arr)
        keep.append(a)
        c.rate_matrix = a.ctypes.data_as(C.POINTER(C.c_double))
    c._keep = keep
    return c


def shard_range(n_traj: int, shard: int, n_shards: int):
    """[lo, hi) of `shard` (pnode_shard_range: floor split, remainder to the low shards)."""
    lo, hi = C.c_int64(0), C.c_int64(0)
    load().pnode_shard_range(n_traj, shard, n_shards, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def solve_sharded(solvers, u0: np.ndarray, p: np.ndarray):
    """One ensemble over several devices (pnode_solve_ensemble_sharded): `solvers[i]` -- one per device, same configuration --
    solves the contiguous range shard_range(n, i, len(solvers)) on its own host thread inside the library.  Returns the
    per-shard Results in shard order."""
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    k = len(solvers)
    hs = (C.c_void_p * k)(*[s._h for s in solvers])
    rs = (C.c_void_p * k)()
    dp = C.POINTER(C.c_double)
    _check(load().pnode_solve_ensemble_sharded(hs, k, u0.shape[0], u0.ctypes.data_as(dp), p.ctypes.data_as(dp) if p.size else None, rs))
    return [Result(C.c_void_p(rs[i]), solvers[i]) for i in range(k)]


def measure_fp64_peak(device: int = 0) -> float:
    """TFLOP/s of a DFMA-saturating probe kernel on `device`."""
    v = C.c_double(0.0)
    _check(load().pnode_measure_fp64_peak(device, C.byref(v)))
    return v.value


def compile_check(cfg: Config) -> int:
    """Validate + NVRTC-compile without touching a device; returns the CUBIN size in bytes."""
    n = C.c_int64(0)
    _check(load().pnode_compile_check(C.byref(cfg), C.byref(n)))
    return n.value


class Result:
    def __init__(self, handle, solver):
        self._h = handle
        self._solver = solver  # keep alive
        info = ResultInfo()
        _check(load().pnode_result_info_get(handle, C.byref(info)))
        self.info = info

    def field(self, f) -> np.ndarray:
        lib = load()
        nb = lib.pnode_result_field_bytes(self._h, f)
        if nb == 0:
            raise KeyError(f"field {f} not produced")
        dt = _FIELD_DTYPE.get(f, np.float64)
        out = np.empty(nb // np.dtype(dt).itemsize, dtype=dt)
        _check(lib.pnode_result_copy(self._h, f, out.ctypes.data_as(C.c_void_p), nb))
        return out

    def field_into(self, f, out: np.ndarray) -> np.ndarray:
        """Copy a field into a caller-owned buffer (e.g. pinned host memory: full-rate D2H)."""
        lib = load()
        nb = lib.pnode_result_field_bytes(self._h, f)
        if nb == 0:
            raise KeyError(f"field {f} not produced")
        if out.nbytes != nb or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("destination must be C-contiguous and exactly the size of the field")
        _check(lib.pnode_result_copy(self._h, f, out.ctypes.data_as(C.c_void_p), nb))
        return out

    def has(self, f) -> bool:
        return load().pnode_result_field_bytes(self._h, f) > 0

    def device_ptr(self, f) -> int:
        return load().pnode_result_device_ptr(self._h, f) or 0

    def free(self):
        if self._h:
            load().pnode_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Solver:
    def __init__(self, cfg: Config):
        self.cfg = cfg
        h = C.c_void_p()
        _check(load().pnode_create(C.byref(cfg), C.byref(h)))
        self._h = h

    @property
    def compile_seconds(self) -> float:
        return load().pnode_compile_seconds(self._h)

    def solve(self, u0: np.ndarray, p: np.ndarray) -> Result:
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        n = u0.shape[0]
        r = C.c_void_p()
        dp = C.POINTER(C.c_double)
        _check(load().pnode_solve_ensemble(self._h, n, u0.ctypes.data_as(dp),
                                           p.ctypes.data_as(dp) if p.size else None, C.byref(r)))
        return Result(r, self)

    def solve_device(self, n_traj: int, d_u0: int, d_p: int) -> Result:
        """Inputs already resident on the device (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        r = C.c_void_p()
        _check(load().pnode_solve_ensemble_device(self._h, n_traj, C.c_void_p(d_u0), C.c_void_p(d_p or 0),
                                                  C.byref(r)))
        return Result(r, self)

    def close(self):
        if self._h:
            load().pnode_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
