"""ctypes binding of libpbf_b200.so (C ABI: include/pbf_b200.h) and a Python mirror of the
reference's `Particles` interface (src/particles.h:105-140) on top of it.

There is no CPU fallback here: if the CUDA library is missing it is built with nvcc, and if no
CUDA device is visible every solver call raises PbfError(PBF_ERR_NO_DEVICE).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

PBF_OK, PBF_ERR_INVALID, PBF_ERR_CUDA, PBF_ERR_NO_DEVICE, PBF_ERR_CAPACITY, PBF_ERR_DOMAIN = range(6)
XSPH_JACOBI, XSPH_REFERENCE_ORDER = 0, 1
ARRAY_XSTAR, ARRAY_LAMBDA, ARRAY_VORTICITY, ARRAY_XPRED = 0, 1, 2, 3
_ERR_NAMES = {1: "PBF_ERR_INVALID", 2: "PBF_ERR_CUDA", 3: "PBF_ERR_NO_DEVICE", 4: "PBF_ERR_CAPACITY", 5: "PBF_ERR_DOMAIN"}


class PbfParams(C.Structure):
    """include/pbf_b200.h::PbfParams (defaults = reference macros, particles.cpp:10-44)."""
    _fields_ = [
        ("h", C.c_double), ("dt", C.c_double), ("rest_density", C.c_double),
        ("eps_relax", C.c_double), ("k_corr", C.c_double), ("dq_ratio", C.c_double),
        ("visc_c", C.c_double), ("vort_eps", C.c_double), ("gravity_y", C.c_double),
        ("n_corr", C.c_int32), ("iterations", C.c_int32),
        ("box_min", C.c_double * 3), ("box_max", C.c_double * 3),
        ("y_light", C.c_double), ("z_front", C.c_double),
        ("xsph_mode", C.c_int32), ("enable_vorticity", C.c_int32), ("enable_xsph", C.c_int32),
        ("reserved", C.c_int32),
    ]


class PbfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library():
    """Loads (building it first if needed) the in-tree CUDA library.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    # PBF_LIB: developer override for A/B runs against another build of the same ABI (never a CPU path)
    path = os.environ.get("PBF_LIB") or _build.build()
    lib = C.CDLL(path)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    sig = {
        "pbf_default_params": (None, [C.POINTER(PbfParams)]),
        "pbf_create": (i32, [C.POINTER(PbfParams), i32, C.POINTER(vp)]),
        "pbf_destroy": (None, [vp]),
        "pbf_last_error": (C.c_char_p, [vp]),
        "pbf_device_count": (i32, []),
        "pbf_upload": (i32, [vp, sz, vp, vp]),
        "pbf_download": (i32, [vp, vp, vp, vp]),
        "pbf_num_particles": (sz, [vp]),
        "pbf_host_register": (i32, [vp, vp, sz]),
        "pbf_host_unregister": (i32, [vp, vp]),
        "pbf_set_readback": (i32, [vp, vp, vp, vp]),
        "pbf_set_obstacle_spheres": (i32, [vp, sz, vp]),
        "pbf_set_obstacle_triangles": (i32, [vp, sz, vp]),
        "pbf_step": (i32, [vp, i32]),
        "pbf_sync": (i32, [vp]),
        "pbf_estimate_densities": (i32, [vp]),
        "pbf_stats": (i32, [vp, vp, vp, vp]),
        "pbf_density_at": (i32, [vp, sz, vp, vp]),
        "pbf_extract_surface": (i32, [vp, vp, vp, C.c_double, C.c_double, C.c_double, sz, vp, vp]),
        "pbf_upload_device": (i32, [vp, sz, vp, vp]),
        "pbf_download_device": (i32, [vp, vp, vp, vp]),
        "pbf_debug_neighbor_digest": (i32, [vp, vp, vp]),
        "pbf_debug_download_neighbors": (i32, [vp, vp, vp, sz]),
        "pbf_debug_download_array": (i32, [vp, i32, vp]),
        "pbf_debug_capture": (i32, [vp, i32]),
        "pbf_launch_count": (u64, [vp]),
        "pbf_profile_enable": (i32, [vp, i32]),
        "pbf_profile_get": (i32, [vp, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def default_params(**kw):
    p = PbfParams()
    load_library().pbf_default_params(C.byref(p))
    for k, v in kw.items():
        if k in ("box_min", "box_max"):
            for a in range(3):
                getattr(p, k)[a] = float(v[a])
        else:
            setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Solver:
    """One pbf_handle: one CUDA device, one stream."""

    def __init__(self, params=None, device=0):
        self.lib = load_library()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        rc = self.lib.pbf_create(C.byref(self.params), device, C.byref(h))
        if rc != PBF_OK:
            raise PbfError(rc, "pbf_create failed (no CUDA device?)" if rc == PBF_ERR_NO_DEVICE else "pbf_create failed")
        self.h = h
        self.n = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.pbf_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != PBF_OK:
            raise PbfError(rc, self.lib.pbf_last_error(self.h).decode())

    def upload(self, pos, vel):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        assert pos.shape == vel.shape and (pos.size == 0 or pos.shape[1] == 3)
        self.n = pos.shape[0]
        self._ck(self.lib.pbf_upload(self.h, self.n, _ptr(pos), _ptr(vel)))

    def pin(self, *arrays):
        """Page-lock numpy arrays the caller will reuse for upload()/download_into() (pbf_host_register)."""
        for a in arrays:
            assert a.flags["C_CONTIGUOUS"]
            self._ck(self.lib.pbf_host_register(self.h, _ptr(a), a.nbytes))

    def unpin(self, *arrays):
        for a in arrays:
            self._ck(self.lib.pbf_host_unregister(self.h, _ptr(a)))

    def set_readback(self, P=None, V=None, R=None):
        """Stream the results of every step() into these pinned arrays (complete after sync())."""
        self._rb = (P, V, R)     # keep them alive
        self._ck(self.lib.pbf_set_readback(self.h, _ptr(P), _ptr(V), _ptr(R)))

    def set_obstacle_spheres(self, spheres):
        """Obstacle spheres of the collision scene, rows (cx, cy, cz, r); an empty list removes them."""
        sp = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        self._ck(self.lib.pbf_set_obstacle_spheres(self.h, sp.shape[0], _ptr(sp)))

    def set_obstacle_triangles(self, tris):
        """Obstacle triangles, rows of 18 (p1, p2, p3, n1, n2, n3); an empty list removes them."""
        t = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 18)
        self._ck(self.lib.pbf_set_obstacle_triangles(self.h, t.shape[0], _ptr(t)))

    def upload_device(self, n, d_pos_ptr, d_vel_ptr):
        self.n = n
        self._ck(self.lib.pbf_upload_device(self.h, n, C.c_void_p(d_pos_ptr), C.c_void_p(d_vel_ptr)))

    def download_device(self, d_pos_ptr=None, d_vel_ptr=None, d_rho_ptr=None):
        self._ck(self.lib.pbf_download_device(self.h, C.c_void_p(d_pos_ptr), C.c_void_p(d_vel_ptr), C.c_void_p(d_rho_ptr)))

    def step(self, n_steps=1, sync=True):
        self._ck(self.lib.pbf_step(self.h, n_steps))
        if sync:
            self.sync()

    def sync(self):
        self._ck(self.lib.pbf_sync(self.h))

    def estimate_densities(self):
        self._ck(self.lib.pbf_estimate_densities(self.h))

    def download(self, pos=True, vel=True, rho=True):
        P = np.empty((self.n, 3)) if pos else None
        V = np.empty((self.n, 3)) if vel else None
        R = np.empty(self.n) if rho else None
        self._ck(self.lib.pbf_download(self.h, _ptr(P), _ptr(V), _ptr(R)))
        return P, V, R

    def download_into(self, P, V, R):
        self._ck(self.lib.pbf_download(self.h, _ptr(P), _ptr(V), _ptr(R)))

    def density_at(self, query):
        """Particles::estimateDensityAt for an [m,3] array of points."""
        q = np.ascontiguousarray(query, dtype=np.float64)
        out = np.empty(q.shape[0])
        self._ck(self.lib.pbf_density_at(self.h, q.shape[0], _ptr(q), _ptr(out)))
        return out

    def extract_surface(self, rho0, lo=(-1.0, 0.0, -1.0), hi=(1.0, 1.5, 1.0), iso_ratio=0.95, step=0.3 * 0.5, eps=0.001):
        """Particles::getSurfacePrims as updateSurface calls it (particles.cpp:352-402; the defaults are the reference's
        hard-coded lattice and macros): [T,18] array, rows p1 p2 p3 n1 n2 n3, in the reference's triangle order."""
        lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
        nt = C.c_size_t(0)
        self._ck(self.lib.pbf_extract_surface(self.h, _ptr(lo), _ptr(hi), iso_ratio * rho0, step, eps, 0, None, C.byref(nt)))
        out = np.empty((nt.value, 18))
        if nt.value:
            self._ck(self.lib.pbf_extract_surface(self.h, _ptr(lo), _ptr(hi), iso_ratio * rho0, step, eps, nt.value, _ptr(out), C.byref(nt)))
        return out

    def stats(self):
        a, b, ms = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.pbf_stats(self.h, C.byref(a), C.byref(b), C.byref(ms)))
        return a.value, b.value, ms.value

    def neighbor_digest(self):
        d = np.empty(self.n, dtype=np.uint64); c = np.empty(self.n, dtype=np.uint32)
        self._ck(self.lib.pbf_debug_neighbor_digest(self.h, _ptr(d), _ptr(c)))
        return d, c

    def neighbors(self):
        row = np.empty(self.n + 1, dtype=np.uint32)
        self._ck(self.lib.pbf_debug_download_neighbors(self.h, _ptr(row), None, 0))
        col = np.empty(int(row[-1]), dtype=np.uint32)
        self._ck(self.lib.pbf_debug_download_neighbors(self.h, _ptr(row), _ptr(col), col.size))
        return row, col

    def array(self, which):
        out = np.empty(self.n if which == ARRAY_LAMBDA else (self.n, 3))
        self._ck(self.lib.pbf_debug_download_array(self.h, which, _ptr(out)))
        return out

    def capture(self, on=True):
        self._ck(self.lib.pbf_debug_capture(self.h, int(on)))

    def launch_count(self):
        return int(self.lib.pbf_launch_count(self.h))

    def profile_enable(self, on=True):
        self._ck(self.lib.pbf_profile_enable(self.h, int(on)))

    def profile(self):
        names = (C.c_char_p * 32)(); ms = (C.c_double * 32)(); ln = (C.c_uint64 * 32)()
        k = self.lib.pbf_profile_get(self.h, 32, names, ms, ln)
        return {names[i].decode(): (ms[i], int(ln[i])) for i in range(k)}


def _bind_multi(lib):
    if getattr(lib, "_multi_bound", False):
        return lib
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    sig = {
        "pbf_create_multi": (i32, [C.POINTER(PbfParams), i32, vp, C.POINTER(vp)]),
        "pbf_multi_destroy": (None, [vp]),
        "pbf_multi_last_error": (C.c_char_p, [vp]),
        "pbf_multi_num_devices": (i32, [vp]),
        "pbf_multi_set_obstacle_spheres": (i32, [vp, sz, vp]),
        "pbf_multi_set_obstacle_triangles": (i32, [vp, sz, vp]),
        "pbf_multi_upload": (i32, [vp, sz, vp, vp]),
        "pbf_multi_step": (i32, [vp, i32]),
        "pbf_multi_sync": (i32, [vp]),
        "pbf_multi_download": (i32, [vp, vp, vp, vp]),
        "pbf_multi_num_particles": (sz, [vp]),
        "pbf_multi_stats": (i32, [vp, vp, vp, vp]),
        "pbf_multi_set_rebalance": (i32, [vp, i32, C.c_double]),
        "pbf_multi_plan": (i32, [vp, vp, vp, vp]),
        "pbf_multi_neighbor_digest": (i32, [vp, vp, vp]),
        "pbf_multi_launch_count": (u64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._multi_bound = True
    return lib


class MultiSolver:
    """One pbf_multi handle (include/pbf_b200_multi.h): x-slabs on several GPUs driven by this process, peer-mode halos.
    `devices` may name a device more than once (several slabs on one GPU: how the 1-GPU test box exercises the path)."""

    def __init__(self, params=None, devices=(0,)):
        self.lib = _bind_multi(load_library())
        self.params = params if params is not None else default_params()
        ids = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.pbf_create_multi(C.byref(self.params), len(devices), ids, C.byref(h))
        if rc != PBF_OK:
            raise PbfError(rc, "pbf_create_multi failed (no CUDA device?)" if rc == PBF_ERR_NO_DEVICE else "pbf_create_multi failed")
        self.h = h
        self.n = 0
        self.world = len(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pbf_multi_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != PBF_OK:
            raise PbfError(rc, self.lib.pbf_multi_last_error(self.h).decode())

    def set_obstacle_spheres(self, spheres):
        sp = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        self._ck(self.lib.pbf_multi_set_obstacle_spheres(self.h, sp.shape[0], _ptr(sp)))

    def set_obstacle_triangles(self, tris):
        t = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 18)
        self._ck(self.lib.pbf_multi_set_obstacle_triangles(self.h, t.shape[0], _ptr(t)))

    def set_rebalance(self, every_k_steps=8, threshold=1.05):
        self._ck(self.lib.pbf_multi_set_rebalance(self.h, every_k_steps, threshold))

    def upload(self, pos, vel):
        pos = np.ascontiguousarray(pos, dtype=np.float64); vel = np.ascontiguousarray(vel, dtype=np.float64)
        self.n = pos.shape[0]
        self._ck(self.lib.pbf_multi_upload(self.h, self.n, _ptr(pos), _ptr(vel)))

    def step(self, n_steps=1, sync=True):
        self._ck(self.lib.pbf_multi_step(self.h, n_steps))
        if sync:
            self.sync()

    def sync(self):
        self._ck(self.lib.pbf_multi_sync(self.h))

    def download(self):
        P = np.empty((self.n, 3)); V = np.empty((self.n, 3)); R = np.empty(self.n)
        self._ck(self.lib.pbf_multi_download(self.h, _ptr(P), _ptr(V), _ptr(R)))
        return P, V, R

    def stats(self):
        a, b, ms = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.pbf_multi_stats(self.h, C.byref(a), C.byref(b), C.byref(ms)))
        return a.value, b.value, ms.value

    def neighbor_digest(self):
        d = np.empty(self.n, dtype=np.uint64); c = np.empty(self.n, dtype=np.uint32)
        self._ck(self.lib.pbf_multi_neighbor_digest(self.h, _ptr(d), _ptr(c)))
        return d, c

    def plan(self):
        """(column boundaries, owned particles per device, number of re-balancing moves so far)"""
        b = np.zeros(self.world + 1, dtype=np.int32); o = np.zeros(self.world, dtype=np.uint64); k = C.c_uint64()
        self._ck(self.lib.pbf_multi_plan(self.h, _ptr(b), _ptr(o), C.byref(k)))
        return b, o, int(k.value)

    def launch_count(self):
        return int(self.lib.pbf_multi_launch_count(self.h))


class Particles:
    """Python mirror of the reference's `struct Particles` (src/particles.h:105-140) backed by the
    GPU solver: same member names, argument meaning and call order as the reference host code
    (Application::load_particles, application.cpp:302-344; PathTracer::fluid_simulate_*,
    pathtracer.cpp:444-480)."""

    DEFAULT_DELTA_T = 0.016   # particles.cpp:24

    def __init__(self, rest_density=1000.0, params=None, device=0):
        self.rest_density = float(rest_density)
        self._params = params if params is not None else default_params()
        self._params.rest_density = self.rest_density
        self._device = device
        self._pos, self._vel = [], []
        self._solver = None
        self.simulate_time = 0.0
        self.surfaceUpToTimestep = False
        self.position = np.zeros((0, 3)); self.velocity = np.zeros((0, 3)); self.density = np.zeros(0)

    def addParticle(self, pos, v):            # particles.h:118-120
        if self._solver is not None:
            raise PbfError(PBF_ERR_INVALID, "addParticle after the first step is not supported")
        self._pos.append(tuple(pos)); self._vel.append(tuple(v))

    def _ensure(self):
        if self._solver is None:
            self._solver = Solver(self._params, self._device)
            self.position = np.array(self._pos, dtype=np.float64).reshape(-1, 3)
            self.velocity = np.array(self._vel, dtype=np.float64).reshape(-1, 3)
            self.density = np.zeros(self.position.shape[0])
            self._solver.upload(self.position, self.velocity)
        return self._solver

    def estimateDensities(self):              # particles.cpp:440-444
        s = self._ensure()
        s.estimate_densities()
        _, _, self.density = s.download(pos=False, vel=False, rho=True)

    def timeStep(self, delta_t=None):         # particles.cpp:250-301
        if delta_t is not None and abs(delta_t - self._params.dt) > 0:
            raise PbfError(PBF_ERR_INVALID, "dt is fixed at creation (PbfParams.dt)")
        s = self._ensure()
        self.simulate_time += self._params.dt
        s.step(1)
        self.position, self.velocity, self.density = s.download()
        self.surfaceUpToTimestep = False
        return s.stats()[:2]                  # the reference prints "avg rho: a => b"

    def getDensityBasedColor(self):           # particles.h:48-52
        ratio = (self.density - 0.8 * self.rest_density) / (0.4 * self.rest_density)
        c = np.clip(ratio, 0.0, 1.0)
        return np.stack([np.ones_like(c), 1.0 - c, 1.0 - c, np.ones_like(c)], axis=1)
