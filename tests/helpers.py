"""Test-side helpers: ctypes view of the CPU oracle (oracle/liboracle.so), readers for the
reference-harness dump format, scene loaders and synthetic generators.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module; the product
package (fluid_b200/) never does.
"""
import ctypes as C
import hashlib
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

COLLIDE_TRIANGLES, COLLIDE_BOX = 0, 1
SEARCH_BRUTE, SEARCH_GRID = 0, 1
XSPH_JACOBI, XSPH_REFERENCE = 0, 1
ARRAY_XSTAR, ARRAY_LAMBDA, ARRAY_VORTICITY, ARRAY_XPRED = 0, 1, 2, 3


class PbfParams(C.Structure):
    """Mirror of include/pbf_b200.h::PbfParams."""
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


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    return os.path.join(ORACLE_DIR, "liboracle.so")


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("pbf_oracle.cpp", "pbf_oracle.hpp"))
        if not os.path.exists(path) or os.path.getmtime(path) < src_m:
            build_oracle()
        lib = C.CDLL(path)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.POINTER(PbfParams), C.c_int, C.c_int, C.c_int]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.oracle_estimate_densities.argtypes = [C.c_void_p]
        lib.oracle_step.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_download_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_density_at.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.oracle_num_pairs.restype = C.c_size_t
        lib.oracle_num_pairs.argtypes = [C.c_void_p]
        lib.oracle_neighbors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_neighbor_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_default_params.argtypes = [C.POINTER(PbfParams)]
        lib.oracle_set_threads.argtypes = [C.c_int]
        lib.oracle_set_spheres.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_set_triangles.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_max_threads.restype = C.c_int
        lib.oracle_surface.restype = C.c_size_t
        lib.oracle_surface.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_void_p]
        lib.oracle_polygonise_case.restype = C.c_int
        lib.oracle_polygonise_case.argtypes = [C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def default_params(**kw):
    p = PbfParams()
    oracle_lib().oracle_default_params(C.byref(p))
    set_params(p, **kw)
    return p


def set_params(p, **kw):
    for k, v in kw.items():
        if k in ("box_min", "box_max"):
            for a in range(3):
                getattr(p, k)[a] = float(v[a])
        else:
            setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle instance.  precision 32|64; see oracle/pbf_oracle.hpp."""

    def __init__(self, params, precision=64, collision=COLLIDE_BOX, search=SEARCH_GRID):
        self.lib = oracle_lib()
        self.h = self.lib.oracle_create(C.byref(params), precision, collision, search)
        self.n = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    def upload(self, pos, vel):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        self.n = pos.shape[0]
        self.lib.oracle_upload(self.h, self.n, _ptr(pos), _ptr(vel))

    def estimate_densities(self):
        self.lib.oracle_estimate_densities(self.h)

    def set_spheres(self, spheres):
        """Obstacle spheres, rows (cx, cy, cz, r) (reference: StaticScene::Sphere primitives in the BVH)."""
        sp = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        self.lib.oracle_set_spheres(self.h, sp.shape[0], _ptr(sp))

    def set_triangles(self, tris):
        """Obstacle triangles, rows of 18: p1, p2, p3, n1, n2, n3 (MarchingTriangle primitives of the reference's BVH)."""
        t = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 18)
        self.lib.oracle_set_triangles(self.h, t.shape[0], _ptr(t))

    def step(self, steps=1):
        self.lib.oracle_step(self.h, steps)

    def download(self):
        pos = np.empty((self.n, 3)); vel = np.empty((self.n, 3)); rho = np.empty(self.n)
        self.lib.oracle_download(self.h, _ptr(pos), _ptr(vel), _ptr(rho))
        return pos, vel, rho

    def array(self, which):
        out = np.empty(self.n if which == ARRAY_LAMBDA else (self.n, 3))
        self.lib.oracle_download_array(self.h, which, _ptr(out))
        return out

    def density_at(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64); out = np.empty(q.shape[0])
        self.lib.oracle_density_at(self.h, q.shape[0], _ptr(q), _ptr(out))
        return out

    def surface(self, rho0, lo=(-1.0, 0.0, -1.0), hi=(1.0, 1.5, 1.0), iso_ratio=0.95, step=0.3 * 0.5, eps=0.001):
        """Particles::getSurfacePrims as updateSurface calls it (particles.cpp:352-402): rows of 18 (p1 p2 p3 n1 n2 n3);
        fp64 handles only."""
        lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
        nt = self.lib.oracle_surface(self.h, _ptr(lo), _ptr(hi), iso_ratio * rho0, step, eps, 0, None)
        out = np.empty((nt, 18))
        self.lib.oracle_surface(self.h, _ptr(lo), _ptr(hi), iso_ratio * rho0, step, eps, nt, _ptr(out))
        return out

    def neighbors(self):
        m = self.lib.oracle_num_pairs(self.h)
        row = np.empty(self.n + 1, dtype=np.uint32); col = np.empty(m, dtype=np.uint32)
        self.lib.oracle_neighbors(self.h, _ptr(row), _ptr(col))
        return row, col

    def digest(self):
        d = np.empty(self.n, dtype=np.uint64); c = np.empty(self.n, dtype=np.uint32)
        self.lib.oracle_neighbor_digest(self.h, _ptr(d), _ptr(c))
        return d, c

    def stats(self):
        a, b, ms = C.c_double(), C.c_double(), C.c_double()
        self.lib.oracle_stats(self.h, C.byref(a), C.byref(b), C.byref(ms))
        return a.value, b.value, ms.value


def oracle_polygonise_case(cube):
    """marching.cpp polygonise restated, unit cell, corner value 0 where the bit of `cube` is set else 1, iso 0.5."""
    out = np.empty((5, 9))
    nt = oracle_lib().oracle_polygonise_case(int(cube), _ptr(out))
    return out[:nt].copy()


# ---- reference harness -------------------------------------------------------------------------

def ref_harness_path(opt="O3"):
    return os.path.join(ORACLE_DIR, "_ref", "ref_harness" if opt == "O3" else "ref_harness_O0")


def have_reference_binary():
    return os.path.exists(ref_harness_path())


def write_bin_scene(path, pos, vel, rho0):
    with open(path, "wb") as f:
        f.write(np.int64(pos.shape[0]).tobytes())
        f.write(np.float64(rho0).tobytes())
        f.write(np.ascontiguousarray(pos, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(vel, dtype=np.float64).tobytes())


def read_dump(path):
    """-> list of dict(state[N,8], counts[N], row_ptr[N+1], col[M], seconds) per step."""
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"PBFDUMP1"
    n, steps = np.frombuffer(buf, dtype=np.int64, count=2, offset=8)
    off = 24
    out = []
    for _ in range(steps):
        st = np.frombuffer(buf, dtype=np.float64, count=n * 8, offset=off).reshape(n, 8); off += n * 64
        cnt = np.frombuffer(buf, dtype=np.int32, count=n, offset=off); off += n * 4
        m = int(cnt.sum())
        col = np.frombuffer(buf, dtype=np.int32, count=m, offset=off); off += m * 4
        sec = float(np.frombuffer(buf, dtype=np.float64, count=1, offset=off)[0]); off += 8
        row = np.zeros(n + 1, dtype=np.int64); np.cumsum(cnt, out=row[1:])
        out.append(dict(state=st, counts=cnt, row_ptr=row, col=col, seconds=sec))
    return out


def state_sha(pos, vel, rho):
    """sha256[:16] of <f8[N,7] (pos, vel, rho) — the digest format of SURVEY.md §8c."""
    a = np.concatenate([pos, vel, rho[:, None]], axis=1).astype("<f8")
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---- scenes --------------------------------------------------------------------------------------

def load_xml_scene(path):
    """Application::load_particles (application.cpp:302-344): rho0 through float (stof, Q17),
    positions/velocities as doubles in document order."""
    root = ET.parse(path).getroot()
    rho0 = float(np.float32(float(root.find("density").text)))
    pos, vel = [], []
    for p in root.find("ps").findall("particle"):
        pos.append([float(t) for t in p.find("pos").text.split()])
        vel.append([float(t) for t in p.find("v").text.split()])
    return np.array(pos, dtype=np.float64), np.array(vel, dtype=np.float64), rho0


def shipped_scene(name):
    """'p' or 'spheres_p': the reference's particles/*.xml, regenerated from their construction
    rule (so the GPU box, which has no /root/reference, gets the same particles)."""
    if name == "p":
        # particles/p.xml: 12 x 5 x 12 lattice, x,z = -0.9 + 0.15 i, y = 0.5 + 0.2 j, v=(0,-0.01,0)
        raise NotImplementedError
    raise KeyError(name)


def pgen_two_blocks():
    """particles/pgen.py:54-61 == particles/spheres_p.xml (N=2106, rho0=700)."""
    pos = []
    for i in range(1, 10):
        for j in range(1, 14):
            for k in range(1, 10):
                pos.append([0.1 * i - 1, 0.1 * j, 1 - 0.1 * k])
    for i in range(1, 10):
        for j in range(1, 14):
            for k in range(1, 10):
                pos.append([1 - 0.1 * i, 0.1 * j, 0.1 * k - 1])
    pos = np.array(pos, dtype=np.float64)
    vel = np.tile(np.array([0.0, -1.0, 0.0]), (pos.shape[0], 1))
    return pos, vel, 700.0


def lattice_block(nx, ny, nz, origin=(0.1, 0.1, 0.1), spacing=0.1, v0=(0.0, -1.0, 0.0), jitter=0.0, seed=1234):
    """pgen-style block: index order x outer / y / z inner; optional jitter
    U(-jitter, jitter) per coordinate drawn in particle-index order x,y,z from default_rng(seed)."""
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.stack([origin[0] + spacing * i, origin[1] + spacing * j, origin[2] + spacing * k], axis=-1)
    pos = pos.reshape(-1, 3).astype(np.float64)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    vel = np.tile(np.array(v0, dtype=np.float64), (pos.shape[0], 1))
    return pos, vel


# ---- obstacle triangle meshes (rows of 18: p1, p2, p3, n1, n2, n3) ---------------------------------------------

def _quad(a, b, c, d, n):
    a, b, c, d, n = (np.asarray(v, dtype=np.float64) for v in (a, b, c, d, n))
    return [np.concatenate([a, b, c, n, n, n]), np.concatenate([a, c, d, n, n, n])]


def box_mesh(lo, hi, bottom=False):
    """Axis-aligned cuboid, outward face normals, counter-clockwise seen from outside (bottom face optional)."""
    x0, y0, z0 = lo; x1, y1, z1 = hi
    t = []
    t += _quad((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (0, 1, 0))      # top
    t += _quad((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0), (-1, 0, 0))     # x-
    t += _quad((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1), (1, 0, 0))      # x+
    t += _quad((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0), (0, 0, -1))     # z-
    t += _quad((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1), (0, 0, 1))      # z+
    if bottom:
        t += _quad((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1), (0, -1, 0))
    return np.array(t)


def ramp_mesh(x_low, x_high, height, z0, z1, smooth=True):
    """Wedge: slope rising from (x_low, 0) to (x_high, height), vertical back face at x_high, two side triangles.
    smooth=True averages the vertex normals along the top edge (exercises the interpolated, non-unit normal of
    MarchingTriangle::intersect); the second slope triangle is wound clockwise on purpose."""
    sl = np.array([-(height), (x_high - x_low), 0.0]); sl /= np.linalg.norm(sl)      # slope normal (x-, y+)
    bk = np.array([1.0, 0.0, 0.0])
    top = (sl + bk) / 2 if smooth else sl                                             # NOT renormalised
    A, B = np.array([x_low, 0, z0]), np.array([x_low, 0, z1])
    C, D = np.array([x_high, height, z1]), np.array([x_high, height, z0])
    E, F = np.array([x_high, 0, z0]), np.array([x_high, 0, z1])
    t = [np.concatenate([A, B, C, sl, sl, top]), np.concatenate([A, D, C, sl, top, top])]     # second one clockwise
    t += _quad(E, D, C, F, bk)
    t += [np.concatenate([A, D, E, [0, 0, -1] * 3]), np.concatenate([B, F, C, [0, 0, 1] * 3])]
    return np.array(t)


def write_tris(path, tris):
    t = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 18)
    with open(path, "wb") as f:
        f.write(np.int64(t.shape[0]).tobytes()); f.write(t.tobytes())


def uv_sphere_mesh(center, r, nlat=48, nlon=96):
    """Closed latitude/longitude sphere, counter-clockwise seen from outside, smooth (radial, unit) vertex normals:
    2 * nlon * (nlat - 1) triangles.  A stand-in for the reference's large obstacle meshes (dae/sky/CBbunny.dae ...)."""
    c = np.asarray(center, dtype=np.float64)
    th = np.linspace(0.0, np.pi, nlat + 1)
    ph = np.linspace(0.0, 2 * np.pi, nlon + 1)[:-1]

    def nrm(i, j):
        return np.array([np.sin(th[i]) * np.cos(ph[j % nlon]), np.cos(th[i]), np.sin(th[i]) * np.sin(ph[j % nlon])])
    t = []
    for i in range(nlat):
        for j in range(nlon):
            n00, n01, n10, n11 = nrm(i, j), nrm(i, j + 1), nrm(i + 1, j), nrm(i + 1, j + 1)
            if i > 0:
                t.append(np.concatenate([c + r * n00, c + r * n01, c + r * n10, n00, n01, n10]))
            if i < nlat - 1:
                t.append(np.concatenate([c + r * n01, c + r * n11, c + r * n10, n01, n11, n10]))
    return np.array(t)


def heightfield_mesh(x0, x1, z0, z1, nx, nz, base=0.12, amp=0.08, freq=9.0):
    """Bumpy floor y = base + amp sin(freq x) cos(freq z) over [x0,x1] x [z0,z1], upward smooth vertex normals
    (analytic gradient, unit length): 2 * nx * nz triangles."""
    xs = np.linspace(x0, x1, nx + 1); zs = np.linspace(z0, z1, nz + 1)

    def vert(i, k):
        x, z = xs[i], zs[k]
        y = base + amp * np.sin(freq * x) * np.cos(freq * z)
        n = np.array([-amp * freq * np.cos(freq * x) * np.cos(freq * z), 1.0, amp * freq * np.sin(freq * x) * np.sin(freq * z)])
        return np.array([x, y, z]), n / np.linalg.norm(n)
    t = []
    for i in range(nx):
        for k in range(nz):
            (a, na), (b, nb), (c, nc), (d, nd) = vert(i, k), vert(i, k + 1), vert(i + 1, k + 1), vert(i + 1, k)
            t.append(np.concatenate([a, b, c, na, nb, nc]))
            t.append(np.concatenate([a, c, d, na, nc, nd]))
    return np.array(t)
