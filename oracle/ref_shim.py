"""numpy-backed stand-ins for the reference's missing imports  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Lets ``/root/reference/mcmc/image_cupy.py`` (+ ``util_cupy.py``, ``extra_linalg.py``) be imported UNMODIFIED in this
container, so ``oracle/make_golden.py`` can run the reference's own classes (``FourierAnalysis_2D``, ``Lmatrix_2D``,
``pCN``, ``Layer``, ``Simulation``) and commit what they compute as golden vectors.  Only ``oracle/make_golden.py``
imports this module; nothing under ``mcmc-multispde_b200/`` does, and it never runs on the GPU box
(``/root/reference`` does not exist there).

What is substituted (everything else is the reference's own Python, executed as is):

* ``cupy`` / ``cupyx``: a module whose attributes fall through to numpy (the reference uses CuPy as "numpy on the
  device"), plus ``asnumpy``, ``get_array_module`` (returns the shim, so the reference takes its GPU branches, not the
  ``xp == np`` hybrid ones), memory-pool no-ops, ``cupy.prof.TimeRangeDecorator``, ``cupy.random`` (below),
  ``cupy.linalg`` = numpy.linalg (LAPACK where the reference calls cuSOLVER), ``cupyx.scipy.fftpack`` = numpy.fft
  (where the reference calls cuFFT), ``cupyx.scipy.sparse`` = scipy.sparse.
* ``cupy.cuda.cublas.?trsm`` (hand-bound in ``mcmc/extra_linalg.py:74-105``): looked up by array address and
  executed by ``scipy.linalg.solve_triangular``.
* the two ``numba.cuda.jit`` kernels run under numba's CUDA simulator (``NUMBA_ENABLE_CUDASIM=1``), verbatim.
* ``skimage`` (``imread`` / ``resize`` / ``radon`` / ``iradon``): image IO through PIL + a rotation-based Radon
  transform.  They only produce *inputs* of the path (phantom, sinogram ``y``), which the goldens store.
* ``h5py``: a recorder (``RecordingFile``) so the key layout written by ``util._save_object`` can be captured.
* ``matplotlib``: empty stub.  ``builtins.profile``: identity (the kernprof builtin used at ``image_cupy.py:618``).
* ``numpy.find_common_type`` (removed in numpy 2; ``extra_linalg.py:60``).

Two precision modes:

* ``alias64=False`` -- native: ``cp.complex64`` / ``cp.float32`` are the 32-bit types, i.e. the reference's own mixed
  precision.
* ``alias64=True``  -- ``cp.complex64 -> complex128`` and ``cp.float32 -> float64``: the reference's *semantics* in
  FP64, which is what the 1e-10 parity target is stated against.  The module constants ``util.PI`` / ``util.SQRT2``
  stay true float32 values (they are part of the input tables, Appendix A.1), as do the ``.astype('float32')``
  roundings of the hyper-parameters (``image_cupy.py:846-847, 869``).

``cupy.random`` draws from ``numpy.random.Generator(PCG64(seed))`` in the reference's own call order and logs every
draw, so the identical noise can be injected into the oracle and the CUDA path.
"""
from __future__ import annotations

import builtins
import os
import sys
import types

import numpy as np
import scipy.linalg as sla
import scipy.sparse as ssparse

REF = "/root/reference"


class _Ptr:
    def __init__(self, ptr):
        self.ptr = ptr


class _DevArray(np.ndarray):
    """ndarray whose ``.data.ptr`` exists (``extra_linalg.py:102-105`` passes raw device pointers to cuBLAS)."""

    _registry: dict = {}

    @property
    def data(self):  # noqa: D401
        p = self.ctypes.data
        _DevArray._registry[p] = self
        return _Ptr(p)


class RandomLog:
    """``cupy.random`` stand-in: PCG64 stream, every draw logged as (kind, array)."""

    def __init__(self, alias64: bool):
        self.alias64 = alias64
        self.rng = np.random.Generator(np.random.PCG64(0))
        self.log: list = []

    def seed(self, s):
        self.rng = np.random.Generator(np.random.PCG64(int(s)))

    def randn(self, *shape, dtype=np.float64):
        x = self.rng.standard_normal(shape if shape else None)
        self.log.append(("randn", np.array(x, dtype=np.float64, copy=True)))
        return np.asarray(x).astype(dtype) if shape else np.dtype(dtype).type(x)

    def rand(self, *shape, dtype=np.float64):
        x = self.rng.random(shape if shape else None)
        self.log.append(("rand", np.array(x, dtype=np.float64, copy=True)))
        return np.asarray(x).astype(dtype) if shape else np.float64(x)

    def take_log(self):
        out, self.log = self.log, []
        return out


class _Pool:
    def used_bytes(self):
        return 0

    def malloc(self, *a):
        return None

    def free_all_blocks(self):
        return None


class _NpzFile:
    """``cp.load`` result: the reference reads ``data['H']`` and ``data.npz_file.files`` (image_cupy.py:517-522)."""

    def __init__(self, f):
        self.npz_file = f

    def __getitem__(self, k):
        return self.npz_file[k]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.npz_file.close()


class RecordingGroup:
    """Minimal ``h5py`` group that records what ``util._save_object`` (util_cupy.py:553-583) writes."""

    def __init__(self, path=""):
        self.path = path
        self.items: dict = {}

    def create_dataset(self, key, data=None, compression=None):
        a = np.asarray(data)
        self.items[key] = dict(kind="dataset", dtype=str(a.dtype), shape=tuple(a.shape), compression=compression)

    def create_group(self, key):
        g = RecordingGroup(self.path + "/" + key)
        self.items[key] = g
        return g

    def layout(self, prefix=""):
        out = {}
        for k, v in self.items.items():
            if isinstance(v, RecordingGroup):
                out.update(v.layout(prefix + k + "/"))
            else:
                out[prefix + k] = v
        return out


class RecordingFile(RecordingGroup):
    last = None

    def __init__(self, name, mode="r"):
        super().__init__("")
        self.name = name
        RecordingFile.last = self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _radon(image, theta=None, circle=True):
    """Rotation-based Radon transform standing in for ``skimage.transform.radon`` (same output shape
    ``(n_r, n_theta)``; rotate about the centre, sum along axis 0).  Produces an *input* (the sinogram)."""
    import scipy.ndimage as ndi
    image = np.asarray(image, dtype=np.float64)
    theta = np.arange(180) if theta is None else np.asarray(theta)
    m = image.shape[0]
    if circle:
        c = (m - 1) / 2.0
        yy, xx = np.mgrid[:m, :m]
        image = np.where((xx - c) ** 2 + (yy - c) ** 2 <= (m / 2.0) ** 2, image, 0.0)
    out = np.zeros((m, len(theta)))
    for i, a in enumerate(theta):
        out[:, i] = ndi.rotate(image, float(a), reshape=False, order=1).sum(0)
    return out


def _iradon(sino, theta=None, circle=True, **kw):
    """Ramp-filtered back-projection standing in for ``skimage.transform.iradon`` (twoDsim.py:27)."""
    sino = np.asarray(sino, dtype=np.float64)
    m, nt = sino.shape
    theta = np.linspace(0, 180, nt, endpoint=False) if theta is None else np.asarray(theta)
    size = max(64, int(2 ** np.ceil(np.log2(2 * m))))
    f = np.fft.fftfreq(size)
    filt = 2 * np.abs(f)
    proj = np.fft.ifft(np.fft.fft(np.pad(sino, ((0, size - m), (0, 0)), mode="constant"), axis=0) * filt[:, None], axis=0).real[:m]
    c = m // 2
    yy, xx = np.mgrid[:m, :m]
    xx = xx - c
    yy = -(yy - c)
    rec = np.zeros((m, m))
    for i, a in enumerate(np.deg2rad(theta)):
        t = xx * np.cos(a) - yy * np.sin(a)  # skimage: ypr*cos - xpr*sin with its own axes; an input only
        rec += np.interp(t.ravel() + c, np.arange(m), proj[:, i], left=0, right=0).reshape(m, m)
    rec *= np.pi / (2 * nt)
    if circle:
        rec[(xx ** 2 + yy ** 2) > (m / 2.0) ** 2] = 0
    return rec


def _imread(path, as_gray=False):
    from PIL import Image
    p = str(path)
    if not os.path.exists(p):
        p = os.path.join(REF, "phantom_images", os.path.basename(p))
    im = np.asarray(Image.open(p).convert("RGB"), dtype=np.float64) / 255.0
    if as_gray:
        return im @ np.array([0.2125, 0.7154, 0.0721])  # skimage.color.rgb2gray weights
    return im


def _resize(img, shape, anti_aliasing=False, preserve_range=True, order=1, mode="symmetric"):
    """Bilinear resize with pixel-centre alignment (skimage ``resize(order=1)`` convention); an input only."""
    import scipy.ndimage as ndi
    img = np.asarray(img, dtype=np.float64)
    r = (np.arange(shape[0]) + 0.5) * img.shape[0] / shape[0] - 0.5
    c = (np.arange(shape[1]) + 0.5) * img.shape[1] / shape[1] - 0.5
    rr, cc = np.meshgrid(r, c, indexing="ij")
    return ndi.map_coordinates(img, [rr, cc], order=order, mode="reflect")


def install(alias64: bool, seed: int = 1):
    """Builds the stand-in modules, imports the reference and returns a namespace
    ``(im, util, cp, random, RecordingFile)``.  Call once per process and per mode (re-importing with the other mode
    needs a fresh interpreter, because ``cp.complex64`` is bound at class-definition time in a few defaults)."""
    os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
    assert os.path.isdir(REF), "the reference tree is only available in the build container"
    if not hasattr(np, "find_common_type"):  # extra_linalg.py:60
        np.find_common_type = lambda array_types, scalar_types: np.result_type(*array_types, *scalar_types)
    builtins.profile = lambda f: f  # image_cupy.py:618

    rnd = RandomLog(alias64)
    rnd.seed(seed)

    cp = types.ModuleType("cupy")

    def _fallthrough(name):
        return getattr(np, name)
    cp.__getattr__ = _fallthrough
    cp.ndarray = np.ndarray
    cp.complex64 = np.complex128 if alias64 else np.complex64
    cp.float32 = np.float64 if alias64 else np.float32
    cp.asnumpy = lambda a, stream=None, order="C": np.array(a, order=order)
    cp.get_array_module = lambda *a: cp
    cp.get_default_memory_pool = lambda: _Pool()
    cp.get_default_pinned_memory_pool = lambda: _Pool()
    cp._default_memory_pool = _Pool()

    def _array(obj, dtype=None, copy=True, order="K", subok=False, ndmin=0):
        a = np.array(obj, dtype=dtype, order=order, ndmin=ndmin) if copy else np.asarray(obj, dtype=dtype, order=order)
        return a.view(_DevArray)
    cp.array = _array

    def _load(file, **kw):
        return _NpzFile(np.load(file, **kw))
    cp.load = _load

    cp.random = types.ModuleType("cupy.random")
    cp.random.randn, cp.random.rand, cp.random.seed = rnd.randn, rnd.rand, rnd.seed

    cp.prof = types.ModuleType("cupy.prof")

    def TimeRangeDecorator(*a, **k):
        return lambda f: f
    cp.prof.TimeRangeDecorator = TimeRangeDecorator

    cp.linalg = types.ModuleType("cupy.linalg")
    for nm in ("solve", "cholesky", "norm", "slogdet", "qr", "inv", "eigh", "eigvalsh", "svd", "det", "lstsq"):
        setattr(cp.linalg, nm, getattr(np.linalg, nm))
    cp.linalg.util = types.ModuleType("cupy.linalg.util")
    cp.linalg.util._assert_cupy_array = lambda *a: None
    cp.linalg.decomposition = types.ModuleType("cupy.linalg.decomposition")

    cp.fft = np.fft

    cp.core = types.ModuleType("cupy.core")
    cp.core.core = types.ModuleType("cupy.core.core")
    cp.core.core.ndarray = np.ndarray

    cuda = types.ModuleType("cupy.cuda")
    cuda.cusolver_enabled = False
    cuda.malloc_managed = None
    cuda.MemoryPool = lambda *a: _Pool()
    cuda.set_allocator = lambda *a: None
    cublas = types.ModuleType("cupy.cuda.cublas")
    cublas.CUBLAS_FILL_MODE_LOWER, cublas.CUBLAS_FILL_MODE_UPPER = 0, 1
    cublas.CUBLAS_OP_N, cublas.CUBLAS_OP_T, cublas.CUBLAS_OP_C = 0, 1, 2
    cublas.CUBLAS_DIAG_NON_UNIT, cublas.CUBLAS_DIAG_UNIT = 0, 1
    cublas.CUBLAS_SIDE_LEFT, cublas.CUBLAS_SIDE_RIGHT = 0, 1

    def _trsm(handle, side, uplo, trans, diag, m, n, alpha, a_ptr, lda, b_ptr, ldb):
        """cuBLAS ?trsm, side-left, column-major operands addressed by pointer (extra_linalg.py:102-105)."""
        assert side == cublas.CUBLAS_SIDE_LEFT and alpha == 1
        a, b = _DevArray._registry[a_ptr], _DevArray._registry[b_ptr]
        assert a.flags.f_contiguous and (b.ndim == 1 or b.flags.f_contiguous)
        x = sla.solve_triangular(np.asarray(a), np.asarray(b), trans=int(trans), lower=(uplo == cublas.CUBLAS_FILL_MODE_LOWER),
                                 unit_diagonal=(diag == cublas.CUBLAS_DIAG_UNIT))
        b[...] = x
    cublas.strsm = cublas.dtrsm = cublas.ctrsm = cublas.ztrsm = _trsm
    device = types.ModuleType("cupy.cuda.device")
    device.get_cublas_handle = lambda: 0
    cuda.cublas, cuda.device = cublas, device
    cp.cuda = cuda

    cpx = types.ModuleType("cupyx")
    cpx.scipy = types.ModuleType("cupyx.scipy")
    cpx.scipy.sparse = ssparse
    fftpack = types.ModuleType("cupyx.scipy.fftpack")
    fftpack.get_fft_plan = lambda x, shape=None, axes=None, value_type="C2C": None
    fftpack.fft2 = lambda x, shape=None, axes=(-2, -1), overwrite_x=False, plan=None: np.fft.fft2(x).astype(
        np.result_type(x.dtype, np.complex64))
    fftpack.ifft2 = lambda x, shape=None, axes=(-2, -1), overwrite_x=False, plan=None: np.fft.ifft2(x).astype(
        np.result_type(x.dtype, np.complex64))
    cpx.scipy.fftpack = fftpack

    mods = {"cupy": cp, "cupy.random": cp.random, "cupy.prof": cp.prof, "cupy.linalg": cp.linalg,
            "cupy.linalg.util": cp.linalg.util, "cupy.linalg.decomposition": cp.linalg.decomposition,
            "cupy.core": cp.core, "cupy.core.core": cp.core.core, "cupy.cuda": cuda, "cupy.cuda.cublas": cublas,
            "cupy.cuda.device": device, "cupy.fft": np.fft, "cupyx": cpx, "cupyx.scipy": cpx.scipy,
            "cupyx.scipy.sparse": ssparse, "cupyx.scipy.fftpack": fftpack}

    h5 = types.ModuleType("h5py")
    h5.File = RecordingFile
    mods["h5py"] = h5
    sk = types.ModuleType("skimage")
    sk.io = types.ModuleType("skimage.io")
    sk.io.imread = _imread
    sk.transform = types.ModuleType("skimage.transform")
    sk.transform.resize, sk.transform.radon, sk.transform.iradon = _resize, _radon, _iradon
    mods.update({"skimage": sk, "skimage.io": sk.io, "skimage.transform": sk.transform})
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    mods.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})
    for k, v in mods.items():
        sys.modules[k] = v

    if REF not in sys.path:
        sys.path.insert(0, REF)
    import mcmc.image_cupy as im   # the reference, unmodified
    import mcmc.util_cupy as util
    if alias64:
        # input-table constants stay float32 values: util_cupy.py:21-22
        util.PI = np.float32(np.pi)
        util.SQRT2 = np.float32(1.41421356)
    return types.SimpleNamespace(im=im, util=util, cp=cp, random=rnd, RecordingFile=RecordingFile)
