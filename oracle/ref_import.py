"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference in the dev container.

The reference (``/root/reference``, alondj/Mesh_R-CNN_Computer_Vision_project) is pure Python / PyTorch 1.2.
It only imports on torch 2.x behind the compatibility shims below.  This module is
used by ``oracle/make_golden.py`` to (i) validate the oracle restatement and (ii) produce the fixtures under
``tests/golden/``.  ``/root/reference`` does not exist on the GPU box: tests never import this file there; the one run-time user is
the reference arm of ``bench.py`` (``--impl reference``), which loads the byte-for-byte copy of the six hot-path files
under ``oracle/_ref`` (see ``oracle/build_ref.py``).  Product code never imports it.

Shims (none of them changes arithmetic):
  * stub packages ``meshRCNN`` / ``utils`` / ``data`` that skip the reference ``__init__`` files (those pull in
    matplotlib and removed torchvision symbols) -- only the hot-path files are loaded:
    ``meshRCNN/layers.py``, ``meshRCNN/loss_functions.py``, ``meshRCNN/utils.py``,
    ``utils/mesh_sampling.py``, ``utils/process.py``, ``utils/rotation.py``.
  * ``torch.symeig`` -> ``torch.linalg.eigh(UPLO='U')`` (removed in torch 2.x; used at loss_functions.py:161).
  * ``Tensor.argsort`` made stable (layers.py:438 relies on the torch-1.2 stable behaviour; with the unstable
    sort of torch >= 1.9 Cubify produces garbage -- with the stable one it reproduces the shipped golden
    ``shapenet_ex/00_mesh_stage0_obj_0.obj`` exactly).
"""
import collections
import importlib.util
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """MESHRCNN_REFERENCE if set; else the reference tree of the dev container; else ``oracle/_ref`` -- the byte-for-byte
    copy of the six hot-path files that ``oracle/build_ref.py`` makes so that bench.py's reference arm can run on the GPU box."""
    env = os.environ.get("MESHRCNN_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/meshRCNN"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "meshRCNN"))


def _stub_package(name: str, path: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load(modname: str, relpath: str):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_LOADED = None


def load_reference():
    """Returns a namespace with the reference hot-path modules: layers, loss_functions, mesh_utils,
    mesh_sampling, process."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    # --- shim: torch.symeig ---------------------------------------------------------------------------
    if not hasattr(torch, "symeig") or getattr(torch.symeig, "_mrb_shim", False) is False:
        _Eig = collections.namedtuple("symeig", ["eigenvalues", "eigenvectors"])

        def _symeig(S, eigenvectors=True, upper=True):
            w, v = torch.linalg.eigh(S, UPLO="U" if upper else "L")
            return _Eig(w, v)

        _symeig._mrb_shim = True
        torch.symeig = _symeig

    # --- shim: stable argsort -------------------------------------------------------------------------
    if not getattr(torch.Tensor.argsort, "_mrb_shim", False):
        _orig_argsort = torch.Tensor.argsort

        def _stable_argsort(self, *a, **k):
            if not a and "stable" not in k:
                k["stable"] = True
            return _orig_argsort(self, *a, **k)

        _stable_argsort._mrb_shim = True
        torch.Tensor.argsort = _stable_argsort

    # --- stub packages (skip reference __init__ files) --------------------------------------------------
    saved = {k: sys.modules.get(k) for k in ("utils", "data", "meshRCNN")}
    utils_pkg = _stub_package("utils", os.path.join(REF_ROOT, "utils"))
    data_pkg = _stub_package("data", os.path.join(REF_ROOT, "data"))
    mesh_pkg = _stub_package("meshRCNN", os.path.join(REF_ROOT, "meshRCNN"))

    class Batch:  # duck-typed stand-in for data/dataloader.py:11-36 (fields used by loss_functions.py:54-55)
        def __init__(self, meshes, vertice_index, face_index):
            self.meshes = meshes
            self.vertice_index = vertice_index
            self.face_index = face_index

    data_pkg.Batch = Batch

    rotation = _load("utils.rotation", "utils/rotation.py")
    process = _load("utils.process", "utils/process.py")
    utils_pkg.rotation = rotation
    utils_pkg.process = process
    utils_pkg.normalize_mesh = process.normalize_mesh
    mesh_sampling = _load("utils.mesh_sampling", "utils/mesh_sampling.py")
    utils_pkg.mesh_sampling = mesh_sampling
    utils_pkg.sample = mesh_sampling.sample
    mesh_utils = _load("meshRCNN.utils", "meshRCNN/utils.py")
    mesh_pkg.utils = mesh_utils
    layers = _load("meshRCNN.layers", "meshRCNN/layers.py")
    loss_functions = _load("meshRCNN.loss_functions", "meshRCNN/loss_functions.py")

    ns = types.SimpleNamespace(layers=layers, loss_functions=loss_functions, mesh_utils=mesh_utils,
                               mesh_sampling=mesh_sampling, process=process, Batch=Batch)
    # leave the stub packages registered under private names only; restore whatever was there before so the
    # repo's own ``utils``-like imports are not shadowed.
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    _LOADED = ns
    return ns
