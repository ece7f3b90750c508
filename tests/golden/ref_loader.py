"""Loads pieces of the (Python) reference from /root/reference WITHOUT importing its package
(kornia / timm / jaxtyping are absent here).  Only used by the golden-vector generators in this
directory, which run in the build container; nothing here is read on the GPU box."""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REF = os.environ.get("FREESPLAT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "model", "encoder"))


def _exec_lines(path, first, last, ns):
    src = open(path).read().splitlines()
    # keep the file's line numbers (torch.jit.script re-reads the source through inspect)
    code = "\n" * (first - 1) + "\n".join(src[first - 1:last])
    exec(compile(code, path, "exec"), ns)


def load_cost_volume_module():
    """Returns the reference's cost_volume module (AVGFeatureVolumeManager etc.), unmodified."""
    import torch  # noqa: F401
    # stub package sr_utils.geometry_utils built from the reference's own source lines 11-89
    gu = types.ModuleType("sr_utils.geometry_utils")
    ns = gu.__dict__
    exec("import numpy as np\nimport torch\nimport torch.jit as jit\nimport torch.nn.functional as F\nfrom torch import Tensor\n", ns)
    _exec_lines(os.path.join(REF, "sr_utils", "geometry_utils.py"), 11, 89, ns)
    pkg = types.ModuleType("sr_utils"); pkg.__path__ = []
    pkg.geometry_utils = gu
    # generic_utils.upsample (sr_utils/generic_utils.py:97-106), the one symbol networks.py imports
    ge = types.ModuleType("sr_utils.generic_utils")
    exec("import torch\nfrom torch import nn\n", ge.__dict__)
    _exec_lines(os.path.join(REF, "sr_utils", "generic_utils.py"), 97, 106, ge.__dict__)
    pkg.generic_utils = ge
    sys.modules["sr_utils"] = pkg; sys.modules["sr_utils.geometry_utils"] = gu
    sys.modules["sr_utils.generic_utils"] = ge
    try:
        import torchvision  # noqa: F401
    except Exception:   # networks.py does `from torchvision import models` but never uses it on this path
        tv = types.ModuleType("torchvision"); tv.models = types.ModuleType("torchvision.models")
        sys.modules["torchvision"] = tv; sys.modules["torchvision.models"] = tv.models
    moddir = os.path.join(REF, "src", "model", "encoder", "modules")
    parent = types.ModuleType("refmods"); parent.__path__ = [moddir]
    sys.modules["refmods"] = parent
    for name in ("networks", "cost_volume"):
        # networks imports .layers (plain torch) -- load it too
        pass
    def load(name):
        spec = importlib.util.spec_from_file_location(f"refmods.{name}", os.path.join(moddir, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refmods.{name}"] = m
        spec.loader.exec_module(m)
        return m
    load("layers")
    load("networks")
    return load("cost_volume")


def load_fuse_gaussians():
    """Returns (fuse_gaussians, positional_encoding, GRU) extracted from encoder_freesplat.py /
    networks.py by AST (the module itself needs timm / jaxtyping to import)."""
    import torch
    import torch.nn as nn
    from einops import rearrange, repeat
    path = os.path.join(REF, "src", "model", "encoder", "encoder_freesplat.py")
    tree = ast.parse(open(path).read())
    ns = {"torch": torch, "nn": nn, "rearrange": rearrange, "repeat": repeat}
    wanted = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "positional_encoding":
            wanted.append(node)
        if isinstance(node, ast.ClassDef) and node.name == "EncoderFreeSplat":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "fuse_gaussians":
                    wanted.append(sub)
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), ns)
    load_cost_volume_module()
    GRU = sys.modules["refmods.networks"].GRU
    return ns["fuse_gaussians"], ns["positional_encoding"], GRU


def load_gaussian_adapter():
    """Returns the reference's GaussianAdapter / GaussianAdapterCfg classes
    (src/model/encoder/common/gaussian_adapter.py), with stubs for the two imports this path never calls
    (get_world_rays, rotate_sh: only used when `coords is None`) and for cv2."""
    import types as _t
    root = "refpkg"
    def mod(name, path=None, is_pkg=False):
        m = _t.ModuleType(name)
        if is_pkg:
            m.__path__ = []
        sys.modules[name] = m
        return m
    for pkg in (root, f"{root}.src", f"{root}.src.geometry", f"{root}.src.misc", f"{root}.src.model", f"{root}.src.model.encoder",
                f"{root}.src.model.encoder.common"):
        if pkg not in sys.modules:
            mod(pkg, is_pkg=True)
    proj = mod(f"{root}.src.geometry.projection"); proj.get_world_rays = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("unused"))
    shr = mod(f"{root}.src.misc.sh_rotation"); shr.rotate_sh = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("unused"))
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = _t.ModuleType("cv2")
    base = os.path.join(REF, "src", "model", "encoder", "common")
    for name in ("gaussians", "gaussian_adapter"):
        full = f"{root}.src.model.encoder.common.{name}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(base, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
    ga = sys.modules[f"{root}.src.model.encoder.common.gaussian_adapter"]
    return ga.GaussianAdapter, ga.GaussianAdapterCfg


def load_export_ply():
    """Returns (export_ply, captured): the reference's src/model/ply_export.py::export_ply with stub `plyfile` /
    `jaxtyping` modules (both absent here); the structured vertex array it hands to plyfile lands in captured["elements"]."""
    captured = {}
    if "jaxtyping" not in sys.modules:
        jt = types.ModuleType("jaxtyping")

        class _Ann:
            def __class_getitem__(cls, item):
                return cls
        jt.Float = _Ann
        sys.modules["jaxtyping"] = jt
    pf = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(el, name):
            captured["elements"] = el.copy(); captured["name"] = name
            return (name, el)

    class PlyData:
        def __init__(self, els):
            self.els = els

        def write(self, path):
            captured["path"] = path
    pf.PlyData, pf.PlyElement = PlyData, PlyElement
    sys.modules["plyfile"] = pf
    spec = importlib.util.spec_from_file_location("ref_ply_export", os.path.join(REF, "src", "model", "ply_export.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.export_ply, captured


def _jaxtyping_stub():
    if "jaxtyping" in sys.modules:
        jt = sys.modules["jaxtyping"]
    else:
        jt = types.ModuleType("jaxtyping")
        sys.modules["jaxtyping"] = jt

    class _Ann:
        def __class_getitem__(cls, item):
            return cls
    for n in ("Float", "Bool", "Int64", "Int", "Shaped", "UInt8"):
        if not hasattr(jt, n):
            setattr(jt, n, _Ann)
    return jt


def load_decoder(rasterizer_module):
    """Loads the reference's raster adapter and decoder UNMODIFIED from their files --
    src/model/decoder/{cuda_splatting,decoder,decoder_splatting_cuda}.py, src/geometry/projection.py, src/model/types.py --
    under a private package name, with `rasterizer_module` standing in for the (absent) third-party extension
    `diff_gaussian_rasterization_depth` they import (cuda_splatting.py:5-8).  Returns (render_cuda, DecoderSplattingCUDA,
    Gaussians)."""
    _jaxtyping_stub()
    sys.modules["diff_gaussian_rasterization_depth"] = rasterizer_module
    root = "refdec"
    def pkg(name):
        m = types.ModuleType(name); m.__path__ = []
        sys.modules[name] = m
        return m
    for name in (root, f"{root}.geometry", f"{root}.model", f"{root}.model.decoder", f"{root}.model.encoder",
                 f"{root}.model.encoder.epipolar"):
        pkg(name)
    ds = types.ModuleType(f"{root}.dataset")
    class DatasetCfg:          # src/dataset/__init__.py: only `background_color` is read on this path
        def __init__(self, background_color):
            self.background_color = background_color
    ds.DatasetCfg = DatasetCfg
    sys.modules[f"{root}.dataset"] = ds
    def load(full, rel):
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF, "src", *rel.split("/")))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        return m
    load(f"{root}.geometry.projection", "geometry/projection.py")
    load(f"{root}.model.encoder.epipolar.conversions", "model/encoder/epipolar/conversions.py")
    ty = load(f"{root}.model.types", "model/types.py")
    cs = load(f"{root}.model.decoder.cuda_splatting", "model/decoder/cuda_splatting.py")
    load(f"{root}.model.decoder.decoder", "model/decoder/decoder.py")
    dsc = load(f"{root}.model.decoder.decoder_splatting_cuda", "model/decoder/decoder_splatting_cuda.py")
    return cs.render_cuda, dsc.DecoderSplattingCUDA, ty.Gaussians, DatasetCfg
