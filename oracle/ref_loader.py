"""TEST INFRASTRUCTURE ONLY -- loads the *real* BiaPy reference modules for the hot path.

This module imports the reference's own files from ``/root/reference`` by file path, behind stub
parent packages, so that the oracle restatement in ``oracle/port_*.py`` can be validated against the
reference itself and golden vectors can be generated (``oracle/make_golden.py``).

It only works in the build container (``/root/reference`` does not exist on the GPU box); nothing in
``biapy_b200`` may import it.  Recipe follows SURVEY.md section 8c:

* stub packages ``biapy``, ``biapy.models``, ``biapy.data``, ``biapy.utils`` and stub modules
  ``h5py``, ``zarr``, ``biapy.utils.misc`` (only ``is_main_process`` is needed by the hot functions);
* one documented patch: ``get_norm_3d/2d('gn', C)`` raise ``TypeError`` in the reference
  (``biapy/models/blocks.py:2124-2125`` and ``:2162-2163`` pass ``num_groups`` twice).  The intended
  semantics ``GroupNorm(8, C)`` (3D) / ``GroupNorm(16, C)`` (2D) are patched in *before* the model files
  bind the names.  Everything else is the unmodified reference code.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BIAPY_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "biapy", "models", "blocks.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


def _load(modname: str, relpath: str) -> types.ModuleType:
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with the reference modules (blocks, unet, resunet, attention_unet, d2, d3, dataset)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch.nn as nn

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "biapy" or k.startswith("biapy.")}
    for k in saved:
        del sys.modules[k]
    try:
        _stub("biapy")
        _stub("biapy.models")
        _stub("biapy.data")
        _stub("biapy.utils")
        _stub("biapy.utils.misc", is_main_process=lambda: True)
        if "h5py" not in sys.modules:
            _stub("h5py", File=type("File", (), {}), Dataset=type("Dataset", (), {}), Group=type("Group", (), {}))
        if "zarr" not in sys.modules:
            _stub("zarr", Array=type("Array", (), {}), Group=type("Group", (), {}))

        blocks = _load("biapy.models.blocks", "biapy/models/blocks.py")

        # --- the one documented patch (reference defect, see module docstring) -------------------------
        _orig3, _orig2 = blocks.get_norm_3d, blocks.get_norm_2d

        def get_norm_3d(norm, out_channels, bn_momentum=0.1):
            if norm == "gn":
                return nn.GroupNorm(8, out_channels)
            return _orig3(norm, out_channels, bn_momentum)

        def get_norm_2d(norm, out_channels, bn_momentum=0.1):
            if norm == "gn":
                return nn.GroupNorm(16, out_channels)
            return _orig2(norm, out_channels, bn_momentum)

        blocks.get_norm_3d, blocks.get_norm_2d = get_norm_3d, get_norm_2d
        # -----------------------------------------------------------------------------------------------
        heads = _load("biapy.models.heads", "biapy/models/heads.py")
        unet = _load("biapy.models.unet", "biapy/models/unet.py")
        resunet = _load("biapy.models.resunet", "biapy/models/resunet.py")
        attention_unet = _load("biapy.models.attention_unet", "biapy/models/attention_unet.py")
        dataset = _load("biapy.data.dataset", "biapy/data/dataset.py")
        d2 = _load("biapy.data.data_2D_manipulation", "biapy/data/data_2D_manipulation.py")
        d3 = _load("biapy.data.data_3D_manipulation", "biapy/data/data_3D_manipulation.py")
        ns = types.SimpleNamespace(
            blocks=blocks, heads=heads, unet=unet, resunet=resunet, attention_unet=attention_unet,
            dataset=dataset, d2=d2, d3=d3,
        )
    finally:
        # do not leave the stub `biapy` package in sys.modules (a real install may coexist)
        for k in [k for k in sys.modules if k == "biapy" or k.startswith("biapy.")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _loaded = ns
    return ns


_chunked = None


def _functions_from_source(relpath: str, names, env: dict) -> dict:
    """Execute only the named top-level functions of a reference file (the file itself imports packages that are not
    installed here: tifffile, imageio, ...).  The code that runs is the reference's own source, read where it lies."""
    import ast
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in body}
    if missing:
        raise RuntimeError(f"{relpath}: functions not found: {sorted(missing)}")
    mod = ast.Module(body=body, type_ignores=[])
    glb = dict(env)
    exec(compile(mod, path, "exec"), glb)
    return {n: glb[n] for n in names}


def load_chunked():
    """The reference's by-chunks generator class (`chunked_test_pair_data_generator`), importable here
    because its grid / extract / insert methods only need numpy: `biapy.data.data_manipulation` is replaced by a stub module
    holding the reference's own `extract_patch_within_image` (AST-extracted), `norm` / `roi_mask` by inert stubs."""
    global _chunked
    if _chunked is not None:
        return _chunked
    ns = load()
    import numpy as np
    from numpy.typing import NDArray
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "biapy" or k.startswith("biapy.")}
    for k in saved:
        del sys.modules[k]
    try:
        _stub("biapy")
        _stub("biapy.data")
        _stub("biapy.utils")
        _stub("biapy.utils.misc", is_main_process=lambda: False, get_world_size=lambda: 1, get_rank=lambda: 0)
        sys.modules["biapy.data.dataset"] = ns.dataset
        sys.modules["biapy.data.data_3D_manipulation"] = ns.d3
        fns = _functions_from_source("biapy/data/data_manipulation.py", ["extract_patch_within_image"],
                                     dict(np=np, NDArray=NDArray, PatchCoords=ns.dataset.PatchCoords))
        _stub("biapy.data.data_manipulation", sample_satisfy_conds=None, save_tif=None, **fns)
        _stub("biapy.data.norm", normalize_image=None, normalize_mask=None)
        _stub("biapy.data.roi_mask", load_roi_mask=None)
        _stub("biapy.data.generators")
        gen = _load("biapy.data.generators.chunked_test_pair_data_generator",
                    "biapy/data/generators/chunked_test_pair_data_generator.py")
    finally:
        for k in [k for k in sys.modules if k == "biapy" or k.startswith("biapy.")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _chunked = types.SimpleNamespace(gen=gen, cls=gen.chunked_test_pair_data_generator, d3=ns.d3,
                                     PatchCoords=ns.dataset.PatchCoords)
    return _chunked


_tta = None


def load_tta():
    """The reference's scalar TTA: `tta.py` is numpy-only and loads as it is; `ensemble_predictions` and its
    three helpers are AST-extracted from `post_processing.py` (whose module imports cv2 / skimage / fill_voids ...), together
    with `to_numpy_format` / `to_pytorch_format` from `biapy/utils/misc.py`."""
    global _tta
    if _tta is not None:
        return _tta
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import math
    import numpy as np
    import torch
    from typing import Callable, Dict, List, Optional, Tuple
    from numpy.typing import NDArray
    tta = _load("_ref_biapy_tta", "biapy/data/post_processing/tta.py")
    sys.modules.pop("_ref_biapy_tta", None)
    misc = _functions_from_source("biapy/utils/misc.py", ["to_numpy_format", "to_pytorch_format"],
                                  dict(torch=torch, np=np, NDArray=NDArray, Tuple=Tuple))
    env = dict(np=np, torch=torch, math=math, NDArray=NDArray, Callable=Callable, Dict=Dict, List=List, Optional=Optional, Tuple=Tuple,
               AxisTransform=tta.AxisTransform, TTASpec=tta.TTASpec, TTA_GROUPS=tta.TTA_GROUPS,
               build_axis_transform_group=tta.build_axis_transform_group, **misc)
    fns = _functions_from_source("biapy/data/post_processing/post_processing.py",
                                 ["_pad_for_orientations", "_crop_padding", "_reduce_orientations", "ensemble_predictions"], env)
    _tta = types.SimpleNamespace(tta=tta, **fns)
    return _tta


_norm = None


def load_norm():
    """The reference's image normalisation functions, AST-extracted from `biapy/data/norm.py` (the module
    imports the dataset machinery); `torch_numpy_dtype_dict` is rebuilt for the dtypes the fixtures use."""
    global _norm
    if _norm is not None:
        return _norm
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np
    import torch
    from typing import Dict, Optional, Tuple
    from numpy.typing import NDArray
    names = ["_is_binary_channel", "normalize_image", "percentile_clip", "torch_percentile", "norm_range01",
             "zero_mean_unit_variance_normalization", "undo_image_norm", "undo_norm_range01",
             "undo_zero_mean_unit_variance_normalization"]
    env = dict(np=np, torch=torch, NDArray=NDArray, Dict=Dict, Optional=Optional, Tuple=Tuple,
               torch_numpy_dtype_dict={"uint8": [torch.uint8, np.uint8], "uint16": [torch.uint16, np.uint16],
                                       "float32": [torch.float32, np.float32]})
    _norm = types.SimpleNamespace(**_functions_from_source("biapy/data/norm.py", names, env))
    return _norm
