"""``BiaPy`` facade (``biapy/_biapy.py:107-121, 387-405, 886-1013``) for the two workflows of the hot path, on in-memory arrays.

The reference resolves ``biapy.engine.<problem type>`` and takes the class whose name contains ``Workflow``; the same lookup
runs here against ``biapy_b200.engine``.  File discovery, data generators, logging and result files are out of scope: `train`
and `test` take arrays instead of reading ``DATA.*.PATH``."""
from __future__ import annotations

import copy
import importlib
from typing import Optional

import torch

from .config import load_config

_MODULE_OF = {"SEMANTIC_SEG": "semantic_seg", "DENOISING": "denoising"}


class BiaPy:
    def __init__(self, config, result_dir: str = "", name: str = "unknown_job", run_id: int = 1, gpu: Optional[str] = "0",
                 world_size: int = 1, local_rank: int = -1, dist_on_itp: bool = False, dist_url: str = "env://",
                 dist_backend: str = "nccl", verbose: bool = False, save_files: bool = False, engine_dtype=torch.bfloat16):
        self.cfg = load_config(config)
        self.job_identifier = "{}_{}".format(name, run_id)
        if not torch.cuda.is_available():
            raise RuntimeError("biapy_b200 needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", int(str(gpu).split(",")[0]) if gpu not in (None, "") and local_rank < 0 else max(local_rank, 0))
        ptype = self.cfg.PROBLEM.TYPE
        if ptype not in _MODULE_OF:
            raise NotImplementedError(f"PROBLEM.TYPE={ptype!r} is outside the B200 hot path (supported: {sorted(_MODULE_OF)})")
        mod = importlib.import_module("biapy_b200.engine." + _MODULE_OF[ptype])
        cls = next(getattr(mod, n) for n in dir(mod) if "Workflow" in n and n != "Base_Workflow")     # _biapy.py:396-405
        self.workflow = cls(self.cfg, self.job_identifier, self.device, {"world_size": world_size}, None)
        self.workflow.prepare_model()
        self.workflow.set_engine(engine_dtype)

    def train(self, X, Y, steps: Optional[int] = None):
        """Iterate ``TRAIN.BATCH_SIZE`` batches of ``X, Y`` (``(N, [Z,] Y, X, C)`` arrays) for `steps` iterations (default: one
        pass); returns the per-step losses."""
        bs = int(self.cfg.TRAIN.BATCH_SIZE)
        n = X.shape[0]
        steps = steps if steps is not None else max(1, n // bs)
        losses = []
        for s in range(steps):
            lo = (s * bs) % max(1, n - bs + 1)
            losses.append(self.workflow.train_step(X[lo:lo + bs], Y[lo:lo + bs]))
        return [float(l.item()) for l in losses]

    def predict(self, image, gt=None, return_prediction: bool = True, verbose: bool = False):
        """``image``: one ``([z,] y, x, C)`` array.  Returns the merged prediction (``_biapy.py:1909``)."""
        pred, _ = self.workflow.process_test_sample(image)
        return pred if return_prediction else None

    def test(self, images):
        """Predict a list of images; returns ``[(prediction, post-processed)]``."""
        return [self.workflow.process_test_sample(im) for im in images]

    def run_job(self, train_data=None, test_images=None):
        """``TRAIN.ENABLE`` -> :meth:`train`, ``TEST.ENABLE`` -> :meth:`test` (``_biapy.py:998-1013``), on in-memory data:
        `train_data` = ``(X, Y)``, `test_images` = list of ``([z,] y, x, C)`` arrays.  Returns ``(losses, predictions)``."""
        losses = preds = None
        if self.cfg.TRAIN.ENABLE:
            if train_data is None:
                raise ValueError("TRAIN.ENABLE is set: pass train_data=(X, Y)")
            losses = self.train(*train_data)
        if self.cfg.TEST.ENABLE:
            if test_images is None:
                raise ValueError("TEST.ENABLE is set: pass test_images=[...]")
            preds = self.test(test_images)
        return losses, preds


VALID_WORKFLOWS = ["SEMANTIC_SEG", "INSTANCE_SEG", "CLASSIFICATION", "DETECTION", "DENOISING", "SUPER_RESOLUTION", "SELF_SUPERVISED",
                   "IMAGE_TO_IMAGE"]


def _deep_merge(base: dict, over: dict) -> dict:
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _deep_merge(base[k], v)
        else:
            base[k] = v
    return base


def build_config(workflow: str, dims: str, phase: str = "both", patch_size: Optional[tuple] = None, model: Optional[dict] = None,
                 train_data: Optional[dict] = None, val_data: Optional[dict] = None, test_data: Optional[dict] = None,
                 extra_config: Optional[dict] = None) -> dict:
    """Configuration overrides from high-level arguments, ready for :class:`BiaPy` (reference ``_biapy.py:1995-2085``): same
    arguments, validation messages and resulting dict.  All eight workflow names are accepted here as in the reference; the
    :class:`BiaPy` constructor is what restricts them to the two workflows of the hot path."""
    workflow = str(workflow).upper()
    if workflow not in VALID_WORKFLOWS:
        raise ValueError("'workflow' must be one of {}. Provided: {}".format(VALID_WORKFLOWS, workflow))
    dims = str(dims).upper()
    if dims not in ["2D", "3D"]:
        raise ValueError("'dims' must be either '2D' or '3D'. Provided: {}".format(dims))
    phase = str(phase).lower()
    if phase not in ["train", "test", "both"]:
        raise ValueError("'phase' must be one of ['train', 'test', 'both']. Provided: {}".format(phase))
    upper = lambda d: {str(k).upper(): v for k, v in d.items()}        # noqa: E731
    cfg: dict = {"PROBLEM": {"TYPE": workflow, "NDIM": dims}, "TRAIN": {"ENABLE": phase in ("train", "both")},
                 "TEST": {"ENABLE": phase in ("test", "both")}}
    if patch_size is not None:
        cfg.setdefault("DATA", {})["PATCH_SIZE"] = tuple(patch_size)
    if model:
        cfg["MODEL"] = upper(model)
    for key, part in (("TRAIN", train_data), ("VAL", val_data), ("TEST", test_data)):
        if part:
            cfg.setdefault("DATA", {})[key] = upper(part)
    if extra_config:
        _deep_merge(cfg, copy.deepcopy(extra_config))
    return cfg
