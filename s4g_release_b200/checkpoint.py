"""Checkpoint files in the reference's on-disk format (the format is the contract; reference utils/checkpoint.py:26-63).

A checkpoint is ``torch.save`` of one dict: ``"model"`` -> state_dict under the reference's module-tree keys (this
package mirrors that tree 1:1), optionally ``"optimizer"`` / ``"scheduler"`` state_dicts, and whatever extra entries
the caller adds (epoch, best metric ...).  Beside the ``<name>.pth`` files a one-line text file ``last_checkpoint``
names the newest one.  State dicts written through ``nn.DataParallel`` carry a ``module.`` prefix on every key; it is
removed when reading (reference :81-89).

Implementation notes (ours): files are written to a temporary name and renamed, so an interrupted save never leaves a
truncated ``.pth`` behind the pointer; module functions do the work, ``CheckPointer`` is the reference-shaped handle
(``save`` / ``load`` / ``has_checkpoint`` / ``get_check_point_path``) that scripts written for the reference call.
"""
import logging
import os
import tempfile
from pathlib import Path

import torch

POINTER_NAME = "last_checkpoint"
_DP_PREFIX = "module."


def strip_data_parallel_prefix(state):
    """Keys as an unwrapped model expects them (order preserved)."""
    return {(k[len(_DP_PREFIX):] if k.startswith(_DP_PREFIX) else k): v for k, v in state.items()}


def _atomic_write(path, writer):
    path = Path(path)
    fd, tmp = tempfile.mkstemp(dir=str(path.parent), prefix=path.name + ".", suffix=".tmp")
    os.close(fd)
    try:
        writer(tmp)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def write_checkpoint(directory, name, model, optimizer=None, scheduler=None, **extra):
    """Writes ``<directory>/<name>.pth`` and points ``last_checkpoint`` at it; returns the file path."""
    payload = dict(extra)
    payload["model"] = model.state_dict()
    for key, part in (("optimizer", optimizer), ("scheduler", scheduler)):
        if part is not None:
            payload[key] = part.state_dict()
    target = Path(directory) / (str(name) + ".pth")
    _atomic_write(target, lambda tmp: torch.save(payload, tmp))
    _atomic_write(Path(directory) / POINTER_NAME, lambda tmp: Path(tmp).write_text(str(target)))
    return str(target)


def latest_checkpoint(directory):
    """Path stored in ``<directory>/last_checkpoint`` or None."""
    pointer = Path(directory) / POINTER_NAME if directory else None
    if pointer is None or not pointer.is_file():
        return None
    return pointer.read_text().strip() or None


def read_checkpoint(path, model, optimizer=None, scheduler=None):
    """Loads ``path`` into ``model`` (strict) and, when present on both sides, optimizer / scheduler; returns the
    remaining entries of the file.  The model's cached fused inference engine is dropped: it holds BN-folded copies
    of the old weights."""
    blob = torch.load(path, map_location="cpu")
    model.load_state_dict(strip_data_parallel_prefix(blob.pop("model")), strict=True)
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
    for key, part in (("optimizer", optimizer), ("scheduler", scheduler)):
        if part is not None and key in blob:
            part.load_state_dict(blob.pop(key))
    return blob


class CheckPointer:
    """Reference-shaped handle: ``CheckPointer(model, optimizer, scheduler, save_dir, logger)``."""

    def __init__(self, model, optimizer=None, scheduler=None, save_dir="", logger=None):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.save_dir = save_dir
        self.logger = logger or logging.getLogger(__name__)

    def save(self, name, **kwargs):
        if not self.save_dir:
            self.logger.warning("checkpoint %r not written: this CheckPointer has no save_dir", name)
            return None
        path = write_checkpoint(self.save_dir, name, self.model, self.optimizer, self.scheduler, **kwargs)
        self.logger.info("checkpoint written: %s", path)
        return path

    def has_checkpoint(self):
        return bool(self.save_dir) and (Path(self.save_dir) / POINTER_NAME).exists()

    def get_check_point_path(self):
        found = latest_checkpoint(self.save_dir)
        if found is None:
            self.logger.warning("no readable %s under %r", POINTER_NAME, self.save_dir)
            return ""
        return found

    def load(self, filename=None, resume=True):
        """``resume`` and a ``last_checkpoint`` pointer in save_dir -> that file wins over ``filename``; nothing to
        load -> ``{}`` and the model keeps its initialisation."""
        source = (latest_checkpoint(self.save_dir) if resume else None) or filename
        if not source:
            self.logger.info("nothing to load (no pointer file, no filename): model left as initialised")
            return {}
        self.logger.info("reading checkpoint %s", source)
        return read_checkpoint(source, self.model, self.optimizer, self.scheduler)
