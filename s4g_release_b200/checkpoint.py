"""CheckPointer — on-disk compatibility with the reference's checkpoints (utils/checkpoint.py:9-91).

File format (``torch.save`` of a dict): ``"model"`` = the model's state_dict (keys of the reference's module tree,
which this package mirrors 1:1; a ``"module."`` prefix left by nn.DataParallel is stripped on load, :80-90),
optional ``"optimizer"`` / ``"scheduler"`` state_dicts, plus any extra keys the caller passed to ``save``; a text file
``last_checkpoint`` in the save directory holds the path of the newest ``<name>.pth`` and takes precedence on
``load(resume=True)``.  Loading into a model invalidates its fused inference engine (the folded BatchNorm weights are
rebuilt on the next eval forward).
"""
import collections
import logging
import os

import torch


class CheckPointer(object):
    def __init__(self, model, optimizer=None, scheduler=None, save_dir="", logger=None):
        self.model = model
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.save_dir = save_dir
        self.logger = logger if logger is not None else logging.getLogger(__name__)

    def save(self, name, **kwargs):
        if not self.save_dir:
            self.logger.warning("No save directory specified. Can not save check point")
            return
        data = {"model": self.model.state_dict()}
        if self.optimizer is not None:
            data["optimizer"] = self.optimizer.state_dict()
        if self.scheduler is not None:
            data["scheduler"] = self.scheduler.state_dict()
        data.update(kwargs)
        save_file = os.path.join(self.save_dir, "{}.pth".format(name))
        self.logger.info("Saving checkpoint to {}".format(save_file))
        torch.save(data, save_file)
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(save_file)

    def load(self, filename=None, resume=True):
        if resume and self.has_checkpoint():
            filename = self.get_check_point_path()  # an existing checkpoint overrides the argument
        if not filename:
            self.logger.info("No checkpoint found. Initializing model from scratch")
            return {}
        self.logger.info("Loading checkpoint from {}".format(filename))
        checkpoint = torch.load(filename, map_location=torch.device("cpu"))
        self.model.load_state_dict(self._compatible_from_old_version(checkpoint.pop("model")), True)
        if hasattr(self.model, "_engine"):
            self.model._engine = None  # the fused path re-folds BN from the new parameters
        if "optimizer" in checkpoint and self.optimizer:
            self.optimizer.load_state_dict(checkpoint.pop("optimizer"))
        if "scheduler" in checkpoint and self.scheduler:
            self.scheduler.load_state_dict(checkpoint.pop("scheduler"))
        return checkpoint

    def has_checkpoint(self):
        return os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_check_point_path(self):
        save_file = os.path.join(self.save_dir, "last_checkpoint")
        try:
            with open(save_file, "r") as f:
                return f.read().strip()
        except IOError:
            self.logger.warning("Last check point indicator file not exist, please check {}".format(save_file))
            return ""

    @staticmethod
    def _compatible_from_old_version(old_model):
        new_model = collections.OrderedDict()
        for key, value in old_model.items():
            new_model[key[7:] if key.startswith("module.") else key] = value
        return new_model
