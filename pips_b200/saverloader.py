"""Checkpoint save / load with the reference's file naming and call signatures, plus the weight-pack cache.

Mirrors ``saverloader.save`` / ``saverloader.load`` of the reference (saverloader.py:5-23 and :25-67):
checkpoints are ``<ckpt_dir>/<model_name>-<step:09d>.pth`` holding ``model_state_dict`` (and optimizer /
scheduler / EMA state), ``load`` picks the highest step unless one is given, loads with ``strict=False``
and returns the step.  Additions:

  * after the parameters are in place on a CUDA device, ``load`` binds ``<model_name>-<step:09d>.pack``
    when it exists and its fingerprint matches the loaded parameters (pips_b200/pack.py), otherwise it
    packs from the module and writes that file (``write_pack=False`` disables the write);
  * non-checkpoint files in ``ckpt_dir`` (the .pack files) are ignored when scanning for steps, and
    ``save`` prunes a checkpoint's pack together with the checkpoint.
"""
from __future__ import annotations

import os
import struct
import pathlib
import re

import torch

from . import pack as _pack

_CKPT = re.compile(r"^(?P<name>.+)-(?P<step>\d+)\.pth$")


def save(ckpt_dir, optimizer, model, global_step, scheduler=None, model_ema=None, keep_latest=5, model_name="model"):
    os.makedirs(ckpt_dir, exist_ok=True)
    prev = sorted(pathlib.Path(ckpt_dir).glob("%s-*.pth" % model_name), key=lambda p: p.stat().st_mtime, reverse=True)
    for old in prev[max(keep_latest - 1, 0):]:
        old.unlink()
        old.with_suffix(".pack").unlink(missing_ok=True)
    path = "%s/%s-%09d.pth" % (ckpt_dir, model_name, global_step)
    ckpt = {"optimizer_state_dict": optimizer.state_dict(), "model_state_dict": model.state_dict()}
    if scheduler is not None:
        ckpt["scheduler_state_dict"] = scheduler.state_dict()
    if model_ema is not None:
        ckpt["ema_model_state_dict"] = model_ema.state_dict()
    torch.save(ckpt, path)
    pathlib.Path(path).with_suffix(".pack").unlink(missing_ok=True)        # a stale pack must not outlive new weights
    print("saved a checkpoint: %s" % path)
    return path


def _steps(ckpt_dir, model_name):
    out = []
    for f in os.listdir(ckpt_dir):
        m = _CKPT.match(f)
        if m and m.group("name") == model_name:
            out.append(int(m.group("step")))
    return out


def load(ckpt_dir, model, optimizer=None, scheduler=None, model_ema=None, step=0, model_name="model",
         ignore_load=None, write_pack=True):
    print("reading ckpt from %s" % ckpt_dir)
    if not os.path.exists(ckpt_dir):
        print("...there is no full checkpoint here!")
        return step
    steps = _steps(ckpt_dir, model_name)
    if not steps:
        print("...there is no full checkpoint here!")
        return step
    if step == 0:
        step = max(steps)
    path = os.path.join(ckpt_dir, "%s-%09d.pth" % (model_name, step))
    print("...found checkpoint %s" % path)
    dev = next(model.parameters()).device
    checkpoint = torch.load(path, map_location=dev)
    state = checkpoint["model_state_dict"]
    if ignore_load is not None:
        print("ignoring", ignore_load)
        merged = model.state_dict()
        merged.update({k: v for k, v in state.items() if not any(ign in k for ign in ignore_load)})
        state = merged
    model.load_state_dict(state, strict=False)
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    if scheduler is not None:
        scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    if model_ema is not None:
        model_ema.load_state_dict(checkpoint["ema_model_state_dict"])
    if dev.type == "cuda" and hasattr(model, "engine"):
        _bind_pack(model, os.path.splitext(path)[0] + ".pack", write_pack)
    return step


def _bind_pack(model, pack_path, write_pack):
    fp = _pack.fingerprint(model)
    if os.path.exists(pack_path):
        try:
            if _pack.load_pack(model, pack_path, expect_fingerprint=fp):
                print("...bound weight pack %s" % pack_path)
                return
            print("...weight pack %s was made from other parameters; re-packing" % pack_path)
        except (ValueError, KeyError, OSError, RuntimeError, struct.error) as e:   # PipsCudaError is a RuntimeError
            print("...weight pack %s is unreadable (%s); re-packing" % (pack_path, e))
    if write_pack:
        try:
            _pack.save_pack(model, pack_path)
            print("...wrote weight pack %s" % pack_path)
        except OSError as e:                    # read-only checkpoint directory: run from the module's parameters
            print("...could not write %s (%s)" % (pack_path, e))
