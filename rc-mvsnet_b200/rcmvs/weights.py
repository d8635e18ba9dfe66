"""Checkpoint helpers: the reference's .ckpt ('model' entry) or the numpy fixture made from it."""
import numpy as np
import torch


def load_npz_state_dict(path):
    z = np.load(path)
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def load_model(model, path, strict=True):
    if path.endswith(".npz"):
        sd = load_npz_state_dict(path)
    else:
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("model", sd)
    model.load_state_dict(sd, strict=strict)
    return model
