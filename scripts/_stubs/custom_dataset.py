"""Synthetic stand-in for the reference's custom_dataset.Dataset_ (custom_dataset.py:10-100): same
constructor and item contract - (image, geometry_change, appearance_change) float tensors in [-1,1],
shape [3,R,R] - without ImageFolder / PIL / albumentations (no dataset and no network here).

Imported by worker.py in every spawned rank before the models are built, so it is also where the
launcher's run-wide settings take effect: LCGAN_SEED seeds torch (the reference never does, which is
fine for training and useless for a parity test), LCGAN_NO_TF32=1 switches TF32 off for the reference's
cuDNN / cuBLAS calls so that its fp32 run is an fp32 oracle."""
import os

import torch
from torch.utils.data import Dataset

if os.environ.get("LCGAN_SEED"):
    torch.manual_seed(int(os.environ["LCGAN_SEED"]))
if os.environ.get("LCGAN_NO_TF32") == "1":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


class Dataset_(Dataset):
    def __init__(self, data_dir, resized_size, is_train):
        self.resized_size, self.is_train, self.n = resized_size, is_train, 4096

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(index)
        r = self.resized_size
        img = [torch.rand(3, r, r, generator=g) * 2 - 1 for _ in range(3)]
        return (img[0], img[1], img[2]) if self.is_train else (img[0], 0)
