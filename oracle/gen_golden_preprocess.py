"""Generate tests/golden/preprocess.npz with the UNMODIFIED reference crop function and the torchvision transform the
reference composes (src/datasets/ho3d.py:35-40, 139-147; src/datasets/utils.py:40-77).  Build container only.

The dataset class itself cannot be instantiated offline (webdataset shards, MANO assets), so the script applies
exactly what its __getitem__ applies to an image: `crop_and_pad_image(frame, bbox)` imported from the reference,
then `transforms.Compose([ToTensor(), Resize((256, 256), antialias=True), Normalize(mean, std)])` from torchvision.
"""
import os
import sys

import numpy as np
import torch
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/src")
import handmvnet_oracle as O  # noqa: E402
from datasets.utils import crop_and_pad_image  # noqa: E402  (the reference's own function)


def main():
    img_transform = transforms.Compose([                     # ho3d.py:35-40, verbatim arguments
        transforms.ToTensor(),
        transforms.Resize((256, 256), antialias=True),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])
    frames, bboxes = O.make_frames(8, seed=0)
    outs = torch.stack([img_transform(crop_and_pad_image(f, b)) for f, b in zip(frames, bboxes)])
    flat = outs.reshape(-1).double()
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, flat.numel(), (4096,), generator=g)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, shape=np.array(outs.shape), sub=outs[:, :, ::8, ::8].numpy(), idx=idx.numpy(),
                        val=flat[idx].float().numpy(), mean=np.float64(flat.mean()), std=np.float64(flat.std()),
                        bboxes=bboxes)
    print("wrote", path, tuple(outs.shape), "mean %.6f std %.6f" % (flat.mean(), flat.std()))
    mine = O.preprocess(frames, bboxes)
    print("oracle restatement vs reference pipeline: max abs diff", float((mine - outs).abs().max()))


if __name__ == "__main__":
    main()
