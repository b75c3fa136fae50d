"""Golden vectors for the input-side transform, produced by the composition the reference's dataset uses
(multiview_detector/datasets/frameDataset.py:66-67: T.ToTensor -> T.Normalize -> T.Resize) with the torchvision of the
build container, on small seeded uint8 images (PIL in, as the reference feeds it).
   python tests/golden/make_golden_preprocess.py"""
import os

import numpy as np
import torch
import torchvision.transforms as T
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"down_1p5": ((54, 96), (36, 64)),       # the reference's ratio: 1080x1920 -> 720x1280
         "down_odd": ((45, 70), (20, 27)),
         "down_3x": ((60, 90), (20, 30)),
         "up": ((10, 12), (25, 31)),
         "same": ((16, 20), (16, 20))}


def main():
    out = {}
    for name, (src, dst) in CASES.items():
        rng = np.random.RandomState(len(name) + src[0])
        img = rng.randint(0, 256, size=(*src, 3)).astype(np.uint8)
        tr = T.Compose([T.ToTensor(), T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), T.Resize(list(dst))])
        res = tr(Image.fromarray(img)).numpy()
        x = torch.from_numpy(img).permute(2, 0, 1).float().div(255)
        x = T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))(x)
        plain = torch.nn.functional.interpolate(x[None], size=dst, mode="bilinear", align_corners=False)[0].numpy()
        out[f"{name}.img"], out[f"{name}.out"], out[f"{name}.out_noaa"] = img, res, plain
        print(name, img.shape, res.shape)
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)


if __name__ == "__main__":
    main()
