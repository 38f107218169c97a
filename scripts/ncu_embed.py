"""ncu target: the generator-input embedding kernel at the C2 batch's packed shape (64 x 512 tokens, t5-base sizes).

    ncu --set full --clock-control none --import-source on -k regex:vt5_embed -o /tmp/r2_embed python scripts/ncu_embed.py
"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


class Ctx:
    dev = torch.device("cuda:0")


out = bench.embed_leg(Ctx, 64, 512, 6523.7, with_cpu=False)
print({k: v for k, v in out.items() if k != "what"})
