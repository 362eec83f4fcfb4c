# Padded batch of molecules (mirrors the reference's examples/batch.py): numbers == 0 is padding.
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tad_dftd4_b200 as d4  # noqa: E402

dev = torch.device("cuda:0")
numbers = d4.pack(
    (
        torch.tensor([6, 6, 7, 7, 1, 1, 1, 1, 1, 1, 8, 8]),  # formamide dimer
        torch.tensor([6, 8, 7, 1, 1, 1]),  # formamide
    )
).to(dev)
positions = d4.pack(
    (
        torch.tensor(
            [
                [-3.81469488143921, +0.09993441402912, 0.0], [+3.81469488143921, -0.09993441402912, 0.0],
                [-2.66030049324036, -2.15898251533508, 0.0], [+2.66030049324036, +2.15898251533508, 0.0],
                [-0.73178529739380, -2.28237795829773, 0.0], [-5.89039325714111, -0.02589114569128, 0.0],
                [-3.71254944801331, -3.73605775833130, 0.0], [+3.71254944801331, +3.73605775833130, 0.0],
                [+0.73178529739380, +2.28237795829773, 0.0], [+5.89039325714111, +0.02589114569128, 0.0],
                [-2.74426102638245, +2.16115570068359, 0.0], [+2.74426102638245, -2.16115570068359, 0.0],
            ],
            dtype=torch.float64,
        ),
        torch.tensor(
            [
                [-0.55569743203406, +1.09030425468557, 0.0], [+0.51473634678469, +3.15152550263611, 0.0],
                [+0.59869690244446, -1.16861263789477, 0.0], [-0.45355203669134, -2.74568780438064, 0.0],
                [+2.52721209544999, -1.29200800956867, 0.0], [-2.63139587595376, +0.96447869452240, 0.0],
            ],
            dtype=torch.float64,
        ),
    )
).to(dev)  # fmt: skip
q = torch.zeros(numbers.shape, dtype=torch.float64, device=dev)
param = d4.get_params(method="d4", functional="tpssh")
energy = torch.sum(d4.dftd4(numbers, positions, torch.zeros(2), param, q=q), -1)
print(energy, "interaction:", energy[0] - 2 * energy[1])
