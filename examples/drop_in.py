# Unchanged tad_dftd4 user code on the B200 kernels: install() routes `import tad_dftd4` to this
# package (or rebinds the entry points of an importable reference), device=... moves the CPU tensors
# such scripts build to the GPU and the results back.  With the reference tree at hand:
#
#     python examples/drop_in.py /root/reference/examples/single.py
#
# runs that script as it is (it needs `tad_mctc` for symbol_to_number / pack: any installation of it,
# or the test shim under oracle/mctc_shim, on PYTHONPATH).
import runpy
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tad_dftd4_b200  # noqa: E402

how = tad_dftd4_b200.install(device="cuda:0")
print(f"tad_dftd4 -> tad_dftd4_b200 ({how})", file=sys.stderr)
if len(sys.argv) > 1:
    runpy.run_path(sys.argv[1], run_name="__main__")
else:
    import torch

    import tad_dftd4 as d4  # the alias

    numbers = torch.tensor([3, 1])
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 3.0157]], dtype=torch.float64)
    param = d4.get_params(method="d4", functional="tpssh")
    pos = positions.clone().requires_grad_(True)
    energy = d4.dftd4(numbers, pos, torch.tensor(0.0), param)  # q=None: EEQ charges on device
    (grad,) = torch.autograd.grad(energy.sum(), pos)
    print(energy, grad, sep="\n")
