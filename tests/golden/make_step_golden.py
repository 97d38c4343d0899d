"""Record the training-shell log of the UNMODIFIED reference Tri_CLIP (build container only: needs /root/reference).

    python tests/golden/make_step_golden.py

Writes tests/golden/step_shell.pt: the initial state dict of the tiny model and the losses the reference logs over five
micro-batches (accumulation 2: two optimizer steps + the left-over step) and two evaluation batches, CPU fp32."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from tests import step_shell  # noqa: E402


def main():
    import transformers

    saved = {}

    def setattr_keep(obj, name, value):
        saved[(obj, name)] = getattr(obj, name)
        setattr(obj, name, value)

    step_shell.patch_tiny_encoders(setattr_keep)
    ref = ref_import.load_reference_model()
    torch.manual_seed(1234)
    model = ref.Tri_CLIP(step_shell.tiny_config())
    init = {k: v.clone() for k, v in model.state_dict().items()}
    train = step_shell.synthetic_batches(5, 12, seed=7)
    valid = step_shell.synthetic_batches(2, 12, seed=8)
    log = step_shell.run_shell(model, train, valid, torch.device("cpu"))
    out = os.path.join(ROOT, "tests", "golden", "step_shell.pt")
    log = {k: [list(map(float, row)) if isinstance(row, (tuple, list)) else float(row) for row in v] for k, v in log.items()}
    torch.save({"init": init, "log": log, "torch": str(torch.__version__), "transformers": str(transformers.__version__)},
               out)
    print(out, os.path.getsize(out), "bytes")
    for k in ("train", "valid", "scales"):
        print(k, log[k])


if __name__ == "__main__":
    main()
