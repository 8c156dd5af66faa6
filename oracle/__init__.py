"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy / torch-CPU, fp64 unless a function says otherwise) of the reference's
eigenvalue-analysis hot path (`analysis/eval_eig.py` and the layers in `models/` it drives).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this package, and only as the checker or the timed CPU baseline -- never as the product path.
The product (`eigb200`) must never import `oracle`; `tests/test_cabi_cpu.py::test_product_never_imports_oracle` enforces it.

Pinning status (see DESIGN.md "Oracle"):
  * extractors (Mamba-2, LTI, linear / normalised / softmax attention), threshold statistics, DPLR discretisation, LRU/S5 eigenvalues,
    linear / normalised / softmax attention layer forwards with the none / mlp / glu / hybrid mixers, Mamba block glue: PINNED against the
    reference's own functions/classes, executed
    in the authoring container by `tests/golden/make_golden.py` (AST-extraction / importlib of the files
    under /root/reference; vectors committed under `tests/golden/`).
  * SSD recurrence behind `mamba_chunk_scan_combined` (third-party mamba-ssm==2.1.0, not in the
    reference tree, not installed) and the XLA `associative_scan` (jax==0.4.25, not installed):
    PARITY UNPINNED by the reference itself; restated from the published recurrences and cross-checked
    against two independent implementations that are in the image (HF transformers Mamba2 torch path,
    fla naive simple-GLA) by `tests/golden/make_golden.py`.
  * `SSD_LTI.forward` (`pseudoLTI`): PARITY UNPINNED -- the reference class does not construct with current torch
    (`nn.Parameter(A, device=...)` raises), so `ssd_lti_mixer_forward` follows the source text only.
"""
from .ref_port import *  # noqa: F401,F403
