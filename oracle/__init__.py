"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the CLC latent hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import anything from this package.  The product (clc_b200/) never does and has no CPU fallback.

Contents
  shim/            minimal restatement of the CompressAI + timm surface the reference imports, so
                   /root/reference/models/{CLC_run,tcm}.py run UNMODIFIED here (and so the
                   entropy-model arithmetic has a pure-PyTorch oracle that travels to the GPU box).
  clc_oracle.py    standalone restatement (plain PyTorch fp32/fp64 on CPU) of every hot-path
                   reference function, each citing the reference file:line it follows.
  ref_loader.py    loads the reference's OWN functions from /root/reference (this container only)
                   to pin the restatements and to generate tests/golden/*.
  make_golden.py   the script that generated tests/golden/* from the reference itself.

PARITY STATUS: the reference holds no tests / golden vectors, and CompressAI
is an un-vendored, un-pinned dependency -> the CompressAI boundary is "parity unpinned".  What
IS pinned: (i) the in-tree Gaussian-likelihood restatement CLC_run.py:718-736, (ii) ste_round,
get_scale_table, the bpp formula, (iii) the Patch_Matching / CLM functions and the full
CLC.forward, all executed from the reference's own source by make_golden.py (through the shim
for the CompressAI layers) with outputs committed under tests/golden/.
"""
import os
import sys

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def enable_shim():
    """Put the CompressAI/timm shim on sys.path (idempotent)."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
