"""TEST INFRASTRUCTURE ONLY -- minimal restatement of the CompressAI surface the reference imports.

CompressAI (InterDigitalInc/CompressAI, BSD-3-Clause-Clear) is an un-vendored,
un-pinned dependency of the reference (README.md:41,:70; imports at
models/CLC_run.py:1-11).  It is not installed in this image and there is no
network, so this package restates -- from the published algorithm -- exactly the
classes the reference touches, in plain PyTorch, so that
``/root/reference/models/{CLC_run,tcm}.py`` import *unmodified*.

PARITY UNPINNED: the reference holds no tests or golden vectors at this
boundary; the only in-tree pin is the Gaussian likelihood
restatement at CLC_run.py:718-736, which tests/test_oracle_golden.py checks this
shim against.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
may import this package.  The product (clc_b200/) never does.
"""
__version__ = "0+clc_b200.oracle.shim"
