"""TEST INFRASTRUCTURE ONLY -- access to the reference's OWN code (this container only;
/root/reference does not exist on the GPU box).  Nothing is copied: models are imported
unmodified through the shim, and the Patch_Matching functions (whose module cannot be imported:
`from turtle import shape`, `import pylab`, `compressai_local`, hard-coded `.cuda()`) are
exec'd from source slices read at run time with `.cuda()` stripped, as the survey did."""
import os
import sys
import types

from . import enable_shim

REFERENCE = os.environ.get("CLC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "models"))


def import_models():
    """-> (CLC, TCM) classes from /root/reference/models, unmodified."""
    enable_shim()
    if REFERENCE not in sys.path:
        sys.path.insert(1, REFERENCE)
    from models import CLC, TCM  # type: ignore
    return CLC, TCM


def _slice(path, ranges):
    lines = open(path).read().split("\n")
    return "\n".join("\n".join(lines[a - 1:b]) for a, b in ranges)


def load_patch_matching():
    """Namespace with SI_Finder_at_Decoder_Feature_Domain, SI_Wraper, create_gaussian_masks,
    L2_or_pearson_corr executed from the reference source (Patch_Matching.py:157-240, :779-807,
    :854-910), `.cuda()` removed."""
    src = _slice(os.path.join(REFERENCE, "models", "Patch_Matching.py"), [(157, 240), (779, 807), (854, 910)])
    src = src.replace(".cuda()", "")
    ns = types.ModuleType("ref_patch_matching")
    exec("import time\nimport numpy as np\nimport torch\nfrom torch import nn\nimport torch.nn.functional as F\n",
         ns.__dict__)
    exec(compile(src, "Patch_Matching.py[sliced]", "exec"), ns.__dict__)
    return ns


def load_clm():
    """The reference's models/CLM.py (importable as-is)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_clm", os.path.join(REFERENCE, "models", "CLM.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_rd_loss():
    """RateDistortionLoss executed from train_CLC.py:36-59."""
    src = _slice(os.path.join(REFERENCE, "train_CLC.py"), [(36, 59)])
    ns = {}
    exec("import math\nimport torch\nimport torch.nn as nn\n", ns)
    exec(compile(src, "train_CLC.py[36:59]", "exec"), ns)
    return ns["RateDistortionLoss"]
