"""TEST INFRASTRUCTURE ONLY -- stub of the two timm symbols the reference imports (CLC_run.py:20)."""
