"""Last GPU seconds of a round: the ESM parity tests (all attention kernels, determinism) and the attention time of the
default kernel inside a real ESM2-650M encode, in ONE process (one torch import)."""
import os, sys
import pytest
rc = pytest.main(["tests/test_gpu_esm.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"])
print("pytest rc", int(rc), flush=True)
sys.argv = ["diag"]
os.environ["PCY_ESM_ATTN"] = "5,6"
sys.path.insert(0, "scripts")
import diag_attn_time
diag_attn_time.main()
