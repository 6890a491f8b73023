"""Diagnostic: tensor-pipe and TMA issue-rate microbenchmarks (b200_umma_selftest) -- prints cycles per MMA / per box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biapy_b200 import _lib

if __name__ == "__main__":
    torch.zeros(1, device="cuda")
    rc = _lib.lib().b200_umma_selftest(1, None)
    print("rc", rc, _lib.lib().b200_last_error() if rc else "")
