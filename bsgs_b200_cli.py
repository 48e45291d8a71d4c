#!/usr/bin/env python3
"""Entry point of the command line (same flags as the reference, see bsgs-cuda_b200/cli.py)."""
import sys
from bsgs_cuda_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
