#!/usr/bin/env python3
"""Drop-in entry point: same command line as the reference's PolyFastA.py, computed on the GPU.
Everything lives in polyfasta_b200/cli.py."""
import sys

from polyfasta_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
