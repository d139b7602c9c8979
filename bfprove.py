#!/usr/bin/env python3
"""bfprove — execute / prove / verify / size from the command line (see zkvm-brainfuck_b200/cli.py)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if __name__ == "__main__":
    sys.exit(importlib.import_module("zkvm-brainfuck_b200.cli").main())
