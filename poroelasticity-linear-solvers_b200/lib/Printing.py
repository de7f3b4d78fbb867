"""parprint: print on rank 0 only (reference lib/Printing.py:4-6).  With PORO_LOG_STDERR=1 the log lines go to
stderr (bench.py sets it so that its stdout carries only the JSON line)."""
import os
import sys


def parprint(*args, **kwargs):
    if int(os.environ.get("RANK", "0")) == 0:
        out = sys.stderr if os.environ.get("PORO_LOG_STDERR") else sys.stdout
        print(*args, **kwargs, file=out, flush=True)
