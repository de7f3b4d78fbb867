"""parprint: print on rank 0 only (reference lib/Printing.py:4-6)."""
import os


def parprint(*args, **kwargs):
    if int(os.environ.get("RANK", "0")) == 0:
        print(*args, **kwargs, flush=True)
