"""Options handling with the reference's grammar (lib/Parser.py:61-73): strip each line, skip
it if it contains '#' or is empty, key = first token, value = last token, bare flag -> None.
Values go to the device library's options database instead of PETSc.Options()."""
from __future__ import annotations

from optparse import OptionParser


def parse_petsc_options(text: str):
    out = []
    for _line in text.splitlines():
        line = _line.rstrip().lstrip()
        if "#" in line or len(line) == 0:
            continue
        split = line.split(" ")
        if len(split) > 1:
            out.append((split[0], split[-1]))
        else:
            out.append((line, None))
    return out


def load_petsc_options(ctx, path_or_text: str, is_text: bool = False):
    text = path_or_text if is_text else open(path_or_text).read()
    opts = parse_petsc_options(text)
    for k, v in opts:
        ctx.set_option(k, v)
    return opts


class Parser:
    """Same flags as the reference's Parser (lib/Parser.py:15-60)."""

    def __init__(self, argv=None, ctx=None):
        parser = OptionParser(add_help_option=False)
        parser.add_option("-h", "--help", action="help")
        parser.add_option("-N", "--Nelements", type="int", dest="N")
        parser.add_option("--N-refinements", type="int", dest="refinements")
        parser.add_option("--solver-type", type="str", dest="solver_type")
        parser.add_option("--pc-type", type="str", dest="pc_type")
        parser.add_option("--fe-solid", type="int", dest="fe_s")
        parser.add_option("--monitor", action="store_true", dest="monitor")
        parser.add_option("--inner-monitor", action="store_true", dest="inner_monitor")
        parser.add_option("--inner-accel-order", type="int", dest="inner_accel_order")
        parser.add_option("--output", action="store_true", dest="output")
        parser.add_option("--time-final", type="float", dest="tf")
        parser.add_option("--petsc-options", type="str", dest="options_file")
        options, _ = parser.parse_args(argv)
        d = {}
        for key, val in (("N", options.N), ("mesh refinements", options.refinements),
                         ("solver type", options.solver_type), ("pc type", options.pc_type),
                         ("fe degree solid", options.fe_s), ("inner accel order", options.inner_accel_order),
                         ("tf", options.tf)):
            if val:                      # truthiness test like the reference (drops zeros)
                d[key] = val
        if options.monitor:
            d["solver monitor"] = True
        if options.inner_monitor:
            d["inner monitor"] = True
        if options.output:
            d["output solutions"] = True
        if options.options_file:
            from .backend import get_context
            load_petsc_options(ctx or get_context(), options.options_file)
        self.options_dict = d
        self.options = options
