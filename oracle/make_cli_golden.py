"""Golden record of the REFERENCE'S OWN command-line schema (run in this container only; needs /root/reference).

    python oracle/make_cli_golden.py            # writes tests/golden/ref_cli_args.json

/root/reference/twoDsim.py builds its argparse parser inside ``if __name__ == '__main__'`` (lines 68-99) next to imports of
plotting code that do not exist here, so the statements that build the parser (and the n_theta == 18 seed rule) are taken
from the file's syntax tree and executed alone, with the reference's own ``parser_help`` module, on a list of command lines.
``tests/test_cli.py`` parses the same command lines with ``mcmc-multispde_b200/twoDsim.py`` and compares every option.

Test infrastructure only."""
from __future__ import annotations

import ast
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_cli_args.json")

COMMAND_LINES = [
    [],
    ["--n-layers", "3", "--n", "32", "--n-theta", "90", "--n-samples", "1000", "--beta", "0.3"],
    ["--n-theta", "18"],                                     # the reference switches the seed for this angle count
    ["--n-theta", "18", "--seed", "7"],
    ["--seq-no", "4", "--init-folder", "result-01-Jan-2020_00_00_00", "--iter", "3"],
    ["--kappa", "1e17", "--sigma-0", "5e6", "--sigma-v", "10", "--sigma-scaling", "1e-3", "--meas-std", "0.1",
     "--chol-epsilon", "1e-4", "--burn-percentage", "5", "--evaluation-interval", "5", "--n-extended", "128"],
    ["--variant", "sigma", "--phantom-name", "blobs.png", "--meas-type", "denoising"],
    ["--naive"], ["--naive", "true"], ["--naive", "false"], ["--naive", "True"], ["--naive", "FALSE"], ["--nonaive"],
    ["--verbose", "false"], ["--noverbose"], ["--hybrid"], ["--nohybrid"],
    ["--enable-step-adaptation", "false"], ["--enable-step-adaptation"], ["--print-progress", "false"],
    ["--adapt-sqrtbeta"], ["--adapt-sqrtbeta", "true", "--enable-step-adaptation", "false", "--print-progress", "false"],
]
REJECTED = [["--naive", "maybe"], ["--n", "x"], ["--unknown-flag"], ["--naive", "--nonaive"]]


def reference_parser_source():
    """The statements of twoDsim.py's main block up to (and including) the seed rule that follows parse_args."""
    tree = ast.parse(open(os.path.join(REF, "twoDsim.py")).read())
    main_if = [n for n in tree.body if isinstance(n, ast.If) and "__main__" in ast.dump(n.test)][-1]
    keep = []
    for node in main_if.body:
        keep.append(node)
        if isinstance(node, ast.If) and "n_theta" in ast.dump(node.test):
            break
    mod = ast.Module(body=keep, type_ignores=[])
    return compile(ast.fix_missing_locations(mod), "twoDsim.py[main block]", "exec")


def main():
    import argparse
    spec = importlib.util.spec_from_file_location("ref_parser_help", os.path.join(REF, "parser_help.py"))
    ph = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ph)
    code = reference_parser_source()
    out = {"source": "/root/reference/twoDsim.py:68-99 + parser_help.py, executed from the syntax tree", "accepted": [], "rejected": []}
    argv0 = sys.argv
    for argv in COMMAND_LINES + REJECTED:
        env = {"argparse": argparse, "ph": ph}
        sys.argv = ["twoDsim.py"] + argv
        try:
            devnull = open(os.devnull, "w")
            stderr0, sys.stderr = sys.stderr, devnull
            try:
                exec(code, env)
            finally:
                sys.stderr = stderr0
                devnull.close()
        except SystemExit:
            out["rejected"].append(argv)
            continue
        ns = vars(env["args"])
        out["accepted"].append({"argv": argv, "args": {k: ns[k] for k in sorted(ns)}})
    sys.argv = argv0
    assert [r for r in out["rejected"]] == REJECTED, out["rejected"]
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print(len(out["accepted"]), "accepted,", len(out["rejected"]), "rejected ->", OUT)
    print(out["accepted"][0]["args"])


if __name__ == "__main__":
    main()
