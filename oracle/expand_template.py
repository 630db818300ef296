#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): expand one of the reference's design templates.

The reference generates an NxN (or MxN) design from `<design>/template/T_*.{h,cpp}` with an
in-house tool called `src_config` (see `<design>/common/runit.csh` and `common/T_gen.cpp`,
e.g. /root/reference/Cholesky/cholesky_v1.3/common/T_gen.cpp:1-12).  That tool is not in the
repository, but its template language is evident from the templates themselves:

  * a line whose first non-blank characters are `//;` is a line of Perl;
  * every other line is emitted verbatim, except that `` `expr` `` is replaced by the value of
    the Perl expression `expr`;
  * `param_define("NAME", default)` yields the value configured in `<design>/common/*.cfg.xml`
    (here: passed on the command line), else the default.

This script turns a template into a Perl program and runs it.  It reads the template where it
lies under /root/reference and writes the expansion wherever it is told to (oracle/_ref/...,
never into the tracked tree), so no reference source is copied into this repository.
"""
import re
import subprocess
import sys


def perl_quote(text: str) -> str:
    return "'" + text.replace("\\", "\\\\").replace("'", "\\'") + "'"


def template_to_perl(template: str, params: dict) -> str:
    out = ["my %P = (" + ", ".join(f"{perl_quote(k)} => {int(v)}" for k, v in params.items()) + ");",
           "sub param_define { my ($n, $d) = @_; return exists $P{$n} ? $P{$n} : $d; }"]
    for line in template.splitlines():
        m = re.match(r"^\s*//;(.*)$", line)
        if m:
            out.append(m.group(1))
            continue
        pieces = line.split("`")
        if len(pieces) % 2 == 0:
            raise ValueError("unbalanced backquote in template line: " + line)
        terms = []
        for idx, piece in enumerate(pieces):
            if idx % 2 == 0:
                if piece:
                    terms.append(perl_quote(piece))
            else:
                terms.append("(" + piece + ")")
        terms.append(perl_quote("\n"))
        out.append("print " + " . ".join(terms) + ";")
    return "\n".join(out) + "\n"


def expand(template_path: str, out_path: str, params: dict) -> None:
    with open(template_path, "r", errors="replace") as fh:
        program = template_to_perl(fh.read(), params)
    res = subprocess.run(["perl", "-e", program], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"perl failed on {template_path}: {res.stderr}")
    with open(out_path, "w") as fh:
        fh.write(res.stdout)


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit("usage: expand_template.py TEMPLATE OUT [NAME=VALUE ...]")
    kv = dict(a.split("=", 1) for a in sys.argv[3:])
    expand(sys.argv[1], sys.argv[2], kv)
