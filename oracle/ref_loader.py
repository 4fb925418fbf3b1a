"""Load the *real* keryil/kereHMM reference for fixture generation (test infrastructure only).

The reference under /root/reference is Python-2 source (``print`` statements,
``except X, e``), so it cannot be imported on Python 3 as it lies.  This module
applies a token-level, arithmetic-neutral rewrite

  * statement-level ``print <expr>``      ->  ``print(<expr>)``
  * ``except <Name>, <name>:``            ->  ``except <Name> as <name>:``

to an in-memory copy of every ``kerehmm/*.py`` file, writes the result to a
throw-away directory under the system temp dir (never into this repository --
reference sources are not copied into the repo) and imports it as package
``kerehmm``.  No arithmetic line is touched; see SURVEY.md section 8(c).

Only ``oracle/make_golden.py`` (run in the build container, where /root/reference
exists), optional CPU tests and the ``cpu_baseline`` leg of ``bench.py`` use this.

``materialize()`` (called by ``__graft_entry__.build()`` where /root/reference exists) writes the converted package to
``oracle/_ref/`` -- git-ignored like a compiled reference binary would be, so it never enters the history, but part of
the ``gpurun`` snapshot -- so that ``bench.py`` can time the REAL reference next to the NumPy port on the GPU box's own
host cores.  ``load_reference()`` falls back to that directory when /root/reference is absent.
"""
from __future__ import annotations

import importlib
import io
import os
import re
import sys
import tempfile
import tokenize

REFERENCE_ROOT = os.environ.get("KHMM_REFERENCE_ROOT", "/root/reference")

_SKIP = {tokenize.NL, tokenize.COMMENT, tokenize.INDENT, tokenize.DEDENT, tokenize.ENCODING}


def py2_to_py3(source: str) -> str:
    """Token-level print/except rewrite.  Returns Python-3 source."""
    toks = list(tokenize.generate_tokens(io.StringIO(source).readline))
    inserts = []  # (row, col, text)
    at_stmt_start = True
    i = 0
    while i < len(toks):
        tok = toks[i]
        if tok.type in _SKIP:
            i += 1
            continue
        if at_stmt_start and tok.type == tokenize.NAME and tok.string == "print":
            # find the end of the logical line
            j = i + 1
            while toks[j].type != tokenize.NEWLINE and toks[j].type != tokenize.ENDMARKER:
                j += 1
            nxt = toks[i + 1]
            already_call = nxt.type == tokenize.OP and nxt.string == "(" and toks[j - 1].string == ")"
            if not already_call:
                inserts.append((tok.end[0], tok.end[1], "("))
                last = toks[j - 1]
                # a trailing comma ("print x,") means "no newline" in py2
                if last.type == tokenize.OP and last.string == ",":
                    inserts.append((last.start[0], last.start[1], ""))  # keep; harmless tuple-ish
                inserts.append((last.end[0], last.end[1], ")"))
            i = j
            continue
        at_stmt_start = tok.type == tokenize.NEWLINE or (tok.type == tokenize.OP and tok.string in (":", ";"))
        i += 1
    lines = source.splitlines(keepends=True)
    for row, col, text in sorted(inserts, reverse=True):
        line = lines[row - 1]
        lines[row - 1] = line[:col] + text + line[col:]
    out = "".join(lines)
    out = re.sub(r"except\s+(\w+)\s*,\s*(\w+)\s*:", r"except \1 as \2:", out)
    return out


_loaded = None
MATERIALIZED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _convert_into(pkg_src: str, dst: str):
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(pkg_src)):
        if not name.endswith(".py"):
            continue
        with open(os.path.join(pkg_src, name)) as f:
            src = f.read()
        with open(os.path.join(dst, name), "w") as f:
            f.write(py2_to_py3(src))


def materialize(root: str | None = None, dest_root: str | None = None) -> str | None:
    """Write the converted reference package to oracle/_ref/kerehmm (build step; no-op without /root/reference)."""
    pkg_src = os.path.join(root or REFERENCE_ROOT, "kerehmm")
    if not os.path.isdir(pkg_src):
        return None
    dest_root = dest_root or MATERIALIZED
    _convert_into(pkg_src, os.path.join(dest_root, "kerehmm"))
    return dest_root


def load_reference(root: str | None = None):
    """Import the converted reference; returns the ``kerehmm`` package module.

    Raises FileNotFoundError when the reference tree is not present (e.g. on the
    GPU box)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = root or REFERENCE_ROOT
    pkg_src = os.path.join(root, "kerehmm")
    if os.path.isdir(pkg_src):
        dst_root = tempfile.mkdtemp(prefix="khmm_ref_py3_")
        _convert_into(pkg_src, os.path.join(dst_root, "kerehmm"))
    elif os.path.isdir(os.path.join(MATERIALIZED, "kerehmm")):
        dst_root = MATERIALIZED          # the GPU box: the copy that build() wrote before the snapshot was taken
    else:
        raise FileNotFoundError(pkg_src)
    sys.path.insert(0, dst_root)
    for mod in [m for m in sys.modules if m == "kerehmm" or m.startswith("kerehmm.")]:
        del sys.modules[mod]
    pkg = importlib.import_module("kerehmm")
    for sub in ("util", "distribution", "abstract_hmm", "discrete_hmm", "gaussian_hmm"):
        importlib.import_module("kerehmm." + sub)
    _loaded = pkg
    return pkg


def reference_available(root: str | None = None) -> bool:
    return os.path.isdir(os.path.join(root or REFERENCE_ROOT, "kerehmm")) or os.path.isdir(os.path.join(MATERIALIZED, "kerehmm"))
