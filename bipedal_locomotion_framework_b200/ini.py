"""Reader for the reference's on-disk configuration format (YARP `.ini` text, as loaded by
src/ParametersHandler/YarpImplementation/src/YarpImplementation.cpp:115-144; fixture
src/ParametersHandler/tests/config.ini:1-9) into a StdImplementation, and the per-contact parameter
table built from it.  Python mirror of cpp/.../IniFile.h -- same
grammar, same strict typing; host logic only.
"""
from __future__ import annotations

import sys

from .contact_models import StdImplementation

PARAM_KEYS = ("length", "width", "spring_coeff", "damper_coeff")


def _classify(word: str, quoted: bool):
    if quoted or word == "":
        return word
    if word in ("true", "false"):
        return word == "true"
    try:
        v = int(word, 10)
        if -2 ** 31 <= v < 2 ** 31:
            return v
    except ValueError:
        pass
    try:
        return float(word)
    except ValueError:
        return word


def _tokenize(line: str):
    """-> (tokens, saw_list) or None on an unterminated quote / unbalanced or nested parenthesis."""
    out, depth, saw_list, p = [], 0, False, 0
    blanks = " \t\r,"
    while p < len(line):
        c = line[p]
        if c in blanks:
            p += 1
        elif c == "#" or line.startswith("//", p):
            break
        elif c == "(":
            depth += 1
            if depth > 1:
                return None
            saw_list = True
            p += 1
        elif c == ")":
            depth -= 1
            if depth < 0:
                return None
            p += 1
        elif c == '"':
            close = line.find('"', p + 1)
            if close < 0:
                return None
            out.append(_classify(line[p + 1:close], True))
            p = close + 1
        else:
            q = p
            while q < len(line) and line[q] not in blanks and line[q] not in '()"':
                q += 1
            out.append(_classify(line[p:q], False))
            p = q
    return (out, saw_list) if depth == 0 else None


def _store(handler, key, values, is_list):
    if not is_list and len(values) == 1:
        handler.setParameter(key, values[0])
        return
    kinds = {type(v) for v in values}
    if kinds <= {int}:
        handler.setParameter(key, [int(v) for v in values])
    elif kinds <= {int, float}:
        handler.setParameter(key, [float(v) for v in values])
    elif kinds <= {bool}:
        handler.setParameter(key, list(values))
    else:
        handler.setParameter(key, [v if isinstance(v, str) else str(v) for v in values])


def load_ini_string(text: str, handler: StdImplementation | None = None):
    """Returns the filled handler, or None (and a message on stderr) on a malformed line."""
    handler = handler if handler is not None else StdImplementation()
    current = handler
    for number, line in enumerate(text.splitlines(), 1):
        s = line.strip(" \t\r,")
        if not s:
            continue
        if s[0] == "[":
            close = s.find("]")
            name = s[1:close].strip() if close > 0 else ""
            if not name:
                print(f"[loadIniString] Malformed group header at line {number}.", file=sys.stderr)
                return None
            current = StdImplementation()
            handler.setGroup(name, current)
            continue
        tk = _tokenize(line)
        if tk is None:
            print(f"[loadIniString] Unterminated quote or unbalanced parenthesis at line {number}.",
                  file=sys.stderr)
            return None
        tokens, saw_list = tk
        if not tokens:
            continue
        if len(tokens) == 1 and not saw_list:
            print(f"[loadIniString] The key {tokens[0]} has no value (line {number}).", file=sys.stderr)
            return None
        key, values = str(tokens[0]), tokens[1:]
        _store(current, key, values, saw_list or len(values) > 1)
    return handler


def load_ini_file(path: str, handler: StdImplementation | None = None):
    try:
        with open(path) as f:
            return load_ini_string(f.read(), handler)
    except OSError:
        print(f"[loadIniFile] Unable to open {path}.", file=sys.stderr)
        return None


def parameter_table(handler):
    """The four equally long float lists of a per-contact table -> (4, n) numpy array, or None."""
    import numpy as np
    if handler is None:
        print("[loadParameterTable] The parameter handler is corrupted.", file=sys.stderr)
        return None
    cols = []
    for key in PARAM_KEYS:
        ok, v = handler.getParameter(key, list)
        if not ok or not all(type(x) is float for x in v):
            print(f"[loadParameterTable] Unable to get the vector named {key}.", file=sys.stderr)
            return None
        if cols and len(v) != len(cols[0]):
            print(f"[loadParameterTable] The vector named {key} has {len(v)} elements, "
                  f"{PARAM_KEYS[0]} has {len(cols[0])}.", file=sys.stderr)
            return None
        cols.append(v)
    return np.ascontiguousarray(np.array(cols, dtype=np.float64).reshape(4, -1))
