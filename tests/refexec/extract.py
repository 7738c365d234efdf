"""Pull pieces of source text out of the reference's HTML page (read at run time, never copied)."""
from __future__ import annotations

import re


def read_page(path="/root/reference/pages/airfoil_flow_lbm_aerolab.html") -> str:
    with open(path, "r") as fh:
        return fh.read()


def _match_brace(text: str, start: int) -> int:
    depth = 0
    i = start
    while i < len(text):
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced braces")


def js_function(page: str, name: str) -> str:
    m = re.search(r"function\s+" + re.escape(name) + r"\s*\(", page)
    if not m:
        raise KeyError(name)
    return page[m.start():_match_brace(page, page.index("{", m.end()))]


def js_statement(page: str, prefix: str) -> str:
    """The `const ...;` / `let ...;` statement that starts with `prefix` (up to the matching ';'
    outside brackets)."""
    i = page.index(prefix)
    depth = 0
    j = i
    while j < len(page):
        c = page[j]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        elif c == ";" and depth == 0:
            return page[i:j + 1]
        j += 1
    raise ValueError(prefix)


def shader(page: str, name: str) -> str:
    """Body of the template literal  const NAME=`...`;  without the #version line."""
    m = re.search(r"const\s+" + re.escape(name) + r"\s*=\s*`", page)
    end = page.index("`", m.end())
    src = page[m.end():end]
    return "\n".join(line for line in src.splitlines() if not line.lstrip().startswith("#"))
