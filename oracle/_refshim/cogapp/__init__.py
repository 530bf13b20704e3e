"""Minimal stand-in for the `cogapp` package (not installed here, no network), used ONLY by
`oracle/make_golden.py` to let the reference expand its kernel templates.

Implements the subset the reference uses (lightplane/triton_src/__init__.py:266-307):
`Cog().main([None, "-o", out, "-d", "-D", "K=V", ..., template])` with `[[[cog ... ]]]` /
`[[[end]]]` blocks and `cog.outl` / `cog.out`.  Written from cog's documented behaviour:
generator lines lose their common prefix and are dedented; output is dedented and re-indented
to the marker's indentation; `-d` drops markers and generator source from the output.
"""

import os
import sys
import textwrap
import types


def _common_prefix(lines):
    lines = [l for l in lines]
    if not lines:
        return ""
    p = os.path.commonprefix(lines)
    return p


class Cog:
    def main(self, argv):
        args = list(argv[1:])
        out_path, defines, template = None, {}, None
        i = 0
        while i < len(args):
            a = args[i]
            if a == "-o":
                out_path = args[i + 1]
                i += 2
            elif a == "-D":
                k, v = args[i + 1].split("=", 1)
                defines[k] = v
                i += 2
            elif a == "-d":
                i += 1
            else:
                template = a
                i += 1
        with open(template) as f:
            src = f.read().split("\n")
        result = []
        glob = dict(defines)
        li = 0
        while li < len(src):
            line = src[li]
            if "[[[cog" not in line:
                result.append(line)
                li += 1
                continue
            marker = line
            code = []
            li += 1
            while "]]]" not in src[li]:
                code.append(src[li])
                li += 1
            end_gen = src[li]
            li += 1
            while "[[[end]]]" not in src[li]:
                li += 1  # drop stale output
            li += 1
            pref = _common_prefix([marker, end_gen] + code)
            # the prefix must not eat into the marker text itself
            pref = pref[: len(pref)] if "[[[" not in pref else pref[: pref.index("[[[")]
            body = "\n".join(c[len(pref):] if c.startswith(pref) else c.lstrip("# ") for c in code)
            body = textwrap.dedent(body)
            chunks = []
            cogmod = types.ModuleType("cog")
            cogmod.out = lambda s="": chunks.append(s)
            cogmod.outl = lambda s="": chunks.append(s + "\n")
            saved = sys.modules.get("cog")
            sys.modules["cog"] = cogmod
            try:
                exec(compile(body, template + ":cog", "exec"), glob)
            finally:
                if saved is not None:
                    sys.modules["cog"] = saved
                else:
                    sys.modules.pop("cog", None)
            text = "".join(chunks)
            if text:
                indent = marker[: len(marker) - len(marker.lstrip(" \t"))]
                text = textwrap.dedent(text)
                for ol in text.split("\n")[:-1] if text.endswith("\n") else text.split("\n"):
                    result.append((indent + ol) if ol.strip() else ol)
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, "w") as f:
            f.write("\n".join(result))
        return 0
