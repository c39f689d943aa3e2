"""Flattens a SPICE deck for the headless reference harness (oracle/refbuild/harness.cc parses
with IP.parseDeck only: subcircuit expansion and .param substitution live in the reference's
front end, which the harness does not build).  Handles what examples/JJexamples/nor_mc.cir needs:
`.param name=value`, arithmetic of parameters and suffixed numbers inside element values and
source arguments, `.subckt/.ends`, `X` instances (nested allowed).  Script blocks (.control /
.exec ... .endc), .stop, .save and patterned sources are dropped; elements inside an instance
are named the way WRspice names them: `<element>.<instance path>` (b0.x0), internal nodes
`<node>.<instance path>`.

    flatten(text, overrides={"b0ics": 150e-6}) -> flat deck text
"""
import re

SUFFIX = {"t": 1e12, "g": 1e9, "meg": 1e6, "k": 1e3, "m": 1e-3, "u": 1e-6, "n": 1e-9, "p": 1e-12, "f": 1e-15}
NUM = re.compile(r"(?<![\w.])(\d+\.?\d*(?:e[+-]?\d+)?|\.\d+(?:e[+-]?\d+)?)(meg|[tgkmunpf])?[a-z]*", re.I)


def _num(tok):
    m = NUM.fullmatch(tok)
    if not m:
        return None
    v = float(m.group(1))
    if m.group(2):
        v *= SUFFIX[m.group(2).lower()]
    return v


TOK = re.compile(r"\s*(?:(\d+\.?\d*(?:e[+-]?\d+)?|\.\d+(?:e[+-]?\d+)?)(meg|[tgkmunpf])?[a-z]*|([A-Za-z_]\w*)|([-+*/()]))", re.I)


def _eval(expr, params):
    """Number, parameter name, or arithmetic of both; None if it is something else."""
    v = _num(expr)
    if v is not None:
        return v
    pos, py, named = 0, [], False
    while pos < len(expr):
        m = TOK.match(expr, pos)
        if not m or m.end() == pos:
            return None
        if m.group(1) is not None:
            x = float(m.group(1))
            if m.group(2):
                x *= SUFFIX[m.group(2).lower()]
            py.append(repr(x))
        elif m.group(3) is not None:
            if m.group(3).lower() not in params:
                return None
            py.append(repr(params[m.group(3).lower()]))
            named = True
        else:
            py.append(m.group(4))
        pos = m.end()
    if not named:
        return None
    return float(eval(" ".join(py), {"__builtins__": {}}, {}))


def _subst(tok, params):
    """Substitute one whitespace-free token (possibly `key=value`)."""
    if "=" in tok:
        k, _, v = tok.partition("=")
        e = _eval(v, params)
        return f"{k}={e!r}" if e is not None and _num(v) is None else tok
    e = _eval(tok, params)
    return repr(e) if e is not None and _num(tok) is None else tok


_STRINGS = {}


def _subst_line(ln, params):
    # source arguments: name(args) groups are substituted argument by argument
    def grp(m):
        return m.group(1) + "(" + " ".join(_STRINGS.get(a.lower(), None) or _subst(a, params)
                                           for a in m.group(2).replace(",", " ").split()) + ")"
    ln = re.sub(r"(\b[a-z]+)\(([^()]*)\)", grp, ln, flags=re.I)
    toks = ln.split()
    return " ".join([toks[0]] + [t if "(" in t or ")" in t else _subst(t, params) for t in toks[1:]])


def parameters(text):
    """The numeric .param values of a deck."""
    params = {}
    for ln in text.splitlines():
        if ln.strip().lower().startswith(".param"):
            for k, v in re.findall(r"(\w+)\s*=\s*(\"[^\"]*\"|\S+)", ln.strip()[6:]):
                e = _eval(v, params)
                if e is not None:
                    params[k.lower()] = e
    return params


def flatten(text, overrides=None):
    lines = []
    for raw in text.splitlines():
        if raw.startswith("+") and lines:
            lines[-1] += " " + raw[1:].strip()
        else:
            lines.append(raw.rstrip())
    title, body = lines[0], lines[1:]
    params, subs, top = {}, {}, []
    strings = {}                # .param name="text": substituted verbatim (bit patterns of sources)
    cur = None
    skip = False
    for ln in body:
        s = ln.strip()
        low = s.lower()
        if not s or s.startswith("*"):
            continue
        if low.startswith((".control", ".exec")):
            skip = True
        if skip:
            if low.startswith(".endc"):
                skip = False
            continue
        if low.startswith((".stop", ".save", ".end ")) or low == ".end":
            continue
        if low.startswith(".param"):
            for k, v in re.findall(r"(\w+)\s*=\s*(\"[^\"]*\"|\S+)", s[6:]):
                e = _eval(v, params)
                if e is not None:
                    params[k.lower()] = e
                elif v.startswith('"'):
                    strings[k.lower()] = v.strip('"')
            continue
        if low.startswith(".subckt"):
            t = s.split()
            cur = (t[1].lower(), [x.lower() for x in t[2:]], [])
            continue
        if low.startswith(".ends"):
            subs[cur[0]] = cur
            cur = None
            continue
        (cur[2] if cur else top).append(s)
    for k, v in (overrides or {}).items():
        params[k.lower()] = float(v)
    out = []
    _STRINGS.clear(); _STRINGS.update(strings)

    def emit(elems, path, nodemap):
        for s in elems:
            t = s.split()
            name = t[0].lower()

            def node(nm):
                nm = nm.lower()
                if nm == "0" or not path:
                    return nm
                return nodemap.get(nm, f"{nm}.{path}")
            if name[0] == "x":
                sub = subs[t[-1].lower()]
                actual = [node(x) for x in t[1:-1]]
                sp = f"{name}.{path}" if path else name
                emit(sub[2], sp, dict(zip(sub[1], actual)))
                continue
            if name[0] == ".":
                out.append(_subst_line(s, params))
                continue
            nn = {"b": 3 if len(t) > 4 and "=" not in t[3] and not _is_model_at(t, 3) else 2}.get(name[0], 2)
            full = f"{name}.{path}" if path else name
            rest = t[1 + nn:]
            out.append(_subst_line(" ".join([full] + [node(x) for x in t[1:1 + nn]] + rest), params))

    def _is_model_at(t, i):
        return False

    emit(top, "", {})
    return "\n".join([title] + out) + "\n"
