"""Parse the public Fortran surface the north-star's primary seam has to keep:

  * every `public ::` name of the reference modules the shim replaces, with the dummy-argument names of
    the public procedures (the drivers pass keyword arguments: nf=, magic=, step=, foldername=);
  * the components of `type lattice_grid` (src/fvm_bardow.F90:37-69) and its type-bound procedures;
  * what the two shipped drivers use: `use` lines, `grid%component` references, `call proc(grid...)`.

`python tools/ref_fortran_surface.py` (in the development container, where /root/reference exists) writes
tests/golden/ref_fortran_surface.json; tests/test_fortran_shim_surface.py checks the shim against the live
reference when it is present and against that snapshot otherwise.  The same parser reads the shim."""
from __future__ import annotations

import json
import os
import re
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODULE_FILES = [
    "src/precision.F90", "src/fvm_bardow.F90", "src/periodic_lbm.f90", "src/collision_bgk.F90", "src/collision_trt.F90",
    "src/collision_regularized.F90", "src/collision_bgk_improved.f90", "src/periodic_dugks.F90", "src/vorticity.f90",
]
REF_DRIVERS = ["app/main_taylor_green.f90", "app/main_vortex.f90"]


def logical_lines(text):
    """Fortran free-form source -> list of logical lines: comments stripped, continuations joined, lower case.
    cpp lines are dropped (both branches of an #if are kept: a name has to exist in every build)."""
    out, cur = [], ""
    for raw in text.splitlines():
        if raw.lstrip().startswith("#"):
            continue
        line, quote = "", None
        for ch in raw:  # strip a trailing comment, respecting character literals
            if quote:
                if ch == quote:
                    quote = None
            elif ch in "'\"":
                quote = ch
            elif ch == "!":
                break
            line += ch
        line = line.strip()
        if not line:
            continue
        if cur:
            line = line[1:].lstrip() if line.startswith("&") else line
            cur += " " + line
        else:
            cur = line
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        out.append(cur.lower())
        cur = ""
    if cur:
        out.append(cur.lower())
    return out


PROC_RE = re.compile(r"^(?:(?:pure|elemental|impure|recursive|module|real\s*\([^)]*\)|integer(?:\s*\([^)]*\))?|logical)\s+)*"
                     r"(subroutine|function)\s+(\w+)\s*(?:\(([^)]*)\))?")


def parse_modules(text):
    """{module: {"public": [names], "public_all": bool, "procs": {name: [dummy args]}, "types": {type: {"components": [...], "bound": [...]}}}}"""
    mods, cur, in_type, in_interface = {}, None, None, 0
    for ln in logical_lines(text):
        m = re.match(r"^module\s+(\w+)\s*$", ln)
        if m and m.group(1) != "procedure":
            cur = mods.setdefault(m.group(1), {"public": [], "public_all": False, "procs": {}, "types": {}, "uses": {}, "decls": []})
            in_type, in_interface = None, 0
            continue
        if cur is None:
            continue
        if re.match(r"^end\s*module", ln):
            cur = None
            continue
        if re.match(r"^(abstract\s+)?interface\b", ln):
            in_interface += 1
            continue
        if re.match(r"^end\s*interface", ln):
            in_interface = max(0, in_interface - 1)
            continue
        if ln == "public":
            cur["public_all"] = True
            continue
        m = re.match(r"^use\s*(?:,\s*intrinsic\s*)?(?:::)?\s*(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", ln)
        if m and in_type is None:
            only = None if m.group(2) is None else [n.split("=>")[0].strip() for n in m.group(2).split(",") if n.strip()]
            prev = cur["uses"].get(m.group(1), [])
            cur["uses"][m.group(1)] = None if only is None or prev is None else prev + only
            continue
        m = re.match(r"^public\s*::\s*(.*)$", ln)
        if m:
            cur["public"] += [n.strip() for n in m.group(1).split(",") if n.strip()]
            continue
        m = re.match(r"^type\s*(?:,[^:]*)?::\s*(\w+)\s*$", ln) or re.match(r"^type\s+(\w+)\s*$", ln)
        if m and in_type is None:
            in_type = m.group(1)
            cur["types"][in_type] = {"components": [], "bound": []}
            bound_part = False
            continue
        if in_type is not None:
            if re.match(r"^end\s*type", ln):
                in_type = None
                continue
            if ln == "contains":
                bound_part = True
                continue
            if "::" in ln:
                names = [re.match(r"\s*(\w+)", n).group(1) for n in split_top(ln.split("::", 1)[1])]
                cur["types"][in_type]["bound" if bound_part else "components"] += names
            continue
        if in_interface:
            continue
        if "::" in ln and not PROC_RE.match(ln) and not cur.get("_contains"):
            cur["decls"] += [re.match(r"\s*(\w+)", n).group(1) for n in split_top(ln.split("::", 1)[1]) if re.match(r"\s*([a-z_]\w*)", n)]
            continue
        if ln == "contains":
            cur["_contains"] = True
            continue
        m = PROC_RE.match(ln)
        if m and not ln.startswith("end"):
            args = [a.strip() for a in (m.group(3) or "").split(",") if a.strip()]
            cur["procs"][m.group(2)] = args
    return mods


def split_top(s):
    """split on commas that are not inside parentheses"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def parse_driver(text):
    uses, comps, calls, bound_calls, kwargs = {}, set(), set(), set(), {}
    for ln in logical_lines(text):
        m = re.match(r"^use\s+(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", ln)
        if m:
            names = [n.strip() for n in (m.group(2) or "").split(",") if n.strip()]
            uses.setdefault(m.group(1), [])
            uses[m.group(1)] += names
            continue
        for c in re.findall(r"\bgrid%(\w+)", ln):
            comps.add(c)
        m = re.match(r"^(?:if\s*\(.*\)\s*)?call\s+grid%(\w+)\s*\((.*)\)\s*$", ln)
        if m:
            bound_calls.add(m.group(1))
            kwargs.setdefault("grid%" + m.group(1), set()).update(re.findall(r"(\w+)\s*=(?!=)", m.group(2)))
            continue
        m = re.match(r"^(?:if\s*\(.*\)\s*)?call\s+(\w+)\s*\((.*)\)\s*$", ln)
        if m and re.search(r"\bgrid\b", m.group(2)):
            calls.add(m.group(1))
            kwargs.setdefault(m.group(1), set()).update(k for k in re.findall(r"(\w+)\s*=(?!=)", m.group(2)))
    return {"uses": {k: sorted(set(v)) for k, v in uses.items()}, "grid_components": sorted(comps), "calls_with_grid": sorted(calls),
            "type_bound_calls": sorted(bound_calls), "keyword_args": {k: sorted(v) for k, v in kwargs.items() if v}}


def reference_surface(ref=REF):
    mods = {}
    for rel in REF_MODULE_FILES:
        with open(os.path.join(ref, rel)) as fh:
            for name, m in parse_modules(fh.read()).items():
                pub = m["public"]
                mods[name] = {"file": rel, "public": sorted(set(pub)),
                              "procs": {p: a for p, a in m["procs"].items() if p in pub or any(p in t["bound"] for t in m["types"].values())},
                              "types": {t: v for t, v in m["types"].items() if t in pub}}
    drivers = {}
    for rel in REF_DRIVERS:
        with open(os.path.join(ref, rel)) as fh:
            drivers[rel] = parse_driver(fh.read())
    return {"modules": mods, "drivers": drivers}


def shim_surface(shim_dir=None):
    shim_dir = shim_dir or os.path.join(ROOT, "periodic_lbm_b200", "fortran")
    mods = {}
    for fn in sorted(os.listdir(shim_dir)):
        if fn.lower().endswith((".f90",)):
            with open(os.path.join(shim_dir, fn)) as fh:
                for name, m in parse_modules(fh.read()).items():
                    m["file"] = fn
                    mods[name] = m
    return mods


if __name__ == "__main__":
    surf = reference_surface()
    out = os.path.join(ROOT, "tests", "golden", "ref_fortran_surface.json")
    with open(out, "w") as fh:
        json.dump(surf, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print(f"wrote {out}: {sum(len(m['public']) for m in surf['modules'].values())} public names in {len(surf['modules'])} modules", file=sys.stderr)
